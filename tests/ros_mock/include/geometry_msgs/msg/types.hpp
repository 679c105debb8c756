#pragma once
#include <array>
#include <memory>
#include <string>
#include <vector>
#include "std_msgs/msg/header.hpp"
namespace geometry_msgs { namespace msg {
struct Point { double x = 0, y = 0, z = 0; };
struct Point32 { float x = 0, y = 0, z = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped {
  std_msgs::msg::Header header; Pose pose;
  using SharedPtr = std::shared_ptr<PoseStamped>; using ConstSharedPtr = std::shared_ptr<const PoseStamped>;
};
struct PoseWithCovariance { Pose pose; std::array<double, 36> covariance{}; };
struct PoseWithCovarianceStamped {
  std_msgs::msg::Header header; PoseWithCovariance pose;
  using SharedPtr = std::shared_ptr<PoseWithCovarianceStamped>;
  using ConstSharedPtr = std::shared_ptr<const PoseWithCovarianceStamped>;
};
struct PointStamped { std_msgs::msg::Header header; Point point; using SharedPtr = std::shared_ptr<PointStamped>; };
struct Polygon { std::vector<Point32> points; };
struct Transform { Vector3 translation; Quaternion rotation; };
struct TransformStamped { std_msgs::msg::Header header; std::string child_frame_id; Transform transform; };
} }
