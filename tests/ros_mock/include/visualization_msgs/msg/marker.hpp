#pragma once
#include <memory>
#include <string>
#include <vector>
#include "geometry_msgs/msg/types.hpp"
namespace rclcpp { class Duration; }
namespace visualization_msgs { namespace msg {
struct Marker {
  enum : int32_t { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6,
                   SPHERE_LIST = 7, POINTS = 8, TEXT_VIEW_FACING = 9, MESH_RESOURCE = 10, TRIANGLE_LIST = 11 };
  enum : int32_t { ADD = 0, MODIFY = 0, DELETE = 2, DELETEALL = 3 };
  std_msgs::msg::Header header; std::string ns; int32_t id = 0; int32_t type = 0; int32_t action = 0;
  geometry_msgs::msg::Pose pose; geometry_msgs::msg::Vector3 scale; std_msgs::msg::ColorRGBA color;
  builtin_interfaces::msg::Duration lifetime; bool frame_locked = false;
  std::vector<geometry_msgs::msg::Point> points; std::vector<std_msgs::msg::ColorRGBA> colors; std::string text;
  using SharedPtr = std::shared_ptr<Marker>;
};
struct MarkerArray { std::vector<Marker> markers; using SharedPtr = std::shared_ptr<MarkerArray>; };
} }
