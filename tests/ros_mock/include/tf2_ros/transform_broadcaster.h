#pragma once
#include <vector>
#include "geometry_msgs/msg/types.hpp"
#include "rclcpp/rclcpp.hpp"
namespace tf2_ros {
class TransformBroadcaster {
 public:
  template <class NodeT> explicit TransformBroadcaster(NodeT &&) {}
  void sendTransform(const geometry_msgs::msg::TransformStamped &t) { sent.push_back(t); }
  void sendTransform(const std::vector<geometry_msgs::msg::TransformStamped> &v) { sent.insert(sent.end(), v.begin(), v.end()); }
  std::vector<geometry_msgs::msg::TransformStamped> sent;
};
}
