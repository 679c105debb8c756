#pragma once
#include <string>
#include "builtin_interfaces/msg/time.hpp"
namespace rclcpp { class Time; }
namespace std_msgs { namespace msg {
struct Stamp : builtin_interfaces::msg::Time {
  Stamp() = default;
  Stamp(const builtin_interfaces::msg::Time &t) : builtin_interfaces::msg::Time(t) {}
};
struct Header { builtin_interfaces::msg::Time stamp; std::string frame_id; };
struct ColorRGBA { float r = 0, g = 0, b = 0, a = 0; };
} }
