// Minimal stand-ins for the ROS 2 headers the reference's node sources include (tests/ros_mock).  NOT ROS: just enough
// types, members and calls for main_path_planning.cpp of both planner packages to compile unmodified against the
// hpp-b200 host shells, and for a test harness to feed messages to the node's callbacks and read what it publishes.
#pragma once
#include <cstdint>
namespace builtin_interfaces { namespace msg {
struct Time { int32_t sec = 0; uint32_t nanosec = 0; };
struct Duration { int32_t sec = 0; uint32_t nanosec = 0; };
} }
