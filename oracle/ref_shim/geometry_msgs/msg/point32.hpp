// Shim for geometry_msgs/msg/Point32 (float32 x,y,z) -- oracle build only.
#pragma once
#include "../../ros_std_includes.hpp"
namespace geometry_msgs { namespace msg {
struct Point32 { float x = 0.0f; float y = 0.0f; float z = 0.0f; };
} }
