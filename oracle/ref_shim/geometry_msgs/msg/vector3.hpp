// Shim for geometry_msgs/msg/Vector3 (float64 x,y,z) -- oracle build only.
#pragma once
#include "../../ros_std_includes.hpp"
namespace geometry_msgs { namespace msg {
struct Vector3 { double x = 0.0; double y = 0.0; double z = 0.0; };
} }
