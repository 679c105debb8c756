// Shim for geometry_msgs/msg/Polygon (Point32[] points) -- oracle build only.
#pragma once
#include "point32.hpp"
namespace geometry_msgs { namespace msg {
struct Polygon { std::vector<Point32> points; };
} }
