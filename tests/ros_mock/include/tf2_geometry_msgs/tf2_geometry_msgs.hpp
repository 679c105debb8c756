#pragma once
#include "geometry_msgs/msg/types.hpp"
#include "tf2/LinearMath/Quaternion.h"
namespace tf2 {
inline geometry_msgs::msg::Quaternion toMsg(const Quaternion &q) { geometry_msgs::msg::Quaternion m; m.x = q.x(); m.y = q.y(); m.z = q.z(); m.w = q.w(); return m; }
inline void fromMsg(const geometry_msgs::msg::Quaternion &m, Quaternion &q) { q = Quaternion(m.x, m.y, m.z, m.w); }
inline double getYaw(const geometry_msgs::msg::Quaternion &m) { double r, p, y; Matrix3x3(Quaternion(m.x, m.y, m.z, m.w)).getRPY(r, p, y); return y; }
}
