// Shim for nav_msgs/msg/Path -- only included, never used, by the reference hot path.
#pragma once
#include "../../ros_std_includes.hpp"
namespace nav_msgs { namespace msg { struct Path {}; } }
