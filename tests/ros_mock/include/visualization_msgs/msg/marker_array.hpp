#pragma once
#include "visualization_msgs/msg/marker.hpp"
