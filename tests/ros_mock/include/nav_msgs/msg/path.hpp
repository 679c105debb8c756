#pragma once
#include "nav_msgs/msg/occupancy_grid.hpp"
