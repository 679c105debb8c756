#pragma once
#include <cstdint>
#include <memory>
#include <vector>
#include "geometry_msgs/msg/types.hpp"
namespace nav_msgs { namespace msg {
struct MapMetaData {
  builtin_interfaces::msg::Time map_load_time; float resolution = 0; uint32_t width = 0, height = 0;
  geometry_msgs::msg::Pose origin;
};
struct OccupancyGrid {
  std_msgs::msg::Header header; MapMetaData info; std::vector<int8_t> data;
  using SharedPtr = std::shared_ptr<OccupancyGrid>; using ConstSharedPtr = std::shared_ptr<const OccupancyGrid>;
};
struct Path {
  std_msgs::msg::Header header; std::vector<geometry_msgs::msg::PoseStamped> poses;
  using SharedPtr = std::shared_ptr<Path>;
};
} }
