// Shim for nav_msgs/msg/OccupancyGrid -- oracle build only. Field types follow the
// ROS 2 message definition: MapMetaData.resolution is float32, width/height uint32,
// origin is a Pose with float64 position, data is int8[].
#pragma once
#include "../../ros_std_includes.hpp"
namespace nav_msgs { namespace msg {
struct OccupancyGrid {
  struct Header { std::string frame_id; } header;
  struct Info {
    float resolution = 0.0f;
    uint32_t width = 0, height = 0;
    struct Pose {
      struct Pt { double x = 0, y = 0, z = 0; } position;
      struct Qt { double x = 0, y = 0, z = 0, w = 1; } orientation;
    } origin;
  } info;
  std::vector<int8_t> data;
  using SharedPtr = std::shared_ptr<OccupancyGrid>;
};
} }
