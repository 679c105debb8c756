// mi_node_harness.cpp -- runs the reference's UNMODIFIED map_inflate node (MI/src/map_inflate_node.cpp, included from where
// it lies under the reference checkout with -Dmain=reference_node_main: the node class has no header) against the cv
// shim of hybrid_path_planning_b200/host/dropin/mi, inside the stand-in ROS of tests/ros_mock.  The harness plays the map
// publisher and writes what the node publishes on /grid_map_inflated.
//   usage: mi_node_harness map.bin inflation_radius out.bin
//   map.bin: int32 W, int32 H, float32 resolution, float64 origin x, float64 origin y, then W*H int8 (row-major [y][x])
// Test infrastructure; nothing of the reference is copied here.
#include "map_inflate_node.cpp"  // found through -I<reference>/map_inflate/src at build time
#undef main

#include <cstdio>
#include <cstdlib>

int main(int argc, char **argv) {
  if (argc != 4) {
    std::fprintf(stderr, "usage: %s map.bin inflation_radius out.bin\n", argv[0]);
    return 2;
  }
  FILE *f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  int32_t W = 0, H = 0;
  float res = 0;
  double ox = 0, oy = 0;
  if (std::fread(&W, 4, 1, f) != 1 || std::fread(&H, 4, 1, f) != 1 || std::fread(&res, 4, 1, f) != 1 ||
      std::fread(&ox, 8, 1, f) != 1 || std::fread(&oy, 8, 1, f) != 1)
    return 2;
  nav_msgs::msg::OccupancyGrid map;
  map.header.frame_id = "map";
  map.info.width = (uint32_t)W;
  map.info.height = (uint32_t)H;
  map.info.resolution = res;
  map.info.origin.position.x = ox;
  map.info.origin.position.y = oy;
  map.data.resize((size_t)W * H);
  if (std::fread(map.data.data(), 1, map.data.size(), f) != map.data.size()) return 2;
  std::fclose(f);
  rclcpp::Node::overrides()["inflation_radius"].d = std::atof(argv[2]);
  auto node = std::make_shared<map_inflate_node>();
  node->deliver<nav_msgs::msg::OccupancyGrid>("/grid_map", map);
  node->deliver<nav_msgs::msg::OccupancyGrid>("/grid_map", map);  // same dimensions: the stored inflated map is re-published
  auto &out = node->published<nav_msgs::msg::OccupancyGrid>("/grid_map_inflated");
  std::printf("published %zu\n", out.size());
  if (out.size() != 2 || out[0].data != out[1].data || out[0].data.size() != map.data.size()) return 1;
  if (out[0].info.width != map.info.width || out[0].info.height != map.info.height || out[0].header.frame_id != "map") return 1;
  FILE *o = std::fopen(argv[3], "wb");
  if (!o) return 2;
  std::fwrite(out[0].data.data(), 1, out[0].data.size(), o);
  std::fclose(o);
  return 0;
}
