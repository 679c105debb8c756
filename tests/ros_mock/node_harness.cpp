// node_harness.cpp -- runs the reference's UNMODIFIED planner node (main_path_planning.cpp of reeds_shepp_hybrid_astar or
// dubins_hybrid_star, compiled from where it lies under the reference checkout with -Dmain=reference_node_main)
// against the hpp-b200 host shells, inside the stand-in ROS of tests/ros_mock: the harness plays the three publishers
// the node subscribes to (map, /initialpose, /goal_pose) and prints what the node publishes on /hybrid_astar_path.
//   usage: node_harness map.bin sx sy syaw gx gy gyaw
//   map.bin: int32 W, int32 H, float32 resolution, float64 origin x, float64 origin y, then W*H int8 (row-major [y][x])
// Test infrastructure; nothing of the reference is copied here.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "main_path_planning.h"  // the reference's node header (found through the build's temporary include directory)

int main(int argc, char **argv) {
  if (argc != 8) {
    std::fprintf(stderr, "usage: %s map.bin sx sy syaw gx gy gyaw\n", argv[0]);
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

  // the launch file's parameters (global_launcher/config: default vehicle of both planners)
  auto &ov = rclcpp::Node::overrides();
  const char *names[7] = {"maxSteerAngle", "wheelBase", "axleToFront", "axleToBack", "width", "pathLength", "step_car"};
  const double values[7] = {0.6, 2.75, 1.1, 1.2, 1.9, 4.0, 1.0};
  for (int i = 0; i < 7; i++) ov[names[i]].d = values[i];

  auto node = std::make_shared<main_path_planning>();
  node->deliver<nav_msgs::msg::OccupancyGrid>("/grid_map", map);

  auto quat = [](double yaw) {
    tf2::Quaternion q;
    q.setRPY(0, 0, yaw);
    geometry_msgs::msg::Quaternion m;
    m.x = q.x(); m.y = q.y(); m.z = q.z(); m.w = q.w();
    return m;
  };
  auto yaw_of = [](const geometry_msgs::msg::Quaternion &m) {  // what the node's callbacks compute
    double r, p, y;
    tf2::Matrix3x3(tf2::Quaternion(m.x, m.y, m.z, m.w)).getRPY(r, p, y);
    return y;
  };
  geometry_msgs::msg::PoseWithCovarianceStamped start;
  start.header.frame_id = "map";
  start.pose.pose.position.x = std::atof(argv[2]);
  start.pose.pose.position.y = std::atof(argv[3]);
  start.pose.pose.orientation = quat(std::atof(argv[4]));
  geometry_msgs::msg::PoseStamped goal;
  goal.header.frame_id = "map";
  goal.pose.position.x = std::atof(argv[5]);
  goal.pose.position.y = std::atof(argv[6]);
  goal.pose.orientation = quat(std::atof(argv[7]));
  std::printf("start %.17g %.17g %.17g\n", start.pose.pose.position.x, start.pose.pose.position.y,
              yaw_of(start.pose.pose.orientation));
  std::printf("goal %.17g %.17g %.17g\n", goal.pose.position.x, goal.pose.position.y, yaw_of(goal.pose.orientation));
  node->deliver<geometry_msgs::msg::PoseWithCovarianceStamped>("/initialpose", start);
  node->deliver<geometry_msgs::msg::PoseStamped>("/goal_pose", goal);  // the node plans when the goal arrives

  auto &paths = node->published<visualization_msgs::msg::MarkerArray>("/hybrid_astar_path");
  std::printf("published %zu\n", paths.size());
  if (!paths.empty())
    for (const auto &mk : paths.back().markers) std::printf("state %.17g %.17g\n", mk.pose.position.x, mk.pose.position.y);
  return 0;
}
