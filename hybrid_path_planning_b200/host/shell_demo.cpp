// shell_demo.cpp -- exercises the host shells the way the reference's ROS 2 nodes use their classes
// (RS/src/main_path_planning.cpp:63-69 gridMapdata, :283-300 the query) and prints the results as text.
// Built and run by tests/test_gpu_host_shells.py, which compares the output with the oracle.
//   shell_demo <occ.raw> <W> <H> <variant 0|1> <sx> <sy> <sh> <gx> <gy> <gh>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "HybridAstarShell.h"

// shaped like nav_msgs::msg::OccupancyGrid (the real message type is used in the ROS 2 node)
struct OccupancyGridLike {
  struct {
    float resolution = 0.2f;
    uint32_t width = 0, height = 0;
    struct {
      struct {
        double x = 0, y = 0, z = 0;
      } position;
    } origin;
  } info;
  std::vector<int8_t> data;
};

static std::vector<double> costs_of(const std::vector<double> &h) { return h; }
static std::vector<double> costs_of(const std::vector<std::shared_ptr<planner::HolonomicNode>> &nodes) {
  std::vector<double> h;
  for (const auto &n : nodes) h.push_back(n->cost);
  return h;
}

static void drive_paths(const State &start, const State &goal);

template <class Planner>
static int drive(Grid_map &grid, const CarData &car, const State &start, const State &goal) {
  Planner planner(grid, car, 4.0, 1.0);  // pathLength, step_car defaults of the node parameters
  const std::vector<State> path = planner.run(start, goal);
  std::printf("path %zu\n", path.size());
  for (const State &s : path) std::printf("%.17g %.17g %.17g %.0f %.0f\n", s.x, s.y, s.heading, s.gridx, s.gridy);
  State g = goal;
  int gx, gy;
  std::tie(gx, gy) = grid.toCellID(goal);
  g.gridx = gx;
  g.gridy = gy;
  auto goal_node = std::make_shared<planner::Node>(g, std::vector<State>{g}, 0.0, 0.0, 1, std::weak_ptr<planner::Node>());
  // the reference's method name with the reference's return type of the package: vector<double> (RS) or one
  // HolonomicNode per cell (Dubins)
  const auto hres = planner.holonomicCostsWithObstacles_planning(goal_node);
  const std::vector<double> h = costs_of(hres);
  double sum = 0;
  std::size_t finite = 0;
  for (double v : h)
    if (v < 1e300) sum += v, finite++;
  std::printf("heuristic %zu %zu %.17g\n", h.size(), finite, sum);
  State s0 = start;
  std::tie(gx, gy) = grid.toCellID(start);
  s0.gridx = gx;
  s0.gridy = gy;
  auto start_node = std::make_shared<planner::Node>(s0, std::vector<State>{s0}, 0.0, 0.0, 1, std::weak_ptr<planner::Node>());
  const auto kids = planner.GetnextNeighbours(start_node);
  std::printf("children %zu\n", kids.size());
  for (const auto &k : kids)
    std::printf("%.17g %.17g %.17g %.17g %.17g %d\n", k->Current_state.x, k->Current_state.y, k->Current_state.heading,
                k->Cost_path, k->steeringAngle, k->direction);
  const auto shot = planner.analyticShot(start_node, goal_node);
  if (shot) std::printf("shot 1 %.17g %zu\n", shot->Cost_path, shot->Trajectory.size());
  else std::printf("shot 0\n");
  // motionCommands / PathCost (RS/include/HybridAstar.h:51-53): every command's cost from the start node
  planner.motionCommands();
  std::printf("commands %zu\n", planner.commands().size());
  for (const auto &cmd : planner.commands())
    std::printf("%.17g %.0f %.17g\n", cmd[0], cmd[1], planner.PathCost(start_node, cmd));
  return 0;
}

// the analytic path classes of the public interface (RS/include/reeds_shepp_path.h, DB/include/dubins_path.h)
static void drive_paths(const State &start, const State &goal) {
  const double maxc = std::tan(0.6) / 2.75;
  reeds_shepp_path rs;
  const std::vector<PATH> all = rs.calc_all_paths(start.x, start.y, start.heading, goal.x, goal.y, goal.heading, maxc, 1.0);
  std::printf("rs_paths %zu\n", all.size());
  for (const PATH &p : all) {
    std::string t;
    for (const auto &c : p.ctypes) t += c;
    double sx = 0, sy = 0;
    long sd = 0;
    for (std::size_t i = 0; i < p.x.size(); i++) sx += p.x[i], sy += p.y[i], sd += p.directions[i];
    std::printf("%s %.17g %zu %.17g %.17g %ld\n", t.c_str(), p.L, p.x.size(), sx, sy, sd);
  }
  const PATH best = rs.calc_optimal_path(start.x, start.y, start.heading, goal.x, goal.y, goal.heading, maxc, 1.0);
  std::printf("rs_optimal %.17g %zu\n", best.L, best.x.size());
  dubins_path db;
  const Path d = db.calc_dubins_path(start.x, start.y, start.heading, goal.x, goal.y, goal.heading, maxc, 0.1);
  std::string m;
  for (const auto &c : d.mode) m += c;
  std::printf("dubins %s %.17g %zu %.17g\n", m.c_str(), d.length, d.x.size(), db.mod2pi(-1.0));
}

int main(int argc, char **argv) {
  if (argc < 11) return 2;
  OccupancyGridLike msg;
  msg.info.width = (uint32_t)std::atoi(argv[2]);
  msg.info.height = (uint32_t)std::atoi(argv[3]);
  msg.info.origin.position.x = -(double)msg.info.width * (double)msg.info.resolution / 2.0;
  msg.info.origin.position.y = -(double)msg.info.height * (double)msg.info.resolution / 2.0;
  msg.data.resize((std::size_t)msg.info.width * msg.info.height);
  std::ifstream f(argv[1], std::ios::binary);
  f.read((char *)msg.data.data(), (std::streamsize)msg.data.size());
  if (!f) return 3;
  const int variant = std::atoi(argv[4]);
  const State start(std::atof(argv[5]), std::atof(argv[6]), std::atof(argv[7]));
  const State goal(std::atof(argv[8]), std::atof(argv[9]), std::atof(argv[10]));
  try {
    Grid_map grid(msg);
    CarData car(0.6, 2.75, 1.1, 1.2, 1.9);  // maxSteerAngle, wheelBase, axleToFront, axleToBack, width
    grid.setcarData(car);
    std::printf("collide %d %d\n", (int)grid.isSingleStateCollisionFree(start), (int)grid.checkCollisionSat(start));
    const std::vector<int8_t> infl = InflateOccupancy(msg.data, msg.info.width, msg.info.height, msg.info.resolution, 0.2);
    long occupied = 0;
    for (int8_t v : infl) occupied += v > 0;
    std::printf("inflated %ld\n", occupied);
    const int rc = variant == 0 ? drive<HybridAstar>(grid, car, start, goal) : drive<HybridAstarDubins>(grid, car, start, goal);
    drive_paths(start, goal);
    return rc;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
