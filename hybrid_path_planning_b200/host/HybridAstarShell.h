// HybridAstarShell.h -- drop-in for the reference's HybridAstar of both packages (RS/include/HybridAstar.h:44-96,
// DB/include/HybridAstar.h:33-83): same constructor and public methods, bodies call the C ABI.
//   HybridAstar        reeds_shepp_hybrid_astar  (26 motion commands, circle footprint, Reeds-Shepp shots)
//   HybridAstarDubins  dubins_hybrid_star        (13 forward commands, SAT footprint, Dubins shots)
// A package selects its class with `using HybridAstar = HybridAstarDubins;`.
// run() hands the whole query to hpp_plan_batch: heuristic field, open list, shots and expansions stay on the device;
// pop order, tie-breaking, the 3500-iteration cap and the failure convention (a 1-state path = the goal) follow
// RS/src/HybridAstar.cpp:553-729.  run_batch() is the multi-query form (one call, n queries).
#pragma once
#include <memory>
#include <vector>

#include <type_traits>

#include "GridMapShell.h"
#include "PathShells.h"

template <int VARIANT>
class HybridAstarT {
 public:
  HybridAstarT(Grid_map grid_map, CarData car_data, double simulationLength, double step_car);

  std::vector<State> run(State start_state, State goal_state);
  std::vector<std::vector<State>> run_batch(const std::vector<State> &starts, const std::vector<State> &goals);

  // RS/include/HybridAstar.h:51 / DB/include/HybridAstar.h:40: the reference fills its private command list here
  // (RS/src/HybridAstar.cpp:14-23); the commands of the shell live with the map on the device (hpp_map_set_params)
  // and this call fetches a host copy of them: {steering angle, direction} per command, in the reference's order
  void motionCommands();
  const std::vector<std::vector<double>> &commands() const { return motionCommand; }
  // RS/include/HybridAstar.h:52-53 / DB/include/HybridAstar.h:41-42
  std::vector<std::shared_ptr<planner::Node>> GetnextNeighbours(const std::shared_ptr<planner::Node> &current_node);
  // RS/src/HybridAstar.cpp:80-107: Cost_path of the child that motionCommand = {steering angle, direction} produces
  double PathCost(const std::shared_ptr<planner::Node> &current_node, const std::vector<double> &motionCommand);
  // RS/include/HybridAstar.h:56 (reeds_shepp_Path) / DB/include/HybridAstar.h:47 (dubins_Path): nullptr = no shot
  std::shared_ptr<planner::Node> analyticShot(const std::shared_ptr<planner::Node> &current_Node,
                                              const std::shared_ptr<planner::Node> &goal_Node);
  std::shared_ptr<planner::Node> reeds_shepp_Path(const std::shared_ptr<planner::Node> &c,
                                                  const std::shared_ptr<planner::Node> &g) {
    return analyticShot(c, g);
  }
  std::shared_ptr<planner::Node> dubins_Path(const std::shared_ptr<planner::Node> &c,
                                             const std::shared_ptr<planner::Node> &g) {
    return analyticShot(c, g);
  }
  // The heuristic under the reference's name and with the reference's return type of each package:
  //   RS/include/HybridAstar.h:62  std::vector<double>, cost per cell, row-major [y*W + x], +inf where unreachable
  //   DB/include/HybridAstar.h:51  std::vector<std::shared_ptr<planner::HolonomicNode>>, one node per cell (->cost 1000
  //                                where unreachable; only ->cost and the cell of ->Current_state are meaningful)
  using HeuristicResult = typename std::conditional<VARIANT == HPP_VARIANT_RS, std::vector<double>,
                                                    std::vector<std::shared_ptr<planner::HolonomicNode>>>::type;
  HeuristicResult holonomicCostsWithObstacles_planning(const std::shared_ptr<planner::Node> &GoalNode);
  // both forms for either planner
  std::vector<double> holonomicCosts(const std::shared_ptr<planner::Node> &GoalNode);
  std::vector<std::shared_ptr<planner::HolonomicNode>> holonomicCostsWithObstacles_nodes(
      const std::shared_ptr<planner::Node> &GoalNode);
  // set when the last heuristic call hit an input on which the reference's sequential quirks (H1/H2) may fire
  const hpp_quirk_report &lastQuirkReport() const { return report_; }

 private:
  State withCell(State s) const;
  Grid_map grid_map_;
  CarData car_data_;
  double simulationLength, step_car;
  State start_state_;
  std::vector<std::vector<double>> motionCommand;
  hpp_quirk_report report_{};
};

// `HybridAstar` is the Reeds-Shepp planner; the dubins_hybrid_star package defines HPP_SHELL_DEFAULT_DUBINS before
// including this header (its forwarding HybridAstar.h does), which makes `HybridAstar` the Dubins planner there.
using HybridAstarReedsShepp = HybridAstarT<HPP_VARIANT_RS>;
using HybridAstarDubins = HybridAstarT<HPP_VARIANT_DUBINS>;
#ifdef HPP_SHELL_DEFAULT_DUBINS
using HybridAstar = HybridAstarDubins;
#else
using HybridAstar = HybridAstarReedsShepp;
#endif

// drop-in for map_inflate_node::InflateMap + CreateInflatedOccupancyGrid (MI/src/map_inflate_node.cpp:17,:118-161):
// occupancy in, inflated occupancy out (same geometry), through hpp_inflate.
std::vector<int8_t> InflateOccupancy(const std::vector<int8_t> &occ, uint32_t width, uint32_t height, double resolution,
                                     double inflation_radius);

// drop-in for grid_map_publisher's image -> OccupancyGrid conversion (GP/src/grid_publisher.cpp:32-50,78-86): 8-bit gray
// image (rows x cols, row-major) in, occupancy data out; the grid has width = rows and height = cols.
std::vector<int8_t> OccupancyFromImage(const std::vector<uint8_t> &img, uint32_t rows, uint32_t cols);
