// hpp_types.h -- value types of the planners' public interface, as the host shells exchange them with callers.
// Same member names and meaning as the reference's State / Circle / StateHash (RS/include/State.h:5-65) and
// planner::Node / planner::HolonomicNode (RS/include/Node.h:12-49), so code written against the reference classes
// compiles unchanged against the shells.  gridx / gridy are doubles that hold truncated cell ids (hazard H6/H7).
#pragma once
#include <cmath>
#include <cstddef>
#include <functional>
#include <memory>
#include <vector>

class State {
 public:
  State() = default;
  State(double x_, double y_, double heading_ = 0) : x(x_), y(y_), heading(heading_) {}
  double x{}, y{}, heading{};
  double gridx{}, gridy{}, gheading{};
  // equality with the reference's 1e-9 tolerance on the continuous pose; ordering lexicographic on (x, y, heading)
  bool operator==(const State &o) const {
    return std::abs(x - o.x) < 1e-9 && std::abs(y - o.y) < 1e-9 && std::abs(heading - o.heading) < 1e-9;
  }
  bool operator<(const State &o) const {
    if (x != o.x) return x < o.x;
    if (y != o.y) return y < o.y;
    return heading < o.heading;
  }
};

struct Circle {
  Circle() = default;
  Circle(double x_, double y_, double r_) : x(x_), y(y_), r(r_) {}
  double x{}, y{}, r{};
};

struct StateHash {
  std::size_t operator()(const State &s) const {
    const std::size_t hx = std::hash<double>()(s.x), hy = std::hash<double>()(s.y), hh = std::hash<double>()(s.heading);
    return ((hx ^ (hy << 1)) >> 1) ^ (hh << 1);
  }
};

namespace planner {

class Node {
 public:
  Node(const State &state, const std::vector<State> &trajectory, double cost_path, double steering, int dir,
       std::weak_ptr<Node> parent)
      : Current_state(state), Trajectory(trajectory), Cost_path(cost_path), steeringAngle(steering), direction(dir),
        Parent(parent) {}
  State Current_state;
  std::vector<State> Trajectory;
  double Cost_path;
  double steeringAngle;
  int direction;
  std::weak_ptr<Node> Parent;
};

class HolonomicNode {
 public:
  HolonomicNode() : cost(1000) {}
  HolonomicNode(const State &state, double c, std::weak_ptr<HolonomicNode> parent)
      : Current_state(state), cost(c), Parent(parent) {}
  State Current_state;
  double cost;
  std::weak_ptr<HolonomicNode> Parent;
};

}  // namespace planner
