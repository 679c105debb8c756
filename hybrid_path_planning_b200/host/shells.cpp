// shells.cpp -- bodies of the host shells (GridMapShell.h, HybridAstarShell.h): marshalling only, every computation
// is a C-ABI call into libhpp_b200.so.  Errors of the library become std::runtime_error (the reference's own
// conventions -- nullptr shot, 1-state failure path -- are kept as return values).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "HybridAstarShell.h"

namespace hpp_host {

std::shared_ptr<CtxHolder> thread_ctx_ref() {
  static thread_local std::shared_ptr<CtxHolder> h;
  if (!h) {
    auto fresh = std::make_shared<CtxHolder>();
    const char *d = std::getenv("HPP_DEVICE");
    const hpp_status st = hpp_ctx_create(d ? std::atoi(d) : 0, &fresh->ctx);
    if (st != HPP_OK) throw std::runtime_error("hpp_ctx_create failed: no usable B200 (there is no CPU fallback)");
    h = fresh;
  }
  return h;
}

hpp_ctx *thread_ctx() { return thread_ctx_ref()->ctx; }

void fail(hpp_ctx *ctx, hpp_status st, const char *what) {
  throw std::runtime_error(std::string(what) + ": status " + std::to_string((int)st) + ": " + hpp_last_error(ctx));
}

}  // namespace hpp_host

using hpp_host::check;

// ---------------------------------------------------------------------------------------------- Grid_map
void Grid_map::init(const int8_t *occ, std::size_t n, uint32_t w, uint32_t h, float res, double ox, double oy) {
  if (n != (std::size_t)w * h) throw std::runtime_error("Grid_map: data size does not match width*height");
  resolution = res;  // float -> double, like the reference (hazard H6)
  originX = ox;
  originY = oy;
  width = w;
  height = h;
  Max_origin_x = originX + width * resolution;
  Max_origin_y = originY + height * resolution;
  occ_ = std::make_shared<std::vector<int8_t>>(occ, occ + n);
  auto handle = std::make_shared<hpp_host::MapHandle>();
  handle->keep = hpp_host::thread_ctx_ref();
  handle->ctx = handle->keep->ctx;
  // default vehicle until setcarData (the reference's Grid_map also starts with a default-constructed CarData)
  car_ = CarData(0.6, 2.75, 1.1, 1.2, 1.9);
  const hpp_params p = car_.params(path_length_, step_car_);
  check(handle->ctx, hpp_map_create(handle->ctx, occ, w, h, res, ox, oy, &p, &handle->map), "hpp_map_create");
  handle->installed = p;
  handle->has_params = true;
  h_ = handle;
}

void Grid_map::setcarData(const CarData &car_data) { setPlannerParams(car_data, path_length_, step_car_); }

void Grid_map::setPlannerParams(const CarData &car_data, double pathLength, double step_car) {
  car_ = car_data;
  path_length_ = pathLength;
  step_car_ = step_car;
  if (h_) syncParams();  // reports unsupported parameters (e.g. pathLength > 8) here, where the reference's setter is called
}

void Grid_map::syncParams() const {
  if (!h_) return;
  const hpp_params p = car_.params(path_length_, step_car_);
  if (h_->has_params && std::memcmp(&h_->installed, &p, sizeof p) == 0) return;
  check(h_->ctx, hpp_map_set_params(h_->map, &p), "hpp_map_set_params");
  h_->installed = p;
  h_->has_params = true;
}

bool Grid_map::isSingleStateCollisionFree(const State &s) {
  syncParams();
  const double xyh[3] = {s.x, s.y, s.heading};
  uint8_t v = 1;
  check(h_->ctx, hpp_collide_circles(h_->map, xyh, 1, &v), "hpp_collide_circles");
  return v != 0;
}

bool Grid_map::isSingleStateCollisionFreeImproved(const State &s) {
  syncParams();
  const double xyh[3] = {s.x, s.y, s.heading};
  uint8_t v = 0;
  check(h_->ctx, hpp_collide_improved(h_->map, xyh, 1, &v), "hpp_collide_improved");
  return v != 0;
}

bool Grid_map::checkCollisionSat(const State &s) {
  syncParams();
  const double xyh[3] = {s.x, s.y, s.heading};
  uint8_t v = 1;
  check(h_->ctx, hpp_collide_sat(h_->map, xyh, 1, &v), "hpp_collide_sat");
  return v != 0;
}

bool Grid_map::isInCollision(int x, int y) const {
  if (!isPointInBounds(x, y)) return true;
  return (*occ_)[(std::size_t)toCellIndex(x, y)] > 0;
}

void Grid_map::layers(float *obstacle, float *distance) const {
  check(h_->ctx, hpp_map_layers(h_->map, obstacle, distance), "hpp_map_layers");
}

// ---------------------------------------------------------------------------------------------- HybridAstar
template <int V>
HybridAstarT<V>::HybridAstarT(Grid_map grid_map, CarData car_data, double simulationLength_, double step_car_)
    : grid_map_(grid_map), car_data_(car_data), simulationLength(simulationLength_), step_car(step_car_) {
  grid_map_.setPlannerParams(car_data_, simulationLength, step_car);
}

template <int V>
State HybridAstarT<V>::withCell(State s) const {
  int cx, cy;
  std::tie(cx, cy) = grid_map_.toCellID(s);
  s.gridx = cx;
  s.gridy = cy;
  return s;
}

static std::vector<State> states_from(const double *t, int n) {
  std::vector<State> out((std::size_t)n);
  for (int i = 0; i < n; i++) {
    out[i].x = t[5 * i];
    out[i].y = t[5 * i + 1];
    out[i].heading = t[5 * i + 2];
    out[i].gridx = t[5 * i + 3];
    out[i].gridy = t[5 * i + 4];
  }
  return out;
}

template <int V>
std::vector<std::vector<State>> HybridAstarT<V>::run_batch(const std::vector<State> &starts,
                                                           const std::vector<State> &goals) {
  const std::size_t n = starts.size();
  if (goals.size() != n) throw std::runtime_error("run_batch: starts and goals differ in length");
  std::vector<std::vector<State>> out(n);
  if (n == 0) return out;
  std::vector<double> s(3 * n), g(3 * n);
  for (std::size_t i = 0; i < n; i++) {
    s[3 * i] = starts[i].x, s[3 * i + 1] = starts[i].y, s[3 * i + 2] = starts[i].heading;
    g[3 * i] = goals[i].x, g[3 * i + 1] = goals[i].y, g[3 * i + 2] = goals[i].heading;
  }
  std::vector<hpp_plan> plans(n);
  std::vector<double> traj;
  std::size_t cap = 4096 * n, used = 0;
  grid_map_.syncParams();
  for (;;) {
    traj.resize(cap * 5);
    const hpp_status st =
        hpp_plan_batch(grid_map_.handle(), V, s.data(), g.data(), n, plans.data(), traj.data(), cap, &used);
    if (st == HPP_ERR_CAPACITY) {
      cap = used > cap ? used : cap * 2;
      continue;
    }
    check(grid_map_.ctx(), st, "hpp_plan_batch");
    break;
  }
  for (std::size_t i = 0; i < n; i++) out[i] = states_from(traj.data() + 5 * plans[i].traj_offset, plans[i].n_states);
  return out;
}

template <int V>
std::vector<State> HybridAstarT<V>::run(State start_state, State goal_state) {
  start_state_ = withCell(start_state);
  return run_batch({start_state}, {goal_state})[0];
}

template <int V>
std::vector<std::shared_ptr<planner::Node>> HybridAstarT<V>::GetnextNeighbours(
    const std::shared_ptr<planner::Node> &cur) {
  hpp_node in;
  std::memset(&in, 0, sizeof in);
  in.x = cur->Current_state.x;
  in.y = cur->Current_state.y;
  in.heading = cur->Current_state.heading;
  in.cost = cur->Cost_path;
  in.steer = cur->steeringAngle;
  in.direction = cur->direction;
  const hpp_params p = car_data_.params(simulationLength, step_car);
  const int nc = hpp_num_commands(&p, V);
  std::vector<hpp_child> kids((std::size_t)nc);
  grid_map_.syncParams();
  check(grid_map_.ctx(), hpp_expand(grid_map_.handle(), V, &in, 1, kids.data()), "hpp_expand");
  int sim = 0;
  while (sim < simulationLength && sim < HPP_MAX_SIM) sim++;
  std::vector<std::shared_ptr<planner::Node>> out;
  for (const hpp_child &k : kids) {
    if (!k.valid) continue;  // the reference drops a rollout when any of its states collides
    std::vector<State> traj = states_from(&k.traj[0][0], sim);
    out.push_back(std::make_shared<planner::Node>(traj.back(), traj, k.cost, k.steer, k.direction, cur));
  }
  return out;
}

template <int V>
std::shared_ptr<planner::Node> HybridAstarT<V>::analyticShot(const std::shared_ptr<planner::Node> &cur,
                                                             const std::shared_ptr<planner::Node> &goal) {
  hpp_node in;
  std::memset(&in, 0, sizeof in);
  in.x = cur->Current_state.x;
  in.y = cur->Current_state.y;
  in.heading = cur->Current_state.heading;
  in.cost = cur->Cost_path;
  in.steer = cur->steeringAngle;
  in.direction = cur->direction;
  const double g[3] = {goal->Current_state.x, goal->Current_state.y, goal->Current_state.heading};
  hpp_shot shot;
  std::vector<double> samples;
  std::size_t cap = 8192, used = 0;
  grid_map_.syncParams();
  for (;;) {
    samples.resize(cap * 5);
    const hpp_status st = (V == HPP_VARIANT_RS)
                              ? hpp_rs_shot(grid_map_.handle(), &in, 1, g, 0, &shot, samples.data(), cap, &used)
                              : hpp_dubins_shot(grid_map_.handle(), &in, 1, g, 0, &shot, samples.data(), cap, &used);
    if (st == HPP_ERR_CAPACITY) {
      cap = used > cap ? used : cap * 2;
      continue;
    }
    check(grid_map_.ctx(), st, "analytic shot");
    break;
  }
  if (!shot.found) return nullptr;
  std::vector<State> traj = states_from(samples.data() + 5 * shot.sample_offset, shot.n_samples);
  // RS/src/HybridAstar.cpp:196-206: the shot node carries direction 1, steering 0
  return std::make_shared<planner::Node>(traj.back(), traj, shot.cost, 0.0, 1, cur);
}

template <int V>
void HybridAstarT<V>::motionCommands() {
  // one expansion of a free-standing node returns every command's {steering angle, direction} in command order
  hpp_node in;
  std::memset(&in, 0, sizeof in);
  in.direction = 1;
  const hpp_params p = car_data_.params(simulationLength, step_car);
  const int nc = hpp_num_commands(&p, V);
  std::vector<hpp_child> kids((std::size_t)nc);
  grid_map_.syncParams();
  check(grid_map_.ctx(), hpp_expand(grid_map_.handle(), V, &in, 1, kids.data()), "hpp_expand");
  motionCommand.clear();
  for (const hpp_child &k : kids) motionCommand.push_back({k.steer, (double)k.direction});
}

template <int V>
double HybridAstarT<V>::PathCost(const std::shared_ptr<planner::Node> &cur, const std::vector<double> &cmd) {
  if (cmd.size() < 2) throw std::runtime_error("PathCost: a motion command is {steering angle, direction}");
  hpp_node in;
  std::memset(&in, 0, sizeof in);
  in.x = cur->Current_state.x;
  in.y = cur->Current_state.y;
  in.heading = cur->Current_state.heading;
  in.cost = cur->Cost_path;
  in.steer = cur->steeringAngle;
  in.direction = cur->direction;
  const hpp_params p = car_data_.params(simulationLength, step_car);
  const int nc = hpp_num_commands(&p, V);
  std::vector<hpp_child> kids((std::size_t)nc);
  grid_map_.syncParams();
  check(grid_map_.ctx(), hpp_expand(grid_map_.handle(), V, &in, 1, kids.data()), "hpp_expand");
  for (const hpp_child &k : kids)  // the cost is filled for every command, colliding rollouts included
    if (k.steer == cmd[0] && (double)k.direction == cmd[1]) return k.cost;
  throw std::runtime_error("PathCost: not one of the planner's motion commands");
}

template <int V>
typename HybridAstarT<V>::HeuristicResult HybridAstarT<V>::holonomicCostsWithObstacles_planning(
    const std::shared_ptr<planner::Node> &goal) {
  if constexpr (V == HPP_VARIANT_RS) return holonomicCosts(goal);
  else return holonomicCostsWithObstacles_nodes(goal);
}

template <int V>
std::vector<double> HybridAstarT<V>::holonomicCosts(const std::shared_ptr<planner::Node> &goal) {
  const State gs = withCell(goal->Current_state);
  std::vector<double> cost((std::size_t)grid_map_.getWidth() * grid_map_.getHeight());
  const hpp_status st =
      (V == HPP_VARIANT_RS)
          ? hpp_heuristic_rs(grid_map_.handle(), (int)gs.gridx, (int)gs.gridy, (int)start_state_.gridx,
                             (int)start_state_.gridy, cost.data(), &report_)
          : hpp_heuristic_dubins(grid_map_.handle(), (int)gs.gridx, (int)gs.gridy, cost.data(), &report_);
  check(grid_map_.ctx(), st, "heuristic");
  return cost;
}

template <int V>
std::vector<std::shared_ptr<planner::HolonomicNode>> HybridAstarT<V>::holonomicCostsWithObstacles_nodes(
    const std::shared_ptr<planner::Node> &goal) {
  const std::vector<double> cost = holonomicCosts(goal);
  std::vector<std::shared_ptr<planner::HolonomicNode>> out(cost.size());
  const int W = (int)grid_map_.getWidth();
  for (std::size_t i = 0; i < cost.size(); i++) {
    State s;
    s.gridx = (double)((int)i % W);
    s.gridy = (double)((int)i / W);
    out[i] = std::make_shared<planner::HolonomicNode>(s, cost[i], std::weak_ptr<planner::HolonomicNode>());
  }
  return out;
}

template class HybridAstarT<HPP_VARIANT_RS>;
template class HybridAstarT<HPP_VARIANT_DUBINS>;

// ---------------------------------------------------------------------------------------------- analytic path classes
static double wrap_pi(double theta) {  // RS/src/reeds_shepp_path.cpp pi_2_pi / DB/src/dubins_path.cpp pi_2_pi
  while (theta > M_PI) theta -= 2.0 * M_PI;
  while (theta < -M_PI) theta += 2.0 * M_PI;
  return theta;
}
double reeds_shepp_path::pi_2_pi(double theta) { return wrap_pi(theta); }
double dubins_path::pi_2_pi(double theta) { return wrap_pi(theta); }
double dubins_path::mod2pi(double theta) { return theta - 2.0 * M_PI * std::floor(theta / M_PI / 2.0); }  // DB :26-29

std::vector<PATH> reeds_shepp_path::calc_all_paths(double sx, double sy, double syaw, double gx, double gy, double gyaw,
                                                   double maxc, double step_size) {
  hpp_ctx *ctx = hpp_host::thread_ctx();
  const double s[3] = {sx, sy, syaw}, g[3] = {gx, gy, gyaw};
  const int max_paths = 48;
  int32_t n_paths = 0, n_seg[48], n_samples[48];
  double lengths[48 * 5], total[48];
  char types[48 * 5];
  std::vector<double> samples;
  std::vector<int32_t> dirs;
  std::size_t cap = 1 << 16;
  for (;;) {
    samples.resize(cap * 3);
    dirs.resize(cap);
    const hpp_status st = hpp_rs_all_paths(ctx, s, g, maxc, step_size, max_paths, &n_paths, n_seg, lengths, types, total,
                                           n_samples, samples.data(), cap, dirs.data());
    if (st == HPP_ERR_CAPACITY && cap < ((std::size_t)1 << 26)) {
      cap *= 4;
      continue;
    }
    check(ctx, st, "hpp_rs_all_paths");
    break;
  }
  std::vector<PATH> out;
  std::size_t off = 0;
  for (int e = 0; e < n_paths; e++) {
    PATH p;
    for (int i = 0; i < n_seg[e]; i++) {
      p.lengths.push_back(lengths[e * 5 + i]);
      p.ctypes.push_back(std::string(1, types[e * 5 + i]));
    }
    p.L = total[e];
    for (int j = 0; j < n_samples[e]; j++) {
      p.x.push_back(samples[3 * (off + j)]);
      p.y.push_back(samples[3 * (off + j) + 1]);
      p.yaw.push_back(samples[3 * (off + j) + 2]);
      p.directions.push_back(dirs[off + j]);
    }
    off += (std::size_t)n_samples[e];
    out.push_back(std::move(p));
  }
  return out;
}

PATH reeds_shepp_path::calc_optimal_path(double sx, double sy, double syaw, double gx, double gy, double gyaw, double maxc,
                                         double step_size) {
  std::vector<PATH> paths = calc_all_paths(sx, sy, syaw, gx, gy, gyaw, maxc, step_size);
  if (paths.empty()) throw std::runtime_error("calc_optimal_path: no path (the reference indexes paths[0] here)");
  std::size_t mini = 0;  // RS/src/reeds_shepp_path.cpp:24-35: `<=`, so the last one among equals
  double minL = paths[0].L;
  for (std::size_t i = 0; i < paths.size(); ++i)
    if (paths[i].L <= minL) {
      minL = paths[i].L;
      mini = i;
    }
  return paths[mini];
}

Path dubins_path::calc_dubins_path(double sx, double sy, double syaw, double gx, double gy, double gyaw, double curv,
                                   double step_size) {
  hpp_ctx *ctx = hpp_host::thread_ctx();
  const double s[3] = {sx, sy, syaw}, g[3] = {gx, gy, gyaw};
  double length = 0;
  char mode[4] = {0, 0, 0, 0};
  int32_t n = 0;
  std::vector<double> samples;
  std::size_t cap = 1 << 14;
  for (;;) {
    samples.resize(cap * 3);
    const hpp_status st = hpp_dubins_path(ctx, s, g, curv, step_size, &length, mode, &n, samples.data(), cap);
    if (st == HPP_ERR_CAPACITY && cap < ((std::size_t)1 << 26)) {
      cap *= 4;
      continue;
    }
    check(ctx, st, "hpp_dubins_path");
    break;
  }
  Path p;
  p.length = length;
  for (int i = 0; i < 3 && mode[i]; i++) p.mode.push_back(std::string(1, mode[i]));
  for (int j = 0; j < n; j++) {
    p.x.push_back(samples[3 * j]);
    p.y.push_back(samples[3 * j + 1]);
    p.yaw.push_back(samples[3 * j + 2]);
  }
  return p;
}

// ---------------------------------------------------------------------------------------------- map_inflate
std::vector<int8_t> InflateOccupancy(const std::vector<int8_t> &occ, uint32_t width, uint32_t height, double resolution,
                                     double inflation_radius) {
  if (occ.size() != (std::size_t)width * height) throw std::runtime_error("InflateOccupancy: size mismatch");
  std::vector<int8_t> out(occ.size());
  hpp_ctx *ctx = hpp_host::thread_ctx();
  check(ctx, hpp_inflate(ctx, occ.data(), width, height, resolution, inflation_radius, out.data()), "hpp_inflate");
  return out;
}

std::vector<int8_t> OccupancyFromImage(const std::vector<uint8_t> &img, uint32_t rows, uint32_t cols) {
  if (img.size() != (std::size_t)rows * cols) throw std::runtime_error("OccupancyFromImage: size mismatch");
  std::vector<int8_t> out(img.size());
  hpp_ctx *ctx = hpp_host::thread_ctx();
  check(ctx, hpp_image_to_occupancy(ctx, img.data(), rows, cols, out.data()), "hpp_image_to_occupancy");
  return out;
}
