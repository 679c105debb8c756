// GridMapShell.h -- drop-in for the reference's Grid_map (RS/include/Grid_map.h:84-110, DB/include/Grid_map.h): same
// public methods, bodies call the C ABI of libhpp_b200.so.  The device-resident map (occupancy, "obstacle"/"distance"
// layers, per-cell tables) is shared by all copies of a Grid_map object -- the reference copies the whole object per
// query (RS/src/main_path_planning.cpp:283); here a copy is a reference-count bump.
//
// MapMsg is any type shaped like nav_msgs::msg::OccupancyGrid (info.resolution float, info.width/height,
// info.origin.position.{x,y}, data = std::vector<int8_t>); the ROS 2 node passes the real message.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>

#include "../../include/hpp_b200.h"
#include "CarDataShell.h"
#include "hpp_types.h"

namespace hpp_host {

// one hpp_ctx per (thread, GPU); device from $HPP_DEVICE (default 0).  Throws when there is no usable B200:
// the product has no CPU path.  The context is reference counted: every map handle keeps it alive, so a Grid_map that
// outlives the thread that created it (or is destroyed during process shutdown) never touches a freed context.  A
// context -- and therefore a Grid_map and its copies -- must be used by one thread at a time.
struct CtxHolder {
  hpp_ctx *ctx = nullptr;
  ~CtxHolder() {
    if (ctx) hpp_ctx_destroy(ctx);
  }
};
std::shared_ptr<CtxHolder> thread_ctx_ref();
hpp_ctx *thread_ctx();
[[noreturn]] void fail(hpp_ctx *ctx, hpp_status st, const char *what);
inline void check(hpp_ctx *ctx, hpp_status st, const char *what) {
  if (st != HPP_OK) fail(ctx, st, what);
}

struct MapHandle {
  std::shared_ptr<CtxHolder> keep;  // released after the map is destroyed
  hpp_ctx *ctx = nullptr;
  hpp_map *map = nullptr;
  // the vehicle / planner parameters the device map currently carries: every Grid_map copy (and every HybridAstar, which
  // holds one) has its OWN parameters, as in the reference where the copies are independent objects; a copy pushes its
  // parameters to the shared device map before a call whenever another copy's are installed (Grid_map::syncParams)
  hpp_params installed{};
  bool has_params = false;
  ~MapHandle() {
    if (map) hpp_map_destroy(map);
  }
};

}  // namespace hpp_host

class Grid_map {
 public:
  Grid_map() = default;
  template <class MapMsg>
  explicit Grid_map(const MapMsg &map_data) {
    init(map_data.data.data(), map_data.data.size(), map_data.info.width, map_data.info.height, map_data.info.resolution,
         map_data.info.origin.position.x, map_data.info.origin.position.y);
  }

  double getResolution() const { return resolution; }
  double getOriginX() const { return originX; }
  double getOriginY() const { return originY; }
  unsigned int getWidth() const { return width; }
  unsigned int getHeight() const { return height; }
  double getMaxOriginX() const { return Max_origin_x; }
  double getMaxOriginY() const { return Max_origin_y; }

  // RS/src/Grid_map.cpp:217-220 (this copy only, like the reference's by-value Grid_map)
  void setcarData(const CarData &car_data);
  void setPlannerParams(const CarData &car_data, double pathLength, double step_car);
  // installs this copy's parameters on the shared device map if another copy's are there (host-side, no device work)
  void syncParams() const;
  // RS/src/Grid_map.cpp:239-284: true = collision or outside the map (hazard H5: the name says the opposite)
  bool isSingleStateCollisionFree(const State &current);
  // RS/src/Grid_map.cpp:286-317 (bounding-circle pre-filter; the reference's own, different, return convention)
  bool isSingleStateCollisionFreeImproved(const State &current);
  // DB/src/Grid_map.cpp:50-95 (the polygon argument is recomputed from the state on the device: it is a pure function
  // of the state and the car, DB/src/CarData.cpp:39-59)
  template <class Polygon>
  bool checkCollision(const State &state, const Polygon &) {
    return checkCollisionSat(state);
  }
  bool checkCollisionSat(const State &state);
  // RS/src/Grid_map.cpp:186-213
  std::tuple<int, int> toCellID(State s) const {
    return std::make_tuple((int)((s.x - originX) / resolution), (int)((s.y - originY) / resolution));
  }
  int toCellIndex(int x, int y) const { return x + y * (int)width; }
  bool isPointInBounds(int x, int y) const { return x >= 0 && x < (int)width && y >= 0 && y < (int)height; }
  bool isInCollision(int x, int y) const;

  // grid_map layers in the reference's memory layout (Eigen::MatrixXf, W rows x H cols, column-major)
  void layers(float *obstacle, float *distance) const;

  hpp_map *handle() const { return h_ ? h_->map : nullptr; }
  hpp_ctx *ctx() const { return h_ ? h_->ctx : nullptr; }

 private:
  void init(const int8_t *occ, std::size_t n, uint32_t w, uint32_t h, float res, double ox, double oy);
  std::shared_ptr<hpp_host::MapHandle> h_;
  std::shared_ptr<std::vector<int8_t>> occ_;  // host copy for the scalar cell predicates
  double resolution = 0, originX = 0, originY = 0, Max_origin_x = 0, Max_origin_y = 0;
  unsigned int width = 0, height = 0;
  CarData car_;
  double path_length_ = 4, step_car_ = 1;
};
