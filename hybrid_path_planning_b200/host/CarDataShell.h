// CarDataShell.h -- drop-in for the reference's CarData (RS/include/CarData.h:12-62, DB/include/CarData.h): the ROS 2
// node parameters maxSteerAngle, wheelBase, axleToFront, axleToBack, width.  The footprint circles / rectangle and the
// motion commands derived from them live on the library side (hpp_map_set_params); this class only carries the
// numbers, the one kinematic helper callers use directly, and the small geometry helpers the ROS nodes call for their
// RViz markers (vehicle outline, footprint circles) -- host arithmetic, a few flops per call.
#pragma once
#include <cmath>
#include <vector>

#include "../../include/hpp_b200.h"
#include "hpp_types.h"

// The outline is a geometry_msgs::msg::Polygon where ROS is present; elsewhere a stand-in with the same members.
#if __has_include(<geometry_msgs/msg/polygon.hpp>)
#include <geometry_msgs/msg/point32.hpp>
#include <geometry_msgs/msg/polygon.hpp>
namespace hpp_shell {
using Point32 = geometry_msgs::msg::Point32;
using Polygon = geometry_msgs::msg::Polygon;
}  // namespace hpp_shell
#else
namespace hpp_shell {
struct Point32 {
  float x = 0, y = 0, z = 0;
};
struct Polygon {
  std::vector<Point32> points;
};
}  // namespace hpp_shell
#endif

class CarData {
 public:
  double maxSteerAngle = 0.0, wheelBase = 0.0, axleToFront = 0.0, axleToBack = 0.0, width = 0.0;
  double steer_ratio = 6;
  double steer_step = 0.0;

  hpp_shell::Polygon vehicle_geometry;  // closed outline in the vehicle frame (float32 vertices), after createVehicleGeometry

  CarData() = default;
  CarData(double maxSteerAngle_, double wheelBase_, double axleToFront_, double axleToBack_, double width_)
      : maxSteerAngle(maxSteerAngle_), wheelBase(wheelBase_), axleToFront(axleToFront_), axleToBack(axleToBack_),
        width(width_), steer_step(maxSteerAngle_ / 6) {
    setCircles();
  }

  // rectangle (af, +-w/2), (-ab, +-w/2) as a closed 5-point polygon (RS/src/CarData.cpp:16-38, DB/src/CarData.cpp:15-37)
  void createVehicleGeometry() {
    const double xs[4] = {axleToFront, axleToFront, -axleToBack, -axleToBack};
    const double ys[4] = {width / 2.0, -width / 2.0, -width / 2.0, width / 2.0};
    for (int k = 0; k < 5; k++) {
      hpp_shell::Point32 p;
      p.x = (float)xs[k % 4];
      p.y = (float)ys[k % 4];
      vehicle_geometry.points.push_back(p);
    }
  }

  // the outline at a pose: rotation and translation in double, vertices rounded to float32
  // (RS/src/CarData.cpp:41-60, DB/src/CarData.cpp:40-59)
  hpp_shell::Polygon getVehicleGeometry_state(const State &state) const {
    hpp_shell::Polygon out;
    const double c = std::cos(state.heading), s = std::sin(state.heading);
    for (const auto &v : vehicle_geometry.points) {
      hpp_shell::Point32 p;
      p.x = (float)(state.x + v.x * c - v.y * s);
      p.y = (float)(state.y + v.x * s + v.y * c);
      out.points.push_back(p);
    }
    return out;
  }

  // bounding circle and the four footprint circles in the vehicle frame (RS/src/CarData.cpp:76-98)
  void setCircles() {
    bounding_c_ = Circle((axleToFront - axleToBack) / 2.0, 0.0,
                         std::sqrt(std::pow((axleToFront + axleToBack) / 2.0, 2) + std::pow(width / 2.0, 2)));
    const double shift = width / 4.0, r = std::sqrt(2 * std::pow(shift, 2));
    circles_.clear();
    circles_.emplace_back(-axleToBack + shift, width / 2.0, r);   // rear left
    circles_.emplace_back(-axleToBack + shift, -width / 2.0, r);  // rear right
    circles_.emplace_back(axleToFront - shift, width / 2.0, r);   // front left
    circles_.emplace_back(axleToFront - shift, -width / 2.0, r);  // front right
  }

  // vehicle frame -> world (RS/src/CarData.cpp:117-124)
  State local2Global(const State &reference, const State &target) const {
    const double c = std::cos(reference.heading), s = std::sin(reference.heading);
    return State(target.x * c - target.y * s + reference.x, target.x * s + target.y * c + reference.y,
                 reference.heading + target.heading);
  }

  // the circles at a pose (RS/src/CarData.cpp:100-115)
  std::vector<Circle> getCircles(const State &pos) const {
    std::vector<Circle> out;
    for (const auto &c : circles_) {
      const State g = local2Global(pos, State(c.x, c.y));
      out.emplace_back(g.x, g.y, c.r);
    }
    return out;
  }
  Circle getBoundingCircle(const State &state) const {
    const State g = local2Global(state, State(bounding_c_.x, bounding_c_.y));
    return Circle(g.x, g.y, bounding_c_.r);
  }

  // bicycle-model Euler step (RS/src/CarData.cpp:63-74): heading is not normalised
  State getVehicleStep(const State &s, double phi, double m = 1, double dt = 1e-2) const {
    State n;
    n.x = s.x + (m * dt) * std::cos(s.heading);
    n.y = s.y + (m * dt) * std::sin(s.heading);
    n.heading = s.heading + m * dt * std::tan(phi) / wheelBase;
    return n;
  }

  hpp_params params(double pathLength, double step_car) const {
    hpp_params p;
    p.max_steer = maxSteerAngle;
    p.wheel_base = wheelBase;
    p.axle_to_front = axleToFront;
    p.axle_to_back = axleToBack;
    p.width = width;
    p.path_length = pathLength;
    p.step_car = step_car;
    return p;
  }

 private:
  Circle bounding_c_;
  std::vector<Circle> circles_;
};
