// CarDataShell.h -- drop-in for the reference's CarData (RS/include/CarData.h:12-62, DB/include/CarData.h): the ROS 2
// node parameters maxSteerAngle, wheelBase, axleToFront, axleToBack, width.  The footprint circles / rectangle and the
// motion commands derived from them live on the library side (hpp_map_set_params); this class only carries the
// numbers and the one kinematic helper callers use directly.
#pragma once
#include <cmath>

#include "../../include/hpp_b200.h"
#include "hpp_types.h"

class CarData {
 public:
  double maxSteerAngle = 0.0, wheelBase = 0.0, axleToFront = 0.0, axleToBack = 0.0, width = 0.0;
  double steer_ratio = 6;
  double steer_step = 0.0;

  CarData() = default;
  CarData(double maxSteerAngle_, double wheelBase_, double axleToFront_, double axleToBack_, double width_)
      : maxSteerAngle(maxSteerAngle_), wheelBase(wheelBase_), axleToFront(axleToFront_), axleToBack(axleToBack_),
        width(width_), steer_step(maxSteerAngle_ / 6) {}

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
};
