// expand.cuh -- device functions for K6: kinematic successor expansion.
// HybridAstar::GetnextNeighbours + PathCost (RS/src/HybridAstar.cpp:25-107, DB/src/HybridAstar.cpp:23-105)
// with CarData::getVehicleStep (RS/src/CarData.cpp:63-74).
#pragma once
#include "collide.cuh"

struct expand_consts {
  int n_cmd;
  int sim;                 // simulationLength (pathLength) as an int loop bound
  double sim_len;          // pathLength
  double step;             // step_car
  double steer[HPP_MAX_COMMANDS];
  double dir[HPP_MAX_COMMANDS];
  double dtheta[HPP_MAX_COMMANDS];  // tan(steer) / wheelBase, evaluated on the host with the reference's libm
  double dir_change;       // 250 (RS) / 200 (Dubins)   (HybridAstar.h:81-85 / :69-73)
};

// one Euler step of the bicycle model (RS/src/CarData.cpp:63-74): heading is NOT normalised
__device__ __forceinline__ void vehicle_step(double &x, double &y, double &h, double m, double dt, double dtheta) {
  double s, c;
  hpp_sincos(h, &s, &c);
  x = x + (m * dt * c);
  y = y + (m * dt * s);
  h = h + (m * dt * dtheta);
}

// PathCost (RS/src/HybridAstar.cpp:81-107)
__device__ __forceinline__ double path_cost(const expand_consts &ec, int cmd, double parent_cost, double parent_steer,
                                            int parent_dir) {
  double cost = parent_cost;
  if (ec.dir[cmd] == 1) cost += ec.sim_len;
  else cost += ec.sim_len * 10.0;  // reverse
  if ((double)parent_dir != ec.dir[cmd]) cost += ec.dir_change;
  cost += fabs(ec.steer[cmd]) * 15.0;                  // steerAngle
  cost += fabs(ec.steer[cmd] - parent_steer) * 20.0;   // steerAngleChange
  return cost;
}
