// PathShells.h -- drop-ins for the two analytic path classes of the reference's public interface:
//   reeds_shepp_path  (RS/include/reeds_shepp_path.h:17-118)  calc_all_paths / calc_optimal_path / pi_2_pi
//   dubins_path       (DB/include/dubins_path.h:13-58)        calc_dubins_path / pi_2_pi / mod2pi
// Same class names, value types (PATH, Path) and signatures; the bodies call the C ABI (hpp_rs_all_paths,
// hpp_dubins_path): the 46 Reeds-Shepp closed forms, the ordered de-duplication, the six Dubins words and the sampling
// of the local course all run on the device (csrc/shots.cuh).  The per-family helpers of the reference class (LSL, LSR,
// ..., generate_local_course, interpolate) are implementation details of calc_all_paths there and are not part of
// what HybridAstar or the nodes call; they are not re-exported.
#pragma once
#include <cmath>
#include <string>
#include <vector>

#include "GridMapShell.h"

const double STEP_SIZE = 0.2;    // RS/include/reeds_shepp_path.h:13
const double MAX_LENGTH = 1000.0;

class PATH {  // RS/include/reeds_shepp_path.h:17-37
 public:
  std::vector<double> lengths;      // lengths of each part of the path (metres, signed)
  std::vector<std::string> ctypes;  // type of each part of the path ("S", "L", "R")
  double L;                         // total path length
  std::vector<double> x, y, yaw;    // sampled course
  std::vector<int> directions;      // forward: 1, backward: -1
  PATH() : L(0) {}
  PATH(const std::vector<double> &lengths_, const std::vector<std::string> &ctypes_, double L_,
       const std::vector<double> &x_, const std::vector<double> &y_, const std::vector<double> &yaw_,
       const std::vector<int> &directions_)
      : lengths(lengths_), ctypes(ctypes_), L(L_), x(x_), y(y_), yaw(yaw_), directions(directions_) {}
};

class reeds_shepp_path {
 public:
  reeds_shepp_path() = default;
  ~reeds_shepp_path() = default;
  // RS/src/reeds_shepp_path.cpp:40-95: every de-duplicated candidate, in the reference's generation order
  std::vector<PATH> calc_all_paths(double sx, double sy, double syaw, double gx, double gy, double gyaw, double maxc,
                                   double step_size = STEP_SIZE);
  // RS/src/reeds_shepp_path.cpp:13-38: the candidate with the smallest |L| (first one among equals)
  PATH calc_optimal_path(double sx, double sy, double syaw, double gx, double gy, double gyaw, double maxc,
                         double step_size = STEP_SIZE);
  double pi_2_pi(double theta);  // RS/src/reeds_shepp_path.cpp (angle wrap of the sampled yaw)
};

struct Path {  // DB/include/dubins_path.h:13-19
  double length = 0;
  std::vector<std::string> mode;
  std::vector<double> x, y, yaw;
};

class dubins_path {
 public:
  dubins_path() = default;
  ~dubins_path() = default;
  double pi_2_pi(double theta);
  double mod2pi(double theta);
  // DB/src/dubins_path.cpp:341-374: shortest of the six words, sampled with the normalised step
  Path calc_dubins_path(double sx, double sy, double syaw, double gx, double gy, double gyaw, double curv,
                        double step_size = 0.1);
};
