// shots.cuh -- device functions for K7 (Reeds-Shepp shot) and K8 (Dubins shot); one warp per start pose.
//
// Reeds-Shepp (reference: RS/src/reeds_shepp_path.cpp:40-860, RS/src/HybridAstar.cpp:110-313)
//   phase 1  the 46 closed-form evaluations of generate_path (:795-817: SCS 2, CSC 8, CCC 8, CCCC 8, CCSC 16,
//            CCSCC 4) run across the lanes of the warp (two rounds), each decoded as
//            (family, reflect/timeflip variant, backwards) from its position in the reference's fixed order;
//   phase 2  set_path's ordered, signed de-duplication (:98-141) is replayed candidate by candidate in that
//            order; the lanes compare one accepted path each and vote with a ballot;
//   phase 3  reeds_Path_Cost (HybridAstar.cpp:276-313) per accepted path, one lane each;
//   phase 4  > 5 paths: the pop order of the reference's std::priority_queue (cost only comparator, libstdc++
//            heap algorithms; equal costs are common) is replayed by one lane; <= 5 paths: generation order;
//   phase 5  candidates are sampled (generate_local_course/interpolate, :684-793) 32 samples per step across
//            the lanes, transformed to the world frame (:76-81) and collision-checked; the reference's sequential
//            "pd += d" accumulation is reproduced so sample counts and positions agree.
// Dubins (reference: DB/src/dubins_path.cpp:13-374, DB/src/HybridAstar.cpp:108-184): six words on six lanes,
//   strict-min selection, the same sampler, SAT collision per sample by the whole warp.
#pragma once
#include "collide.cuh"

#define RS_NCAND 46
#define RS_PI 3.14159265358979323846 /* M_PI: RS/include/reeds_shepp_path.h:15 */
#define RS_MAX_LENGTH 1000.0

struct rs_warp_smem {
  double len[RS_NCAND][5];  // normalised, signed segment lengths
  double L[RS_NCAND];       // normalised total length
  unsigned long long word[RS_NCAND];  // segment count and types packed into one integer (one compare in set_path)
  double cost[RS_NCAND];
  unsigned char typ[RS_NCAND][5];
  unsigned char nseg[RS_NCAND];
  unsigned char flag[RS_NCAND];
  unsigned char acc[RS_NCAND];    // accepted candidates, generation order
  unsigned char order[RS_NCAND];  // evaluation order: positions into acc[]
  int n_acc;
};

// ------------------------------------------------------------------------------------ scalar helpers
// (rs_M, rs_R and the sampler's sincos are deliberately NOT inlined: the planner kernel walks through its code once
// per A* iteration with a handful of warps, so its speed is set by instruction fetch -- B200's L1.5 instruction cache
// holds 32 KB -- and one shared copy of each libm expansion keeps the footprint down.  Same instructions, same bits.)
static __device__ __noinline__ double rs_M(double theta) {  // :846-860
  double phi = hpp_fmod(theta, 2.0 * RS_PI);
  if (phi < -RS_PI) phi += 2.0 * RS_PI;
  if (phi > RS_PI) phi -= 2.0 * RS_PI;
  return phi;
}
static __device__ __noinline__ void rs_R(double x, double y, double &r, double &theta) {  // :837-843
  r = hypot(x, y);
  theta = hpp_atan2(y, x);
}
__device__ __forceinline__ double pi_2_pi(double theta) {  // :821-834 (and DB/src/dubins_path.cpp:13-24)
  while (theta > RS_PI) theta -= 2.0 * RS_PI;
  while (theta < -RS_PI) theta += 2.0 * RS_PI;
  return theta;
}

enum { F_SLS, F_LSL, F_LSR, F_LRL, F_LRLRN, F_LRLRP, F_LRSL, F_LRSR, F_LRSLR };

// sp, cp: sin(phi), cos(phi), taken once per lane by the caller (every family but SLS starts from them)
static __device__ __noinline__ bool rs_family(int fam, double x, double y, double phi, double sp, double cp, double &t,
                                              double &u, double &v) {
  switch (fam) {
    case F_SLS: {  // :217-239
      phi = rs_M(phi);
      if (y > 0.0 && 0.0 < phi && phi < RS_PI * 0.99) {
        double xd = -y / hpp_tan(phi) + x;
        t = xd - hpp_tan(phi / 2.0);
        u = phi;
        v = sqrt((x - xd) * (x - xd) + y * y) - hpp_tan(phi / 2.0);
        return true;
      } else if (y < 0.0 && 0.0 < phi && phi < RS_PI * 0.99) {
        double xd = -y / hpp_tan(phi) + x;
        t = xd - hpp_tan(phi / 2.0);
        u = phi;
        v = -sqrt((x - xd) * (x - xd) + y * y) - hpp_tan(phi / 2.0);
        return true;
      }
      return false;
    }
    case F_LSL: {  // :143-157
      double uu, tt;
      rs_R(x - sp, y - 1.0 + cp, uu, tt);
      if (tt >= 0.0) {
        double vv = rs_M(phi - tt);
        if (vv >= 0.0) {
          t = tt; u = uu; v = vv;
          return true;
        }
      }
      return false;
    }
    case F_LSR: {  // :159-179
      double u1, t1;
      rs_R(x + sp, y - 1.0 - cp, u1, t1);
      u1 = u1 * u1;
      if (u1 >= 4.0) {
        double uu = sqrt(u1 - 4.0);
        double theta = hpp_atan2(2.0, uu);
        double tt = rs_M(t1 + theta);
        double vv = rs_M(tt - phi);
        if (tt >= 0.0 && vv >= 0.0) {
          t = tt; u = uu; v = vv;
          return true;
        }
      }
      return false;
    }
    case F_LRL: {  // :181-199
      double u1, t1;
      rs_R(x - sp, y - 1.0 + cp, u1, t1);
      if (u1 <= 4.0) {
        double uu = -2.0 * hpp_asin(0.25 * u1);
        double tt = rs_M(t1 + 0.5 * uu + RS_PI);
        double vv = rs_M(phi - tt + uu);
        if (tt >= 0.0 && uu <= 0.0) {
          t = tt; u = uu; v = vv;
          return true;
        }
      }
      return false;
    }
    case F_LRLRN:
    case F_LRLRP: {  // :374-417 with calc_tauOmega :359-372
      double xi = x + sp;
      double eta = y - 1.0 - cp;
      double uu, vq;
      if (fam == F_LRLRN) {
        double rho = 0.25 * (2.0 + sqrt(xi * xi + eta * eta));
        if (!(rho <= 1.0)) return false;
        uu = hpp_acos(rho);
        vq = -uu;
      } else {
        double rho = (20.0 - xi * xi - eta * eta) / 16.0;
        if (!(rho >= 0.0 && rho <= 1.0)) return false;
        uu = -hpp_acos(rho);
        if (!(uu >= -0.5 * RS_PI)) return false;
        vq = uu;
      }
      double delta = rs_M(uu - vq);
      double su, cu, sd, cd;  // each angle's sine and cosine once (the reference calls sin / cos term by term)
      hpp_sincos(uu, &su, &cu);
      hpp_sincos(delta, &sd, &cd);
      const double cvq = (fam == F_LRLRN) ? hpp_cos(vq) : cu;  // LRLRP: vq is uu itself
      double A = su - sd;
      double B = cu - cd - 1.0;
      double t1 = hpp_atan2(eta * A - xi * B, xi * A + eta * B);
      double t2 = 2.0 * (cd - cvq - cu) + 3.0;
      double tau = t2 < 0 ? rs_M(t1 + RS_PI) : rs_M(t1);
      double omega = rs_M(tau - uu + vq - phi);
      bool ok = (fam == F_LRLRN) ? (tau >= 0.0 && omega <= 0.0) : (tau >= 0.0 && omega >= 0.0);
      if (ok) {
        t = tau; u = uu; v = omega;
      }
      return ok;
    }
    case F_LRSR: {  // :475-494
      double xi = x + sp;
      double eta = y - 1.0 - cp;
      double rho, theta;
      rs_R(-eta, xi, rho, theta);
      if (rho >= 2.0) {
        double tt = theta;
        double uu = 2.0 - rho;
        double vv = rs_M(tt + 0.5 * RS_PI - phi);
        if (tt >= 0.0 && uu <= 0.0 && vv <= 0.0) {
          t = tt; u = uu; v = vv;
          return true;
        }
      }
      return false;
    }
    case F_LRSL: {  // :496-516
      double xi = x - sp;
      double eta = y - 1.0 + cp;
      double rho, theta;
      rs_R(xi, eta, rho, theta);
      if (rho >= 2.0) {
        double r = sqrt(rho * rho - 4.0);
        double uu = 2.0 - r;
        double tt = rs_M(theta + hpp_atan2(r, -2.0));
        double vv = rs_M(phi - 0.5 * RS_PI - tt);
        if (tt >= 0.0 && uu <= 0.0 && vv <= 0.0) {
          t = tt; u = uu; v = vv;
          return true;
        }
      }
      return false;
    }
    default: {  // F_LRSLR :626-650
      double xi = x + sp;
      double eta = y - 1.0 - cp;
      double rho, theta;
      rs_R(xi, eta, rho, theta);
      if (rho >= 2.0) {
        double uu = 4.0 - sqrt(rho * rho - 4.0);
        if (uu <= 0.0) {
          double tt = rs_M(hpp_atan2((4.0 - uu) * xi - 2.0 * eta, -2.0 * xi + (uu - 4.0) * eta));
          double vv = rs_M(tt - phi);
          if (tt >= 0.0 && vv >= 0.0) {
            t = tt; u = uu; v = vv;
            return true;
          }
        }
      }
      return false;
    }
  }
}

// formula `idx` (0..45) in the order generate_path issues them
__device__ __forceinline__ void rs_eval_formula(int idx, double x, double y, double phi, double xb, double yb,
                                                rs_warp_smem &S) {
  int fam, var, back = 0;
  if (idx < 2) { fam = F_SLS; var = idx ? 2 : 0; }
  else if (idx < 6) { fam = F_LSL; var = idx - 2; }
  else if (idx < 10) { fam = F_LSR; var = idx - 6; }
  else if (idx < 14) { fam = F_LRL; var = idx - 10; }
  else if (idx < 18) { fam = F_LRL; var = idx - 14; back = 1; }
  else if (idx < 22) { fam = F_LRLRN; var = idx - 18; }
  else if (idx < 26) { fam = F_LRLRP; var = idx - 22; }
  else if (idx < 30) { fam = F_LRSL; var = idx - 26; }
  else if (idx < 34) { fam = F_LRSR; var = idx - 30; }
  else if (idx < 38) { fam = F_LRSL; var = idx - 34; back = 1; }
  else if (idx < 42) { fam = F_LRSR; var = idx - 38; back = 1; }
  else { fam = F_LRSLR; var = idx - 42; }
  const double X = back ? xb : x, Y = back ? yb : y;
  const double ax = (var & 1) ? -X : X;                    // timeflip variants negate x
  const double ay = (var & 2) ? -Y : Y;                    // reflect variants negate y
  const double ap = (var == 1 || var == 2) ? -phi : phi;
  const double sg = (var & 1) ? -1.0 : 1.0;                // timeflip variants negate every length
  const bool swapLR = (var & 2) != 0;                      // reflect variants swap L and R
  double t = 0, u = 0, v = 0;
  double sp = 0.0, cp = 1.0;
  if (fam != F_SLS) hpp_sincos(ap, &sp, &cp);
  const bool ok = rs_family(fam, ax, ay, ap, sp, cp, t, u, v);
  S.flag[idx] = ok;
  if (!ok) return;
  const double hp = -0.5 * RS_PI;
  double l[5] = {0, 0, 0, 0, 0};
  const char *ty;
  int ns;
  switch (fam) {
    case F_SLS: ns = 3; ty = "SLS"; l[0] = t; l[1] = u; l[2] = v; break;
    case F_LSL: ns = 3; ty = "LSL"; l[0] = t; l[1] = u; l[2] = v; break;
    case F_LSR: ns = 3; ty = "LSR"; l[0] = t; l[1] = u; l[2] = v; break;
    case F_LRL: ns = 3; ty = "LRL";
      if (back) { l[0] = v; l[1] = u; l[2] = t; } else { l[0] = t; l[1] = u; l[2] = v; }
      break;
    case F_LRLRN: ns = 4; ty = "LRLR"; l[0] = t; l[1] = u; l[2] = -u; l[3] = v; break;
    case F_LRLRP: ns = 4; ty = "LRLR"; l[0] = t; l[1] = u; l[2] = u; l[3] = v; break;
    case F_LRSL: ns = 4;
      if (back) { ty = "LSRL"; l[0] = v; l[1] = u; l[2] = hp; l[3] = t; }
      else { ty = "LRSL"; l[0] = t; l[1] = hp; l[2] = u; l[3] = v; }
      break;
    case F_LRSR: ns = 4;
      if (back) { ty = "RSRL"; l[0] = v; l[1] = u; l[2] = hp; l[3] = t; }
      else { ty = "LRSR"; l[0] = t; l[1] = hp; l[2] = u; l[3] = v; }
      break;
    default: ns = 5; ty = "LRSLR"; l[0] = t; l[1] = hp; l[2] = u; l[3] = hp; l[4] = v; break;
  }
  S.nseg[idx] = (unsigned char)ns;
  unsigned long long word = (unsigned long long)ns;
  double L = 0.0;
  for (int i = 0; i < 5; i++) {
    const double li = (i < ns) ? l[i] * sg : 0.0;
    S.len[idx][i] = li;
    if (i < ns) L = L + fabs(li);  // set_path :123-127, same order
    char c = (i < ns) ? ty[i] : 0;
    if (swapLR) c = (c == 'L') ? 'R' : (c == 'R') ? 'L' : c;
    S.typ[idx][i] = (unsigned char)c;
    word = (word << 8) | (unsigned char)c;
  }
  S.word[idx] = word;
  S.L[idx] = L;
}

// phase 2: set_path :98-141, replayed in generation order over the formulas' results in S.  Warp-collective.
__device__ __forceinline__ void rs_dedup(rs_warp_smem &S) {
  const int lane = threadIdx.x & 31;
  int n_acc = 0;
  int mine = -1;  // the accepted candidate this lane compares against (acc[lane])
  // formulas that produced a path, as two ballot words (bit k <-> formula k)
  const unsigned f_lo = __ballot_sync(0xffffffffu, S.flag[lane] != 0);
  const unsigned f_hi = __ballot_sync(0xffffffffu, lane + 32 < RS_NCAND && S.flag[lane + 32] != 0);
  for (unsigned long long todo = ((unsigned long long)f_hi << 32) | f_lo; todo; todo &= todo - 1) {
    const int k = __ffsll((long long)todo) - 1;
    bool dup = false;
    if (mine >= 0 && S.word[mine] == S.word[k]) {
      const int ns = S.nseg[k];
      double sum_diff = 0.0;
      for (int i = 0; i < ns; i++) sum_diff += S.len[mine][i] - S.len[k][i];
      dup = sum_diff <= 0.01;
    }
    if (__any_sync(0xffffffffu, dup)) continue;
    if (S.L[k] >= RS_MAX_LENGTH) continue;
    // the reference asserts L >= 0.01 here (aborts when start ~ goal); release-build behaviour is kept
    if (lane == n_acc) mine = k;
    if (lane == 0) S.acc[n_acc] = (unsigned char)k;
    n_acc++;
  }
  if (lane == 0) S.n_acc = n_acc;
  __syncwarp();
}

// the goal in the start's frame, scaled by maxc (generate_path :799-805) and its "backwards" form (:329-330, :572-573)
struct rs_frame {
  double x, y, dth, xb, yb;
};
__device__ __forceinline__ rs_frame rs_make_frame(double sx, double sy, double syaw, double gx, double gy, double gyaw,
                                                  double maxc) {
  rs_frame f;
  const double dx = gx - sx, dy = gy - sy;
  f.dth = gyaw - syaw;
  const double c = hpp_cos(syaw), s = hpp_sin(syaw);
  f.x = (c * dx + s * dy) * maxc;
  f.y = (-s * dx + c * dy) * maxc;
  f.xb = f.x * hpp_cos(f.dth) + f.y * hpp_sin(f.dth);
  f.yb = f.x * hpp_sin(f.dth) - f.y * hpp_cos(f.dth);
  return f;
}

// phases 1+2: all candidates of (start -> goal), de-duplicated.  Warp-collective.
__device__ __forceinline__ void rs_generate(rs_warp_smem &S, double sx, double sy, double syaw, double gx, double gy,
                                            double gyaw, double maxc) {
  const int lane = threadIdx.x & 31;
  const rs_frame f = rs_make_frame(sx, sy, syaw, gx, gy, gyaw, maxc);
  // Lanes of different families run one after the other (different code), so every family is kept inside ONE of the two
  // rounds: LSL, LSR, LRL (8), LRSL (8), LRSR (8) = 32 formulas in the first, SLS, LRLRN, LRLRP, LRSLR = 14 in the second --
  // nine family passes instead of the twelve of "lane = formula index".  Results are filed under the formula index, so
  // the de-duplication sees them in the reference's order whatever lane worked them out.
  {
    int idx;  // first round: formulas 2..17 (LSL, LSR, LRL fwd + back), 26..29 / 34..37 (LRSL), 30..33 / 38..41 (LRSR)
    if (lane < 16) idx = 2 + lane;
    else if (lane < 20) idx = 26 + (lane - 16);
    else if (lane < 24) idx = 34 + (lane - 20);
    else if (lane < 28) idx = 30 + (lane - 24);
    else idx = 38 + (lane - 28);
    rs_eval_formula(idx, f.x, f.y, f.dth, f.xb, f.yb, S);
    int idx2 = -1;  // second round: 0..1 (SLS), 18..25 (LRLRN, LRLRP), 42..45 (LRSLR)
    if (lane < 2) idx2 = lane;
    else if (lane < 10) idx2 = 18 + (lane - 2);
    else if (lane < 14) idx2 = 42 + (lane - 10);
    if (idx2 >= 0) rs_eval_formula(idx2, f.x, f.y, f.dth, f.xb, f.yb, S);
  }
  __syncwarp();
  rs_dedup(S);
}

// ------------------------------------------------------------------------------------ sampler
// interpolate (:750-793 / DB/src/dubins_path.cpp:181-230)
static __device__ __noinline__ void rs_interpolate(double l, char m, double maxc, double ox, double oy, double oyaw,
                                               double &px, double &py, double &pyaw) {
  if (m == 'S') {
    double s, c;
    hpp_sincos(oyaw, &s, &c);
    px = ox + l / maxc * c;
    py = oy + l / maxc * s;
    pyaw = oyaw;
  } else {
    double sl, cl;
    hpp_sincos(l, &sl, &cl);
    const double ldx = sl / maxc;
    const double ldy = (m == 'L') ? (1.0 - cl) / maxc : (1.0 - cl) / (-maxc);
    double sn, cn;
    hpp_sincos(-oyaw, &sn, &cn);
    const double gdx = cn * ldx + sn * ldy;
    const double gdy = -sn * ldx + cn * ldy;
    px = ox + gdx;
    py = oy + gdy;
    pyaw = (m == 'L') ? oyaw + l : oyaw - l;
  }
}

// Walks generate_local_course (:684-748 / DB :232-294) for one path, 32 samples per step across the warp.
// vis(valid, idx, X, Y, YAW, local_x) is called warp-convergently; it returns true (uniformly) to stop early.
// Returns the number of samples the reference keeps after its trailing "px == 0.0" trim.
//   RS:      world = (cos(-q)x + sin(-q)y + sx, -sin(-q)x + cos(-q)y + sy, pi_2_pi(yaw + q))      (:76-81)
//   Dubins:  world = (x cos q - y sin q + sx,  x sin q + y cos q + sy,     pi_2_pi(yaw + q))       (DB :365-371)
// visit0 = false skips the visit of sample 0 (the start pose itself, the same for every path of a shot: the caller has
// looked at it once).  seg_counts (optional, zeroed by the caller): per segment, interior samples | (those with pd > 0) << 16.
template <bool DUBINS, class Visitor>
__device__ __forceinline__ int warp_sample_path(const double *len, const unsigned char *typ, int nseg, double maxc,
                                                double step, double sx, double sy, double syaw, Visitor &vis,
                                                bool visit0 = true, int *seg_counts = nullptr) {
  const int lane = threadIdx.x & 31;
  double rs_, rc_;
  if (DUBINS) hpp_sincos(syaw, &rs_, &rc_);
  else hpp_sincos(-syaw, &rs_, &rc_);
  auto to_world = [&](double px, double py, double pyaw, double &X, double &Y, double &YAW) {
    if (DUBINS) {
      X = px * rc_ - py * rs_ + sx;
      Y = px * rs_ + py * rc_ + sy;
    } else {
      X = rc_ * px + rs_ * py + sx;
      Y = -rs_ * px + rc_ * py + sy;
    }
    YAW = pi_2_pi(pyaw + syaw);
  };
  int last_nonzero = -1;  // largest sample index whose local x != 0.0 (warp-uniform after each step)
  // sample 0: the untouched (0,0,0) slot
  if (visit0) {
    double X, Y, YAW;
    to_world(0.0, 0.0, 0.0, X, Y, YAW);
    if (vis(lane == 0, lane, X, Y, YAW, 0.0)) return -1;
  }
  int ind = 1;  // next free slot after the interior samples so far
  double ox = 0.0, oy = 0.0, oyaw = 0.0, ll = 0.0;
  double ex = 0.0, ey = 0.0, eyaw = 0.0;  // end sample of the previous segment
  for (int i = 0; i < nseg; i++) {
    const double l = len[i];
    const char m = (char)typ[i];
    const double d = (l > 0.0) ? step : -step;
    ox = ex; oy = ey; oyaw = eyaw;  // px[ind] read before --ind (:711)
    double pd = (i >= 1 && (len[i - 1] * l) > 0) ? -d - ll : d - ll;
    const double al = fabs(l);
    const bool last_seg = i == nseg - 1;
    // interior samples: pd, pd + d, (pd + d) + d, ... while |pd| <= |l|
    for (;;) {
      // lane j needs pd advanced j times; every lane replays the chain so the rounding matches the reference
      double mine = pd;
      for (int j = 0; j < lane; j++) mine += d;
      const bool in_range = fabs(mine) <= al;
      // the reference's while loop stops at the first pd that fails, so only the leading run of lanes counts
      const unsigned vm = __ballot_sync(0xffffffffu, in_range);
      const int cnt = (vm == 0xffffffffu) ? 32 : (__ffs(~vm) - 1);
      // the first idle lane of the segment's last step takes the end-of-segment sample (:733-735, length l): it is the
      // next segment's origin, and for the last segment it survives at the slot after the interior samples
      const bool is_end = cnt < 32 && lane == cnt;
      const bool valid = lane < cnt || (is_end && last_seg);
      double px = 0, py = 0, pyaw = 0, X = 0, Y = 0, YAW = 0;
      if (lane < cnt || is_end) rs_interpolate(is_end ? l : mine, m, maxc, ox, oy, oyaw, px, py, pyaw);
      if (valid) to_world(px, py, pyaw, X, Y, YAW);
      const unsigned nz = __ballot_sync(0xffffffffu, valid && px != 0.0);
      if (nz) last_nonzero = ind + (31 - __clz(nz));
      if ((cnt || (cnt < 32 && last_seg)) && vis(valid, ind + lane, X, Y, YAW, px)) return -1;
      ind += cnt;
      if (seg_counts) {  // PATH::directions: interpolate() stores the sign of ITS length argument (:791) -- pd for the
                         // interior samples -- so segment i reports its interior samples and how many have pd > 0
        const unsigned pm = __ballot_sync(0xffffffffu, lane < cnt && mine > 0.0);
        if (lane == 0) seg_counts[i] += cnt + (__popc(pm) << 16);
      }
      // advance pd by cnt steps (all lanes identically)
      for (int j = 0; j < cnt; j++) pd += d;
      if (cnt < 32) {
        ex = __shfl_sync(0xffffffffu, px, cnt);
        ey = __shfl_sync(0xffffffffu, py, cnt);
        eyaw = __shfl_sync(0xffffffffu, pyaw, cnt);
        break;
      }
    }
    ll = l - pd - d;  // :731
  }
  if (nseg > 0) ind++;  // the last segment's end sample
  int n_keep = last_nonzero + 1;
  if (DUBINS) {
    if (n_keep < 1) n_keep = 1;   // while (px.size() > 1 && back == 0)
    if (n_keep <= 1) n_keep = 0;  // if (px.size() <= 1) return empty
  }
  return n_keep;
}

// libstdc++ heap replay (bits/stl_heap.h __push_heap/__adjust_heap) for comp(a,b) = a.cost > b.cost
__device__ __forceinline__ void stdheap_push(double *key, unsigned char *pay, int &n, double k, unsigned char p) {
  int hole = n++;
  int parent = (hole - 1) / 2;
  while (hole > 0 && key[parent] > k) {
    key[hole] = key[parent];
    pay[hole] = pay[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  key[hole] = k;
  pay[hole] = p;
}
__device__ __forceinline__ unsigned char stdheap_pop(double *key, unsigned char *pay, int &n) {
  const unsigned char top = pay[0];
  if (n > 1) {
    const double vk = key[n - 1];
    const unsigned char vp = pay[n - 1];
    const int len = n - 1;
    int hole = 0, second = 0;
    while (second < (len - 1) / 2) {
      second = 2 * (second + 1);
      if (key[second] > key[second - 1]) second--;
      key[hole] = key[second];
      pay[hole] = pay[second];
      hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
      second = 2 * (second + 1);
      key[hole] = key[second - 1];
      pay[hole] = pay[second - 1];
      hole = second - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > 0 && key[parent] > vk) {
      key[hole] = key[parent];
      pay[hole] = pay[parent];
      hole = parent;
      parent = (hole - 1) / 2;
    }
    key[hole] = vk;
    pay[hole] = vp;
  }
  n--;
  return top;
}

// ------------------------------------------------------------------------------------ visitors
struct circ_visitor {  // RS: circle-footprint check, lane-parallel over samples
  const hpp_geom *g;
  const circ_consts *cc;
  const float *dist;
  int hit_zero_idx;  // smallest index of a colliding sample whose local x == 0.0 (may be trimmed later)
  __device__ __forceinline__ bool operator()(bool valid, int idx, double X, double Y, double YAW, double lx) {
    int hit = 0;
    if (valid) hit = circles_collide_ilp(*g, *cc, dist, X, Y, YAW);
    const unsigned hard = __ballot_sync(0xffffffffu, hit && lx != 0.0);
    const unsigned soft = __ballot_sync(0xffffffffu, hit && lx == 0.0);
    if (soft) {
      int first = idx - (int)(threadIdx.x & 31) + (__ffs(soft) - 1);
      if (first < hit_zero_idx) hit_zero_idx = first;
    }
    return hard != 0;
  }
};

struct sat_visitor {  // Dubins: SAT check, the whole warp per sample
  const hpp_geom *g;
  const float *rx, *ry;
  const uint32_t *bits;
  int pitch;
  int hit_zero_idx;
  __device__ __forceinline__ bool operator()(bool valid, int idx, double X, double Y, double YAW, double lx) {
    const unsigned vm = __ballot_sync(0xffffffffu, valid);
    const int lane = threadIdx.x & 31;
    const int base = idx - lane;
    bool stop = false;
    for (unsigned mset = vm; mset && !stop;) {
      const int b = __ffs(mset) - 1;
      mset &= mset - 1;
      const double x = __shfl_sync(0xffffffffu, X, b), y = __shfl_sync(0xffffffffu, Y, b),
                   h = __shfl_sync(0xffffffffu, YAW, b), l0 = __shfl_sync(0xffffffffu, lx, b);
      if (sat_collide_warp(*g, rx, ry, bits, pitch, x, y, h)) {
        if (l0 != 0.0) stop = true;
        else if (base + b < hit_zero_idx) hit_zero_idx = base + b;
      }
    }
    return stop;
  }
};

struct store_visitor {  // writes State rows (x, y, heading, gridx, gridy)
  double *out;
  const hpp_geom *g;
  int cap;  // rows available
  __device__ __forceinline__ bool operator()(bool valid, int idx, double X, double Y, double YAW, double lx) {
    if (valid && idx < cap) {
      double *r = out + (size_t)idx * 5;
      r[0] = X;
      r[1] = Y;
      r[2] = YAW;
      r[3] = (double)(int)((X - g->ox) / g->res);
      r[4] = (double)(int)((Y - g->oy) / g->res);
    }
    return false;
  }
};

struct store3_visitor {  // writes (x, y, yaw) triples (unit-level parity hooks)
  double *out;
  size_t cap;
  size_t base;
  __device__ __forceinline__ bool operator()(bool valid, int idx, double X, double Y, double YAW, double lx) {
    if (valid && base + idx < cap) {
      double *r = out + (base + idx) * 3;
      r[0] = X;
      r[1] = Y;
      r[2] = YAW;
    }
    return false;
  }
};

// reeds_Path_Cost (RS/src/HybridAstar.cpp:276-313); lengths in metres = normalised / maxc (:87-90)
__device__ __forceinline__ double rs_path_cost(const rs_warp_smem &S, int k, double parent_cost, double maxc,
                                               double max_steer) {
  double cost = parent_cost;
  const int ns = S.nseg[k];
  double lm[5];
  for (int i = 0; i < ns; i++) lm[i] = S.len[k][i] / maxc;
  for (int i = 0; i < ns; i++) {
    if (lm[i] >= 0) cost += 1.0;
    else cost += fabs(lm[i]) * 10.0;
  }
  for (int i = 0; i + 1 < ns; i++)
    if (lm[i] * lm[i + 1] < 0) cost += 250.0;
  for (int i = 0; i < ns; i++)
    if (S.typ[k][i] != 'S') cost += max_steer * 15.0;
  return cost;
}

struct rs_shot_result {
  int found, n_paths, chosen, n_samples, cand;
  double cost;
};

// phases 3+4: cost of every accepted path and the order in which they are tried.  Warp-collective.
__device__ __forceinline__ void rs_order_paths(rs_warp_smem &S, double parent_cost, double maxc, double max_steer) {
  const int lane = threadIdx.x & 31;
  const int n_acc = S.n_acc;
  for (int e = lane; e < n_acc; e += 32) S.cost[e] = rs_path_cost(S, S.acc[e], parent_cost, maxc, max_steer);
  __syncwarp();
  if (lane == 0) {
    if (n_acc > 5) {  // RS/src/HybridAstar.cpp:131-142
      double key[RS_NCAND];
      unsigned char pay[RS_NCAND];
      int hn = 0;
      for (int e = 0; e < n_acc; e++) stdheap_push(key, pay, hn, S.cost[e], (unsigned char)e);
      for (int e = 0; e < n_acc; e++) S.order[e] = stdheap_pop(key, pay, hn);
    } else {
      for (int e = 0; e < n_acc; e++) S.order[e] = (unsigned char)e;
    }
  }
  __syncwarp();
}

// phase 5 for the e-th accepted path: samples it and checks the footprint; returns the kept sample count
// (>= 0) when the path is collision-free, -1 when it collides.  Warp-collective.
// hit0: the footprint verdict of sample 0 (the start pose) when the caller has it already, -1 to look here.
__device__ __forceinline__ int rs_check_path(const rs_warp_smem &S, int e, const hpp_geom &g, const circ_consts &cc,
                                             const float *dist, double sx, double sy, double syaw, double maxc,
                                             int hit0 = -1) {
  const int k = S.acc[e];
  circ_visitor vis;
  vis.g = &g; vis.cc = &cc; vis.dist = dist;
  vis.hit_zero_idx = hit0 > 0 ? 0 : 0x7fffffff;  // sample 0 has local x == 0.0: a hit there counts if it is kept
  const int n_keep = warp_sample_path<false>(S.len[k], S.typ[k], S.nseg[k], maxc, maxc /* step_size(1) * maxc */, sx, sy,
                                             syaw, vis, hit0 < 0);
  return ((n_keep < 0) || (vis.hit_zero_idx < n_keep)) ? -1 : n_keep;
}

// Phase 5 for a whole shot, GROUPED: the warp checks 32 / G candidate paths at once, one per group of G lanes, each
// group pulling its next path from *next_path.  rs_check_path walks one path segment by segment with the whole warp:
// a segment is one dependent step whatever its length, and with 1 m between samples most segments fill a few lanes
// (3-5 steps per path, ~10 paths per warp in the planner).  Here a step serves 32 / G paths.  The samples of a path
// are the same whatever the group width: pd advances by the reference's repeated addition (:722-729) in both forms,
// and the verdict of a path (-1 collides / kept sample count) does not depend on the order its samples are looked at.
// verdict[e] receives the verdict of the e-th accepted path, *any_free is set when some path is free.  hit0: footprint
// verdict of sample 0 (the start pose, common to all paths).  Warp-collective; every warp of a CTA may run it at once.
template <int G>
__device__ __forceinline__ void rs_check_paths_grouped(const rs_warp_smem &S, int n_acc, int *next_path, int *verdict,
                                                       int *any_free, const hpp_geom &g, const circ_consts &cc,
                                                       const float *dist, double sx, double sy, double syaw, double maxc,
                                                       int hit0) {
  static_assert(G == 4 || G == 8 || G == 16 || G == 32, "group width");
  constexpr unsigned FULLM = 0xffffffffu;
  constexpr unsigned GMASK = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
  const int lane = threadIdx.x & 31, sub = lane & (G - 1), gbase = lane & ~(G - 1);
  const double step = maxc;  // step_size (1.0) * maxc
  double rs_, rc_;
  hpp_sincos(-syaw, &rs_, &rc_);
  // per-group state (identical in the lanes of a group)
  int e = -1, k = 0, i = 0, nseg = 0, ind = 1, last_nonzero = -1, hit_zero_idx = 0x7fffffff;
  bool exhausted = false, seg_start = true;
  double pd = 0.0, ll = 0.0, ox = 0.0, oy = 0.0, oyaw = 0.0, ex = 0.0, ey = 0.0, eyaw = 0.0;
  for (;;) {
    // groups without a path take the next one
    const bool want = e < 0 && !exhausted;
    int ne = -1;
    if (want && sub == 0) ne = atomicAdd(next_path, 1);
    ne = __shfl_sync(FULLM, ne, gbase);
    if (want) {
      if (ne >= n_acc) {
        exhausted = true;
      } else {
        e = ne;
        k = S.acc[e];
        nseg = S.nseg[k];
        i = 0;
        ind = 1;
        last_nonzero = -1;
        hit_zero_idx = hit0 > 0 ? 0 : 0x7fffffff;  // sample 0 has local x == 0.0: a hit there counts if it is kept
        ll = 0.0;
        ex = ey = eyaw = 0.0;
        seg_start = true;
      }
    }
    if (!__any_sync(FULLM, e >= 0)) break;
    const bool act = e >= 0 && nseg > 0;
    // ---- one step of the group's current segment (generate_local_course, :697-735)
    const double l = act ? S.len[k][i] : 0.0;
    const char m = act ? (char)S.typ[k][i] : 'S';
    const double d = (l > 0.0) ? step : -step;
    if (act && seg_start) {
      ox = ex; oy = ey; oyaw = eyaw;  // px[ind] read before --ind (:711)
      pd = (i >= 1 && (S.len[k][i - 1] * l) > 0) ? -d - ll : d - ll;
      seg_start = false;
    }
    double mine = pd;
    for (int j = 0; j < sub; j++) mine += d;  // the reference's repeated addition, so the rounding matches
    const bool in_range = act && fabs(mine) <= fabs(l);
    const unsigned vm = (__ballot_sync(FULLM, in_range) >> gbase) & GMASK;
    const int cnt = (vm == GMASK) ? G : (__ffs(~vm) - 1);  // the leading run of lanes: the while loop stops at the first miss
    const bool last_seg = i == nseg - 1;
    const bool is_end = act && cnt < G && sub == cnt;       // the end-of-segment sample (:733-735), length l
    const bool valid = act && (sub < cnt || (is_end && last_seg));
    double px = 0, py = 0, pyaw = 0;
    if (act && (sub < cnt || is_end)) rs_interpolate(is_end ? l : mine, m, maxc, ox, oy, oyaw, px, py, pyaw);
    int hit = 0;
    if (valid) {  // world transform (:76-81) and footprint
      const double X = rc_ * px + rs_ * py + sx, Y = -rs_ * px + rc_ * py + sy, YAW = pi_2_pi(pyaw + syaw);
      hit = circles_collide_ilp(g, cc, dist, X, Y, YAW);
    }
    const unsigned nz = (__ballot_sync(FULLM, valid && px != 0.0) >> gbase) & GMASK;
    const unsigned hard = (__ballot_sync(FULLM, hit && px != 0.0) >> gbase) & GMASK;
    const unsigned soft = (__ballot_sync(FULLM, hit && px == 0.0) >> gbase) & GMASK;
    if (nz) last_nonzero = ind + (31 - __clz(nz));
    if (soft) hit_zero_idx = min(hit_zero_idx, ind + __ffs(soft) - 1);
    ind += cnt;
    for (int j = 0; j < cnt; j++) pd += d;
    const bool seg_done = cnt < G;
    const double bx = __shfl_sync(FULLM, px, gbase + (cnt < G ? cnt : 0));
    const double by = __shfl_sync(FULLM, py, gbase + (cnt < G ? cnt : 0));
    const double byaw = __shfl_sync(FULLM, pyaw, gbase + (cnt < G ? cnt : 0));
    if (act && seg_done) {
      ex = bx; ey = by; eyaw = byaw;
      ll = l - pd - d;  // :731
      i++;
      seg_start = true;
    }
    const bool path_done = e >= 0 && (!act || hard != 0 || (seg_done && i >= nseg));
    if (path_done) {
      int res = -1;
      if (!hard) {
        const int n_keep = last_nonzero + 1;  // the trailing "px == 0.0" trim (:738-744)
        res = (hit_zero_idx < n_keep) ? -1 : n_keep;
      }
      if (sub == 0) {
        verdict[e] = res;
        if (res >= 0) *any_free = 1;
      }
      e = -1;
    }
  }
}

// folds the verdict of the o-th path into the result; returns true when the search is over
// (> 5 paths: first collision-free path in pop order, :226-271; else strict minimum, first wins, :188-200)
__device__ __forceinline__ bool rs_fold_path(const rs_warp_smem &S, int o, int n_keep, rs_shot_result &R, double &best) {
  if (n_keep < 0) return false;
  const int e = S.order[o];
  if (S.n_acc > 5) {
    R.found = 1; R.chosen = e; R.cand = S.acc[e]; R.cost = S.cost[e]; R.n_samples = n_keep;
    return true;
  }
  if (S.cost[e] < best) {
    best = S.cost[e];
    R.found = 1; R.chosen = e; R.cand = S.acc[e]; R.cost = best; R.n_samples = n_keep;
  }
  return false;
}

// phases 3-5 for one start pose.  Warp-collective; S must hold rs_generate's output.
__device__ __forceinline__ rs_shot_result rs_select(rs_warp_smem &S, const hpp_geom &g, const circ_consts &cc,
                                                    const float *dist, double sx, double sy, double syaw,
                                                    double parent_cost, double maxc, double max_steer) {
  rs_shot_result R;
  R.found = 0; R.chosen = -1; R.n_samples = 0; R.cand = -1; R.cost = 0.0;
  const int n_acc = S.n_acc;
  R.n_paths = n_acc;
  if (n_acc == 0) return R;
  rs_order_paths(S, parent_cost, maxc, max_steer);
  double best = INFINITY;
  for (int o = 0; o < n_acc; o++)
    if (rs_fold_path(S, o, rs_check_path(S, S.order[o], g, cc, dist, sx, sy, syaw, maxc), R, best)) break;
  return R;
}

// ------------------------------------------------------------------------------------ Dubins
__device__ __forceinline__ double db_mod2pi(double theta) {  // DB/src/dubins_path.cpp:26-29
  return theta - 2.0 * RS_PI * floor(theta / RS_PI / 2.0);
}

// word w in the reference's order LSL, RSR, LSR, RSL, RLR, LRL (:306-312); returns false for its NaN result
static __device__ __noinline__ bool db_word(int w, double alpha, double beta, double dist, double &t, double &p, double &q) {
  double sin_a, cos_a, sin_b, cos_b;
  hpp_sincos(alpha, &sin_a, &cos_a);
  hpp_sincos(beta, &sin_b, &cos_b);
  const double cos_a_b = hpp_cos(alpha - beta);
  switch (w) {
    case 0: {  // LSL :33-57
      double pp = 2 + dist * dist - 2 * cos_a_b + 2 * dist * (sin_a - sin_b);
      if (pp < 0) return false;
      pp = sqrt(pp);
      double den = dist + sin_a - sin_b;
      t = db_mod2pi(-alpha + hpp_atan2(cos_b - cos_a, den));
      p = pp;
      q = db_mod2pi(beta - hpp_atan2(cos_b - cos_a, den));
      return true;
    }
    case 1: {  // RSR :59-83
      double pp = 2 + dist * dist - 2 * cos_a_b + 2 * dist * (sin_b - sin_a);
      if (pp < 0) return false;
      pp = sqrt(pp);
      double den = dist - sin_a + sin_b;
      t = db_mod2pi(alpha - hpp_atan2(cos_a - cos_b, den));
      p = pp;
      q = db_mod2pi(-beta + hpp_atan2(cos_a - cos_b, den));
      return true;
    }
    case 2: {  // LSR :85-109
      double pp = -2 + dist * dist + 2 * cos_a_b + 2 * dist * (sin_a + sin_b);
      if (pp < 0) return false;
      pp = sqrt(pp);
      double rec = hpp_atan2(-cos_a - cos_b, dist + sin_a + sin_b) - hpp_atan2(-2.0, pp);
      t = db_mod2pi(-alpha + rec);
      p = pp;
      q = db_mod2pi(-db_mod2pi(beta) + rec);
      return true;
    }
    case 3: {  // RSL :111-135
      double pp = -2 + dist * dist + 2 * cos_a_b - 2 * dist * (sin_a + sin_b);
      if (pp < 0) return false;
      pp = sqrt(pp);
      double rec = hpp_atan2(cos_a + cos_b, dist - sin_a - sin_b) - hpp_atan2(2.0, pp);
      t = db_mod2pi(alpha - rec);
      p = pp;
      q = db_mod2pi(beta - rec);
      return true;
    }
    case 4: {  // RLR :137-157
      double rec = (6.0 - dist * dist + 2.0 * cos_a_b + 2.0 * dist * (sin_a - sin_b)) / 8.0;
      if (fabs(rec) > 1.0) return false;
      p = db_mod2pi(2 * RS_PI - hpp_acos(rec));
      t = db_mod2pi(alpha - hpp_atan2(cos_a - cos_b, dist - sin_a + sin_b) + db_mod2pi(p / 2.0));
      q = db_mod2pi(alpha - beta - t + db_mod2pi(p));
      return true;
    }
    default: {  // LRL :159-179
      double rec = (6.0 - dist * dist + 2.0 * cos_a_b + 2.0 * dist * (sin_b - sin_a)) / 8.0;
      if (fabs(rec) > 1.0) return false;
      p = db_mod2pi(2 * RS_PI - hpp_acos(rec));
      t = db_mod2pi(-alpha - hpp_atan2(cos_a - cos_b, dist + sin_a - sin_b) + p / 2.0);
      q = db_mod2pi(db_mod2pi(beta) - alpha - t + db_mod2pi(p));
      return true;
    }
  }
}

struct db_best {
  double len[3];
  unsigned char typ[3];
  int nmode;        // 0 = no feasible word
  double best_cost; // normalised |t|+|p|+|q|
};

// planning_from_origin's word selection (:296-338) for start -> goal.  Warp-collective.
__device__ __forceinline__ db_best db_plan(double sx, double sy, double syaw, double gx, double gy, double gyaw,
                                           double curv) {
  const int lane = threadIdx.x & 31;
  // calc_dubins_path :346-356
  gx -= sx;
  gy -= sy;
  double sin_yaw, cos_yaw;
  hpp_sincos(syaw, &sin_yaw, &cos_yaw);
  const double le_x = gx * cos_yaw + gy * sin_yaw;
  const double le_y = -gx * sin_yaw + gy * cos_yaw;
  const double le_yaw = gyaw - syaw;
  const double D = hypot(le_x, le_y);
  const double d = D * curv;
  const double theta = db_mod2pi(hpp_atan2(le_y, le_x));
  const double alpha = db_mod2pi(-theta);
  const double beta = db_mod2pi(le_yaw - theta);
  double t = 0, p = 0, q = 0;
  bool ok = false;
  if (lane < 6) ok = db_word(lane, alpha, beta, d, t, p, q);
  ok = ok && !isnan(t);  // :322
  const double cost = fabs(t) + fabs(p) + fabs(q);
  db_best B;
  B.nmode = 0;
  B.best_cost = INFINITY;
  B.len[0] = B.len[1] = B.len[2] = 0.0;
  B.typ[0] = B.typ[1] = B.typ[2] = 0;
  const char *modes[6] = {"LSL", "RSR", "LSR", "RSL", "RLR", "LRL"};
  for (int w = 0; w < 6; w++) {  // strict '>' : first of equal costs wins (:328)
    const bool okw = __shfl_sync(0xffffffffu, (int)ok, w);
    const double cw = __shfl_sync(0xffffffffu, cost, w);
    const double tw = __shfl_sync(0xffffffffu, t, w), pw = __shfl_sync(0xffffffffu, p, w),
                 qw = __shfl_sync(0xffffffffu, q, w);
    if (okw && B.best_cost > cw) {
      B.best_cost = cw;
      B.len[0] = tw; B.len[1] = pw; B.len[2] = qw;
      B.typ[0] = modes[w][0]; B.typ[1] = modes[w][1]; B.typ[2] = modes[w][2];
      B.nmode = 3;
    }
  }
  return B;
}
