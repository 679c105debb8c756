// collide.cuh -- device functions for the two vehicle-footprint collision checks.
//   circles_collide : Grid_map::isSingleStateCollisionFree (RS/src/Grid_map.cpp:239-284) over the "distance"
//                     layer with grid_map::GridMap::{isInside, atPosition(INTER_LINEAR)} semantics
//                     (grid_map_core GridMapMath.cpp / GridMap.cpp, un-vendored dependency) and
//                     CarData::getCircles/local2Global (RS/src/CarData.cpp:100-124).
//   sat_collide_warp: Grid_map::checkCollision + createObstaclePolygon (DB/src/Grid_map.cpp:50-123),
//                     fop::SATCollisionChecker::check_collision (DB/src/sat_collision_checker.cc:131-161),
//                     CarData::getVehicleGeometry_state (DB/src/CarData.cpp:40-59) and the planner's
//                     "|| !isPointInBounds(gridx, gridy)" (DB/src/HybridAstar.cpp:52).
// All arithmetic is fp64 in the reference's operation order; the library is compiled with -fmad=false so
// nothing is contracted into FMAs (the reference build targets baseline x86-64: no FMA).
#pragma once
#include <math.h>
#include <stdlib.h>

#include "hpp_internal.cuh"

struct circ_consts {
  double cx[4], cy[4], cr;
  // the float32 filter over the float distance layer (circles_filter_f32 below; hpp_make_circ_consts fills them)
  double fx0, fy0;      // centre of cell (0,0): pos + (0.5 len - 0.5 res)
  float fcu[4], fcv[4]; // circle offsets in cells
  float f_lo, f_hi;     // certain collision below f_lo, certainly clear above f_hi (metres)
  int f_on;
};
// CarData::setCircles' bounding circle (RS/src/CarData.cpp:78-81), for isSingleStateCollisionFreeImproved
struct bound_consts {
  double bx, br;
};

// grid_map::getIndexFromPosition: ((p - 0.5*len) - pos) / res, truncated after negation
__device__ __forceinline__ void gm_index(const hpp_geom &g, double px, double py, int &i, int &j) {
  double vx = hpp_div_res(g, (px - 0.5 * g.lenx) - g.posx);
  double vy = hpp_div_res(g, (py - 0.5 * g.leny) - g.posy);
  i = (int)(-vx);
  j = (int)(-vy);
}

// grid_map::getPositionFromIndex (cell centre); leaves x,y untouched when the index is out of range
__device__ __forceinline__ void gm_position(const hpp_geom &g, int i, int j, double &x, double &y) {
  if (i < 0 || j < 0 || i >= g.W || j >= g.H) return;
  x = (g.posx + (0.5 * g.lenx - 0.5 * g.res)) + g.res * (double)(-i);
  y = (g.posy + (0.5 * g.leny - 0.5 * g.res)) + g.res * (double)(-j);
}

// one footprint circle at world position (px,py), in two halves so that the taps of several circles can be in
// flight together: circle_prep resolves the taps, circle_finish interpolates and compares
struct circle_taps {
  int mode;           // 0 = outside the map (collision), 1 = INTER_NEAREST fallback (tap l[0]), 2 = bilinear
  long l[4];
  double rx, ry;
  int ia, ib, ja, jb; // cell indices of the taps: l[0] = (ia, ja), l[1] = (ib, ja), l[2] = (ia, jb), l[3] = (ib, jb)
};

__device__ __forceinline__ void circle_prep(const hpp_geom &g, double px, double py, circle_taps &t) {
  // isInside -> checkIfPositionWithinMap
  double tx = -((px - g.posx) - 0.5 * g.lenx);
  double ty = -((py - g.posy) - 0.5 * g.leny);
  t.mode = 0;
  t.l[0] = t.l[1] = t.l[2] = t.l[3] = 0;
  t.rx = t.ry = 0.0;
  t.ia = t.ib = t.ja = t.jb = 0;
  if (!(tx >= 0.0 && ty >= 0.0 && tx < g.lenx && ty < g.leny)) return;
  // atPositionLinearInterpolated
  int i0, j0;
  gm_index(g, px, py, i0, j0);
  double cx = 0.0, cy = 0.0;
  gm_position(g, i0, j0, cx, cy);
  const bool dirx = px >= cx, diry = py >= cy;
  const int i1 = dirx ? i0 - 1 : i0 + 1;  // second index along x
  const int j2 = diry ? j0 - 1 : j0 + 1;  // third index along y
  // f[0] is the tap with the smallest x,y (largest i,j) of the four
  const int ia = dirx ? i0 : i1, ib = dirx ? i1 : i0;  // ia: larger index (smaller x)
  const int ja = diry ? j0 : j2, jb = diry ? j2 : j0;
  const long W = g.W, nn = (long)g.W * g.H;
  const long l0 = ja * W + ia, l1 = ja * W + ib, l2 = jb * W + ia, l3 = jb * W + ib;
  if (l0 < 0 || l0 >= nn || l1 < 0 || l1 >= nn || l2 < 0 || l2 >= nn || l3 < 0 || l3 >= nn) {
    t.mode = 1;  // INTER_NEAREST fallback
    t.l[0] = t.l[1] = t.l[2] = t.l[3] = (long)j0 * W + i0;
    t.ia = t.ib = i0;
    t.ja = t.jb = j0;
  } else {
    t.mode = 2;
    t.l[0] = l0; t.l[1] = l1; t.l[2] = l2; t.l[3] = l3;
    t.ia = ia; t.ib = ib; t.ja = ja; t.jb = jb;
    double bx = cx, by = cy;
    gm_position(g, ia, ja, bx, by);
    t.rx = hpp_div_res(g, px - bx);
    t.ry = hpp_div_res(g, py - by);
  }
}

__device__ __forceinline__ int circle_finish(const circle_taps &t, float f0, float f1, float f2, float f3, double r) {
  if (t.mode == 0) return 1;
  float value = f0;
  if (t.mode == 2) {
    const double fx = 1.0 - t.rx, fy = 1.0 - t.ry;
    value = (float)((double)f0 * fx * fy + (double)f1 * t.rx * fy + (double)f2 * fx * t.ry + (double)f3 * t.rx * t.ry);
  }
  return (double)value < r;
}

// 1 = outside the map or clearance < r
__device__ __forceinline__ int circle_hits(const hpp_geom &g, const float *__restrict__ dist, double px, double py,
                                           double r) {
  circle_taps t;
  circle_prep(g, px, py, t);
  if (t.mode == 0) return 1;
  const float f0 = __ldg(dist + t.l[0]);
  float f1 = f0, f2 = f0, f3 = f0;
  if (t.mode == 2) {
    f1 = __ldg(dist + t.l[1]);
    f2 = __ldg(dist + t.l[2]);
    f3 = __ldg(dist + t.l[3]);
  }
  return circle_finish(t, f0, f1, f2, f3, r);
}

// The same circle over the COMPACT layer (hpp_map::d_d2u8: squared cell distance clamped to 255, one byte per cell in
// 8 x 4 micro-tiles; lut[v] = sqrtf((float)v) * res_f, the float the distance layer holds for that cell).
// Why the clamp cannot change a verdict: the Euclidean distance transform is 1-Lipschitz, and the four taps lie within
// sqrt(2) cells of each other.  If a tap is clamped (true distance >= sqrt(255) = 15.97 cells) every tap -- clamped or
// not -- is at least 14.55 cells, so the interpolated clearance is at least 14.55 * res either way, and the caller only
// uses this path for circle radii below 14 cells: "no collision" with and without the clamp.  If no tap is clamped the
// values are the layer's, bit for bit.
__device__ __forceinline__ int circle_hits_u8(const hpp_geom &g, const uint8_t *__restrict__ d2, int tiles_x,
                                              const float *lut, double px, double py, double r) {
  circle_taps t;
  circle_prep(g, px, py, t);
  if (t.mode == 0) return 1;
  size_t a0, a1, a2, a3;
  if ((unsigned)t.ia < (unsigned)g.W && (unsigned)t.ib < (unsigned)g.W) {  // then ja, jb are rows of the map too
    a0 = hpp_d2_addr(t.ia, t.ja, tiles_x);
    a1 = hpp_d2_addr(t.ib, t.ja, tiles_x);
    a2 = hpp_d2_addr(t.ia, t.jb, tiles_x);
    a3 = hpp_d2_addr(t.ib, t.jb, tiles_x);
  } else {  // grid_map's linear index aliases a column index of -1 / W into the neighbouring row: follow it
    a0 = hpp_d2_addr((int)(t.l[0] % g.W), (int)(t.l[0] / g.W), tiles_x);
    a1 = hpp_d2_addr((int)(t.l[1] % g.W), (int)(t.l[1] / g.W), tiles_x);
    a2 = hpp_d2_addr((int)(t.l[2] % g.W), (int)(t.l[2] / g.W), tiles_x);
    a3 = hpp_d2_addr((int)(t.l[3] % g.W), (int)(t.l[3] / g.W), tiles_x);
  }
  const float f0 = lut[__ldg(d2 + a0)];
  float f1 = f0, f2 = f0, f3 = f0;
  if (t.mode == 2) {
    f1 = lut[__ldg(d2 + a1)];
    f2 = lut[__ldg(d2 + a2)];
    f3 = lut[__ldg(d2 + a3)];
  }
  return circle_finish(t, f0, f1, f2, f3, r);
}

// ---------------------------------------------------------------------------------------- float32 filter
// Most poses are nowhere near the verdict boundary: their clearance differs from the circle radius by centimetres,
// while the reference's fp64 arithmetic is needed only to place the last 1e-15.  circles_filter_u8 evaluates the same
// bilinear clearance in float32 (a fifth of the instructions) and returns a verdict only when it is CERTAIN to be
// the reference's; everything else (2 = undecided) goes through the exact path above.  Why the answer is certain:
//  * Away from the map border (index 2 .. W-4, checked here) the reference's value is F(p) + e_ref, F = the bilinear
//    interpolation of the layer between cell centres, a CONTINUOUS function of the position p (the same on both sides
//    of a cell boundary, so it does not matter which taps either side picks), and e_ref is its fp64 rounding plus the
//    final cast to float: <= 2^-24 * V, V = the largest tap.
//  * The layer is an exact Euclidean distance transform times the resolution, so neighbouring taps differ by at most
//    res (1 + 2^-19) -- clamped at 255 or not -- and |F(p') - F(p)| <= res (1 + 2^-19) (|du| + |dv|) for a position
//    error (du, dv) in cells.
//  * The float32 position is off by at most E = 2^-22 D + C (e_sc + 2^-21) cells per axis: D = max(W, H) + 4 bounds the
//    index coordinate (its cast and the two fused multiply-adds that add the circle offset round by half an ulp at that
//    magnitude each: 3 * 2^-24 D <= 2^-22 D), C = max_k (|cx_k| + |cy_k|) / res
//    is the circle offset in cells, e_sc = 4e-6 bounds sine and cosine: heading cast (|h| <= 16: 9.6e-7), two-constant
//    range reduction (2.4e-7), clamp to the documented __sinf/__cosf range (5e-7), and their error there (2^-21.41).
//  * The float32 interpolation itself works in cells on sqrt.approx(d2) (relative error 2^-23; the layer holds
//    RN(RN(sqrt d2) * res_f), another 2^-23) and takes 10 operations on values <= V: <= 48 * 2^-24 * V in all.
// So |res * value32 - value_ref| <= M = 2 res (1 + 2^-19) E + 48 * 2^-24 V + 2^-23 r, and the host passes
// r_lo = (r - 1.5 M) / res rounded down and r_hi = (r + 1.5 M) / res rounded up: value32 < r_lo means the reference saw
// value < r, value32 > r_hi means it did not.  A collision on ANY circle is the reference's verdict, so one certain hit decides the pose; a
// free verdict needs all four circles certainly free.  On the cfg4 map 0.03-0.06 % of the poses come back undecided.
struct filt_consts {
  double x0, y0;      // centre of cell (0,0): pos + (0.5 len - 0.5 res)
  double rcp_res;
  float cu[4], cv[4]; // circle offsets in cells
  float r_lo, r_hi;   // in cells
  int W, H, on;       // on = 0: map too small for the border rule, everything is undecided
};

// sqrt.approx.f32: one MUFU, maximum relative error 2^-23 (PTX ISA); the margin below carries it
__device__ __forceinline__ float filt_sqrt(float v) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

template <int FIRST, int LAST>
__device__ __forceinline__ int circles_filter_u8(const filt_consts &fc, const uint8_t *__restrict__ d2, int tiles_x,
                                                 double x, double y, double h) {
  const float hf = (float)h;
  if (!(fabsf(hf) <= 16.f)) return 2;
  const float kf = rintf(hf * 0.15915494f);
  float hr = fmaf(-kf, 6.2831855f, hf);       // 2 pi = 6.2831855f - 1.7484555e-7
  hr = fmaf(kf, 1.7484555e-7f, hr);
  hr = fminf(fmaxf(hr, -3.1415925f), 3.1415925f);
  float s, c;
  __sincosf(hr, &s, &c);
  const float ax = (float)((fc.x0 - x) * fc.rcp_res), ay = (float)((fc.y0 - y) * fc.rcp_res);
  int undecided = 0;
#pragma unroll
  for (int k = FIRST; k < LAST; k++) {
    const float u = fmaf(fc.cv[k], s, fmaf(-fc.cu[k], c, ax));
    const float v = fmaf(-fc.cv[k], c, fmaf(-fc.cu[k], s, ay));
    // floor without the conversion pipe: u + 1.5 * 2^23 rounded DOWN is floor(u) + 1.5 * 2^23 exactly (ulp 1 there), its
    // low bits are the integer.  NaN, infinite and huge u give an index far outside the map: undecided.
    const float tu = __fadd_rd(u, 12582912.f), tv = __fadd_rd(v, 12582912.f);
    const int il = __float_as_int(tu) - 0x4B400000, jl = __float_as_int(tv) - 0x4B400000;
    if (!((unsigned)(il - 2) <= (unsigned)(fc.W - 6) && (unsigned)(jl - 2) <= (unsigned)(fc.H - 6))) return 2;
    const float fu = u - (tu - 12582912.f), fv = v - (tv - 12582912.f);
    // the taps: one 8-byte load per tile row brings (il, jl) and (il + 1, jl) together (the kernel is bound by L1 tag
    // lookups, one per lane and load); only il % 8 == 7 needs the neighbouring tile
    const uint32_t rs = (uint32_t)tiles_x << 5, tc = (uint32_t)(il >> 3) << 5;
    const uint32_t r0 = (uint32_t)(jl >> 2) * rs + ((uint32_t)(jl & 3) << 3) + tc, r1 = r0 + ((jl & 3) == 3 ? rs - 24u : 8u);
    const unsigned sh = (unsigned)(il & 7) << 3;
    const unsigned long long q0 = __ldg((const unsigned long long *)(d2 + r0)) >> sh;
    const unsigned long long q1 = __ldg((const unsigned long long *)(d2 + r1)) >> sh;
    unsigned a00 = (unsigned)q0 & 0xffu, a10 = ((unsigned)q0 >> 8) & 0xffu;
    unsigned a01 = (unsigned)q1 & 0xffu, a11 = ((unsigned)q1 >> 8) & 0xffu;
    if (sh == 56u) {
      a10 = __ldg(d2 + (r0 + 32u));
      a11 = __ldg(d2 + (r1 + 32u));
    }
    const float f00 = filt_sqrt((float)a00), f10 = filt_sqrt((float)a10);
    const float f01 = filt_sqrt((float)a01), f11 = filt_sqrt((float)a11);
    const float lo = fmaf(fu, f10 - f00, f00), hi = fmaf(fu, f11 - f01, f01);
    const float val = fmaf(fv, hi - lo, lo);  // in cells
    if (val < fc.r_lo) return 1;
    undecided |= !(val > fc.r_hi);
  }
  return undecided ? 2 : 0;
}

// The same filter over the FLOAT distance layer (the planner's per-query chains: maps of a few hundred cells that live
// in L1/L2, where the fp64 position arithmetic -- not the taps -- is the dependent chain).  All four circles at once,
// sixteen taps in flight.  The layer is not clamped here, so the interpolation error is relative to the taps at hand
// (all within 1.5 res of the value): with A = 2 res (1 + 2^-19) E + 2^-23 r and B = 48 * 2^-24 the verdict is certain
// when value32 (1 + B) < r - A - 1.5 res B (collision) or value32 (1 - B) > r + A + 1.5 res B (clear); the host folds
// that, with the same 1.5 x slack as above, into f_lo and f_hi.
__device__ __forceinline__ int circles_filter_f32(const hpp_geom &g, const circ_consts &cc, const float *__restrict__ dist,
                                                  double x, double y, double h) {
  const float hf = (float)h;
  if (!(fabsf(hf) <= 16.f)) return 2;
  const float kf = rintf(hf * 0.15915494f);
  float hr = fmaf(-kf, 6.2831855f, hf);
  hr = fmaf(kf, 1.7484555e-7f, hr);
  hr = fminf(fmaxf(hr, -3.1415925f), 3.1415925f);
  float s, c;
  __sincosf(hr, &s, &c);
  const float ax = (float)((cc.fx0 - x) * g.rcp_res), ay = (float)((cc.fy0 - y) * g.rcp_res);
  float fu[4], fv[4];
  long base[4];  // (a 65 534^2 layer has more than 2^31 cells)
  bool inside = true;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const float u = fmaf(cc.fcv[k], s, fmaf(-cc.fcu[k], c, ax));
    const float v = fmaf(-cc.fcv[k], c, fmaf(-cc.fcu[k], s, ay));
    const float tu = __fadd_rd(u, 12582912.f), tv = __fadd_rd(v, 12582912.f);
    const int il = __float_as_int(tu) - 0x4B400000, jl = __float_as_int(tv) - 0x4B400000;
    inside = inside && (unsigned)(il - 2) <= (unsigned)(g.W - 6) && (unsigned)(jl - 2) <= (unsigned)(g.H - 6);
    fu[k] = u - (tu - 12582912.f);
    fv[k] = v - (tv - 12582912.f);
    base[k] = (long)jl * g.W + il;
  }
  if (!inside) return 2;
  float f[4][4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    f[k][0] = __ldg(dist + base[k]);
    f[k][1] = __ldg(dist + base[k] + 1);
    f[k][2] = __ldg(dist + base[k] + g.W);
    f[k][3] = __ldg(dist + base[k] + g.W + 1);
  }
  int hit = 0, undecided = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const float lo = fmaf(fu[k], f[k][1] - f[k][0], f[k][0]), hi = fmaf(fu[k], f[k][3] - f[k][2], f[k][2]);
    const float val = fmaf(fv[k], hi - lo, lo);
    hit |= val < cc.f_lo;
    undecided |= !(val > cc.f_hi);
  }
  return hit ? 1 : (undecided ? 2 : 0);
}

// throughput form: the four circles in the reference's order, later circles are skipped after a hit
__device__ __forceinline__ int circles_collide(const hpp_geom &g, const circ_consts &cc,
                                               const float *__restrict__ dist, double x, double y, double h) {
  double s, c;
  sincos(h, &s, &c);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    double px = cc.cx[k] * c - cc.cy[k] * s + x;
    double py = cc.cx[k] * s + cc.cy[k] * c + y;
    if (circle_hits(g, dist, px, py, cc.cr)) return 1;
  }
  return 0;
}

// latency form (the planner's per-query chains): the taps of all four circles are loaded together; same verdict.
// Not inlined: the planner kernels are instruction-fetch bound, one shared copy keeps their code small.
static __device__ __noinline__ int circles_collide_ilp(const hpp_geom &g, const circ_consts &cc,
                                                   const float *__restrict__ dist, double x, double y, double h) {
  if (cc.f_on) {
    const int r = circles_filter_f32(g, cc, dist, x, y, h);
    if (r != 2) return r;
  }
  double s, c;
  hpp_sincos(h, &s, &c);
  circle_taps t[4];
  float f[4][4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    double px = cc.cx[k] * c - cc.cy[k] * s + x;
    double py = cc.cx[k] * s + cc.cy[k] * c + y;
    circle_prep(g, px, py, t[k]);
  }
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int q = 0; q < 4; q++) f[k][q] = t[k].mode ? __ldg(dist + t[k].l[q]) : 0.f;
  int hit = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) hit |= circle_finish(t[k], f[k][0], f[k][1], f[k][2], f[k][3], cc.cr);
  return hit;
}

// host: the circle constants of a map, exact part and filter part (margins derived at circles_filter_u8 / _f32)
struct filt_margin {
  double E, A, C;  // position error in cells, its share of the margin in metres, circle offset in cells
};
static inline filt_margin hpp_filt_margin(const hpp_map *m) {
  const hpp_geom &g = m->g;
  filt_margin fm;
  fm.C = 0.0;
  for (int k = 0; k < 4; k++) {
    const double ck = (fabs(m->car.cx[k]) + fabs(m->car.cy[k])) / g.res;
    if (ck > fm.C) fm.C = ck;
  }
  const double D = (double)(g.W > g.H ? g.W : g.H) + 4.0;
  fm.E = ldexp(D, -22) + fm.C * (4e-6 + ldexp(1.0, -21));
  fm.A = 2.0 * g.res * (1.0 + ldexp(1.0, -19)) * fm.E + ldexp(fabs(m->car.cr), -23);
  return fm;
}
static inline circ_consts hpp_make_circ_consts(const hpp_map *m) {
  circ_consts cc;
  const hpp_geom &g = m->g;
  for (int k = 0; k < 4; k++) {
    cc.cx[k] = m->car.cx[k];
    cc.cy[k] = m->car.cy[k];
    cc.fcu[k] = (float)(m->car.cx[k] / g.res);
    cc.fcv[k] = (float)(m->car.cy[k] / g.res);
  }
  cc.cr = m->car.cr;
  cc.fx0 = g.posx + (0.5 * g.lenx - 0.5 * g.res);
  cc.fy0 = g.posy + (0.5 * g.leny - 0.5 * g.res);
  const filt_margin fm = hpp_filt_margin(m);
  const double B = 48.0 * ldexp(1.0, -24), r = m->car.cr;
  const double lo = (r - 1.5 * (fm.A + 1.5 * g.res * B)) / (1.0 + 1.5 * B);
  const double hi = (r + 1.5 * (fm.A + 1.5 * g.res * B)) / (1.0 - 1.5 * B);
  cc.f_lo = nextafterf((float)lo, -INFINITY);
  cc.f_hi = nextafterf((float)hi, INFINITY);
  static const bool off = getenv("HPP_NO_FILTER") != nullptr;  // exact arithmetic everywhere (same results, for comparison)
  cc.f_on = !off && g.W >= 8 && g.H >= 8 && fm.C < 1e5 && r > 0.0 && lo > 0.0;
  return cc;
}

// ---------------------------------------------------------------------------------------- SAT
struct sat_rect {
  float x[5], y[5];      // Point32 vertices of the vehicle rectangle (closed)
  double ax[3], ay[3];   // its three axes (orthogonals of the first three edges, float32 edge vectors)
  double mn[3], mx[3];   // its own projections on those axes
};

__device__ __forceinline__ void sat_make_rect(const float *rx, const float *ry, double x, double y, double h,
                                              sat_rect &R) {
  double s, c;
  sincos(h, &s, &c);
#pragma unroll
  for (int k = 0; k < 5; k++) {
    double px = x + (double)rx[k] * c - (double)ry[k] * s;
    double py = y + (double)rx[k] * s + (double)ry[k] * c;
    R.x[k] = (float)px;
    R.y[k] = (float)py;
  }
#pragma unroll
  for (int e = 0; e < 3; e++) {
    float ex = __fsub_rn(R.x[e + 1], R.x[e]), ey = __fsub_rn(R.y[e + 1], R.y[e]);
    R.ax[e] = (double)ey;
    R.ay[e] = -(double)ex;
    double mn = 0, mx = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
      double d = (double)R.x[k] * R.ax[e] + (double)R.y[k] * R.ay[e];
      if (k == 0 || d < mn) mn = d;
      if (k == 0 || d > mx) mx = d;
    }
    R.mn[e] = mn;
    R.mx[e] = mx;
  }
}

__device__ __forceinline__ bool sat_overlap(double min1, double max1, double min2, double max2) {
  return (min1 < min2) ? ((min2 - max1) < 0) : ((min1 - max2) < 0);
}

// SAT test of the vehicle rectangle against the square of occupancy cell (j = column/x, i = row/y)
__device__ __forceinline__ bool sat_cell_hit(const hpp_geom &g, const sat_rect &R, int i, int j) {
  double px = g.ox + j * g.res, py = g.oy + i * g.res;
  const double half = g.res / 2.0;
  px += g.res / 2;
  py += g.res / 2;
  float qx[5], qy[5];
  qx[0] = (float)(px - half); qy[0] = (float)(py - half);
  qx[1] = (float)(px + half); qy[1] = (float)(py - half);
  qx[2] = (float)(px + half); qy[2] = (float)(py + half);
  qx[3] = (float)(px - half); qy[3] = (float)(py + half);
  qx[4] = qx[0];              qy[4] = qy[0];
  // rectangle axes first (same order as the reference: edges1 then edges2)
#pragma unroll
  for (int e = 0; e < 3; e++) {
    double mn = 0, mx = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
      double d = (double)qx[k] * R.ax[e] + (double)qy[k] * R.ay[e];
      if (k == 0 || d < mn) mn = d;
      if (k == 0 || d > mx) mx = d;
    }
    if (!sat_overlap(R.mn[e], R.mx[e], mn, mx)) return false;
  }
#pragma unroll
  for (int e = 0; e < 3; e++) {
    float ex = __fsub_rn(qx[e + 1], qx[e]), ey = __fsub_rn(qy[e + 1], qy[e]);
    double ax = (double)ey, ay = -(double)ex;
    double mn1 = 0, mx1 = 0, mn2 = 0, mx2 = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
      double d1 = (double)R.x[k] * ax + (double)R.y[k] * ay;
      double d2 = (double)qx[k] * ax + (double)qy[k] * ay;
      if (k == 0 || d1 < mn1) mn1 = d1;
      if (k == 0 || d1 > mx1) mx1 = d1;
      if (k == 0 || d2 < mn2) mn2 = d2;
      if (k == 0 || d2 > mx2) mx2 = d2;
    }
    if (!sat_overlap(mn1, mx1, mn2, mx2)) return false;
  }
  return true;
}

// Whole-warp SAT check of one pose (all 32 lanes call with the same pose).  The +-5 m window is scanned
// through the bit-packed occupancy: lane = window row (strided), two or three 32-bit words per row.
// Result is a boolean OR over cells, so the scan order does not matter.  Returns the planner predicate.
__device__ __forceinline__ int sat_collide_warp(const hpp_geom &g, const float *rx, const float *ry,
                                                const uint32_t *__restrict__ bits, int pitch, double x, double y,
                                                double h) {
  const int lane = threadIdx.x & 31;
  sat_rect R;
  sat_make_rect(rx, ry, x, y, h, R);
  const double offset = 5.0;
  const double min_x = fmax(g.ox, x - offset), max_x = fmin(g.ox + g.W * g.res, x + offset);
  const double min_y = fmax(g.oy, y - offset), max_y = fmin(g.oy + g.H * g.res, y + offset);
  const int min_i = max(0, (int)(hpp_div_res(g, min_y - g.oy))), max_i = min(g.H - 1, (int)(hpp_div_res(g, max_y - g.oy)));
  const int min_j = max(0, (int)(hpp_div_res(g, min_x - g.ox))), max_j = min(g.W - 1, (int)(hpp_div_res(g, max_x - g.ox)));
  int hit = 0;
  if (max_j >= min_j) {
    const int w0 = min_j >> 5, w1 = max_j >> 5;
    for (int i = min_i + lane; i <= max_i && !hit; i += 32) {
      for (int w = w0; w <= w1 && !hit; w++) {
        uint32_t word = __ldg(bits + (size_t)i * pitch + w);
        if (w == w0) word &= 0xFFFFFFFFu << (min_j & 31);
        if (w == w1) word &= 0xFFFFFFFFu >> (31 - (max_j & 31));
        while (word) {
          int b = __ffs(word) - 1;
          word &= word - 1;
          if (sat_cell_hit(g, R, i, (w << 5) + b)) {
            hit = 1;
            break;
          }
        }
      }
    }
  }
  hit = __any_sync(0xffffffffu, hit);
  const int gx = (int)(hpp_div_res(g, x - g.ox)), gy = (int)(hpp_div_res(g, y - g.oy));
  const int inb = gx >= 0 && gx < g.W && gy >= 0 && gy < g.H;
  return hit || !inb;
}
