// hpp_internal.cuh -- shared declarations of libhpp_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/hpp_b200.h"

#define HPP_SM_COUNT_B200 148

struct hpp_ctx {
  int device;
  int sm_count;
  cudaStream_t stream;
  cudaStream_t copy_stream;   // D2H copies overlapped with compute (created on first use)
  cudaEvent_t copy_done[2];
  uint64_t launches;
  char err[512];
  // pinned staging for small control words (convergence flags, counters)
  uint32_t *h_flags;   // pinned host, 64 words
  uint32_t *d_flags;   // device, 64 words
  // grow-only scratch (device) + pinned host staging for the host-pointer entry points
  int refs;              // the creator's reference + one per live map / communicator (hpp_ctx_retain / _release)
  void *d_scratch[8];
  size_t d_scratch_bytes[8];
  void *h_pinned[2];
  size_t h_pinned_bytes[2];
  // hpp_plan_batch: values that cost a driver round trip each (tens of microseconds of a 0.3 ms single-query call)
  size_t plan_free_bytes;   // cudaMemGetInfo at the last large call (0 = never asked)
  int plan_occ_blocks[4];   // resident plan_kernel CTAs per SM, per variant and kernel shape (0 = not asked yet)
};

// per-cell class byte (see map.cu: build_cls_kernel)
#define CLS_BLOCKED 0x80  // data > 0  (Grid_map::isInCollision, RS/src/Grid_map.cpp:212)
#define CLS_NEAR4 0x40    // some in-bounds 4-neighbour is blocked (RS/src/HybridAstar.cpp:507-519)
#define CLS_CNT_MASK 0x0F // number of isInCollision cells in the 3x3 block, out of bounds included
                          // (DB/src/HybridAstar.cpp:204-218)

// Maps and communicators keep their context alive: hpp_ctx_destroy drops the creator's reference and the resources
// go with the last one, whatever order a garbage-collected host language finalises the handles in.
void hpp_ctx_retain(hpp_ctx *c);
void hpp_ctx_release(hpp_ctx *c);

struct hpp_geom {
  int W, H;
  float res_f;          // info.resolution as sent (float32)
  double res;           // (double)res_f, what the planners compute with (hazard H6)
  double rcp_res;       // RN(1 / res): exact division by the resolution with one multiply and two FMAs (hpp_div_res)
  double ox, oy;        // origin
  // grid_map::GridMap geometry (RS/src/Grid_map.cpp:15-25, grid_map_core setGeometry)
  double lenx, leny, posx, posy;
};

struct hpp_car_dev {
  hpp_params p;
  // RS circles (RS/src/CarData.cpp:76-98)
  double cx[4], cy[4], cr;
  // Dubins rectangle corners as float32 (DB/src/CarData.cpp:15-37), closed polygon of 5
  float rx[5], ry[5];
  // motion commands (RS/src/HybridAstar.cpp:14-23, DB/src/HybridAstar.cpp:14-21)
  int n_cmd[2];
  double cmd_steer[2][HPP_MAX_COMMANDS];
  double cmd_dir[2][HPP_MAX_COMMANDS];
  double maxc;          // tan(maxSteerAngle) / wheelBase  (RS/src/HybridAstar.cpp:121)
};

struct hpp_map {
  hpp_ctx *ctx;
  hpp_geom g;
  hpp_car_dev car;
  size_t n;
  int8_t *d_occ;        // [y*W + x]
  uint8_t *d_cls;       // [y*W + x]
  uint32_t *d_occ_bits; // bit-packed "data > 0", pitch words per row
  uint32_t *d_near_bits; // bit-packed CLS_NEAR4, same pitch
  int bits_pitch;
  uint16_t *d_g;        // pass-1 row distances of the EDT (scratch, kept for reuse)
  float *d_obstacle;    // [x + y*W]
  float *d_distance;    // [x + y*W] (already flipped: distance[k] = edt[n-1-k]*res)
  // compact copy of the distance layer for the footprint check (K4): the squared cell distance clamped to 255, one byte
  // per cell, in micro-tiles of 8 (i) x 4 (j) cells = one 32-byte sector, same (flipped) cell order as d_distance.
  // distance == sqrtf((float)d2) * res_f exactly for d2 < 255; see collide.cu for why the clamp keeps the verdicts.
  uint8_t *d_d2u8;
  int d2_tiles_x;       // micro-tiles per row: ceil(W / 8)
};

#define HPP_D2_TILE_W 8
#define HPP_D2_TILE_H 4
static inline __host__ __device__ size_t hpp_d2_addr(int i, int j, int tiles_x) {
  return ((size_t)(j >> 2) * (size_t)tiles_x + (size_t)(i >> 3)) * 32u + (size_t)(((j & 3) << 3) + (i & 7));
}

#define HPP_CUDA(ctx, call)                                                                       \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s -> %s", __FILE__, __LINE__, #call,      \
               cudaGetErrorString(e_));                                                           \
      return HPP_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

#define HPP_FAIL(ctx, code, ...)                                  \
  do {                                                            \
    snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__);        \
    return (code);                                                \
  } while (0)

#define HPP_LAUNCHED(ctx)                                                                         \
  do {                                                                                            \
    (ctx)->launches++;                                                                            \
    cudaError_t e_ = cudaGetLastError();                                                          \
    if (e_ != cudaSuccess) {                                                                      \
      snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: kernel launch -> %s", __FILE__, __LINE__,  \
               cudaGetErrorString(e_));                                                           \
      return HPP_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

// grow-only device scratch slot / pinned host slot
hpp_status hpp_scratch(hpp_ctx *ctx, int slot, size_t bytes, void **out);
hpp_status hpp_pinned(hpp_ctx *ctx, int slot, size_t bytes, void **out);

// internal launchers (device pointers, enqueue only)
hpp_status hpp_edt_build(hpp_map *m);                                   // edt.cu
hpp_status hpp_edt_band_rows(hpp_ctx *c, const int8_t *occ_band_dev, int W, int rows, uint16_t *g_band);
hpp_status hpp_edt_band_cols(hpp_ctx *c, const uint16_t *g_ext, int W, int rows_ext, int y_begin, int y_count,
                             int more_above, int more_below, float res_f, float *distance_band, uint32_t *unresolved);
hpp_status hpp_build_cls(hpp_map *m);                                   // map.cu
hpp_status hpp_inflate_launch(hpp_ctx *ctx, const int8_t *occ_dev, uint32_t W, uint32_t H, double resolution,
                              double inflation_radius, int8_t *out_dev); // inflate.cu

static inline int hpp_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// a / g.res, correctly rounded, without the division routine: with y = RN(1/d), q0 = RN(a*y) is a faithful quotient,
// r = a - q0*d is exact in one FMA, and RN(q0 + r*y) = RN(a/d) (Markstein 1990).  The theorem needs y to be the
// correctly rounded reciprocal, which holds unless d's 53-bit significand is all ones -- d is a float32 widened to
// double (24 significant bits), so it always holds here; checked against a/d on 3.2e9 random operands over 22 resolutions
// and exponents +-930.  Zeros, subnormal-range and huge operands take the plain division.
#ifdef __CUDACC__
// libm by CALL, for the planner and shot code.  An inlined fp64 sincos is ~950 SASS instructions (argument reduction with its
// Payne-Hanek slow path, two polynomials), tan / atan2 / acos are of the same order, and the Reeds-Shepp formulas, the path
// samplers and the footprint check hold some sixty call sites: inlined, plan_pipe_kernel was 185 KB of code, more than the
// SM's instruction cache, and ncu showed `no_instruction` as its first stall reason (12.6 cycles per issued instruction
// with two CTAs per SM).  One copy per translation unit, same code as the inlined form, so the same bits.
#ifdef HPP_LIBM_INLINE  // (A/B builds: -DHPP_LIBM_INLINE gives the inlined form back)
#define HPP_LIBM_CALL __forceinline__
#else
#define HPP_LIBM_CALL __noinline__
#endif
struct hpp_sc {
  double s, c;
};
[[maybe_unused]] static __device__ HPP_LIBM_CALL hpp_sc hpp_sincos_call(double a) {
  hpp_sc r;
  sincos(a, &r.s, &r.c);
  return r;
}
__device__ __forceinline__ void hpp_sincos(double a, double *s, double *c) {
  const hpp_sc r = hpp_sincos_call(a);
  *s = r.s;
  *c = r.c;
}
[[maybe_unused]] static __device__ HPP_LIBM_CALL double hpp_sin(double a) { return sin(a); }
[[maybe_unused]] static __device__ HPP_LIBM_CALL double hpp_cos(double a) { return cos(a); }
[[maybe_unused]] static __device__ HPP_LIBM_CALL double hpp_tan(double a) { return tan(a); }
[[maybe_unused]] static __device__ HPP_LIBM_CALL double hpp_atan2(double y, double x) { return atan2(y, x); }
[[maybe_unused]] static __device__ HPP_LIBM_CALL double hpp_acos(double a) { return acos(a); }
[[maybe_unused]] static __device__ HPP_LIBM_CALL double hpp_asin(double a) { return asin(a); }
[[maybe_unused]] static __device__ HPP_LIBM_CALL double hpp_fmod(double a, double b) { return fmod(a, b); }
#endif

#ifdef __CUDACC__
static __device__ __noinline__ double hpp_div_slow(double a, double d) { return a / d; }  // cold: one copy per kernel
__device__ __forceinline__ double hpp_div_res(const hpp_geom &g, double a) {
  const double m = fabs(a);
  if (!(m > 1e-280 && m < 1e280)) return hpp_div_slow(a, g.res);
  const double q0 = a * g.rcp_res;
  const double r = __fma_rn(-q0, g.res, a);
  return __fma_rn(r, g.rcp_res, q0);
}
#endif
