// collide.cu -- K4 (circle footprint over the distance layer) and K5 (SAT footprint over the occupancy)
// batch kernels.  See collide.cuh for the reference citations.
//
// K4: one thread per pose, the 4 circles in the reference's order with its early exit.  Poses stream in as
//     3 doubles (24 B), the verdict streams out as 1 byte; the up-to-16 float taps are gathers into the
//     distance layer served by L2/HBM.  Algorithmic bytes: 24 + 1 + 4*4*4 = 89 B per pose.
// K5: one warp per pose (window rows across lanes, bit-packed occupancy), 8 poses per CTA.
#include <math.h>
#include <stdlib.h>

#include "collide.cuh"

__global__ void __launch_bounds__(256) collide_circles_kernel(hpp_geom g, circ_consts cc,
                                                              const float *__restrict__ dist,
                                                              const double *__restrict__ xyh, size_t n,
                                                              uint8_t *__restrict__ verdict) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double x = xyh[3 * i], y = xyh[3 * i + 1], h = xyh[3 * i + 2];
    verdict[i] = (uint8_t)circles_collide(g, cc, dist, x, y, h);
  }
}

// K4 over the compact layer: the same verdicts from a quarter of the bytes (the 8192^2 layer is 67 MB and stays in L2;
// a 2x2 tap window costs 1.4 sectors on average instead of 2.25)
__global__ void __launch_bounds__(256) collide_circles_u8_kernel(hpp_geom g, circ_consts cc,
                                                                 const uint8_t *__restrict__ d2, int tiles_x,
                                                                 const double *__restrict__ xyh, size_t n,
                                                                 uint8_t *__restrict__ verdict) {
  __shared__ float lut[256];
  lut[threadIdx.x] = __fmul_rn(__fsqrt_rn((float)threadIdx.x), g.res_f);  // what edt_col_kernel stores for d2 = v
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double x = xyh[3 * i], y = xyh[3 * i + 1], h = xyh[3 * i + 2];
    double s, c;
    sincos(h, &s, &c);
    int hit = 0;
#pragma unroll
    for (int k = 0; k < 4 && !hit; k++) {
      const double px = cc.cx[k] * c - cc.cy[k] * s + x;
      const double py = cc.cx[k] * s + cc.cy[k] * c + y;
      hit = circle_hits_u8(g, d2, tiles_x, lut, px, py, cc.cr);
    }
    verdict[i] = (uint8_t)hit;
  }
}

// Large batches: a float32 filter pass, then the exact arithmetic for the poses it could not decide (collide.cuh,
// circles_filter_u8).  ncu on the one-pass kernel above: 1 165 instructions per pose, nearly all of them fp64 position
// arithmetic whose result only matters within 1e-3 m of the verdict boundary; issue slots, not bytes, were the bound
// (DRAM traffic is 0.67x the algorithmic bytes).  Pass 1 decides ~99.6 % of the poses in ~1/5 of the instructions and
// appends the rest to a list (one atomicAdd per warp); pass 2 is the exact kernel over that list.
__global__ void __launch_bounds__(256) collide_filter_kernel(filt_consts fc, const uint8_t *__restrict__ d2, int tiles_x,
                                                             const double *__restrict__ xyh, size_t n,
                                                             uint8_t *__restrict__ verdict,
                                                             uint32_t *__restrict__ undecided,
                                                             unsigned long long *__restrict__ n_undecided) {
  const int lane = threadIdx.x & 31;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n_up = (n + 31) & ~(size_t)31;  // whole warps run the loop (the ballot below)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_up; i += stride) {
    const bool in = i < n;
    int r = 0;
    if (in) {
      const double x = xyh[3 * i], y = xyh[3 * i + 1], h = xyh[3 * i + 2];
      r = fc.on ? circles_filter_u8<0, 4>(fc, d2, tiles_x, x, y, h) : 2;
      if (r != 2) verdict[i] = (uint8_t)r;
    }
    const unsigned m = __ballot_sync(0xffffffffu, r == 2);
    if (m) {
      unsigned long long b = 0;
      if (lane == 0) b = atomicAdd(n_undecided, (unsigned long long)__popc(m));
      b = __shfl_sync(0xffffffffu, b, 0);
      if (r == 2) undecided[b + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
    }
  }
}

__global__ void __launch_bounds__(256) collide_u8_list_kernel(hpp_geom g, circ_consts cc, const uint8_t *__restrict__ d2,
                                                              int tiles_x, const double *__restrict__ xyh,
                                                              uint8_t *__restrict__ verdict,
                                                              const uint32_t *__restrict__ list,
                                                              const unsigned long long *__restrict__ n_list) {
  __shared__ float lut[256];
  lut[threadIdx.x] = __fmul_rn(__fsqrt_rn((float)threadIdx.x), g.res_f);
  __syncthreads();
  const size_t ns = (size_t)*n_list;
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < ns; j += (size_t)gridDim.x * blockDim.x) {
    const size_t i = list[j];
    const double x = xyh[3 * i], y = xyh[3 * i + 1], h = xyh[3 * i + 2];
    double s, c;
    sincos(h, &s, &c);
    int hit = 0;
#pragma unroll
    for (int k = 0; k < 4 && !hit; k++) {
      const double px = cc.cx[k] * c - cc.cy[k] * s + x;
      const double py = cc.cx[k] * s + cc.cy[k] * c + y;
      hit = circle_hits_u8(g, d2, tiles_x, lut, px, py, cc.cr);
    }
    verdict[i] = (uint8_t)hit;
  }
}

// Grid_map::isSingleStateCollisionFreeImproved (RS/src/Grid_map.cpp:286-317), statement by statement, including its
// return convention (which is NOT the one of isSingleStateCollisionFree: `true` when the bounding circle is clear,
// `false` when its centre is outside the map, and the detailed check's "true = collision" in between)
__global__ void __launch_bounds__(256) collide_improved_kernel(hpp_geom g, circ_consts cc, bound_consts bc,
                                                               const float *__restrict__ dist,
                                                               const double *__restrict__ xyh, size_t n,
                                                               uint8_t *__restrict__ result) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double x = xyh[3 * i], y = xyh[3 * i + 1], h = xyh[3 * i + 2];
    // CarData::getBoundingCircle -> local2Global(state, (bx, 0))  (RS/src/CarData.cpp:112-124)
    double sn, cs;
    sincos(h, &sn, &cs);
    const double bpx = bc.bx * cs - 0.0 * sn + x;
    const double bpy = bc.bx * sn + 0.0 * cs + y;
    circle_taps t;
    circle_prep(g, bpx, bpy, t);
    int r;
    if (t.mode == 0) {
      r = 0;                                               // :313-316 out of bounds -> false
    } else if (circle_hits(g, dist, bpx, bpy, bc.br)) {    // :301 clearance < bounding radius
      r = circles_collide(g, cc, dist, x, y, h);           // :304 detailed check (true = collision)
    } else {
      r = 1;                                               // :309 -> true
    }
    result[i] = (uint8_t)r;
  }
}

struct rect_consts {
  float rx[5], ry[5];
};

__global__ void __launch_bounds__(256) collide_sat_kernel(hpp_geom g, rect_consts rc,
                                                          const uint32_t *__restrict__ bits, int pitch,
                                                          const double *__restrict__ xyh, size_t n,
                                                          uint8_t *__restrict__ verdict) {
  const int lane = threadIdx.x & 31;
  const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const double x = xyh[3 * i], y = xyh[3 * i + 1], h = xyh[3 * i + 2];
    int v = sat_collide_warp(g, rc.rx, rc.ry, bits, pitch, x, y, h);
    if (lane == 0) verdict[i] = (uint8_t)v;
  }
}

static circ_consts make_cc(const hpp_map *m) { return hpp_make_circ_consts(m); }
// the float32 filter's constants and its certainty margin M (derivation: collide.cuh, circles_filter_u8)
static filt_consts make_fc(const hpp_map *m) {
  filt_consts fc;
  const hpp_geom &g = m->g;
  fc.x0 = g.posx + (0.5 * g.lenx - 0.5 * g.res);
  fc.y0 = g.posy + (0.5 * g.leny - 0.5 * g.res);
  fc.rcp_res = g.rcp_res;
  for (int k = 0; k < 4; k++) {
    fc.cu[k] = (float)(m->car.cx[k] / g.res);
    fc.cv[k] = (float)(m->car.cy[k] / g.res);
  }
  const filt_margin fm = hpp_filt_margin(m);
  const double V = sqrt(255.0) * g.res * (1.0 + 1e-6);
  const double r = m->car.cr;
  const double M = fm.A + 48.0 * ldexp(V, -24);
  fc.r_lo = nextafterf((float)((r - 1.5 * M) / g.res), -INFINITY);
  fc.r_hi = nextafterf((float)((r + 1.5 * M) / g.res), INFINITY);
  fc.W = g.W;
  fc.H = g.H;
  fc.on = g.W >= 8 && g.H >= 8 && fm.C < 1e5 && r > 0.0 && r - 1.5 * M > 0.0;
  return fc;
}
static rect_consts make_rc(const hpp_map *m) {
  rect_consts rc;
  for (int k = 0; k < 5; k++) {
    rc.rx[k] = m->car.rx[k];
    rc.ry[k] = m->car.ry[k];
  }
  return rc;
}

extern "C" hpp_status hpp_collide_circles_dev(hpp_map *m, const double *xyh_dev, size_t n, uint8_t *verdict_dev) {
  if (!m || (n && (!xyh_dev || !verdict_dev))) return HPP_ERR_ARG;
  if (n == 0) return HPP_OK;
  hpp_ctx *c = m->ctx;
  HPP_CUDA(c, cudaSetDevice(c->device));
  // grid: a multiple of the SM count, 8 resident CTAs of 256 threads per SM
  long long want = hpp_div_up((long long)n, 256);
  int grid = (int)(want < (long long)c->sm_count * 8 ? want : (long long)c->sm_count * 8);
  // compact layer whenever the circle radius is safely below the clamp distance (see circle_hits_u8); HPP_K4_FLOAT=1
  // forces the float layer (same verdicts, for comparison)
  static const bool force_float = getenv("HPP_K4_FLOAT") != nullptr;
  static const bool one_pass = getenv("HPP_K4_ONE_PASS") != nullptr;
  if (!force_float && m->car.cr < 14.0 * m->g.res && n >= ((size_t)1 << 14) && n < 0xffffffffull && !one_pass) {
    void *scr = nullptr;
    hpp_status s = hpp_scratch(c, 6, 256 + n * sizeof(uint32_t), &scr);
    if (s != HPP_OK) return s;
    unsigned long long *d_count = (unsigned long long *)scr;
    uint32_t *d_list = (uint32_t *)((uint8_t *)scr + 256);
    HPP_CUDA(c, cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), c->stream));
    const filt_consts fc = make_fc(m);
    collide_filter_kernel<<<grid, 256, 0, c->stream>>>(fc, m->d_d2u8, m->d2_tiles_x, xyh_dev, n, verdict_dev,
                                                       d_list, d_count);
    HPP_LAUNCHED(c);
    collide_u8_list_kernel<<<c->sm_count * 2, 256, 0, c->stream>>>(m->g, make_cc(m), m->d_d2u8, m->d2_tiles_x, xyh_dev,
                                                                   verdict_dev, d_list, d_count);
  } else if (!force_float && m->car.cr < 14.0 * m->g.res)
    collide_circles_u8_kernel<<<grid, 256, 0, c->stream>>>(m->g, make_cc(m), m->d_d2u8, m->d2_tiles_x, xyh_dev, n,
                                                           verdict_dev);
  else
    collide_circles_kernel<<<grid, 256, 0, c->stream>>>(m->g, make_cc(m), m->d_distance, xyh_dev, n, verdict_dev);
  HPP_LAUNCHED(c);
  return HPP_OK;
}

extern "C" hpp_status hpp_collide_improved_dev(hpp_map *m, const double *xyh_dev, size_t n, uint8_t *result_dev) {
  if (!m || (n && (!xyh_dev || !result_dev))) return HPP_ERR_ARG;
  if (n == 0) return HPP_OK;
  hpp_ctx *c = m->ctx;
  HPP_CUDA(c, cudaSetDevice(c->device));
  long long want = hpp_div_up((long long)n, 256);
  int grid = (int)(want < (long long)c->sm_count * 8 ? want : (long long)c->sm_count * 8);
  bound_consts bc;
  const hpp_params &p = m->car.p;
  bc.bx = (p.axle_to_front - p.axle_to_back) / 2.0;                                        // RS/src/CarData.cpp:79
  bc.br = sqrt(pow((p.axle_to_front + p.axle_to_back) / 2.0, 2) + pow(p.width / 2.0, 2));  // :81
  collide_improved_kernel<<<grid, 256, 0, c->stream>>>(m->g, make_cc(m), bc, m->d_distance, xyh_dev, n, result_dev);
  HPP_LAUNCHED(c);
  return HPP_OK;
}

extern "C" hpp_status hpp_collide_sat_dev(hpp_map *m, const double *xyh_dev, size_t n, uint8_t *verdict_dev) {
  if (!m || (n && (!xyh_dev || !verdict_dev))) return HPP_ERR_ARG;
  if (n == 0) return HPP_OK;
  hpp_ctx *c = m->ctx;
  HPP_CUDA(c, cudaSetDevice(c->device));
  long long want = hpp_div_up((long long)n, 8);
  int grid = (int)(want < (long long)c->sm_count * 8 ? want : (long long)c->sm_count * 8);
  collide_sat_kernel<<<grid, 256, 0, c->stream>>>(m->g, make_rc(m), m->d_occ_bits, m->bits_pitch, xyh_dev, n,
                                                   verdict_dev);
  HPP_LAUNCHED(c);
  return HPP_OK;
}

typedef hpp_status (*collide_dev_fn)(hpp_map *, const double *, size_t, uint8_t *);

// host-pointer path: chunked H2D -> kernel -> D2H so that arbitrarily large pose batches stream through
// a bounded device scratch.
static hpp_status collide_host(hpp_map *m, const double *xyh, size_t n, uint8_t *verdict, collide_dev_fn fn) {
  if (!m || (n && (!xyh || !verdict))) return HPP_ERR_ARG;
  if (n == 0) return HPP_OK;
  hpp_ctx *c = m->ctx;
  HPP_CUDA(c, cudaSetDevice(c->device));
  const size_t chunk = n < ((size_t)1 << 24) ? n : ((size_t)1 << 24);
  void *d_in, *d_out;
  hpp_status s = hpp_scratch(c, 0, chunk * 24, &d_in);
  if (s != HPP_OK) return s;
  s = hpp_scratch(c, 1, chunk, &d_out);
  if (s != HPP_OK) return s;
  for (size_t off = 0; off < n; off += chunk) {
    size_t k = n - off < chunk ? n - off : chunk;
    HPP_CUDA(c, cudaMemcpyAsync(d_in, xyh + 3 * off, k * 24, cudaMemcpyHostToDevice, c->stream));
    s = fn(m, (const double *)d_in, k, (uint8_t *)d_out);
    if (s != HPP_OK) return s;
    HPP_CUDA(c, cudaMemcpyAsync(verdict + off, d_out, k, cudaMemcpyDeviceToHost, c->stream));
  }
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  return HPP_OK;
}

extern "C" hpp_status hpp_collide_circles(hpp_map *m, const double *xyh, size_t n, uint8_t *verdict) {
  return collide_host(m, xyh, n, verdict, hpp_collide_circles_dev);
}
extern "C" hpp_status hpp_collide_sat(hpp_map *m, const double *xyh, size_t n, uint8_t *verdict) {
  return collide_host(m, xyh, n, verdict, hpp_collide_sat_dev);
}
extern "C" hpp_status hpp_collide_improved(hpp_map *m, const double *xyh, size_t n, uint8_t *result) {
  return collide_host(m, xyh, n, result, hpp_collide_improved_dev);
}
