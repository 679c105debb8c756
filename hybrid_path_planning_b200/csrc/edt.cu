// edt.cu -- K1: grid_map "obstacle" and "distance" layers (replaces RS/src/Grid_map.cpp:22-74:
// threshold -> cv::distanceTransform(DIST_L2, DIST_MASK_PRECISE) -> flip(-1) -> * resolution).
//
// Separable exact Euclidean distance transform in integer arithmetic:
//   pass 1 (rows, along x):  g(x,y) = distance along x to the nearest obstacle cell of row y (uint16,
//                            0xFFFF = the row has none).  One CTA per row, row staged in shared memory,
//                            block-wide max/min scans of the nearest obstacle index, 16-byte vector I/O.
//   pass 2 (columns, along y): d2(x,y) = min_y' g(x,y')^2 + (y-y')^2 by an outward search that stops as
//                            soon as k^2 >= best.  64x64 output tiles with a +-32-row halo of g in shared
//                            memory (adjacent threads = adjacent x => coalesced); rows beyond the halo are
//                            read through L2.  Epilogue: sqrtf((float)d2) * (float)res (both correctly
//                            rounded, no FMA) stored at the 180-degree-rotated index n-1-k (hazard H7).
// The float32 lower-envelope pass of OpenCV returns exactly sqrtf(d2) whenever d2 < 2^24 (distances below
// 4096 cells); beyond that its float rounding is not reproduced (documented in DESIGN.md).
#include "hpp_internal.cuh"

#define EDT_INF16 0xFFFFu
#define EDT_INF32 0xFFFFFFFFu

struct pred_lut {
  uint32_t w[8];  // bit v (as uint8) set <=> the occupancy byte v is an obstacle for this transform
};

__device__ __forceinline__ int lut_test(const pred_lut &l, int8_t v) {
  unsigned u = (uint8_t)v;
  return (l.w[u >> 5] >> (u & 31)) & 1u;
}

// ---- block-wide inclusive scans over one int per thread (256 threads) ----
__device__ __forceinline__ int block_excl_max_scan(int v, int *warp_tot /*8*/, int ident) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc = max(inc, o);
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  int carry = ident;
  for (int w = 0; w < wid; w++) carry = max(carry, warp_tot[w]);
  int exc = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) exc = ident;
  __syncthreads();
  return max(carry, exc);
}

__device__ __forceinline__ int block_excl_min_suffix_scan(int v, int *warp_tot /*8*/, int ident) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int o = __shfl_down_sync(0xffffffffu, inc, d);
    if (lane + d < 32) inc = min(inc, o);
  }
  if (lane == 0) warp_tot[wid] = inc;
  __syncthreads();
  int carry = ident;
  for (int w = wid + 1; w < nw; w++) carry = min(carry, warp_tot[w]);
  int exc = __shfl_down_sync(0xffffffffu, inc, 1);
  if (lane == 31) exc = ident;
  __syncthreads();
  return min(carry, exc);
}

// pass 1.  smem: flag[W] bytes then (2-byte aligned) g[W] uint16.
// obstacle may be null (inflate reuses this kernel).  lut_dev: predicate table in device memory.
__global__ void __launch_bounds__(256) edt_row_kernel(const int8_t *__restrict__ occ, int W, int H,
                                                      const pred_lut *__restrict__ lut_dev,
                                                      uint16_t *__restrict__ g, float *__restrict__ obstacle,
                                                      uint32_t *__restrict__ any_obstacle) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int warp_tot[8];
  __shared__ pred_lut lut;
  const int Wpad = (W + 15) & ~15;
  unsigned char *flag = smem;
  uint16_t *gs = (uint16_t *)(smem + Wpad);
  const int y = blockIdx.x;
  const int8_t *row = occ + (size_t)y * W;
  if (threadIdx.x < 8) lut.w[threadIdx.x] = lut_dev->w[threadIdx.x];
  __syncthreads();

  // stage the row: predicate bytes
  if ((W & 15) == 0) {
    const int4 *row4 = (const int4 *)row;
    for (int i = threadIdx.x; i < W / 16; i += blockDim.x) {
      int4 v = __ldg(row4 + i);
      int8_t b[16];
      *(int4 *)b = v;
      unsigned char f[16];
#pragma unroll
      for (int k = 0; k < 16; k++) f[k] = (unsigned char)lut_test(lut, b[k]);
      *(int4 *)(flag + 16 * i) = *(int4 *)f;
    }
  } else {
    for (int i = threadIdx.x; i < W; i += blockDim.x) flag[i] = (unsigned char)lut_test(lut, row[i]);
  }
  __syncthreads();

  // per-thread contiguous chunk
  const int C = (W + blockDim.x - 1) / blockDim.x;
  const int c0 = min(W, (int)threadIdx.x * C), c1 = min(W, c0 + C);
  int last = -(1 << 20), first = (1 << 20);
  for (int x = c0; x < c1; x++)
    if (flag[x]) {
      last = x;
      if (first == (1 << 20)) first = x;
    }
  int carryL = block_excl_max_scan(last, warp_tot, -(1 << 20));
  int carryR = block_excl_min_suffix_scan(first, warp_tot, (1 << 20));
  int l = carryL;
  for (int x = c0; x < c1; x++) {
    if (flag[x]) l = x;
    int d = x - l;
    gs[x] = (uint16_t)(d >= (int)EDT_INF16 ? EDT_INF16 : d);
  }
  int r = carryR;
  for (int x = c1 - 1; x >= c0; x--) {
    if (flag[x]) r = x;
    int d = r - x;
    if (d < (int)gs[x]) gs[x] = (uint16_t)d;
  }
  if (threadIdx.x == 0 && (carryR < (1 << 20) || first < (1 << 20))) *any_obstacle = 1u;
  __syncthreads();

  // store g (and the obstacle layer: same linear index x + y*W as the occupancy)
  uint16_t *grow = g + (size_t)y * W;
  if ((W & 7) == 0) {
    for (int i = threadIdx.x; i < W / 8; i += blockDim.x) ((uint4 *)grow)[i] = ((const uint4 *)gs)[i];
  } else {
    for (int i = threadIdx.x; i < W; i += blockDim.x) grow[i] = gs[i];
  }
  if (obstacle) {
    float *orow = obstacle + (size_t)y * W;
    if ((W & 3) == 0) {
      for (int i = threadIdx.x; i < W / 4; i += blockDim.x) {
        uchar4 f = ((const uchar4 *)flag)[i];
        float4 o = make_float4(f.x ? 0.f : 1.f, f.y ? 0.f : 1.f, f.z ? 0.f : 1.f, f.w ? 0.f : 1.f);
        ((float4 *)orow)[i] = o;
      }
    } else {
      for (int i = threadIdx.x; i < W; i += blockDim.x) orow[i] = flag[i] ? 0.f : 1.f;
    }
  }
}

// pass 2.
#define EDT_TX 64
#define EDT_TY 64
#define EDT_R 32
__global__ void __launch_bounds__(256) edt_col_kernel(const uint16_t *__restrict__ g, int W, int H, float res_f,
                                                      float *__restrict__ distance,
                                                      const uint32_t *__restrict__ any_obstacle) {
  __shared__ uint16_t gs[EDT_TY + 2 * EDT_R][EDT_TX];
  const int x0 = blockIdx.x * EDT_TX, y0 = blockIdx.y * EDT_TY;
  const int tx = threadIdx.x & (EDT_TX - 1), tyb = threadIdx.x / EDT_TX;  // 4 row groups
  const int x = x0 + tx;
  const size_t n = (size_t)W * H;
  for (int r = tyb; r < EDT_TY + 2 * EDT_R; r += 256 / EDT_TX) {
    int yy = y0 - EDT_R + r;
    uint16_t v = EDT_INF16;
    if (x < W && yy >= 0 && yy < H) v = g[(size_t)yy * W + x];
    gs[r][tx] = v;
  }
  __syncthreads();
  if (x >= W) return;
  const bool empty_map = (*any_obstacle == 0u);
  for (int ry = tyb; ry < EDT_TY; ry += 256 / EDT_TX) {
    const int y = y0 + ry;
    if (y >= H) break;
    const int sr = ry + EDT_R;  // row of (x,y) inside gs
    unsigned own = gs[sr][tx];
    unsigned best = (own == EDT_INF16) ? EDT_INF32 : own * own;
    if (!empty_map) {
      const int kmax = max(y, H - 1 - y);
      for (int k = 1; k <= kmax; k++) {
        const unsigned kk = (unsigned)k * (unsigned)k;
        if (kk >= best) break;
        unsigned a, b;
        if (k <= EDT_R) {
          a = gs[sr - k][tx];
          b = gs[sr + k][tx];
        } else {
          a = (y - k >= 0) ? g[(size_t)(y - k) * W + x] : EDT_INF16;
          b = (y + k < H) ? g[(size_t)(y + k) * W + x] : EDT_INF16;
        }
        unsigned m = min(a, b);
        if (m != EDT_INF16) {
          unsigned cand = m * m + kk;
          best = min(best, cand);
        }
      }
    }
    // cv2 returns the constant 65536 for an image without any zero pixel
    float d = (best == EDT_INF32) ? 65536.0f : __fsqrt_rn((float)best);
    distance[n - 1 - ((size_t)y * W + x)] = __fmul_rn(d, res_f);
  }
}

static const pred_lut *edt_lut_dev(hpp_ctx *c) { return (const pred_lut *)(c->d_flags + 32); }

// row pass launcher shared with inflate.cu (lut_dev: 8 words in device memory)
hpp_status hpp_row_dist_launch(hpp_ctx *c, const int8_t *occ_dev, int W, int H, const void *lut_dev, uint16_t *g,
                               float *obstacle, uint32_t *any_obstacle) {
  size_t smem = (size_t)((W + 15) & ~15) + (size_t)W * 2 + 16;
  if (smem > 48 * 1024)
    HPP_CUDA(c, cudaFuncSetAttribute(edt_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  edt_row_kernel<<<H, 256, smem, c->stream>>>(occ_dev, W, H, (const pred_lut *)lut_dev, g, obstacle, any_obstacle);
  HPP_LAUNCHED(c);
  return HPP_OK;
}

hpp_status hpp_edt_build(hpp_map *m) {
  hpp_ctx *c = m->ctx;
  const int W = m->g.W, H = m->g.H;
  if (W > 46340 || H > 46340) HPP_FAIL(c, HPP_ERR_UNSUPPORTED, "EDT supports maps up to 46340 cells per side");
  // obstacle <=> data > free_thres_ (20) || data < 0   (RS/src/Grid_map.cpp:41)
  pred_lut lut;
  memset(&lut, 0, sizeof lut);
  for (int v = -128; v < 128; v++)
    if (v > 20 || v < 0) lut.w[((uint8_t)v) >> 5] |= 1u << (((uint8_t)v) & 31);
  uint32_t *h = c->h_flags + 32;
  memcpy(h, &lut, sizeof lut);
  HPP_CUDA(c, cudaMemcpyAsync(c->d_flags + 32, h, sizeof lut, cudaMemcpyHostToDevice, c->stream));
  HPP_CUDA(c, cudaMemsetAsync(c->d_flags + 40, 0, sizeof(uint32_t), c->stream));
  hpp_status s = hpp_row_dist_launch(c, m->d_occ, W, H, edt_lut_dev(c), m->d_g, m->d_obstacle, c->d_flags + 40);
  if (s != HPP_OK) return s;
  dim3 grid(hpp_div_up(W, EDT_TX), hpp_div_up(H, EDT_TY));
  edt_col_kernel<<<grid, 256, 0, c->stream>>>(m->d_g, W, H, m->g.res_f, m->d_distance, c->d_flags + 40);
  HPP_LAUNCHED(c);
  return HPP_OK;
}
