// edt.cu -- K1: grid_map "obstacle" and "distance" layers (replaces RS/src/Grid_map.cpp:22-74:
// threshold -> cv::distanceTransform(DIST_L2, DIST_MASK_PRECISE) -> flip(-1) -> * resolution).
//
// Separable exact Euclidean distance transform in integer arithmetic:
//   pass 1 (rows, along x):  g(x,y) = distance along x to the nearest obstacle cell of row y (uint16,
//                            0xFFFF = the row has none).  One warp per row, bit-parallel: ballot words of the
//                            obstacle predicate, clz/ffs inside a word, per-word carries across words.
//   pass 2 (columns, along y): d2(x,y) = min_y' g(x,y')^2 + (y-y')^2 by an outward search that stops as
//                            soon as k^2 >= best.  64x128 output tiles with a +-32-row halo of g in shared
//                            memory (adjacent threads = adjacent x => coalesced, 16-byte staging loads); rows
//                            beyond the halo are read through L2.  Epilogue: sqrtf((float)d2) * (float)res (both correctly
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

// pass 1.  One warp per row, bit-parallel: the obstacle predicate of 32 consecutive cells is one ballot word, the
// nearest obstacle inside a word is a clz / ffs on the masked word, and across words it is a per-word carry (the
// last set bit before / the first set bit after the word), found by a 32-lane chunked scan over the row's words.
// Every global access is coalesced (lane = consecutive cell) and there is no sequential dependency along the row.
// Shared memory per warp: mask[NW], prev[NW], next[NW] (NW = ceil(W/32) words).
// obstacle may be null (inflate reuses this kernel).  lut_dev: predicate table in device memory.
#define EDT_ROW_WARPS 4
__global__ void __launch_bounds__(EDT_ROW_WARPS * 32) edt_row_kernel(const int8_t *__restrict__ occ, int W, int H,
                                                                     const pred_lut *__restrict__ lut_dev,
                                                                     uint16_t *__restrict__ g,
                                                                     float *__restrict__ obstacle,
                                                                     uint32_t *__restrict__ any_obstacle) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ pred_lut lut;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int NW = (W + 31) >> 5;
  uint32_t *mask = (uint32_t *)smem + (size_t)wid * 3 * NW;
  int *prev = (int *)(mask + NW), *next = prev + NW;
  if (threadIdx.x < 8) lut.w[threadIdx.x] = lut_dev->w[threadIdx.x];
  __syncthreads();
  const int FAR = 1 << 20;
  for (int y = blockIdx.x * EDT_ROW_WARPS + wid; y < H; y += gridDim.x * EDT_ROW_WARPS) {
    const int8_t *row = occ + (size_t)y * W;
    float *orow = obstacle ? obstacle + (size_t)y * W : nullptr;
    // predicate words (+ the obstacle layer: same linear index x + y*W as the occupancy, 1 = free)
    uint32_t any = 0;
    int j = 0;
    for (; j + 4 <= NW; j += 4) {  // four independent byte loads in flight per lane
      int8_t v[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int x = ((j + u) << 5) + lane;
        v[u] = x < W ? row[x] : (int8_t)0;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int x = ((j + u) << 5) + lane;
        const bool p = x < W && lut_test(lut, v[u]);
        const uint32_t m = __ballot_sync(0xffffffffu, p);
        if (lane == 0) mask[j + u] = m;
        any |= m;
        if (orow && x < W) orow[x] = p ? 0.f : 1.f;
      }
    }
    for (; j < NW; j++) {
      const int x = (j << 5) + lane;
      const bool p = x < W && lut_test(lut, row[x]);
      const uint32_t m = __ballot_sync(0xffffffffu, p);
      if (lane == 0) mask[j] = m;
      any |= m;
      if (orow && x < W) orow[x] = p ? 0.f : 1.f;
    }
    __syncwarp();
    // carries: lane l scans the contiguous chunk of words [l*C, (l+1)*C)
    const int C = (NW + 31) >> 5;
    const int j0 = min(NW, lane * C), j1 = min(NW, j0 + C);
    int last = -FAR, first = FAR;
    for (int j = j0; j < j1; j++) {
      const uint32_t m = mask[j];
      if (m) {
        last = (j << 5) + 31 - __clz(m);
        if (first == FAR) first = (j << 5) + __ffs(m) - 1;
      }
    }
    int inc_l = last, inc_f = first;  // inclusive max-scan of last (up), inclusive min-scan of first (down)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, inc_l, d);
      if (lane >= d) inc_l = max(inc_l, o);
      const int q = __shfl_down_sync(0xffffffffu, inc_f, d);
      if (lane + d < 32) inc_f = min(inc_f, q);
    }
    int carry_l = __shfl_up_sync(0xffffffffu, inc_l, 1), carry_f = __shfl_down_sync(0xffffffffu, inc_f, 1);
    if (lane == 0) carry_l = -FAR;
    if (lane == 31) carry_f = FAR;
    for (int j = j0; j < j1; j++) {  // prev[j]: last obstacle before word j
      prev[j] = carry_l;
      const uint32_t m = mask[j];
      if (m) carry_l = (j << 5) + 31 - __clz(m);
    }
    for (int j = j1 - 1; j >= j0; j--) {  // next[j]: first obstacle after word j
      next[j] = carry_f;
      const uint32_t m = mask[j];
      if (m) carry_f = (j << 5) + __ffs(m) - 1;
    }
    __syncwarp();
    uint16_t *grow = g + (size_t)y * W;
#pragma unroll 4
    for (int j = 0; j < NW; j++) {
      const int x = (j << 5) + lane;
      const uint32_t m = mask[j];
      const uint32_t ml = m & (0xffffffffu >> (31 - lane)), mr = m & (0xffffffffu << lane);
      const int li = ml ? (j << 5) + 31 - __clz(ml) : prev[j];
      const int ri = mr ? (j << 5) + __ffs(mr) - 1 : next[j];
      const int d = min(x - li, ri - x);
      if (x < W) grow[x] = (uint16_t)(d >= (int)EDT_INF16 ? EDT_INF16 : d);
    }
    if (lane == 0 && any) *any_obstacle = 1u;
    __syncwarp();
  }
}

// pass 2.  64 columns x 128 rows of output per CTA, +-32 rows of halo in shared memory (1.5x read amplification of g).
// The kernel works on any window of rows of a g buffer: the whole map (hpp_edt_build: flipped grid_map layout out), or
// one rank's row band inside a band + halo buffer (hpp_edt_band_cols, multi-GPU: plain row-major metres out, and a flag
// when some distance reaches past the rows the buffer holds, so the caller can widen the halo).
#define EDT_TX 64
#define EDT_TY 128
#define EDT_R 32
struct edt_band {
  int y_begin, y_count;   // output rows [y_begin, y_begin + y_count) of the g buffer
  int more_above;         // map rows exist above row 0 of the buffer (not held by it)
  int more_below;         // map rows exist below row H-1 of the buffer
  uint32_t *unresolved;   // set to 1 when a result could be beaten by an obstacle outside the buffer
};

// epilogue shared by the column kernels: sqrtf((float)d2) * (float)res at the 180-degree-rotated index, plus the compact
// copy for K4 (squared cell distance clamped to 255) at the same flipped cell (i, j) = (W-1-x, H-1-y)
__device__ __forceinline__ void edt_store(unsigned best, int x, int y, int W, int H, float res_f,
                                          float *__restrict__ distance, uint8_t *__restrict__ d2u8, int d2_tiles_x) {
  // cv2 returns the constant 65536 for an image without any zero pixel
  const float d = (best == EDT_INF32) ? 65536.0f : __fsqrt_rn((float)best);
  distance[(size_t)W * H - 1 - ((size_t)y * W + x)] = __fmul_rn(d, res_f);
  if (d2u8) d2u8[hpp_d2_addr(W - 1 - x, H - 1 - y, d2_tiles_x)] = (uint8_t)min(best, 255u);
}

// far_flags (whole-map mode): one word per output tile + a global word at far_flags[-1]; far_mask: one bit per cell of a
// tile, [tile][row][half].  A cell whose search needs rows beyond the +-32-row halo parks its best-so-far in its slot of
// the distance layer, sets its bit and flags its tile; edt_far_kernel (block-minimum pyramid) finishes exactly those
// cells instead of walking the rows one by one through L2 -- the cost per cell stays bounded on sparse maps.
template <bool BAND>
__global__ void __launch_bounds__(256) edt_col_kernel(const uint16_t *__restrict__ g, int W, int H, float res_f,
                                                      float *__restrict__ distance,
                                                      const uint32_t *__restrict__ any_obstacle, edt_band bd,
                                                      uint8_t *__restrict__ d2u8, int d2_tiles_x,
                                                      uint32_t *__restrict__ far_flags, uint32_t *__restrict__ far_mask) {
  __shared__ __align__(16) uint16_t gs[EDT_TY + 2 * EDT_R][EDT_TX];
  const int x0 = blockIdx.x * EDT_TX, y0 = (BAND ? bd.y_begin : 0) + blockIdx.y * EDT_TY;
  const int y_end = BAND ? bd.y_begin + bd.y_count : H;
  const int tx = threadIdx.x & (EDT_TX - 1), tyb = threadIdx.x / EDT_TX;  // 4 row groups
  const int x = x0 + tx;
  if ((W & 7) == 0 && x0 + EDT_TX <= W) {
    // 16-byte loads: a tile row is 8 chunks of 8 uint16
    for (int i = threadIdx.x; i < (EDT_TY + 2 * EDT_R) * (EDT_TX / 8); i += 256) {
      const int r = i / (EDT_TX / 8), c8 = i % (EDT_TX / 8);
      const int yy = y0 - EDT_R + r;
      uint4 v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
      if (yy >= 0 && yy < H) v = __ldg((const uint4 *)(g + (size_t)yy * W + x0) + c8);
      *((uint4 *)&gs[r][0] + c8) = v;
    }
  } else {
    for (int r = tyb; r < EDT_TY + 2 * EDT_R; r += 256 / EDT_TX) {
      int yy = y0 - EDT_R + r;
      uint16_t v = EDT_INF16;
      if (x < W && yy >= 0 && yy < H) v = g[(size_t)yy * W + x];
      gs[r][tx] = v;
    }
  }
  __syncthreads();
  if (x >= W) return;
  const bool empty_map = !BAND && (*any_obstacle == 0u);
  bool open_result = false;
  for (int ry = tyb; ry < EDT_TY; ry += 256 / EDT_TX) {
    const int y = y0 + ry;
    if (y >= y_end) break;
    const int sr = ry + EDT_R;  // row of (x,y) inside gs
    unsigned own = gs[sr][tx];
    unsigned best = (own == EDT_INF16) ? EDT_INF32 : own * own;
    bool far = false;
    if (!empty_map) {
      const int kmax = max(y, H - 1 - y);
      for (int k = 1; k <= kmax; k++) {
        const unsigned kk = (unsigned)k * (unsigned)k;
        if (kk >= best) break;
        unsigned a, b;
        if (k <= EDT_R) {
          a = gs[sr - k][tx];
          b = gs[sr + k][tx];
        } else if (!BAND) {
          far = true;  // rows beyond the halo: this tile goes to edt_far_kernel
          break;
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
    if (!BAND) {
      const size_t tile = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
      const unsigned fm = __ballot_sync(__activemask(), far);  // a warp = 32 adjacent columns of one row
      if ((threadIdx.x & 31) == 0) far_mask[(tile * EDT_TY + ry) * 2 + ((threadIdx.x >> 5) & 1)] = fm;
      if (far) {
        distance[(size_t)W * H - 1 - ((size_t)y * W + x)] = __uint_as_float(best);  // parked for edt_far_kernel
        far_flags[tile] = 1u;
        far_flags[-1] = 1u;
      } else {
        edt_store(best, x, y, W, H, res_f, distance, d2u8, d2_tiles_x);
      }
      continue;
    }
    // cv2 returns the constant 65536 for an image without any zero pixel
    float d = (best == EDT_INF32) ? 65536.0f : __fsqrt_rn((float)best);
    if (BAND) {
      // the nearest obstacle outside the buffer is at least `lim` rows away: the result stands iff best <= lim^2
      unsigned long long lim = ~0ull;
      if (bd.more_above) lim = (unsigned long long)(y + 1);
      if (bd.more_below) lim = min(lim, (unsigned long long)(H - y));
      if (lim != ~0ull && (best == EDT_INF32 || (unsigned long long)best > lim * lim)) open_result = true;
      distance[(size_t)(y - bd.y_begin) * W + x] = __fmul_rn(d, res_f);
    }
  }
  if (BAND && open_result) *bd.unresolved = 1u;
}

// ---- far path (sparse maps).  bmin[b][x] = min of g(x, y) over the EDT_BLK rows of block b; a block whose lower bound
// bmin^2 + (rows to its nearest row)^2 cannot beat the best so far is skipped as a whole.
#define EDT_BLK 16
__global__ void __launch_bounds__(256) edt_blockmin_kernel(const uint16_t *__restrict__ g, int W, int H,
                                                           uint16_t *__restrict__ bmin,
                                                           const uint32_t *__restrict__ far_flags) {
  if (far_flags[-1] == 0u) return;  // no cell asked for the far path
  const int x = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (x >= W) return;
  unsigned m = EDT_INF16;
  const int y1 = min(H, (b + 1) * EDT_BLK);
  for (int y = b * EDT_BLK; y < y1; y++) m = min(m, (unsigned)g[(size_t)y * W + x]);
  bmin[(size_t)b * W + x] = (uint16_t)m;
}

__global__ void __launch_bounds__(256) edt_far_kernel(const uint16_t *__restrict__ g, const uint16_t *__restrict__ bmin,
                                                      int W, int H, float res_f, float *__restrict__ distance,
                                                      uint8_t *__restrict__ d2u8, int d2_tiles_x,
                                                      const uint32_t *__restrict__ far_flags,
                                                      const uint32_t *__restrict__ far_mask) {
  const size_t tile = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
  if (far_flags[tile] == 0u) return;  // the column kernel finished this tile
  const int tx = threadIdx.x & (EDT_TX - 1), tyb = threadIdx.x / EDT_TX;
  const int x = blockIdx.x * EDT_TX + tx;
  if (x >= W) return;
  for (int ry = tyb; ry < EDT_TY; ry += 256 / EDT_TX) {
    const int y = blockIdx.y * EDT_TY + ry;
    if (y >= H) break;
    const unsigned fm = far_mask[(tile * EDT_TY + ry) * 2 + ((threadIdx.x >> 5) & 1)];
    if (!((fm >> (threadIdx.x & 31)) & 1u)) continue;  // finished by the column kernel
    const size_t slot = (size_t)W * H - 1 - ((size_t)y * W + x);
    const unsigned parked = __float_as_uint(distance[slot]);  // best over the rows y-32 .. y+32
    unsigned long long best = parked == EDT_INF32 ? ~0ull : (unsigned long long)parked;
#pragma unroll
    for (int dir = -1; dir <= 1; dir += 2) {
      int yy = y + dir * (EDT_R + 1);
      while (yy >= 0 && yy < H) {
        const unsigned long long dy = (unsigned long long)(dir > 0 ? yy - y : y - yy);
        if (dy * dy >= best) break;
        // entering a block at its nearest row: skip it when its minimum cannot win
        if ((dir > 0 && (yy & (EDT_BLK - 1)) == 0) || (dir < 0 && (yy & (EDT_BLK - 1)) == EDT_BLK - 1)) {
          const unsigned mb = bmin[(size_t)(yy / EDT_BLK) * W + x];
          if (mb == EDT_INF16 || (unsigned long long)mb * mb + dy * dy >= best) {
            yy += dir * EDT_BLK;
            continue;
          }
        }
        const unsigned v = g[(size_t)yy * W + x];
        if (v != EDT_INF16) {
          const unsigned long long cand = (unsigned long long)v * v + dy * dy;
          best = cand < best ? cand : best;
        }
        yy += dir;
      }
    }
    edt_store(best == ~0ull ? EDT_INF32 : (unsigned)best, x, y, W, H, res_f, distance, d2u8, d2_tiles_x);
  }
}

static const pred_lut *edt_lut_dev(hpp_ctx *c) { return (const pred_lut *)(c->d_flags + 32); }

// row pass launcher shared with inflate.cu (lut_dev: 8 words in device memory)
hpp_status hpp_row_dist_launch(hpp_ctx *c, const int8_t *occ_dev, int W, int H, const void *lut_dev, uint16_t *g,
                               float *obstacle, uint32_t *any_obstacle) {
  const size_t smem = (size_t)EDT_ROW_WARPS * 3 * ((W + 31) / 32) * sizeof(uint32_t);
  if (smem > 48 * 1024)
    HPP_CUDA(c, cudaFuncSetAttribute(edt_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = hpp_div_up(H, EDT_ROW_WARPS);
  if (grid > c->sm_count * 8) grid = c->sm_count * 8;
  edt_row_kernel<<<grid, EDT_ROW_WARPS * 32, smem, c->stream>>>(occ_dev, W, H, (const pred_lut *)lut_dev, g, obstacle,
                                                                any_obstacle);
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
  edt_band none;
  memset(&none, 0, sizeof none);
  // far-path state (scratch slot 5): [global flag | one flag per tile | block minima]; the two far kernels return at
  // once unless the column kernel flagged a tile (maps with distances beyond the 32-row halo)
  const size_t n_tiles = (size_t)grid.x * grid.y, n_blk = (size_t)hpp_div_up(H, EDT_BLK);
  const size_t flag_bytes = ((n_tiles + 1) * sizeof(uint32_t) + 255) & ~(size_t)255;
  const size_t mask_bytes = n_tiles * EDT_TY * 2 * sizeof(uint32_t);  // written by the column kernel, never cleared
  void *far = nullptr;
  s = hpp_scratch(c, 5, flag_bytes + mask_bytes + n_blk * W * sizeof(uint16_t), &far);
  if (s != HPP_OK) return s;
  uint32_t *far_flags = (uint32_t *)far + 1;
  uint32_t *far_mask = (uint32_t *)((uint8_t *)far + flag_bytes);
  uint16_t *bmin = (uint16_t *)((uint8_t *)far + flag_bytes + mask_bytes);
  HPP_CUDA(c, cudaMemsetAsync(far, 0, flag_bytes, c->stream));
  edt_col_kernel<false><<<grid, 256, 0, c->stream>>>(m->d_g, W, H, m->g.res_f, m->d_distance, c->d_flags + 40, none,
                                                     m->d_d2u8, m->d2_tiles_x, far_flags, far_mask);
  HPP_LAUNCHED(c);
  edt_blockmin_kernel<<<dim3(hpp_div_up(W, 256), (unsigned)n_blk), 256, 0, c->stream>>>(m->d_g, W, H, bmin, far_flags);
  HPP_LAUNCHED(c);
  edt_far_kernel<<<grid, 256, 0, c->stream>>>(m->d_g, bmin, W, H, m->g.res_f, m->d_distance, m->d_d2u8, m->d2_tiles_x,
                                              far_flags, far_mask);
  HPP_LAUNCHED(c);
  return HPP_OK;
}

// ---- row bands (multi-GPU, mgpu.cu): pass 1 of a band of rows, pass 2 of a band inside a band + halo buffer
hpp_status hpp_edt_band_rows(hpp_ctx *c, const int8_t *occ_band_dev, int W, int rows, uint16_t *g_band) {
  pred_lut lut;
  memset(&lut, 0, sizeof lut);
  for (int v = -128; v < 128; v++)
    if (v > 20 || v < 0) lut.w[((uint8_t)v) >> 5] |= 1u << (((uint8_t)v) & 31);
  uint32_t *h = c->h_flags + 32;
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));  // the pinned staging words may still be in flight
  memcpy(h, &lut, sizeof lut);
  HPP_CUDA(c, cudaMemcpyAsync(c->d_flags + 32, h, sizeof lut, cudaMemcpyHostToDevice, c->stream));
  HPP_CUDA(c, cudaMemsetAsync(c->d_flags + 40, 0, sizeof(uint32_t), c->stream));
  return hpp_row_dist_launch(c, occ_band_dev, W, rows, edt_lut_dev(c), g_band, nullptr, c->d_flags + 40);
}

hpp_status hpp_edt_band_cols(hpp_ctx *c, const uint16_t *g_ext, int W, int rows_ext, int y_begin, int y_count,
                             int more_above, int more_below, float res_f, float *distance_band, uint32_t *unresolved) {
  if (y_count <= 0) return HPP_OK;
  edt_band bd;
  bd.y_begin = y_begin;
  bd.y_count = y_count;
  bd.more_above = more_above;
  bd.more_below = more_below;
  bd.unresolved = unresolved;
  dim3 grid(hpp_div_up(W, EDT_TX), hpp_div_up(y_count, EDT_TY));
  edt_col_kernel<true><<<grid, 256, 0, c->stream>>>(g_ext, W, rows_ext, res_f, distance_band, c->d_flags + 40, bd,
                                                    nullptr, 0, nullptr, nullptr);
  HPP_LAUNCHED(c);
  return HPP_OK;
}
