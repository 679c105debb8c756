// heuristic.cu -- K3a/K3b: the holonomic-with-obstacles heuristic of both planners
// (HybridAstar::holonomicCostsWithObstacles_planning, RS/src/HybridAstar.cpp:353-551 and
// DB/src/HybridAstar.cpp:198-313).
//
// Both reference implementations are (bucket / lazy-heap) Dijkstra sweeps from the goal over the
// 8-connected occupancy grid.  Outside their sequential quirks (SURVEY hazards H1/H2) the result is the
// unique least fixed point of
//     RS:      cost[v] = min_u fl(cost[u] + w(dir, v)),  w = 1.0 | 1.414, + 5.0 if an in-bounds 4-neighbour of
//                                                          v is occupied   (RS/src/HybridAstar.cpp:386-390,:500-526)
//     Dubins:  cost[v] = min(1000, min_u fl(cost[u] + fl(hypot(dir) + 10*cnt3x3(u))))  (DB/src/HybridAstar.cpp:198-221,:286-297)
// over free cells v, which does not depend on the relaxation order.  The device therefore runs an
// asynchronous, tile-local label-correcting relaxation:
//   * the field is cut into 64x64 tiles; a CTA owns one active tile per round, stages the tile's costs with a
//     one-cell halo and the per-cell class bytes in shared memory, and relaxes to the tile-local fixed point;
//   * inside the tile only the frontier is touched: "changed" cells are kept as 64-bit row masks, dilated by
//     one cell with shifts, and compacted into a dense work list with popc + a warp shuffle scan;
//   * a tile whose border cells improved flags the neighbouring tiles for the next round; rounds are plain
//     stream-ordered launches (no inter-CTA spinning), the host polls a convergence word every few rounds;
//   * fields (goals) are a batch dimension: blockIdx.y.
// Costs are fp64 with exactly the reference's additions (library built with -fmad=false).
#include <math.h>

#include "hpp_internal.cuh"

#define HT 64              // tile side
#define HP (HT + 2)        // with halo
#define H_THREADS 256

struct heur_weights {
  // RS: [near][0 = cardinal, 1 = diagonal];  Dubins: [dir class][count 0..9]
  double rs[2][2];
  double db[2][10];
  double init;  // +inf (RS) or 1000 (Dubins)
};

__global__ void __launch_bounds__(256) heur_init_kernel(double *__restrict__ cost, size_t n, size_t n_fields,
                                                        const int32_t *__restrict__ goal_xy, int W, double init,
                                                        uint8_t *__restrict__ act0, uint8_t *__restrict__ act1,
                                                        int tiles_x, int n_tiles) {
  const size_t total = n * n_fields;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) cost[i] = init;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n_tiles * n_fields; i += stride) {
    act0[i] = 0;
    act1[i] = 0;
  }
}

__global__ void heur_seed_kernel(double *__restrict__ cost, size_t n, size_t n_fields,
                                 const int32_t *__restrict__ goal_xy, int W, uint8_t *__restrict__ act0,
                                 int tiles_x, int n_tiles) {
  size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_fields) return;
  int gx = goal_xy[2 * f], gy = goal_xy[2 * f + 1];
  cost[f * n + (size_t)gy * W + gx] = 0.0;  // goal = 0 even if occupied (RS/src/HybridAstar.cpp:378)
  act0[f * n_tiles + (size_t)(gy / HT) * tiles_x + gx / HT] = 1;
}

template <int VARIANT>
__global__ void __launch_bounds__(H_THREADS) heur_round_kernel(double *__restrict__ cost_all,
                                                               const uint8_t *__restrict__ cls, int W, int H,
                                                               int tiles_x, int tiles_y, heur_weights wt,
                                                               uint8_t *__restrict__ act_cur,
                                                               uint8_t *__restrict__ act_next,
                                                               uint32_t *__restrict__ work, int first_round,
                                                               uint32_t *__restrict__ visits) {
  const int n_tiles = tiles_x * tiles_y;
  const int tile = blockIdx.x, field = blockIdx.y;
  uint8_t *my_flag = act_cur + (size_t)field * n_tiles + tile;
  __shared__ int s_active;
  if (threadIdx.x == 0) s_active = *my_flag;
  __syncthreads();
  if (!s_active) return;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *cs = (double *)smem_raw;                                  // HP*HP costs
  unsigned long long *need = (unsigned long long *)(cs + HP * HP);  // [HT]
  unsigned long long *freem = need + HT;                            // [HT]
  unsigned long long *chg = freem + HT;                             // [2][HT]
  unsigned short *list = (unsigned short *)(chg + 2 * HT);          // [HT*HT]
  unsigned char *cl = (unsigned char *)(list + HT * HT);            // HP*HP class bytes
  __shared__ int s_total;
  __shared__ unsigned s_border;
  __shared__ int s_any;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int tx = tile % tiles_x, ty = tile / tiles_x;
  const int x0 = tx * HT, y0 = ty * HT;
  double *cost = cost_all + (size_t)field * W * H;

  // ---- stage tile + halo
  for (int e = tid; e < HP * HP; e += H_THREADS) {
    int r = e / HP, c = e - r * HP;
    int gx = x0 + c - 1, gy = y0 + r - 1;
    double v = wt.init;
    unsigned char k = CLS_BLOCKED;
    if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
      v = cost[(size_t)gy * W + gx];
      k = cls[(size_t)gy * W + gx];
    }
    cs[e] = v;
    cl[e] = k;
  }
  if (tid == 0) {
    s_border = 0;
    s_any = 0;
    *my_flag = 0;
  }
  __syncthreads();
  // free masks: warp w builds rows w, w+8, ...
  for (int r = wid; r < HT; r += H_THREADS / 32) {
    unsigned lo = __ballot_sync(0xffffffffu, !(cl[(r + 1) * HP + 1 + lane] & CLS_BLOCKED));
    unsigned hi = __ballot_sync(0xffffffffu, !(cl[(r + 1) * HP + 33 + lane] & CLS_BLOCKED));
    if (lane == 0) {
      unsigned long long fm = (unsigned long long)lo | ((unsigned long long)hi << 32);
      freem[r] = fm;
      // first sweep: everything (first round = goal tile) or the ring next to the halo
      unsigned long long ring = (r == 0 || r == HT - 1) ? ~0ull : (1ull | (1ull << 63));
      need[r] = first_round ? fm : (fm & ring);
      chg[r] = 0;
      chg[HT + r] = 0;
    }
  }
  __syncthreads();

  int sweep = 0;
  for (;; sweep++) {
    unsigned long long *chg_w = chg + (sweep & 1) * HT;        // written by this sweep's relax phase
    const unsigned long long *chg_r = chg + ((sweep + 1) & 1) * HT;  // written by the previous sweep
    // ---- warp 0: dilate the previous sweep's changes into this sweep's need masks and compact them
    if (wid == 0) {
      unsigned long long nd[2];
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int r = lane * 2 + q;
        unsigned long long m;
        if (sweep == 0) {
          m = need[r];
        } else {
          unsigned long long a = (r > 0) ? chg_r[r - 1] : 0ull, b = chg_r[r], c = (r < HT - 1) ? chg_r[r + 1] : 0ull;
          unsigned long long ac = a | c;
          m = (ac | (ac << 1) | (ac >> 1) | (b << 1) | (b >> 1)) & freem[r];
        }
        nd[q] = m;
      }
      __syncwarp();
      // the buffer the coming relax phase writes was last read one sweep ago: clear it now
      chg_w[lane * 2] = 0;
      chg_w[lane * 2 + 1] = 0;
      int cnt = __popcll(nd[0]) + __popcll(nd[1]);
      int inc = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
      }
      int off = inc - cnt;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        unsigned long long m = nd[q];
        const int r = lane * 2 + q;
        while (m) {
          int b = __ffsll((long long)m) - 1;
          m &= m - 1;
          list[off++] = (unsigned short)(r * HT + b);
        }
      }
      if (lane == 31) s_total = inc;
    }
    __syncthreads();
    const int total = s_total;
    if (total == 0) break;
    // ---- relax the compacted frontier
    for (int k = tid; k < total; k += H_THREADS) {
      const int cell = list[k];
      const int r = cell >> 6, c = cell & 63;
      const int e = (r + 1) * HP + (c + 1);
      double best;
      if (VARIANT == HPP_VARIANT_RS) {
        const double mc = fmin(fmin(cs[e - HP], cs[e + HP]), fmin(cs[e - 1], cs[e + 1]));
        const double md = fmin(fmin(cs[e - HP - 1], cs[e - HP + 1]), fmin(cs[e + HP - 1], cs[e + HP + 1]));
        const int nr = (cl[e] & CLS_NEAR4) ? 1 : 0;
        best = fmin(mc + wt.rs[nr][0], md + wt.rs[nr][1]);
      } else {
        best = wt.init;
#pragma unroll
        for (int dy = -1; dy <= 1; dy++)
#pragma unroll
          for (int dx = -1; dx <= 1; dx++) {
            if (dx == 0 && dy == 0) continue;
            const int u = e + dy * HP + dx;
            const double w = wt.db[(dx != 0 && dy != 0) ? 1 : 0][cl[u] & CLS_CNT_MASK];
            best = fmin(best, cs[u] + w);
          }
      }
      if (best < cs[e]) {
        cs[e] = best;
        atomicOr(&chg_w[r], 1ull << c);
        unsigned bf = 0;
        if (r == 0) bf |= 1u | (c == 0 ? 16u : 0u) | (c == HT - 1 ? 32u : 0u);
        if (r == HT - 1) bf |= 2u | (c == 0 ? 64u : 0u) | (c == HT - 1 ? 128u : 0u);
        if (c == 0) bf |= 4u;
        if (c == HT - 1) bf |= 8u;
        if (bf) atomicOr(&s_border, bf);
        s_any = 1;
      }
    }
    __syncthreads();
  }

  if (tid == 0) atomicAdd(visits, 1u);
  if (!s_any) return;
  // ---- write the interior back
  for (int e = tid; e < HT * HT; e += H_THREADS) {
    int r = e >> 6, c = e & 63;
    int gx = x0 + c, gy = y0 + r;
    if (gx < W && gy < H) cost[(size_t)gy * W + gx] = cs[(r + 1) * HP + c + 1];
  }
  // ---- wake the neighbours whose halo changed
  if (tid < 8) {
    const int ddx[8] = {0, 0, -1, 1, -1, 1, -1, 1};
    const int ddy[8] = {-1, 1, 0, 0, -1, -1, 1, 1};
    if (s_border & (1u << tid)) {
      int nx = tx + ddx[tid], ny = ty + ddy[tid];
      if (nx >= 0 && nx < tiles_x && ny >= 0 && ny < tiles_y) {
        __threadfence();
        act_next[(size_t)field * n_tiles + ny * tiles_x + nx] = 1;
        *work = 1u;
      }
    }
  }
}

// reachable count / max finite cost per field
__global__ void __launch_bounds__(256) heur_stats_kernel(const double *__restrict__ cost, size_t n, double limit,
                                                         unsigned long long *__restrict__ out /* [f][2] */) {
  const size_t f = blockIdx.y;
  const double *c = cost + f * n;
  unsigned long long cnt = 0, mx = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double v = c[i];
    if (v < limit) {
      cnt++;
      unsigned long long b = (unsigned long long)__double_as_longlong(v);
      mx = b > mx ? b : mx;  // non-negative doubles order like their bit patterns
    }
  }
  for (int d = 16; d; d >>= 1) {
    cnt += __shfl_down_sync(0xffffffffu, cnt, d);
    unsigned long long o = __shfl_down_sync(0xffffffffu, mx, d);
    mx = o > mx ? o : mx;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out[2 * f], cnt);
    atomicMax(&out[2 * f + 1], mx);
  }
}

static void make_weights(int variant, heur_weights *w) {
  // RS/src/HybridAstar.cpp:389 (1.0 / 1.414) and :523 (movement_cost += 5.0)
  for (int nr = 0; nr < 2; nr++) {
    double c = 1.0, d = 1.414;
    if (nr) {
      c += 5.0;
      d += 5.0;
    }
    w->rs[nr][0] = c;
    w->rs[nr][1] = d;
  }
  // DB/src/HybridAstar.cpp:200-220: hypot(dx,dy) + (10.0 accumulated cnt times)
  for (int k = 0; k < 10; k++) {
    double pen = 0.0;
    for (int i = 0; i < k; i++) pen += 10.0;
    w->db[0][k] = hypot(1.0, 0.0) + pen;
    w->db[1][k] = hypot(1.0, 1.0) + pen;
  }
  w->init = (variant == HPP_VARIANT_RS) ? INFINITY : 1000.0;
}

static size_t heur_smem_bytes() {
  return sizeof(double) * HP * HP + sizeof(unsigned long long) * HT * 4 + sizeof(unsigned short) * HT * HT + HP * HP + 16;
}

extern "C" hpp_status hpp_heuristic_batch_dev(hpp_map *m, int variant, const int32_t *goal_xy, size_t n_fields,
                                              double *cost_dev, hpp_quirk_report *reports) {
  if (!m || !goal_xy || !cost_dev || n_fields == 0 || variant < 0 || variant > 1) return HPP_ERR_ARG;
  hpp_ctx *c = m->ctx;
  const int W = m->g.W, H = m->g.H;
  for (size_t f = 0; f < n_fields; f++)
    if (goal_xy[2 * f] < 0 || goal_xy[2 * f] >= W || goal_xy[2 * f + 1] < 0 || goal_xy[2 * f + 1] >= H)
      HPP_FAIL(c, HPP_ERR_ARG, "goal cell (%d,%d) of field %zu outside the %dx%d map", goal_xy[2 * f], goal_xy[2 * f + 1],
               f, W, H);
  if (n_fields > 65535) HPP_FAIL(c, HPP_ERR_ARG, "at most 65535 fields per call");
  HPP_CUDA(c, cudaSetDevice(c->device));
  const int tiles_x = hpp_div_up(W, HT), tiles_y = hpp_div_up(H, HT), n_tiles = tiles_x * tiles_y;
  // scratch slot 2: act flags (2 x n_fields x n_tiles bytes) + goals + stats
  const size_t act_bytes = ((size_t)n_tiles * n_fields + 255) & ~(size_t)255;
  const size_t goals_bytes = ((n_fields * 8) + 255) & ~(size_t)255;
  const size_t stats_bytes = n_fields * 16;
  void *scr;
  hpp_status s = hpp_scratch(c, 2, 2 * act_bytes + goals_bytes + stats_bytes, &scr);
  if (s != HPP_OK) return s;
  uint8_t *act[2] = {(uint8_t *)scr, (uint8_t *)scr + act_bytes};
  int32_t *d_goals = (int32_t *)((uint8_t *)scr + 2 * act_bytes);
  unsigned long long *d_stats = (unsigned long long *)((uint8_t *)scr + 2 * act_bytes + goals_bytes);
  void *h_goals;
  s = hpp_pinned(c, 1, goals_bytes + stats_bytes, &h_goals);
  if (s != HPP_OK) return s;
  memcpy(h_goals, goal_xy, n_fields * 8);
  HPP_CUDA(c, cudaMemcpyAsync(d_goals, h_goals, n_fields * 8, cudaMemcpyHostToDevice, c->stream));

  heur_weights wt;
  make_weights(variant, &wt);
  const size_t total = m->n * n_fields;
  int grid = (int)((total / 256 + 1) < (size_t)c->sm_count * 16 ? (total / 256 + 1) : (size_t)c->sm_count * 16);
  heur_init_kernel<<<grid, 256, 0, c->stream>>>(cost_dev, m->n, n_fields, d_goals, W, wt.init, act[0], act[1], tiles_x,
                                               n_tiles);
  HPP_LAUNCHED(c);
  heur_seed_kernel<<<hpp_div_up((long long)n_fields, 128), 128, 0, c->stream>>>(cost_dev, m->n, n_fields, d_goals, W,
                                                                               act[0], tiles_x, n_tiles);
  HPP_LAUNCHED(c);

  const size_t smem = heur_smem_bytes();
  HPP_CUDA(c, cudaFuncSetAttribute(heur_round_kernel<HPP_VARIANT_RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  HPP_CUDA(c, cudaFuncSetAttribute(heur_round_kernel<HPP_VARIANT_DUBINS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // convergence words: d_flags[0..15] = "round r woke somebody", d_flags[16] = tile visits
  const int K = 16;
  uint32_t *d_work = c->d_flags, *d_visits = c->d_flags + 16;
  HPP_CUDA(c, cudaMemsetAsync(d_visits, 0, sizeof(uint32_t), c->stream));
  dim3 rgrid(n_tiles, (unsigned)n_fields);
  uint32_t round = 0;
  bool done = false;
  // every cell can only be improved finitely often; the bound below is a safety net against a logic error
  const uint32_t max_rounds = 64u * (uint32_t)(tiles_x + tiles_y) * 64u + 1024u;
  while (!done && round < max_rounds) {
    HPP_CUDA(c, cudaMemsetAsync(d_work, 0, K * sizeof(uint32_t), c->stream));
    for (int k = 0; k < K; k++, round++) {
      if (variant == HPP_VARIANT_RS)
        heur_round_kernel<HPP_VARIANT_RS><<<rgrid, H_THREADS, smem, c->stream>>>(
            cost_dev, m->d_cls, W, H, tiles_x, tiles_y, wt, act[round & 1], act[(round + 1) & 1], d_work + k, round == 0,
            d_visits);
      else
        heur_round_kernel<HPP_VARIANT_DUBINS><<<rgrid, H_THREADS, smem, c->stream>>>(
            cost_dev, m->d_cls, W, H, tiles_x, tiles_y, wt, act[round & 1], act[(round + 1) & 1], d_work + k, round == 0,
            d_visits);
      HPP_LAUNCHED(c);
    }
    HPP_CUDA(c, cudaMemcpyAsync(c->h_flags, d_work, 17 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    HPP_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < K; k++)
      if (c->h_flags[k] == 0) {
        done = true;
        round = round - K + k + 1;
        break;
      }
  }
  if (!done) HPP_FAIL(c, HPP_ERR_CUDA, "heuristic relaxation did not converge in %u rounds", max_rounds);
  const uint32_t visits = c->h_flags[16];

  if (reports) {
    HPP_CUDA(c, cudaMemsetAsync(d_stats, 0, stats_bytes, c->stream));
    dim3 sgrid(c->sm_count * 4, (unsigned)n_fields);
    heur_stats_kernel<<<sgrid, 256, 0, c->stream>>>(cost_dev, m->n, variant == HPP_VARIANT_RS ? INFINITY : 1000.0, d_stats);
    HPP_LAUNCHED(c);
    unsigned long long *h_stats = (unsigned long long *)((uint8_t *)h_goals + goals_bytes);
    HPP_CUDA(c, cudaMemcpyAsync(h_stats, d_stats, stats_bytes, cudaMemcpyDeviceToHost, c->stream));
    HPP_CUDA(c, cudaStreamSynchronize(c->stream));
    for (size_t f = 0; f < n_fields; f++) {
      hpp_quirk_report *r = &reports[f];
      memset(r, 0, sizeof *r);
      r->reachable = h_stats[2 * f];
      memcpy(&r->max_cost, &h_stats[2 * f + 1], 8);
      r->rounds = round;
      r->tile_visits = visits;
      if (variant == HPP_VARIANT_RS) {
        const int MAX_COST = (W + H > 5000) ? W + H : 5000;  // RS/src/HybridAstar.cpp:395
        r->h2_bucket_clamp = r->max_cost >= (double)(MAX_COST - 1);
      }
    }
  }
  return HPP_OK;
}

// H1 (RS/src/HybridAstar.cpp:452-472): every 70000 settled cells the reference stops if the start cell is
// out of bounds, already reached, or more than 75 % of the map is settled.
static void classify_h1(const hpp_map *m, hpp_quirk_report *r, int sx, int sy, double start_cost) {
  const uint64_t settled = r->reachable;  // the goal is counted even when occupied
  const uint64_t interval = 70000;
  if (settled < interval) {
    r->h1_progress_break = 0;
    return;
  }
  const bool start_inb = sx >= 0 && sx < m->g.W && sy >= 0 && sy < m->g.H;
  if (!start_inb) {
    r->h1_progress_break = 1;
    return;
  }
  if (isfinite(start_cost)) {
    r->h1_progress_break = 2;  // depends on when the start is reached relative to a multiple of 70000
    return;
  }
  // start never reached: fires iff some multiple of 70000 that is <= settled exceeds 75 % of the map
  const double thr = (double)m->n * 0.75;
  const uint64_t k = settled / interval;
  r->h1_progress_break = ((double)(k * interval) > thr) ? 1 : 0;
}

static hpp_status heuristic_host(hpp_map *m, int variant, int gx, int gy, int sx, int sy, double *cost,
                                 hpp_quirk_report *report) {
  if (!m || !cost) return HPP_ERR_ARG;
  hpp_ctx *c = m->ctx;
  HPP_CUDA(c, cudaSetDevice(c->device));
  void *d_cost;
  hpp_status s = hpp_scratch(c, 3, m->n * sizeof(double), &d_cost);
  if (s != HPP_OK) return s;
  int32_t goal[2] = {gx, gy};
  hpp_quirk_report rep;
  s = hpp_heuristic_batch_dev(m, variant, goal, 1, (double *)d_cost, &rep);
  if (s != HPP_OK) return s;
  HPP_CUDA(c, cudaMemcpyAsync(cost, d_cost, m->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  if (report) {
    if (variant == HPP_VARIANT_RS) {
      double sc = INFINITY;
      if (sx >= 0 && sx < m->g.W && sy >= 0 && sy < m->g.H) sc = cost[(size_t)sy * m->g.W + sx];
      classify_h1(m, &rep, sx, sy, sc);
    }
    *report = rep;
  }
  return HPP_OK;
}

extern "C" hpp_status hpp_heuristic_rs(hpp_map *m, int32_t gx, int32_t gy, int32_t sx, int32_t sy, double *cost,
                                       hpp_quirk_report *report) {
  return heuristic_host(m, HPP_VARIANT_RS, gx, gy, sx, sy, cost, report);
}
extern "C" hpp_status hpp_heuristic_dubins(hpp_map *m, int32_t gx, int32_t gy, double *cost, hpp_quirk_report *report) {
  return heuristic_host(m, HPP_VARIANT_DUBINS, gx, gy, -1, -1, cost, report);
}
