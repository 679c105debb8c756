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
// over free cells v, which does not depend on the relaxation order (fl(a + w) is monotone in a).
//
// Device algorithm: ONE persistent kernel per batch of fields (heur_warp_kernel) -- delta-stepping over 32x32 tiles,
// ONE WARP per tile (no block-level barrier inside a tile solve), 16 / 20 / 4 warps per CTA.  A tile with its halo is
// staged by one TMA tensor box copy (3-D tensor map over the batch's fields; row-wise bulk copies or plain loads where
// the tensor map cannot be used).  The frontier lives in REGISTERS, one 32-bit "needs evaluation" mask per lane
// (= row).  Sweeps are COMPACTED (sweep_compact): the rows' needed cells go into one shared-memory list, every lane
// takes two entries whatever their rows, improved cells OR their push masks into their rows' words with 32-bit shared
// atomics.  (Round 1 also carried a multi-warp-tile kernel, a row-wise sweep and an asynchronous tile-queue kernel;
// they were slower on every measured case -- profiles/r1_summary.md -- and were removed in round 2.)
//   * A cell is evaluated by PULL: min over its 8 neighbours, exactly the reference's additions (no FMA; min
//     before add is legal because fl(a+w) is monotone).  An improved cell pushes only the neighbours it can still
//     improve, so settled cells behind the front are not re-evaluated.
//   * Cost window: a change whose new cost is >= bound is DEFERRED (bit kept in a per-tile mask in global memory)
//     instead of propagated; a tile is scheduled when a neighbour posted a border change or when its cheapest
//     deferred cell falls below the bound; when nothing is schedulable the bound advances to (global minimum of
//     deferred costs) + delta.  This keeps expensive waves (the +5 near-obstacle cells) from running ahead of the
//     cheap ones and being corrected later: 3 evaluations per cell instead of 27 on the cfg3 map.  Inside a tile visit
//     the same mechanism runs with sub-windows of delta_in.
//   * Scheduling is done on the device.  A batch is split over independent CTA groups, one field per group at a time:
//     a group of several CTAs (cooperative launch) scans the tile table in slices and meets at two group barriers per
//     round; ONE CTA per field (plain launch, large batches) keeps the round's schedule as the list of tiles posted
//     to during the previous round and scans the table only when the window moves.  The host launches once and
//     never polls.
// Costs are fp64 with exactly the reference's additions (library built with -fmad=false).
#include <cooperative_groups.h>
#include <cuda.h>  // CUtensorMap (the encoder itself is fetched through cudaGetDriverEntryPoint: no libcuda link)
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "hpp_internal.cuh"

// per-warp iteration / evaluation counters for the HPP_HEUR_DEBUG report (cost ~3 % of the instructions)
#ifdef HPP_HEUR_COUNTERS
#define HPP_COUNT(x) x
#else
#define HPP_COUNT(x)
#endif


namespace {

constexpr int HB = 32;  // block side = one warp, lane = row
constexpr unsigned FULL = 0xffffffffu;
constexpr unsigned long long INF_BITS = 0x7ff0000000000000ull;

// direction bits of the tile activity flags / out flags
enum { DIR_N = 1, DIR_S = 2, DIR_W = 4, DIR_E = 8, DIR_NW = 16, DIR_NE = 32, DIR_SW = 64, DIR_SE = 128 };

struct heur_weights {
  double rs[2][2];   // [near][0 = cardinal, 1 = diagonal]
  double db[2][10];  // [diagonal][3x3 collision count of the SOURCE cell]
  double init;       // +inf (RS) or 1000 (Dubins)
};

struct heur_args {
  double *cost_all;                // [n_fields][H*W]
  const uint32_t *occ_bits;        // bit-packed data > 0, pitch words per row
  const uint32_t *near_bits;       // bit-packed CLS_NEAR4
  const uint8_t *cls;              // class bytes (Dubins: 3x3 counts)
  int pitch;
  int W, H, tiles_x, tiles_y, n_tiles;
  int n_fields;
  uint32_t *act;                   // [n_fields][3][n_tiles]
  unsigned long long *pmin;        // [n_fields][2][n_tiles]  cheapest deferred cost of the tile (bit pattern)
  uint32_t *dmask;                 // [n_fields][n_tiles][NB*NB*32]  deferred cells
  uint32_t *bar;                   // [n_groups][32]: count, generation
  uint32_t *stats;                 // [n_fields][4]: rounds, tile visits, windows, sweeps of CTA rank 0
  uint32_t *lists;                 // warp-tile kernel: [grid][n_tiles] schedule of the round
  uint32_t *sched;                 // warp-tile kernel: [n_groups][4 * group_size] slice counts + minima of a round
  unsigned long long *prof;        // optional [grid][8] cycle counters (HPP_HEUR_DEBUG): schedule, stage, solve,
                                   // write-back, barrier wait, tiles, phases, longest single solve
  double delta;                    // width of the global cost window
  double delta_in;                 // width of the sub-windows inside a tile visit
  int group_size, n_groups;
  int lpt;                         // one-CTA-per-field mode: longest-first ordering of a round's tiles
  int tmap_ok;                     // warp-tile kernel: the tensor map below is valid (tiles staged by one TMA box copy)
  alignas(64) CUtensorMap tmap;    // cost_all as a (W, H, n_fields) fp64 tensor, box 38 x 34 x 1 = a tile with its halo
  heur_weights wt;
};

template <int NB>
struct tile_cfg {
  static constexpr int T = NB * HB;
  // row stride in doubles: even (16-byte aligned rows) and == 2 (mod 4), which keeps a warp whose lanes sit on 32
  // different rows at (nearly) the same column at <= 2-way bank conflicts, for vertical and both diagonal fronts
  static constexpr int S = T + 6;
  static constexpr int ROWS = T + 2;
  static constexpr int NBLK = NB * NB;
  static constexpr int THREADS = NBLK * 32;
};

__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }

// ---------------------------------------------------------------- group barrier (co-resident CTAs, cooperative launch)
__device__ __forceinline__ void group_barrier(uint32_t *bar, int group_size) {
  __syncthreads();
  if (group_size > 1 && threadIdx.x == 0) {
    volatile uint32_t *gen = bar + 1;
    __threadfence();
    const uint32_t g = *gen;
    if (atomicAdd(bar, 1u) == (uint32_t)group_size - 1u) {
      bar[0] = 0u;
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      while (*gen == g) __nanosleep(32);
    }
    __threadfence();
  }
  __syncthreads();
}

// ---------------------------------------------------------------- per-warp block state
struct blk_state {
  uint32_t need;   // cells of my row that must be (re-)evaluated
  uint32_t defer;  // cells of my row whose latest change is not propagated yet (cost >= bound)
  uint32_t freem;  // free, in-bounds cells of my row
  uint32_t nearm;  // RS: near-obstacle cells of my row
  uint32_t dirty;  // my row changed (needs write-back)
  uint32_t posted; // I posted a border change to another warp of this CTA in this phase
  // counters of a visit (HPP_HEUR_COUNTERS builds): dense-row / sparse iterations, needed cells summed over the
  // sweeps, ceil(needed / 64) summed over the sweeps (= iterations of a perfectly balanced evaluation), sweeps
  uint32_t c_heavy, c_sparse, c_on, c_imp, c_sweeps;
};

// weights held in registers (RS) / shared memory (Dubins table)
struct eval_w {
  double c0, d0;          // RS: cheapest cardinal / diagonal edge (the push filter's lower bound)
  double init;
  const double2 *rs;      // RS: [near] -> (cardinal, diagonal) weights, in shared memory (one LDS.128)
  const double *db;       // Dubins: [2][10] in shared memory
};

// Evaluate one cell by PULL (e = shared index of the cell): returns the candidate cost, and in nmask the neighbours
// that a cell of cost `best` could still improve (a superset: the cheapest possible edge weight is used for RS, the
// exact source-dependent weight for Dubins), so that settled cells behind the wave front are not re-evaluated.
template <int VARIANT, int S>
__device__ __forceinline__ double eval_cell(const volatile double *cs, const volatile uint8_t *cl, int e, bool near,
                                            const eval_w &wt, uint32_t &nmask) {
  const double nw = cs[e - S - 1], n = cs[e - S], ne = cs[e - S + 1], w = cs[e - 1], ea = cs[e + 1];
  const double sw = cs[e + S - 1], s = cs[e + S], se = cs[e + S + 1];
  double best, t1, t2;
  if (VARIANT == HPP_VARIANT_RS) {
    const double mc = dmin(dmin(n, s), dmin(w, ea));
    const double md = dmin(dmin(nw, ne), dmin(sw, se));
    const double2 wcd = wt.rs[near ? 1 : 0];
    best = dmin(mc + wcd.x, md + wcd.y);
    t1 = best + wt.c0;
    t2 = best + wt.d0;
  } else {
    const double *db = wt.db;
    best = wt.init;
    best = dmin(best, nw + db[10 + (cl[e - S - 1] & CLS_CNT_MASK)]);
    best = dmin(best, n + db[cl[e - S] & CLS_CNT_MASK]);
    best = dmin(best, ne + db[10 + (cl[e - S + 1] & CLS_CNT_MASK)]);
    best = dmin(best, w + db[cl[e - 1] & CLS_CNT_MASK]);
    best = dmin(best, ea + db[cl[e + 1] & CLS_CNT_MASK]);
    best = dmin(best, sw + db[10 + (cl[e + S - 1] & CLS_CNT_MASK)]);
    best = dmin(best, s + db[cl[e + S] & CLS_CNT_MASK]);
    best = dmin(best, se + db[10 + (cl[e + S + 1] & CLS_CNT_MASK)]);
    const int k = cl[e] & CLS_CNT_MASK;  // as a SOURCE this cell charges its own 3x3 count
    t1 = best + db[k];
    t2 = best + db[10 + k];
  }
  nmask = (t2 < nw ? 1u : 0u) | (t1 < n ? 2u : 0u) | (t2 < ne ? 4u : 0u) | (t1 < w ? 8u : 0u) | (t1 < ea ? 16u : 0u) |
          (t2 < sw ? 32u : 0u) | (t1 < s ? 64u : 0u) | (t2 < se ? 128u : 0u);
  return best;
}

// ---------------------------------------------------------------- TMA (bulk async copy) + mbarrier helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// global -> shared bulk copy (TMA engine; bytes and both addresses multiples of 16), completion on the mbarrier
__device__ __forceinline__ void tma_load_row(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// global -> shared tensor box copy (TMA, 3-D tiled tensor map): one instruction stages a tile with its halo; elements
// outside the tensor arrive as zeros.  dst must be 128-byte aligned.
__device__ __forceinline__ void tma_load_box3(uint32_t dst, const CUtensorMap *tmap, int x, int y, int z, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"((unsigned long long)tmap), "r"(x), "r"(y), "r"(z), "r"(bar)
               : "memory");
}
// order this thread's generic-proxy accesses with async-proxy (TMA) accesses, per state space (a tile visit needs one of
// each; the unrestricted fence.proxy.async twice per visit was 4.6 % of the stall samples)
__device__ __forceinline__ void fence_proxy_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// ---------------------------------------------------------------- one tile visit by one warp
// Stages tile `tile` (+ halo) of the field `cost` into the warp's shared-memory slot, relaxes it to its fixed point under
// `bound` (sub-windows of delta_in), writes the changed rows and the deferred masks back.  af = ring sides to
// re-evaluate (DIR_* flags posted by the neighbours).  Returns the cheapest cost left deferred (bit pattern, INF_BITS
// if none); *wc.s_out holds the DIR_* flags of the neighbour tiles whose halo this visit changed.
// warp-tile kernel: n_tiles-sized arrays per CTA in a.lists (schedule of the current / next round, last visit times)
#define LISTS_PER_CTA 3

// a warp's tile slot in shared memory: ROWS*S doubles, padded to a multiple of 128 bytes (TMA tensor copies need a
// 128-byte aligned destination)
#define WSLOT_DOUBLES ((((HB + 2) * (HB + 6) * 8 + 127) / 128) * 128 / 8)

// per-warp scratch of the compacted sweep, in shared memory (all zero between sweeps)
#define WAUX_LIST 128
struct warp_aux {
  uint4 row[HB];             // per row: .x = cells a change of this sweep may improve, .y = cells improved below the
                             // bound (their deferred bit is cleared), .z = cells improved at or above it (deferred)
  uint16_t list[WAUX_LIST];  // the sweep's needed cells, row << 5 | column
};

struct warp_ctx {
  double *cs_nv;
  uint8_t *cl_nv;
  uint32_t *s_out;
  warp_aux *aux;
  uint32_t mbar, parity;
  bool tma_ok;   // rows can be bulk-copied (W even, aligned base)
  bool tma_2d;   // ... and the whole tile + halo comes as one tensor box copy
  int field;     // field index (third tensor coordinate)
};

// the three neighbours (columns x-1, x, x+1; m3 bit 0 = column x-1) of one row as a mask of the columns 0..31
__device__ __forceinline__ uint32_t row_mask3(int x, uint32_t m3) { return x >= 1 ? (m3 << (x - 1)) : (m3 >> 1); }

// an improved cell (row r, column x) below the bound: its neighbours that may improve (nmask k: bit 0 NW, 1 N, 2 NE,
// 3 W, 4 E, 5 SW, 6 S, 7 SE) go to their rows' words; what lies outside the tile becomes a flag for that neighbour tile
__device__ __forceinline__ void aux_deliver(warp_aux *ax, uint32_t *s_out, int r, int x, uint32_t k) {
  const uint32_t up = k & 7u, own = ((k >> 3) & 1u) | ((k >> 2) & 4u), dn = (k >> 5) & 7u;
  atomicOr(&ax->row[r].x, row_mask3(x, own));
  atomicOr(&ax->row[r].y, 1u << x);
  if (r > 0) atomicOr(&ax->row[r - 1].x, row_mask3(x, up));
  if (r < HB - 1) atomicOr(&ax->row[r + 1].x, row_mask3(x, dn));
  if (r == 0 || r == HB - 1 || x == 0 || x == 31) {
    uint32_t d = 0;
    if (r == 0) d |= ((up & 2u) ? DIR_N : 0u) | ((up & 1u) ? (x == 0 ? DIR_NW : DIR_N) : 0u) | ((up & 4u) ? (x == 31 ? DIR_NE : DIR_N) : 0u);
    else if (x == 0 && (up & 1u)) d |= DIR_W;
    else if (x == 31 && (up & 4u)) d |= DIR_E;
    if (r == HB - 1) d |= ((dn & 2u) ? DIR_S : 0u) | ((dn & 1u) ? (x == 0 ? DIR_SW : DIR_S) : 0u) | ((dn & 4u) ? (x == 31 ? DIR_SE : DIR_S) : 0u);
    else if (x == 0 && (dn & 1u)) d |= DIR_W;
    else if (x == 31 && (dn & 4u)) d |= DIR_E;
    if (x == 0 && (own & 1u)) d |= DIR_W;
    if (x == 31 && (own & 4u)) d |= DIR_E;
    if (d) atomicOr(s_out, d);
  }
}

// One sweep of a warp's 32x32 tile with the needed cells COMPACTED over the lanes.  The row-wise sweep above keeps
// lane = row, so a thin wave front (one or two needed cells in each of a few rows) occupies a few lanes for several
// iterations; here the rows write their needed cells into one list, every lane takes two list entries whatever
// their rows, and the results go back to the rows through shared-memory ORs: (nearly) all lanes evaluate, and a sweep
// with up to 64 needed cells is a single evaluation step.  The fixed point does not depend on the evaluation order,
// so the cost field is bit-identical.
template <int VARIANT>
__device__ __forceinline__ bool sweep_compact(volatile double *cs, const volatile uint8_t *cl, warp_aux *ax, uint32_t *s_out,
                                              blk_state &st, int lane, double bound, const eval_w &wt) {
  constexpr int S = tile_cfg<1>::S;
  if (!__any_sync(FULL, st.need != 0)) return false;
  const int ebase = S + 2;
  // the rows' needed cells, back to back; rows that do not fit in the list keep theirs for the next sweep
  const int cnt = __popc(st.need);
  int inc = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(FULL, inc, d);
    if (lane >= d) inc += o;
  }
  const bool fits = inc <= WAUX_LIST;
  const unsigned fit_m = __ballot_sync(FULL, fits);  // a prefix of the rows
  const int n_rows = __popc(fit_m);
  int total = __shfl_sync(FULL, inc, n_rows > 0 ? n_rows - 1 : 0);
  if (n_rows == 0) total = 0;
  HPP_COUNT(st.c_sweeps++; st.c_imp += (total + 63) >> 6; st.c_on += total;)
  // rows with many needed cells (a front along the row) are written by the whole warp, one lane per column; the
  // others by their own lane, one cell per step
  uint32_t heavy = __ballot_sync(FULL, fits && cnt > 8);
  while (heavy) {
    const int r = __ffs(heavy) - 1;
    heavy &= heavy - 1;
    const uint32_t m = __shfl_sync(FULL, st.need, r);
    const int base = __shfl_sync(FULL, inc - cnt, r);
    if ((m >> lane) & 1u) ax->list[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)((r << 5) | lane);
  }
  if (fits) {
    if (cnt <= 8) {
      int pos = inc - cnt;
      for (uint32_t m = st.need; m; m &= m - 1) ax->list[pos++] = (uint16_t)((lane << 5) | (__ffs(m) - 1));
    }
    st.need = 0;
  }
  __syncwarp();
  for (int q0 = 0; q0 < total; q0 += 64) {
    HPP_COUNT(st.c_sparse++;)
    const int qa = q0 + lane, qb = q0 + 32 + lane;
    const bool on0 = qa < total, on1 = qb < total;
    const int c0i = on0 ? ax->list[qa] : 0, c1i = on1 ? ax->list[qb] : 0;
    const int r0 = c0i >> 5, x0 = c0i & 31, r1 = c1i >> 5, x1 = c1i & 31;
    const uint32_t nm0 = __shfl_sync(FULL, st.nearm, r0), nm1 = __shfl_sync(FULL, st.nearm, r1);
    const int e0 = ebase + r0 * S + x0, e1 = ebase + r1 * S + x1;
    const double c0 = cs[e0], c1 = cs[e1];
    uint32_t k0, k1;
    const double b0 = eval_cell<VARIANT, S>(cs, cl, e0, (nm0 >> x0) & 1u, wt, k0);
    double b1 = c1;
    k1 = 0;
    if (q0 + 32 < total) b1 = eval_cell<VARIANT, S>(cs, cl, e1, (nm1 >> x1) & 1u, wt, k1);  // warp-uniform
    if (on0 && b0 < c0) {
      cs[e0] = b0;
      if (b0 < bound) {
        aux_deliver(ax, s_out, r0, x0, k0);
      } else {
        atomicOr(&ax->row[r0].z, 1u << x0);
      }
    }
    if (on1 && b1 < c1) {
      cs[e1] = b1;
      if (b1 < bound) {
        aux_deliver(ax, s_out, r1, x1, k1);
      } else {
        atomicOr(&ax->row[r1].z, 1u << x1);
      }
    }
    __syncwarp();
  }
  // back to the rows
  const uint4 w = ax->row[lane];
  __syncwarp();
  if (w.x | w.y | w.z) {
    ax->row[lane] = make_uint4(0u, 0u, 0u, 0u);
    if (w.y | w.z) {
      st.dirty = 1;
      st.defer = (st.defer | w.z) & ~w.y;
    }
    st.need |= w.x & st.freem;
  }
  __syncwarp();
  return true;
}

template <int VARIANT>
__device__ __forceinline__ unsigned long long warp_visit_tile(const heur_args &a, double *__restrict__ cost,
                                                              uint32_t *__restrict__ dmask, int tile, uint32_t af,
                                                              double bound, const eval_w &wt, warp_ctx &wc, int lane) {
  using C = tile_cfg<1>;
  constexpr int T = C::T, S = C::S, ROWS = C::ROWS;
  const int W = a.W, H = a.H;
  double *cs_nv = wc.cs_nv;
  uint8_t *cl_nv = wc.cl_nv;
  volatile double *cs = cs_nv;
  const volatile uint8_t *cl = cl_nv;
  const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
  const int x0 = tx * T, y0 = ty * T;
    // ---- stage the tile + halo.  TMA path: one bulk copy per row (columns x0-2 .. x0+33, 16-byte aligned),
    // all rows in flight at once, completion counted on the warp's mbarrier.  Cells outside the map keep the
    // initial value: edge tiles are pre-filled before the copies land.
    if (wc.tma_2d) {
      // one box copy: columns x0-2 .. x0+35, rows y0-1 .. y0+32 of field wc.field; what lies outside the map comes as
      // zeros and is overwritten with the initial value afterwards (edge tiles only)
      fence_proxy_async_shared();  // my generic-proxy accesses to this buffer before the async-proxy writes
      if (lane == 0) {
        mbar_expect_tx(wc.mbar, (uint32_t)(ROWS * S * sizeof(double)));
        tma_load_box3(smem_u32(cs_nv), &a.tmap, x0 - 2, y0 - 1, wc.field, wc.mbar);
      }
      __syncwarp();
      if (VARIANT == HPP_VARIANT_DUBINS) {
        for (int r = 0; r < ROWS; r++) {
          const int gy = y0 - 1 + r;
#pragma unroll
          for (int k = 0; k < 2; k++) {
            const int c = lane + 32 * k, gx = x0 - 1 + c;
            if (c < T + 2) {
              uint8_t k8 = CLS_BLOCKED | 9;
              if (gy >= 0 && gy < H && gx >= 0 && gx < W) k8 = a.cls[(size_t)gy * W + gx];
              cl_nv[r * S + c + 1] = k8;
            }
          }
        }
      }
      mbar_wait(wc.mbar, wc.parity);
      wc.parity ^= 1u;
      if (x0 == 0 || x0 + T + 1 > W || y0 == 0 || y0 + T + 1 > H) {
        for (int i = lane; i < ROWS * S; i += 32) {
          const int gy = y0 - 1 + i / S, gx = x0 - 2 + i % S;
          if (gy < 0 || gy >= H || gx < 0 || gx >= W) cs_nv[i] = wt.init;
        }
        __syncwarp();
      }
    } else if (wc.tma_ok) {
      const int cx0 = x0 - 2 < 0 ? 0 : x0 - 2, cx1 = x0 + T + 2 > W ? W : x0 + T + 2;
      const int ry0 = y0 - 1 < 0 ? 0 : y0 - 1, ry1 = y0 + T + 1 > H ? H : y0 + T + 1;  // valid global rows
      const bool edge = x0 == 0 || x0 + T + 2 > W || y0 == 0 || y0 + T + 1 > H;
      if (edge) {
        for (int i = lane; i < ROWS * S; i += 32) cs_nv[i] = wt.init;
        __syncwarp();
      }
      fence_proxy_async_shared();  // my generic-proxy accesses to this buffer before the async-proxy writes
      const uint32_t row_bytes = (uint32_t)(cx1 - cx0) * 8u;
      if (lane == 0) mbar_expect_tx(wc.mbar, row_bytes * (uint32_t)(ry1 - ry0));
      __syncwarp();
      for (int gy = ry0 + lane; gy < ry1; gy += 32) {
        const int r = gy - (y0 - 1);
        tma_load_row(smem_u32(cs_nv + r * S + (cx0 - (x0 - 2))), cost + (size_t)gy * W + cx0, row_bytes, wc.mbar);
      }
      if (VARIANT == HPP_VARIANT_DUBINS) {
        for (int r = 0; r < ROWS; r++) {
          const int gy = y0 - 1 + r;
#pragma unroll
          for (int k = 0; k < 2; k++) {
            const int c = lane + 32 * k, gx = x0 - 1 + c;
            if (c < T + 2) {
              uint8_t k8 = CLS_BLOCKED | 9;
              if (gy >= 0 && gy < H && gx >= 0 && gx < W) k8 = a.cls[(size_t)gy * W + gx];
              cl_nv[r * S + c + 1] = k8;
            }
          }
        }
      }
      mbar_wait(wc.mbar, wc.parity);
      wc.parity ^= 1u;
    } else {
      for (int r = 0; r < ROWS; r++) {
        const int gy = y0 - 1 + r;
        const bool rin = gy >= 0 && gy < H;
#pragma unroll
        for (int k = 0; k < 2; k++) {
          const int c = lane + 32 * k;
          if (c < T + 2) {
            const int gx = x0 - 1 + c;
            double v = wt.init;
            uint8_t k8 = CLS_BLOCKED | 9;
            if (rin && gx >= 0 && gx < W) {
              v = __ldcg(cost + (size_t)gy * W + gx);
              if (VARIANT == HPP_VARIANT_DUBINS) k8 = a.cls[(size_t)gy * W + gx];
            }
            cs_nv[r * S + c + 1] = v;
            if (VARIANT == HPP_VARIANT_DUBINS) cl_nv[r * S + c + 1] = k8;
          }
        }
      }
    }
    if (lane == 0) (*wc.s_out) = 0;
    blk_state st;
    {
      const int gy = y0 + lane, wx = x0 >> 5;
      uint32_t ob = 0xffffffffu, nb = 0;
      if (gy < H && wx < a.pitch) {
        ob = a.occ_bits[(size_t)gy * a.pitch + wx];
        nb = a.near_bits[(size_t)gy * a.pitch + wx];
        const int rem = W - (wx << 5);
        if (rem < 32) ob |= ~((1u << rem) - 1u);
      }
      st.freem = ~ob;
      st.nearm = nb;
      st.need = 0;
      st.dirty = 0;
      st.posted = 0;
      st.c_heavy = st.c_sparse = st.c_on = st.c_imp = st.c_sweeps = 0;
      st.defer = __ldcg(dmask + (size_t)tile * 32 + lane);
    }
    __syncwarp();
    if (af) {
      if ((af & DIR_N) && lane == 0) st.need |= st.freem;
      if ((af & DIR_S) && lane == 31) st.need |= st.freem;
      if (af & DIR_W) st.need |= st.freem & 1u;
      if (af & DIR_E) st.need |= st.freem & 0x80000000u;
      if ((af & DIR_NW) && lane == 0) st.need |= st.freem & 1u;
      if ((af & DIR_NE) && lane == 0) st.need |= st.freem & 0x80000000u;
      if ((af & DIR_SW) && lane == 31) st.need |= st.freem & 1u;
      if ((af & DIR_SE) && lane == 31) st.need |= st.freem & 0x80000000u;
    }
    // relax under the bound in sub-windows; everything is warp-local
    const int ebase = S + 2;
    auto run_sweeps = [&](double lb) {
      while (sweep_compact<VARIANT>(cs, cl, wc.aux, wc.s_out, st, lane, lb, wt)) {
      }
    };
    // One pass over my row's deferred cells: those below lb are released (returned mask), the cheapest of the others
    // is returned in `rest` (bit pattern; costs are non-negative, so the integer order is the numeric order).
    auto scan_deferred = [&](double lb, unsigned long long &rest) {
      uint32_t rel = 0;
      rest = INF_BITS;
      for (uint32_t m = st.defer; m; m &= m - 1) {
        const int x = __ffs(m) - 1;
        const double v = cs[ebase + lane * S + x];
        if (v < lb) rel |= 1u << x;
        else {
          const unsigned long long vb = (unsigned long long)__double_as_longlong(v);
          rest = vb < rest ? vb : rest;
        }
      }
      return rel;
    };
    auto warp_min = [&](unsigned long long mn) {
      for (int d = 16; d; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(FULL, mn, d);
        mn = o < mn ? o : mn;
      }
      return mn;
    };
    // released cells may improve all eight neighbours (the exact set is not known any more): need bits of my row and
    // of the rows above / below, flags for the neighbouring tiles
    auto release = [&](uint32_t rel) {
      st.defer &= ~rel;
      const uint32_t h = rel | (rel << 1) | (rel >> 1);
      uint32_t from_below = __shfl_down_sync(FULL, h, 1), from_above = __shfl_up_sync(FULL, h, 1);
      if (lane == 31) from_below = 0;
      if (lane == 0) from_above = 0;
      st.need |= ((rel << 1) | (rel >> 1) | from_below | from_above) & st.freem;
      uint32_t d = 0;
      if (rel & 1u) d |= DIR_W;
      if (rel >> 31) d |= DIR_E;
      if (lane == 0 && rel) d |= DIR_N | ((rel & 1u) ? DIR_NW : 0u) | ((rel >> 31) ? DIR_NE : 0u);
      if (lane == 31 && rel) d |= DIR_S | ((rel & 1u) ? DIR_SW : 0u) | ((rel >> 31) ? DIR_SE : 0u);
      d = __reduce_or_sync(FULL, d);
      if (d && lane == 0) atomicOr(wc.s_out, d);
    };
    // The first sub-window starts at the cheapest deferred cost; after that the lower bound moves in steps of delta_in
    // (one pass over the deferred cells per step), jumping over empty ranges.
    unsigned long long left = INF_BITS;
    run_sweeps(0.0);
    double lb = 0.0;
    for (;;) {
      unsigned long long rest;
      uint32_t rel = scan_deferred(lb, rest);
      if (!__any_sync(FULL, rel != 0)) {
        // nothing below lb: every deferred cost is in `rest` (no sweep ran since), so its minimum is exact
        const unsigned long long mn = warp_min(rest);
        left = mn;
        const double gm = __longlong_as_double((long long)mn);
        if (mn == INF_BITS || !(gm < bound)) break;
        lb = gm + a.delta_in;
        if (lb > bound) lb = bound;
        rel = scan_deferred(lb, rest);
      }
      release(rel);
      __syncwarp();
      run_sweeps(lb);
      lb += a.delta_in;
      if (lb > bound) lb = bound;
    }
    // write back
    __stcg(dmask + (size_t)tile * 32 + lane, st.defer);
    uint32_t dr = __ballot_sync(FULL, st.dirty != 0);
    const int gx = x0 + lane;
    while (dr) {
      const int r = __ffs(dr) - 1;
      dr &= dr - 1;
      const int gy = y0 + r;
      if (gx < W && gy < H) __stcg(cost + (size_t)gy * W + gx, cs[ebase + r * S + lane]);
    }
    fence_proxy_async_global();  // the rows just stored are read by other warps' bulk copies (async proxy) later
    __syncwarp();
#ifdef HPP_HEUR_COUNTERS
    if (a.prof && lane == 0) {  // warp-uniform counters of this visit (debug builds only)
      unsigned long long *pw = a.prof + (size_t)gridDim.x * 8;
      atomicAdd(pw + 0, (unsigned long long)st.c_heavy);
      atomicAdd(pw + 1, (unsigned long long)st.c_sparse);
      atomicAdd(pw + 2, (unsigned long long)st.c_on);   // needed cells summed over the sweeps
      atomicAdd(pw + 3, (unsigned long long)st.c_imp);  // ceil(needed / 64) summed over the sweeps: ideal iterations
      atomicAdd(pw + 4, (unsigned long long)st.c_sweeps);
    }
#endif
    return left;
}

// ---------------------------------------------------------------- the warp-tile kernel (T = 32: one warp per tile)
// Scheduling protocol: cost windows, activity flags, deferred minima, one group barrier per round; a tile is a single
// 32x32 block solved by ONE warp with no block-level barrier at all: the 16 warps
// of a CTA pull tiles of the round's schedule from a shared counter and never wait for each other inside a round.
// This is the throughput configuration: with several fields in flight every warp of the GPU stays busy.
// Warps (= tiles in flight) per CTA: 16 (123 registers) is the latency configuration; 20 (96 registers, 207 KB of
// shared memory) gives +6 % throughput when every field has its own CTA -- RS only, the Dubins tiles also hold the
// class bytes and do not fit 20 times.

// (the 4-warp CTAs are meant to sit five to an SM: without the second launch bound the compiler took 200 registers and the
// driver a small shared-memory carve-out, and ncu showed TWO resident CTAs = 8 warps per SM, issue slots 27 % busy)
template <int VARIANT, int WT_WARPS>
__global__ void __launch_bounds__(WT_WARPS * 32, WT_WARPS == 4 ? 5 : 1) heur_warp_kernel(const __grid_constant__ heur_args a) {
  using C = tile_cfg<1>;
  constexpr int S = C::S, ROWS = C::ROWS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *cs_all = (double *)smem_raw;                                  // WT_WARPS x WSLOT_DOUBLES (128-byte aligned slots)
  uint8_t *cl_all = (uint8_t *)(cs_all + WT_WARPS * WSLOT_DOUBLES);     // WT_WARPS x ROWS*S (Dubins only)
  warp_aux *aux_all = (warp_aux *)(cl_all + (VARIANT == HPP_VARIANT_DUBINS ? WT_WARPS * ROWS * S : 0));
  __shared__ unsigned long long s_gmin;
  __shared__ uint32_t s_mine, s_next, s_lnext, s_dsum, s_dcnt, s_warp_cnt[32], s_outw[WT_WARPS], s_pref[WT_WARPS * 32];
  __shared__ double s_db[20];
  __shared__ __align__(8) unsigned long long s_mbar[WT_WARPS];  // one TMA completion barrier per warp

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int group = blockIdx.x / a.group_size, grank = blockIdx.x % a.group_size;
  if (group >= a.n_groups) return;
  uint32_t *bar = a.bar + group * 32;
  uint32_t *my_list = a.lists + (size_t)blockIdx.x * LISTS_PER_CTA * a.n_tiles;  // two lists (current / next round) per CTA and,
                                                                     // one-CTA-per-field mode, the tiles' last visit times
  const int W = a.W, H = a.H, n_tiles = a.n_tiles;
  const size_t n = (size_t)W * H;
  if (tid < 20) s_db[tid] = a.wt.db[tid / 10][tid % 10];
  __shared__ __align__(16) double2 s_rs[2];
  if (tid < 2) s_rs[tid] = make_double2(a.wt.rs[tid][0], a.wt.rs[tid][1]);
  eval_w wt;
  wt.c0 = a.wt.rs[0][0];
  wt.d0 = a.wt.rs[0][1];
  wt.init = a.wt.init;
  wt.rs = s_rs;
  wt.db = s_db;
  warp_ctx wc;
  wc.cs_nv = cs_all + wid * WSLOT_DOUBLES;
  wc.cl_nv = cl_all + wid * ROWS * S;
  wc.aux = aux_all + wid;
  for (int i = lane; i < (int)(sizeof(warp_aux) / 4); i += 32) ((uint32_t *)wc.aux)[i] = 0u;
  wc.s_out = &s_outw[wid];
  wc.mbar = smem_u32(&s_mbar[wid]);
  wc.parity = 0;
  if (lane == 0) mbar_init(wc.mbar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  // rows can be bulk-copied when every row start (x0 - 2 doubles) is 16-byte aligned: W even and an aligned base
  wc.tma_ok = (W % 2 == 0) && ((reinterpret_cast<uintptr_t>(a.cost_all) & 15) == 0);
  wc.tma_2d = wc.tma_ok && a.tmap_ok && ((smem_u32(wc.cs_nv) & 127u) == 0u);
  wc.field = 0;

  for (int field = group; field < a.n_fields; field += a.n_groups) {
    wc.field = field;
    double *cost = a.cost_all + (size_t)field * n;
    uint32_t *act = a.act + (size_t)field * 3 * n_tiles;
    unsigned long long *pmin = a.pmin + (size_t)field * 2 * n_tiles;
    uint32_t *dmask = a.dmask + (size_t)field * n_tiles * 32;
    double bound = 0.0;
    uint32_t rounds = 0, visits = 0, windows = 0;
    unsigned long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tk = clock64();
#define PF_MARK(i)                            \
  if (a.prof) {                               \
    const long long now_ = clock64();         \
    pf[i] += (unsigned long long)(now_ - tk); \
    tk = now_;                                \
  }

    if (a.group_size == 1) {
      // ---- one CTA owns the whole field: nobody else reads its tables, so the schedule of a round is simply the list
      // of tiles that were posted to during the previous round (a post appends the tile when it sets the first
      // flag), and the tile table is only scanned when that list runs empty and the cost window has to move.
      uint32_t *lcur = my_list, *lnext = my_list + n_tiles;
      uint32_t *acur = act, *anext = act + n_tiles;
      // Longest first: the warps pull the round's tiles from a shared counter and the round ends with its slowest
      // visit, so tiles whose previous visit was longer than the previous round's average go to the front of the list.
      uint32_t *dur = my_list + 2 * n_tiles;
      for (int t = tid; t < n_tiles; t += WT_WARPS * 32) dur[t] = 0u;
      if (tid == 0) {
        s_mine = 0;
        s_dsum = 0;
        s_dcnt = 0;
      }
      __syncthreads();
      for (;;) {
        uint32_t ncur = s_mine;
        __syncthreads();
        if (ncur == 0) {
          if (tid == 0) s_gmin = INF_BITS;
          __syncthreads();
          unsigned long long mn = INF_BITS;
          for (int t = tid; t < n_tiles; t += WT_WARPS * 32) {
            const unsigned long long pm = __ldcg(pmin + t);
            mn = pm < mn ? pm : mn;
          }
          for (int d = 16; d; d >>= 1) {
            const unsigned long long o = __shfl_down_sync(FULL, mn, d);
            mn = o < mn ? o : mn;
          }
          if (lane == 0 && mn != INF_BITS) atomicMin(&s_gmin, mn);
          __syncthreads();
          const unsigned long long g = s_gmin;
          if (g == INF_BITS) break;  // nothing posted, nothing deferred: fixed point reached
          bound = __longlong_as_double((long long)g) + a.delta;
          windows++;
          for (int t = tid; t < n_tiles; t += WT_WARPS * 32)
            if (__longlong_as_double((long long)__ldcg(pmin + t)) < bound) lcur[atomicAdd(&s_mine, 1u)] = (uint32_t)t;
          __syncthreads();
          ncur = s_mine;
          __syncthreads();
        }
        rounds++;
        if (a.lpt && ncur > WT_WARPS && ncur <= WT_WARPS * 32) {
          const uint32_t thr = s_dcnt ? s_dsum / s_dcnt : 0u;
          __syncthreads();
          if (tid == 0) {
            s_next = 0;   // front cursor
            s_lnext = 0;  // back cursor
            s_dsum = 0;
            s_dcnt = 0;
          }
          __syncthreads();
          if (tid < (int)ncur) {
            const uint32_t t = __ldcg(lcur + tid);
            const bool longer = __ldcg(dur + t) > thr;
            const uint32_t pos = longer ? atomicAdd(&s_next, 1u) : ncur - 1u - atomicAdd(&s_lnext, 1u);
            s_pref[pos] = t;
          }
          __syncthreads();
          if (tid < (int)ncur) __stcg(lcur + tid, s_pref[tid]);
          __syncthreads();
        } else if (tid == 0) {
          s_dsum = 0;
          s_dcnt = 0;
        }
        if (tid == 0) {
          s_next = 0;
          s_lnext = 0;
        }
        __syncthreads();
        PF_MARK(0);
        for (;;) {
          uint32_t i = 0;
          if (lane == 0) i = atomicAdd(&s_next, 1u);
          i = __shfl_sync(FULL, i, 0);
          if (i >= ncur) break;
          const int tile = (int)__ldcg(lcur + i);
          const uint32_t af = __ldcg(acur + tile) & 0xffu;
          visits++;
          const long long t_visit = clock64();
          const unsigned long long left = warp_visit_tile<VARIANT>(a, cost, dmask, tile, af, bound, wt, wc, lane);
          pf[5]++;
          if (lane == 0) {
            const uint32_t dt = (uint32_t)min((long long)0xfffff, (clock64() - t_visit) >> 4);
            __stcg(dur + tile, dt);
            atomicAdd(&s_dsum, dt);
            atomicAdd(&s_dcnt, 1u);
            __stcg(acur + tile, 0u);  // this buffer receives the posts of the next round
            __stcg(pmin + tile, left);
            const uint32_t o = ((volatile uint32_t *)s_outw)[wid];
            if (o) {
              const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
#pragma unroll
              for (int d = 0; d < 8; d++) {
                constexpr int ddx[8] = {0, 0, -1, 1, -1, 1, -1, 1};
                constexpr int ddy[8] = {-1, 1, 0, 0, -1, -1, 1, 1};
                constexpr uint32_t facing[8] = {DIR_S, DIR_N, DIR_E, DIR_W, DIR_SE, DIR_SW, DIR_NE, DIR_NW};
                if (o & (1u << d)) {
                  const int nx = tx + ddx[d], ny = ty + ddy[d];
                  if (nx >= 0 && nx < a.tiles_x && ny >= 0 && ny < a.tiles_y) {
                    const int nb = ny * a.tiles_x + nx;
                    if (atomicOr(anext + nb, facing[d]) == 0u) __stcg(lnext + atomicAdd(&s_lnext, 1u), (uint32_t)nb);
                  }
                }
              }
            }
          }
          __syncwarp();
          PF_MARK(2);
        }
        PF_MARK(3);
        __syncthreads();  // round end: every visit of the round is written back
        PF_MARK(4);
        uint32_t *tl = lcur;
        lcur = lnext;
        lnext = tl;
        uint32_t *ta = acur;
        acur = anext;
        anext = ta;
        if (tid == 0) s_mine = s_lnext;
        __syncthreads();
      }
    } else
    for (uint32_t round = 0;; round++) {
      uint32_t *act_cur = act + (round % 3) * n_tiles, *act_next = act + ((round + 1) % 3) * n_tiles,
               *act_clr = act + ((round + 2) % 3) * n_tiles;
      unsigned long long *pm_cur = pmin + (round & 1) * n_tiles, *pm_next = pmin + ((round + 1) & 1) * n_tiles;
      // ---- schedule: one scan of the tile table builds my list of the round (ranks == grank mod group size,
      // rotated per round), counts the schedulable tiles and finds the global minimum of the deferred costs; when
      // nothing is schedulable the bound moves to that minimum + delta and the scan is repeated.  Housekeeping for
      // the tiles t == grank (mod group size): carry an unscheduled tile's minimum over
      // to the next round's table, zero the activity buffer of the round after (both idempotent under a re-scan:
      // a tile that becomes schedulable overwrites its minimum when it is processed).
      const bool solo = a.group_size == 1;
      // Scan of the tile table.  A lone CTA (one field per CTA) scans all of it.  In a group, CTA c scans only the
      // slice [c, c+1) * n_tiles / group_size, publishes (count, minimum, compacted slice list) to global memory, the
      // group meets at one extra barrier, and every CTA derives its share of the round from the group_size counts --
      // instead of 148 CTAs each pulling the whole table through L2 every round.
      const int s0 = solo ? 0 : (int)((long long)n_tiles * grank / a.group_size);
      const int s1 = solo ? n_tiles : (int)((long long)n_tiles * (grank + 1) / a.group_size);
      const int range = ((((s1 - s0) + WT_WARPS - 1) / WT_WARPS) + 31) & ~31;
      const int r0 = s0 + wid * range, r1 = min(s1, r0 + range);
      uint32_t *slice_cnt = a.sched + (size_t)group * 4 * a.group_size;        // [group_size] counts
      unsigned long long *slice_min = (unsigned long long *)(slice_cnt + 2 * a.group_size);  // [group_size] minima
      uint32_t total;
      for (;;) {
        if (tid == 0) {
          s_gmin = INF_BITS;
          s_mine = 0;
        }
        __syncthreads();
        // pass A: count my range, minimum of the deferred costs, housekeeping (whoever scans a tile owns it)
        unsigned long long mn = INF_BITS;
        uint32_t cnt = 0, onbits = 0;  // onbits: my tiles' verdicts, kept for pass B when the range has <= 32 steps
        const bool cached = range <= 32 * 32;
        for (int t0 = r0, k = 0; t0 < r1; t0 += 32, k++) {
          const int t = t0 + lane;
          bool on = false;
          if (t < r1) {
            const uint32_t af = __ldcg(act_cur + t);
            const unsigned long long pm = __ldcg(pm_cur + t);
            on = af != 0 || __longlong_as_double((long long)pm) < bound;
            mn = pm < mn ? pm : mn;
            if (!on) __stcg(pm_next + t, pm);
            __stcg(act_clr + t, 0u);
          }
          if (cached && on) onbits |= 1u << k;
          cnt += __popc(__ballot_sync(FULL, on));
        }
        for (int d = 16; d; d >>= 1) {
          const unsigned long long o = __shfl_down_sync(FULL, mn, d);
          mn = o < mn ? o : mn;
        }
        if (lane == 0) {
          s_warp_cnt[wid] = cnt;
          if (mn != INF_BITS) atomicMin(&s_gmin, mn);
        }
        __syncthreads();
        // exclusive prefix of the warp counts by shuffles (every warp computes the same scan)
        uint32_t v = lane < WT_WARPS ? s_warp_cnt[lane] : 0u, inc = v;
#pragma unroll
        for (int d = 1; d < WT_WARPS; d <<= 1) {
          const uint32_t o = __shfl_up_sync(FULL, inc, d);
          if (lane >= d) inc += o;
        }
        uint32_t rank0 = __shfl_sync(FULL, inc - v, wid);
        const uint32_t my_total = __shfl_sync(FULL, inc, WT_WARPS - 1);
        unsigned long long g = s_gmin;
        // pass B: compact the schedulable tiles of my range (in tile order) into my list
        if (my_total != 0) {
          for (int t0 = r0, k = 0; t0 < r1; t0 += 32, k++) {
            const int t = t0 + lane;
            bool on = false;
            if (cached) {
              on = (onbits >> k) & 1u;
            } else if (t < r1) {
              const uint32_t af = __ldcg(act_cur + t);
              const unsigned long long pm = __ldcg(pm_cur + t);
              on = af != 0 || __longlong_as_double((long long)pm) < bound;
            }
            const uint32_t b = __ballot_sync(FULL, on);
            if (on) my_list[rank0 + __popc(b & ((1u << lane) - 1u))] = (uint32_t)t;
            rank0 += __popc(b);
          }
        }
        total = my_total;
        if (!solo) {
          // publish my slice, meet, then read everybody's counts and minima
          if (tid == 0) {
            __stcg(slice_cnt + grank, my_total);
            __stcg(slice_min + grank, g);
          }
          group_barrier(bar, a.group_size);
          if (tid == 0) {
            s_mine = 0;
            s_gmin = INF_BITS;
          }
          uint32_t my_c = 0;
          if (tid < a.group_size) {
            my_c = __ldcg(slice_cnt + tid);
            s_pref[tid] = my_c;
          }
          __syncthreads();
          if (tid < a.group_size) {
            const unsigned long long m2 = __ldcg(slice_min + tid);
            if (m2 != INF_BITS) atomicMin(&s_gmin, m2);
            uint32_t pre = 0;
            for (int c = 0; c < tid; c++) pre += s_pref[c];
            if (tid == a.group_size - 1) s_mine = pre + my_c;
            my_c = pre;
          }
          __syncthreads();
          if (tid < a.group_size) s_pref[tid] = my_c;  // exclusive prefix: first global rank of slice tid
          __syncthreads();
          total = s_mine;
          g = s_gmin;
          __syncthreads();
          if (total == 0 && g != INF_BITS) group_barrier(bar, a.group_size);  // before the slices are rewritten
        }
        if (total != 0 || g == INF_BITS) break;
        bound = __longlong_as_double((long long)g) + a.delta;  // open the next cost window
        windows++;
      }
      if (total == 0) break;  // nothing active, nothing deferred: fixed point reached
      rounds++;
      if (tid == 0) s_next = 0;
      __syncthreads();
      // my share of the round: a lone CTA takes its whole list; in a group the schedulable tiles are ranked in slice
      // order and I take the ranks == grank + round (mod group size): rank -> (slice, offset) through the counts
      const uint32_t mine = solo ? total : (total + (uint32_t)a.group_size - 1u - (uint32_t)((grank + round) % a.group_size)) /
                                               (uint32_t)a.group_size;

      // ---- drain: every warp pulls tiles until the list is empty
      for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(&s_next, 1u);
        i = __shfl_sync(FULL, i, 0);
        if (i >= mine) break;
        int tile;
        if (solo) {
          tile = (int)my_list[i];
        } else {
          // global rank of my i-th tile, then the slice that holds it (the counts are few: a serial walk by lane 0)
          const uint32_t rk = (uint32_t)((grank + round) % a.group_size) + i * (uint32_t)a.group_size;
          int tl = 0;
          if (lane == 0) {
            int lo = 0, hi = a.group_size - 1;  // last slice whose first rank is <= rk
            while (lo < hi) {
              const int mid = (lo + hi + 1) >> 1;
              if (s_pref[mid] <= rk) lo = mid;
              else hi = mid - 1;
            }
            tl = (int)__ldcg(a.lists + (size_t)(group * a.group_size + lo) * LISTS_PER_CTA * a.n_tiles + (rk - s_pref[lo]));
          }
          tile = __shfl_sync(FULL, tl, 0);
        }
        const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
        visits++;
        PF_MARK(0);
        const uint32_t af = __ldcg(act_cur + tile);
        const unsigned long long left = warp_visit_tile<VARIANT>(a, cost, dmask, tile, af, bound, wt, wc, lane);
        pf[5]++;
        PF_MARK(2);
        if (lane == 0) {
          __stcg(pm_next + tile, left);
          const uint32_t o = ((volatile uint32_t *)s_outw)[wid];
          if (o) {
            __threadfence();
#pragma unroll
            for (int d = 0; d < 8; d++) {
              constexpr int ddx[8] = {0, 0, -1, 1, -1, 1, -1, 1};
              constexpr int ddy[8] = {-1, 1, 0, 0, -1, -1, 1, 1};
              constexpr uint32_t facing[8] = {DIR_S, DIR_N, DIR_E, DIR_W, DIR_SE, DIR_SW, DIR_NE, DIR_NW};
              if (o & (1u << d)) {
                const int nx = tx + ddx[d], ny = ty + ddy[d];
                if (nx >= 0 && nx < a.tiles_x && ny >= 0 && ny < a.tiles_y)
                  atomicOr(act_next + ny * a.tiles_x + nx, facing[d]);
              }
            }
          }
        }
        __syncwarp();
        PF_MARK(3);
      }
      PF_MARK(0);
      group_barrier(bar, a.group_size);
      PF_MARK(4);
    }
#undef PF_MARK
    if (a.prof && tid == 0)
      for (int i = 0; i < 8; i++) a.prof[blockIdx.x * 8 + i] = pf[i];
    if (grank == 0 && tid == 0) {
      uint32_t *sp = a.stats + field * 4;
      sp[0] = rounds;
      sp[2] = windows;
    }
    if (lane == 0 && visits) atomicAdd(a.stats + field * 4 + 1, visits);
  }
}
// ---------------------------------------------------------------- init / seed / stats
__global__ void __launch_bounds__(256) heur_init_kernel(double *__restrict__ cost, size_t total, double init,
                                                        unsigned long long *__restrict__ pmin, size_t n_pmin) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = i0; i < total; i += stride) cost[i] = init;
  for (size_t i = i0; i < n_pmin; i += stride) pmin[i] = INF_BITS;
}

template <int NB>
__global__ void heur_seed_kernel(double *__restrict__ cost, size_t n, size_t n_fields, const int32_t *__restrict__ goal_xy,
                                 int W, unsigned long long *__restrict__ pmin, uint32_t *__restrict__ dmask, int tiles_x,
                                 int n_tiles) {
  constexpr int T = tile_cfg<NB>::T, NBLK = tile_cfg<NB>::NBLK;
  const size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_fields) return;
  const int gx = goal_xy[2 * f], gy = goal_xy[2 * f + 1];
  cost[f * n + (size_t)gy * W + gx] = 0.0;  // goal = 0 even if occupied (RS/src/HybridAstar.cpp:378)
  // the goal's "change" is a deferred cell of cost 0: the first window releases it
  const int tile = (gy / T) * tiles_x + gx / T;
  const int lx = gx % T, ly = gy % T;
  const int blk = (ly / HB) * NB + lx / HB;
  dmask[((f * n_tiles + tile) * NBLK + blk) * 32 + (ly % HB)] = 1u << (lx % HB);
  pmin[f * 2 * n_tiles + tile] = 0ull;
}

// async kernel: control words, the goal tile queued, the first window open (state and ring are memset by the host)
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

void make_weights(int variant, heur_weights *w) {
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

double env_double(const char *name, double dflt) {
  const char *s = getenv(name);
  return (s && *s) ? atof(s) : dflt;
}

template <int NB, int VARIANT>
hpp_status launch_persist(hpp_ctx *c, heur_args &a, int ctas_per_field_hint) {
  using C = tile_cfg<NB>;
  // warp-tile kernel: 20 warps per CTA when every field gets its own CTA (throughput), else 16 (latency).  Batches of at
  // least two fields per SM on maps up to 2048^2 get 4-warp CTAs, five per SM, so that five fields share an SM instead
  // of one: the round of a 4-warp CTA waits for the slowest of 4 tiles, not of 20, and a small map's wave front (a few
  // tiles) leaves most of a big CTA's warps idle.
  const bool many = NB == 1 && VARIANT == HPP_VARIANT_RS && ctas_per_field_hint <= 1 && a.n_fields >= c->sm_count / 2;
  const int warps_env = (int)env_double("HPP_HEUR_WARPS", 0);
  const bool small = NB == 1 && ctas_per_field_hint <= 1 &&
                     (warps_env ? warps_env == 4 : (a.n_fields >= 2 * c->sm_count && a.n_tiles <= 4096));
  // (measured against the 20-warp CTAs with five resident 4-warp CTAs per SM: 2 048 fields of 256^2 6.5 vs 14.7 ms;
  // 400 x 512^2 7.6 vs 9.1; 300 x 1024^2 21.1 vs 25.5; 600 x 1024^2 33.3 vs 43.0; 600 x 2048^2 131.6 vs 145.8;
  // 600 x 4096^2 534 vs 540 ms -- the same, so the big maps keep the 20-warp CTAs; 148 x 4096^2: 139 vs 112)
  const int WT_WARPS = small ? 4 : (many ? 20 : 16);
  const int threads = NB == 1 ? WT_WARPS * 32 : C::THREADS;
  size_t smem;
  const void *kern;
  smem = (size_t)WT_WARPS * (WSLOT_DOUBLES * sizeof(double) + (VARIANT == HPP_VARIANT_DUBINS ? C::ROWS * C::S : 0) + sizeof(warp_aux)) + 128;
  kern = small ? (const void *)heur_warp_kernel<VARIANT, 4>
               : (many ? (const void *)heur_warp_kernel<VARIANT, (VARIANT == HPP_VARIANT_RS ? 20 : 16)>
                       : (const void *)heur_warp_kernel<VARIANT, 16>);
  HPP_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // all of the SM's shared memory: five 4-warp CTAs (or one 20-warp CTA) need 207 KB of it
  HPP_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int per_sm = 0;
  HPP_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  if (per_sm < 1) HPP_FAIL(c, HPP_ERR_CUDA, "heuristic kernel does not fit on an SM (%zu B shared)", smem);
  const int per_sm_cap = (int)env_double("HPP_HEUR_CTAS_PER_SM", small ? 5 : 1);
  if (per_sm > per_sm_cap) per_sm = per_sm_cap;
  const int G = per_sm * c->sm_count;  // co-resident CTAs
  // CTA groups: one field at a time per group, the groups never synchronise with each other.  A tile solve is a
  // latency chain on a few warps, so a single field cannot fill the GPU (its CTAs mostly wait for the slowest tile
  // of the round); with several fields the GPU is shared evenly, down to one CTA per field (no barrier at all).
  int gs = ctas_per_field_hint > 0 ? ctas_per_field_hint : G / a.n_fields;
  const int useful = NB == 1 ? (a.n_tiles + WT_WARPS - 1) / WT_WARPS : a.n_tiles;
  if (gs > useful) gs = useful;  // never more CTAs than tiles (warp-tile kernel: than tiles / warps per CTA)
  if (gs > G) gs = G;
  if (NB == 1 && gs > WT_WARPS * 32) gs = WT_WARPS * 32;  // the group's slice counts are scanned by one CTA
  if (gs < 1) gs = 1;
  int n_groups = G / gs;
  if (n_groups > a.n_fields) n_groups = a.n_fields;
  // One CTA per field never waits for another CTA, so it needs no co-residency guarantee: a plain launch with one CTA
  // per field lets the hardware scheduler balance the fields and lets the kernel share the GPU with other streams
  // (a cooperative launch would wait until every one of its CTAs fits at once).
  const bool solo = gs == 1;
  if (solo) n_groups = a.n_fields;
  a.group_size = gs;
  a.n_groups = n_groups;
  const int grid = gs * n_groups;
  if constexpr (NB == 1) {
    void *lp;
    const size_t lists_b = ((sizeof(uint32_t) * (size_t)grid * LISTS_PER_CTA * a.n_tiles) + 255) & ~(size_t)255;
    hpp_status s = hpp_scratch(c, 4, lists_b + sizeof(uint32_t) * 4 * (size_t)grid + 256, &lp);
    if (s != HPP_OK) return s;
    a.lists = (uint32_t *)lp;
    a.sched = (uint32_t *)((uint8_t *)lp + lists_b);
  }
  void *params[] = {&a};
  if (solo)
    HPP_CUDA(c, cudaLaunchKernel(kern, dim3(grid), dim3(threads), params, smem, c->stream));
  else
    HPP_CUDA(c, cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(threads), params, smem, c->stream));
  HPP_LAUNCHED(c);
  return HPP_OK;
}

template <int NB>
hpp_status run_batch(hpp_map *m, int variant, const int32_t *goal_xy, size_t n_fields, double *cost_dev,
                     hpp_quirk_report *reports) {
  using C = tile_cfg<NB>;
  hpp_ctx *c = m->ctx;
  const int W = m->g.W, H = m->g.H;
  const int tiles_x = hpp_div_up(W, C::T), tiles_y = hpp_div_up(H, C::T), n_tiles = tiles_x * tiles_y;
  // scratch slot 2: act (3 x) | pmin (2 x) | dmask | barriers | stats | goals | field stats
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t act_b = al(sizeof(uint32_t) * 3 * n_tiles * n_fields);
  const size_t pmin_b = al(sizeof(unsigned long long) * 2 * n_tiles * n_fields);
  const size_t dmask_b = al(sizeof(uint32_t) * (size_t)n_tiles * C::NBLK * 32 * n_fields);
  const size_t bar_b = al(sizeof(uint32_t) * 32 * ((size_t)c->sm_count * 8 + n_fields));
  const size_t stats_b = al(sizeof(uint32_t) * 4 * n_fields);
  const size_t goals_b = al(n_fields * 8);
  const size_t fstats_b = al(n_fields * 16);
  void *scr;
  hpp_status s = hpp_scratch(c, 2, act_b + pmin_b + dmask_b + bar_b + stats_b + goals_b + fstats_b, &scr);
  if (s != HPP_OK) return s;
  uint8_t *p = (uint8_t *)scr;
  heur_args a;
  memset(&a, 0, sizeof a);
  a.act = (uint32_t *)p;                    p += act_b;
  a.pmin = (unsigned long long *)p;         p += pmin_b;
  a.dmask = (uint32_t *)p;                  p += dmask_b;
  a.bar = (uint32_t *)p;                    p += bar_b;
  a.stats = (uint32_t *)p;                  p += stats_b;
  int32_t *d_goals = (int32_t *)p;          p += goals_b;
  unsigned long long *d_fstats = (unsigned long long *)p;
  void *h_stage;
  s = hpp_pinned(c, 1, goals_b + stats_b + fstats_b, &h_stage);
  if (s != HPP_OK) return s;
  memcpy(h_stage, goal_xy, n_fields * 8);
  HPP_CUDA(c, cudaMemcpyAsync(d_goals, h_stage, n_fields * 8, cudaMemcpyHostToDevice, c->stream));
  // zero: act, dmask, barriers, stats (pmin is filled with +inf by the init kernel)
  HPP_CUDA(c, cudaMemsetAsync(a.act, 0, act_b, c->stream));
  HPP_CUDA(c, cudaMemsetAsync(a.dmask, 0, dmask_b + bar_b + stats_b, c->stream));

  make_weights(variant, &a.wt);
  a.cost_all = cost_dev;
  a.occ_bits = m->d_occ_bits;
  a.near_bits = m->d_near_bits;
  a.cls = m->d_cls;
  a.pitch = m->bits_pitch;
  a.W = W;
  a.H = H;
  a.tiles_x = tiles_x;
  a.tiles_y = tiles_y;
  a.n_tiles = n_tiles;
  a.n_fields = (int)n_fields;
  // window width: wide windows = fewer rounds (latency of a lone field), narrow = fewer re-evaluations (throughput)
  // Re-measured on the final kernel (tools/heur_big_sweep.sh, tools/heur_small_sweep.sh): 148 fields of 4096^2, one CTA per
  // field: 117.4 (48) / 113.9 (64) / 112.1 (96) / 111.4 (128) / 113.1 (160) / 118.6 (256) / 141.7 ms (768); 2 048 fields of
  // 256^2: parking map 11.4 (64) / 10.7 (128) / 10.3 (256), cluttered map 12.1 / 12.0 / 13.0 ms.
  a.delta = env_double("HPP_HEUR_DELTA", NB == 1 ? (n_fields < 8 ? 768.0 : 128.0) : 2.0 * C::T);
  a.delta_in = env_double("HPP_HEUR_DELTA_IN", 32.0);
  a.lpt = (int)env_double("HPP_HEUR_LPT", 1);
  // the fields as one (W, H, n_fields) fp64 tensor for the TMA: a tile with its halo is one 38 x 34 box.  Needs 16-byte
  // multiples for the row pitch (W even); anything else keeps the row-wise bulk copies / plain loads.
  a.tmap_ok = 0;
  memset(&a.tmap, 0, sizeof a.tmap);
  if (NB == 1 && (W % 2) == 0 && ((uintptr_t)cost_dev & 15) == 0 && env_double("HPP_HEUR_TMA2D", 1) != 0) {
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    // looked up once per process (thread-safe static initialisation: plan_batch calls come from several host threads)
    static const encode_fn encode = []() -> encode_fn {
      void *fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
          qres == cudaDriverEntryPointSuccess)
        return (encode_fn)fn;
      (void)cudaGetLastError();
      return nullptr;
    }();
    if (encode) {
      const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_fields};
      const cuuint64_t gstride[2] = {(cuuint64_t)W * 8, (cuuint64_t)W * H * 8};
      const cuuint32_t box[3] = {(cuuint32_t)C::S, (cuuint32_t)C::ROWS, 1};
      const cuuint32_t estr[3] = {1, 1, 1};
      if (encode(&a.tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)cost_dev, gdim, gstride, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
        a.tmap_ok = 1;
    }
  }
  const bool debug = getenv("HPP_HEUR_DEBUG") != nullptr;
  const size_t prof_n = ((size_t)c->sm_count * 8 + n_fields) * 8 + 8;
  if (debug) {
    void *pp;
    s = hpp_scratch(c, 5, prof_n * 8, &pp);
    if (s != HPP_OK) return s;
    a.prof = (unsigned long long *)pp;
    HPP_CUDA(c, cudaMemsetAsync(a.prof, 0, prof_n * 8, c->stream));
  }

  const size_t total = m->n * n_fields;
  size_t want = total / 1024 + 1;
  const int grid = (int)(want < (size_t)c->sm_count * 16 ? want : (size_t)c->sm_count * 16);
  heur_init_kernel<<<grid, 256, 0, c->stream>>>(cost_dev, total, a.wt.init, a.pmin, (size_t)2 * n_tiles * n_fields);
  HPP_LAUNCHED(c);
  heur_seed_kernel<NB><<<hpp_div_up((long long)n_fields, 128), 128, 0, c->stream>>>(cost_dev, m->n, n_fields, d_goals, W,
                                                                                   a.pmin, a.dmask, tiles_x, n_tiles);
  HPP_LAUNCHED(c);
  const int hint = (int)env_double("HPP_HEUR_GROUP", 0);
  // (an asynchronous tile-queue kernel for a lone field was tried in round 1: it removes the wait for the slowest tile
  // of a round, but visits tiles eagerly -- 1.7x the visits -- and ended up level with the round-based kernel: both are
  // bound by the chain of dependent sweeps)
  s = (variant == HPP_VARIANT_RS) ? launch_persist<NB, HPP_VARIANT_RS>(c, a, hint)
                                  : launch_persist<NB, HPP_VARIANT_DUBINS>(c, a, hint);
  if (s != HPP_OK) return s;

  if (debug) {
    const int grid = a.group_size * a.n_groups;
    std::vector<unsigned long long> hp((size_t)grid * 8 + 8);
    HPP_CUDA(c, cudaMemcpyAsync(hp.data(), a.prof, hp.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    HPP_CUDA(c, cudaStreamSynchronize(c->stream));
    const char *nm[8] = {"schedule", "stage", "solve", "writeback", "barrier", "tiles", NB == 1 ? "sweeps" : "phases",
                         NB == 1 ? "iters" : "max_solve"};
    for (int i = 0; i < 8; i++) {
      unsigned long long sum = 0, mx = 0;
      for (int b = 0; b < grid; b++) {
        sum += hp[(size_t)b * 8 + i];
        if (hp[(size_t)b * 8 + i] > mx) mx = hp[(size_t)b * 8 + i];
      }
      fprintf(stderr, "[hpp heur prof] %-9s mean/CTA=%12.0f max=%12llu cta0=%12llu\n", nm[i], (double)sum / grid, mx, hp[i]);
    }
    {
      const unsigned long long *pw = hp.data() + (size_t)grid * 8;
      fprintf(stderr, "[hpp heur prof] warp totals: sweeps=%llu heavy_iter=%llu sparse_iter=%llu needed_cells=%llu ideal_iter=%llu\n", pw[4],
              pw[0], pw[1], pw[2], pw[3]);
    }
  }
  if (debug) {
    uint32_t *h_stats = (uint32_t *)((uint8_t *)h_stage + goals_b);
    HPP_CUDA(c, cudaMemcpyAsync(h_stats, a.stats, n_fields * 16, cudaMemcpyDeviceToHost, c->stream));
    HPP_CUDA(c, cudaStreamSynchronize(c->stream));
    for (size_t f = 0; f < n_fields && f < 4; f++)
      fprintf(stderr, "[hpp heur] delta=%g groups=%d x %d CTAs, field %zu: rounds=%u visits=%u windows=%u\n", a.delta,
              a.n_groups, a.group_size, f, h_stats[4 * f], h_stats[4 * f + 1], h_stats[4 * f + 2]);
  }
  if (reports) {
    uint32_t *h_stats = (uint32_t *)((uint8_t *)h_stage + goals_b);
    unsigned long long *h_fstats = (unsigned long long *)((uint8_t *)h_stage + goals_b + stats_b);
    HPP_CUDA(c, cudaMemsetAsync(d_fstats, 0, n_fields * 16, c->stream));
    dim3 sgrid(c->sm_count * 4, (unsigned)n_fields);
    heur_stats_kernel<<<sgrid, 256, 0, c->stream>>>(cost_dev, m->n, variant == HPP_VARIANT_RS ? INFINITY : 1000.0, d_fstats);
    HPP_LAUNCHED(c);
    HPP_CUDA(c, cudaMemcpyAsync(h_stats, a.stats, n_fields * 16, cudaMemcpyDeviceToHost, c->stream));
    HPP_CUDA(c, cudaMemcpyAsync(h_fstats, d_fstats, n_fields * 16, cudaMemcpyDeviceToHost, c->stream));
    HPP_CUDA(c, cudaStreamSynchronize(c->stream));
    for (size_t f = 0; f < n_fields; f++) {
      hpp_quirk_report *r = &reports[f];
      memset(r, 0, sizeof *r);
      r->reachable = h_fstats[2 * f];
      memcpy(&r->max_cost, &h_fstats[2 * f + 1], 8);
      r->rounds = h_stats[4 * f];
      r->tile_visits = h_stats[4 * f + 1];
      if (variant == HPP_VARIANT_RS) {
        const int MAX_COST = (W + H > 5000) ? W + H : 5000;  // RS/src/HybridAstar.cpp:395
        r->h2_bucket_clamp = r->max_cost >= (double)(MAX_COST - 1);
      }
    }
  } else {
    // the relaxation is one stream-ordered launch; completion (and any device fault) is reported here
    HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return HPP_OK;
}

}  // namespace

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
  return run_batch<1>(m, variant, goal_xy, n_fields, cost_dev, reports);
}

// host-buffer batch: n_fields goal cells on one map, fields returned in cost (n_fields x W*H doubles, pinned memory
// recommended).  Large batches are solved in two or more chunks so that the device-to-host copy of one chunk (a
// field is 8 B/cell: the copy is PCIe-bound) overlaps the relaxation of the next.
extern "C" hpp_status hpp_heuristic_batch(hpp_map *m, int variant, const int32_t *goal_xy, size_t n_fields, double *cost,
                                          hpp_quirk_report *reports) {
  if (!m || !goal_xy || !cost || n_fields == 0) return HPP_ERR_ARG;
  hpp_ctx *c = m->ctx;
  HPP_CUDA(c, cudaSetDevice(c->device));
  if (!c->copy_stream) {
    HPP_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    HPP_CUDA(c, cudaEventCreateWithFlags(&c->copy_done[0], cudaEventDisableTiming));
    HPP_CUDA(c, cudaEventCreateWithFlags(&c->copy_done[1], cudaEventDisableTiming));
  }
  size_t free_b = 0, total_b = 0;
  HPP_CUDA(c, cudaMemGetInfo(&free_b, &total_b));
  const size_t field_b = m->n * sizeof(double);
  // chunks of >= 16 fields: the relaxation of a chunk takes about as long as its copy-out (2.4 ms per 134 MB field over
  // PCIe), and the first chunk's relaxation is the only part that no copy overlaps
  size_t chunk = n_fields;
  const size_t n_chunks = (size_t)env_double("HPP_HEUR_CHUNKS", 8);
  if (n_fields >= 64 && n_chunks > 1) {
    chunk = (n_fields + n_chunks - 1) / n_chunks;
    if (chunk < 16) chunk = 16;
  }
  const size_t have = c->d_scratch_bytes[3] + free_b / 2;
  const int nbuf = chunk < n_fields ? 2 : 1;
  if (nbuf * chunk * field_b > have) chunk = have / (2 * field_b);
  if (chunk < 1) HPP_FAIL(c, HPP_ERR_NOMEM, "not enough device memory for two heuristic fields");
  void *d_cost;
  hpp_status s = hpp_scratch(c, 3, (chunk < n_fields ? 2 : 1) * chunk * field_b, &d_cost);
  if (s != HPP_OK) return s;
  size_t k = 0;
  for (size_t f0 = 0; f0 < n_fields; f0 += chunk, k++) {
    const size_t nf = n_fields - f0 < chunk ? n_fields - f0 : chunk;
    double *buf = (double *)d_cost + (k & 1) * chunk * m->n;
    if (k >= 2) HPP_CUDA(c, cudaEventSynchronize(c->copy_done[k & 1]));  // the copy that last read this buffer
    s = hpp_heuristic_batch_dev(m, variant, goal_xy + 2 * f0, nf, buf, reports ? reports + f0 : nullptr);
    if (s != HPP_OK) return s;  // (returns after the relaxation kernel completed)
    HPP_CUDA(c, cudaMemcpyAsync(cost + f0 * m->n, buf, nf * field_b, cudaMemcpyDeviceToHost, c->copy_stream));
    HPP_CUDA(c, cudaEventRecord(c->copy_done[k & 1], c->copy_stream));
  }
  HPP_CUDA(c, cudaStreamSynchronize(c->copy_stream));
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
