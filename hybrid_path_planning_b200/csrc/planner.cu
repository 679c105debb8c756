// planner.cu -- whole Hybrid A* queries on the device: HybridAstar::run (RS/src/HybridAstar.cpp:553-729,
// DB/src/HybridAstar.cpp:315-496) for a batch of independent start/goal pairs on one map.
//
// One warp owns one query from start to finish (and then takes the next one): no host round trip per A*
// iteration.  Per query the warp keeps, in its own slice of global memory,
//   * a node pool (state, Cost_path, steering angle, direction, parent, command index),
//   * the open list as a binary heap of (f, node) that replays libstdc++'s std::push_heap / std::pop_heap for
//     the reference's cost-only comparator, so equal-f nodes pop in the reference's order,
//   * one open-addressing hash table keyed by the State doubles that plays both openSet and closedSet
//     (value = node index, low bit = "closed").
// Each iteration follows the reference statement by statement: pop -> openSet lookup -> closed test -> analytic
// shot (K7/K8 device code) -> goal test (same cell and |dheading| <= 0.05 with PI = 3.14159) -> expansion
// (K6 device code: lanes = motion commands for RS, warp-wide SAT per state for Dubins) -> insert/replace in
// command order.  The 3500-iteration cap, the 1-state failure trajectory and the back-track order are kept.
// Child trajectories are not stored: they are re-rolled from (parent state, command) at back-track time.
// The heuristic field of every distinct goal cell comes from K3 (hpp_heuristic_batch_dev) beforehand.
#include <algorithm>
#include <vector>

#include "expand.cuh"
#include "shots.cuh"

#define PLAN_WARPS 4
#define PLAN_ITER_CAP 3500
#define PI_ASTAR 3.14159 /* RS/include/HybridAstar.h:24 */

struct pnode {
  double x, y, h, cost, steer;
  int parent;
  signed char dir;
  signed char cmd;  // -1 = start, -2 = shot node
  short pad;
};

struct plan_consts {
  hpp_geom g;
  expand_consts ec;
  circ_consts cc;
  float rx[5], ry[5];
  double maxc, max_steer;
  int max_nodes;     // node pool / heap capacity per query
  int table_mask;    // hash slots - 1
};

void hpp_make_expand_consts(const hpp_map *m, int variant, expand_consts *ec);  // expand.cu

__device__ __forceinline__ unsigned long long dkey(double d) {
  if (d == 0.0) d = 0.0;  // -0.0 == 0.0
  return (unsigned long long)__double_as_longlong(d);
}
__device__ __forceinline__ unsigned state_hash(double x, double y, double h) {
  unsigned long long a = dkey(x) * 0x9E3779B97F4A7C15ull;
  a ^= (dkey(y) + 0x7F4A7C15ull + (a << 6) + (a >> 2));
  a ^= (dkey(h) * 0xC2B2AE3D27D4EB4Full + (a << 6) + (a >> 2));
  a ^= a >> 29;
  return (unsigned)a;
}

// returns the table value (node << 1 | closed) or -1; slot = where it was found / the first empty slot
__device__ __forceinline__ int table_find(const int *table, const pnode *nodes, int mask, double x, double y, double h,
                                          int &slot) {
  int i = (int)(state_hash(x, y, h) & (unsigned)mask);
  for (;;) {
    const int v = table[i];
    if (v < 0) {
      slot = i;
      return -1;
    }
    const pnode &nd = nodes[v >> 1];
    if (nd.x == x && nd.y == y && nd.h == h) {
      slot = i;
      return v;
    }
    i = (i + 1) & mask;
  }
}

// ---- libstdc++ heap replay on global arrays (one lane) ----
__device__ __forceinline__ void gheap_push(double *key, int *idx, int &n, double k, int p) {
  int hole = n++;
  int parent = (hole - 1) / 2;
  while (hole > 0 && key[parent] > k) {
    key[hole] = key[parent];
    idx[hole] = idx[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  key[hole] = k;
  idx[hole] = p;
}
__device__ __forceinline__ int gheap_pop(double *key, int *idx, int &n) {
  const int top = idx[0];
  if (n > 1) {
    const double vk = key[n - 1];
    const int vp = idx[n - 1];
    const int len = n - 1;
    int hole = 0, second = 0;
    while (second < (len - 1) / 2) {
      second = 2 * (second + 1);
      if (key[second] > key[second - 1]) second--;
      key[hole] = key[second];
      idx[hole] = idx[second];
      hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
      second = 2 * (second + 1);
      key[hole] = key[second - 1];
      idx[hole] = idx[second - 1];
      hole = second - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > 0 && key[parent] > vk) {
      key[hole] = key[parent];
      idx[hole] = idx[parent];
      hole = parent;
      parent = (hole - 1) / 2;
    }
    key[hole] = vk;
    idx[hole] = vp;
  }
  n--;
  return top;
}

template <int VARIANT>
__global__ void __launch_bounds__(PLAN_WARPS * 32) plan_kernel(plan_consts pc, const float *__restrict__ dist,
                                                               const uint32_t *__restrict__ bits, int pitch,
                                                               const double *__restrict__ starts,
                                                               const double *__restrict__ goals,
                                                               const int *__restrict__ query_ids, int n_queries,
                                                               const int *__restrict__ field_of_query,
                                                               const double *__restrict__ fields,
                                                               pnode *__restrict__ node_pool, double *__restrict__ key_pool,
                                                               int *__restrict__ idx_pool, int *__restrict__ table_pool,
                                                               hpp_plan *__restrict__ plans, double *__restrict__ traj,
                                                               unsigned long long traj_cap,
                                                               unsigned long long *__restrict__ cursor) {
  __shared__ rs_warp_smem S_all[PLAN_WARPS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  rs_warp_smem &S = S_all[wid];
  const int gwarp = blockIdx.x * PLAN_WARPS + wid, nwarps = gridDim.x * PLAN_WARPS;
  const hpp_geom &g = pc.g;
  const int W = g.W, H = g.H;
  const long map_size = (long)W * H;
  pnode *nodes = node_pool + (size_t)gwarp * pc.max_nodes;
  double *hkey = key_pool + (size_t)gwarp * pc.max_nodes;
  int *hidx = idx_pool + (size_t)gwarp * pc.max_nodes;
  int *table = table_pool + (size_t)gwarp * (pc.table_mask + 1);
  const int ncmd = pc.ec.n_cmd;

  for (int qi = gwarp; qi < n_queries; qi += nwarps) {
    const int q = query_ids[qi];
    const double *hmap = fields + (size_t)field_of_query[qi] * map_size;
    const double sx = starts[3 * q], sy = starts[3 * q + 1], sh = starts[3 * q + 2];
    const double gx = goals[3 * q], gy = goals[3 * q + 1], gh = goals[3 * q + 2];
    const int s_cx = (int)(hpp_div_res(g, sx - g.ox)), s_cy = (int)(hpp_div_res(g, sy - g.oy));
    const int g_cx = (int)(hpp_div_res(g, gx - g.ox)), g_cy = (int)(hpp_div_res(g, gy - g.oy));
    for (int i = lane; i <= pc.table_mask; i += 32) table[i] = -1;
    __syncwarp();
    int n_nodes = 0, hn = 0, iterations = 0;
    int goal_node = -1, via_shot = 0, status = 0;
    rs_shot_result SR;
    SR.found = 0; SR.n_samples = 0; SR.cand = -1; SR.chosen = -1; SR.cost = 0; SR.n_paths = 0;
    db_best DB;
    DB.nmode = 0;
    int shot_samples = 0;
    double shot_cost = 0.0;

    const long start_index = (long)s_cy * W + s_cx;
    if (start_index < 0 || start_index >= map_size) {
      status = 2;  // the reference indexes goal_map_ out of range here (undefined behaviour)
    } else {
      if (lane == 0) {
        pnode nd;
        nd.x = sx; nd.y = sy; nd.h = sh; nd.cost = 0.0; nd.steer = 0.0; nd.parent = -1; nd.dir = 1; nd.cmd = -1; nd.pad = 0;
        nodes[0] = nd;
        int slot;
        table_find(table, nodes, pc.table_mask, sx, sy, sh, slot);
        table[slot] = 0;  // node 0, open
        gheap_push(hkey, hidx, hn, 0.0 + 30.0 * hmap[start_index], 0);
      }
      n_nodes = 1;
      hn = __shfl_sync(0xffffffffu, hn, 0);
      __syncwarp();
    }

    while (status == 0 && hn > 0) {
      iterations++;
      if (iterations > PLAN_ITER_CAP) break;
      // ---- pop
      int cur = -1, slot = 0, tv = -1;
      if (lane == 0) {
        const int top = gheap_pop(hkey, hidx, hn);
        const pnode &tn = nodes[top];
        tv = table_find(table, nodes, pc.table_mask, tn.x, tn.y, tn.h, slot);
      }
      hn = __shfl_sync(0xffffffffu, hn, 0);
      tv = __shfl_sync(0xffffffffu, tv, 0);
      slot = __shfl_sync(0xffffffffu, slot, 0);
      if (tv < 0 || (tv & 1)) continue;  // already in the closed set
      cur = tv >> 1;                     // openSet[currentState] (may be a cheaper replacement of the popped node)
      if (lane == 0) table[slot] = tv | 1;  // openSet.erase + closedSet[currentState] = currentNode
      __syncwarp();
      const pnode cn = nodes[cur];

      // ---- analytic shot
      if (VARIANT == HPP_VARIANT_RS) {
        rs_generate(S, cn.x, cn.y, cn.h, gx, gy, gh, pc.maxc);
        SR = rs_select(S, g, pc.cc, dist, cn.x, cn.y, cn.h, cn.cost, pc.maxc, pc.max_steer);
        if (SR.found) {
          shot_samples = SR.n_samples;
          shot_cost = SR.cost;
        }
      } else {
        DB = db_plan(cn.x, cn.y, cn.h, gx, gy, gh, pc.maxc);
        sat_visitor vis;
        vis.g = &g; vis.rx = pc.rx; vis.ry = pc.ry; vis.bits = bits; vis.pitch = pitch; vis.hit_zero_idx = 0x7fffffff;
        const int n_keep = warp_sample_path<true>(DB.len, DB.typ, DB.nmode, pc.maxc, 0.1, cn.x, cn.y, cn.h, vis);
        const bool collides = (n_keep < 0) || (vis.hit_zero_idx < n_keep);
        SR.found = (n_keep > 0 && !collides);
        if (SR.found) {
          double cost = cn.cost;
          cost += fabs(DB.best_cost);
          for (int i = 0; i < 3; i++)
            if (DB.typ[i] != 'S') cost += pc.max_steer * 15.0;
          shot_samples = n_keep;
          shot_cost = cost;
        }
      }
      if (SR.found) {
        goal_node = cur;  // the shot node hangs off `cur`
        via_shot = 1;
        break;
      }

      // ---- goal test (RS/src/HybridAstar.cpp:641-655)
      {
        double hd = fabs(cn.h - gh);
        if (hd > PI_ASTAR) hd = fabs(hd - 2 * PI_ASTAR);
        const int c_cx = (int)(hpp_div_res(g, cn.x - g.ox)), c_cy = (int)(hpp_div_res(g, cn.y - g.oy));
        if (c_cx == g_cx && c_cy == g_cy && hd <= 0.05) {
          goal_node = cur;
          break;
        }
      }

      // ---- expansion: lane = command
      double cxs = 0, cys = 0, chs = 0, ccost = 0;
      int valid = 0;
      if (VARIANT == HPP_VARIANT_RS) {
        if (lane < ncmd) {
          double x = cn.x, y = cn.y, h = cn.h;
          int hit = 0;
          for (int i = 0; i < pc.ec.sim; i++) {
            vehicle_step(x, y, h, pc.ec.dir[lane], pc.ec.step, pc.ec.dtheta[lane]);
            if (!hit) hit = circles_collide(g, pc.cc, dist, x, y, h);
          }
          valid = !hit;
          cxs = x; cys = y; chs = h;
        }
      } else {
        for (int c = 0; c < ncmd; c++) {
          double x = cn.x, y = cn.y, h = cn.h;
          int hit = 0;
          for (int i = 0; i < pc.ec.sim; i++) {
            vehicle_step(x, y, h, pc.ec.dir[c], pc.ec.step, pc.ec.dtheta[c]);
            if (!hit) hit = sat_collide_warp(g, pc.rx, pc.ry, bits, pitch, x, y, h);
          }
          if (lane == c) {
            valid = !hit;
            cxs = x; cys = y; chs = h;
          }
        }
      }
      double fkey = 0.0;
      int ins_slot = -1, replace = 0;
      if (valid) {
        ccost = path_cost(pc.ec, lane, cn.cost, cn.steer, cn.dir);
        int sl;
        const int v = table_find(table, nodes, pc.table_mask, cxs, cys, chs, sl);
        if (v >= 0 && (v & 1)) valid = 0;  // in the closed set
        else {
          const int ccx = (int)(hpp_div_res(g, cxs - g.ox)), ccy = (int)(hpp_div_res(g, cys - g.oy));
          const long nidx = (long)ccy * W + ccx;
          if (VARIANT == HPP_VARIANT_RS && (nidx < 0 || nidx >= map_size)) valid = 0;  // :674-677
          else {
            fkey = ccost + 30.0 * hmap[nidx];
            if (v >= 0) {
              if (ccost < nodes[v >> 1].cost) replace = 1;  // strictly cheaper than the open node
              else valid = 0;
            }
            ins_slot = sl;
          }
        }
      }
      const unsigned vm = __ballot_sync(0xffffffffu, valid);
      const int n_new = __popc(vm);
      if (n_nodes + n_new > pc.max_nodes) {
        status = 1;
        break;
      }
      if (valid) {
        const int my = n_nodes + __popc(vm & ((1u << lane) - 1u));
        pnode nd;
        nd.x = cxs; nd.y = cys; nd.h = chs; nd.cost = ccost; nd.steer = pc.ec.steer[lane]; nd.parent = cur;
        nd.dir = (signed char)(int)pc.ec.dir[lane]; nd.cmd = (signed char)lane; nd.pad = 0;
        nodes[my] = nd;
        __threadfence_block();
        if (replace) table[ins_slot] = my << 1;
        else {
          int sl = ins_slot;
          while (atomicCAS(&table[sl], -1, my << 1) != -1) sl = (sl + 1) & pc.table_mask;
        }
      }
      __syncwarp();
      // heap pushes in command order (one lane; keys fetched with shuffles)
      for (unsigned mset = vm; mset;) {
        const int b = __ffs(mset) - 1;
        mset &= mset - 1;
        const double k = __shfl_sync(0xffffffffu, fkey, b);
        const int my = n_nodes + __popc(vm & ((1u << b) - 1u));
        if (lane == 0) gheap_push(hkey, hidx, hn, k, my);
      }
      hn = __shfl_sync(0xffffffffu, hn, 0);
      n_nodes += n_new;
      __syncwarp();
    }

    // ---- back-track (RS/src/HybridAstar.cpp:690-710)
    int total = 1;  // failure: the goal state alone (hazard H10)
    if (goal_node >= 0) {
      total = via_shot ? shot_samples : 0;
      for (int b = goal_node; b >= 0; b = nodes[b].parent) total += (nodes[b].cmd == -1) ? 1 : pc.ec.sim;
    }
    unsigned long long off = 0;
    if (lane == 0) off = atomicAdd(cursor, (unsigned long long)total);
    off = __shfl_sync(0xffffffffu, off, 0);
    const bool fits = off + (unsigned long long)total <= traj_cap;
    if (fits) {
      double *out = traj + off * 5;
      if (goal_node < 0) {
        if (lane == 0) {
          out[0] = gx; out[1] = gy; out[2] = gh; out[3] = (double)g_cx; out[4] = (double)g_cy;
        }
      } else {
        int pos = total;
        if (via_shot) {
          pos -= shot_samples;
          store_visitor sv;
          sv.out = out + (size_t)pos * 5;
          sv.g = &g;
          sv.cap = shot_samples;
          const pnode cn = nodes[goal_node];
          if (VARIANT == HPP_VARIANT_RS) {
            const int k = SR.cand;
            warp_sample_path<false>(S.len[k], S.typ[k], S.nseg[k], pc.maxc, pc.maxc, cn.x, cn.y, cn.h, sv);
          } else {
            warp_sample_path<true>(DB.len, DB.typ, DB.nmode, pc.maxc, 0.1, cn.x, cn.y, cn.h, sv);
          }
        }
        if (lane == 0) {
          for (int b = goal_node; b >= 0; b = nodes[b].parent) {
            const pnode nd = nodes[b];
            if (nd.cmd == -1) {
              pos -= 1;
              double *r = out + (size_t)pos * 5;
              r[0] = nd.x; r[1] = nd.y; r[2] = nd.h;
              r[3] = (double)(int)(hpp_div_res(g, nd.x - g.ox));
              r[4] = (double)(int)(hpp_div_res(g, nd.y - g.oy));
            } else {
              pos -= pc.ec.sim;
              const pnode pr = nodes[nd.parent];
              double x = pr.x, y = pr.y, h = pr.h;
              for (int i = 0; i < pc.ec.sim; i++) {
                vehicle_step(x, y, h, pc.ec.dir[nd.cmd], pc.ec.step, pc.ec.dtheta[nd.cmd]);
                double *r = out + (size_t)(pos + i) * 5;
                r[0] = x; r[1] = y; r[2] = h;
                r[3] = (double)(int)(hpp_div_res(g, x - g.ox));
                r[4] = (double)(int)(hpp_div_res(g, y - g.oy));
              }
            }
          }
        }
      }
    }
    if (lane == 0) {
      hpp_plan P;
      P.n_states = total;
      P.iterations = iterations;
      P.via_shot = via_shot;
      P.status = status;
      P.cost = (goal_node < 0) ? 0.0 : (via_shot ? shot_cost : nodes[goal_node].cost);
      P.traj_offset = off;
      plans[q] = P;
    }
    __syncwarp();
  }
}

extern "C" hpp_status hpp_plan_batch(hpp_map *m, int variant, const double *starts, const double *goals, size_t n,
                                     hpp_plan *plans, double *traj, size_t traj_cap, size_t *traj_used) {
  if (!m || variant < 0 || variant > 1 || (n && (!starts || !goals || !plans)) || (traj_cap && !traj)) return HPP_ERR_ARG;
  if (traj_used) *traj_used = 0;
  if (n == 0) return HPP_OK;
  if (n > 0x7fffffff) return HPP_ERR_ARG;
  hpp_ctx *c = m->ctx;
  HPP_CUDA(c, cudaSetDevice(c->device));
  const int W = m->g.W, H = m->g.H;

  // group the queries by goal cell: one heuristic field per distinct cell
  std::vector<long> cell(n);
  for (size_t q = 0; q < n; q++) {
    const int gx = (int)((goals[3 * q] - m->g.ox) / m->g.res), gy = (int)((goals[3 * q + 1] - m->g.oy) / m->g.res);
    if (gx < 0 || gx >= W || gy < 0 || gy >= H)
      HPP_FAIL(c, HPP_ERR_ARG, "goal of query %zu lies outside the map (cell %d,%d)", q, gx, gy);
    cell[q] = (long)gy * W + gx;
  }
  std::vector<int> order(n);
  for (size_t q = 0; q < n; q++) order[q] = (int)q;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cell[a] < cell[b]; });

  plan_consts pc;
  memset(&pc, 0, sizeof pc);
  pc.g = m->g;
  hpp_make_expand_consts(m, variant, &pc.ec);
  for (int k = 0; k < 4; k++) {
    pc.cc.cx[k] = m->car.cx[k];
    pc.cc.cy[k] = m->car.cy[k];
  }
  pc.cc.cr = m->car.cr;
  for (int k = 0; k < 5; k++) {
    pc.rx[k] = m->car.rx[k];
    pc.ry[k] = m->car.ry[k];
  }
  pc.maxc = m->car.maxc;
  pc.max_steer = m->car.p.max_steer;
  pc.max_nodes = PLAN_ITER_CAP * pc.ec.n_cmd + 2;
  int slots = 1;
  while (slots < 2 * pc.max_nodes) slots <<= 1;
  pc.table_mask = slots - 1;

  // device buffers: inputs, per-warp pools, outputs, heuristic fields for a chunk of goal cells
  size_t free_b = 0, total_b = 0;
  HPP_CUDA(c, cudaMemGetInfo(&free_b, &total_b));
  const size_t per_warp = (size_t)pc.max_nodes * (sizeof(pnode) + 12) + (size_t)slots * 4;
  int occ_blocks = 0;
  if (variant == HPP_VARIANT_RS)
    HPP_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_blocks, plan_kernel<HPP_VARIANT_RS>, PLAN_WARPS * 32, 0));
  else
    HPP_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_blocks, plan_kernel<HPP_VARIANT_DUBINS>, PLAN_WARPS * 32, 0));
  if (occ_blocks < 1) occ_blocks = 1;
  size_t max_warps = (size_t)c->sm_count * occ_blocks * PLAN_WARPS;  // one resident wave of persistent warps
  if (max_warps > n) max_warps = n;
  const size_t field_bytes = m->n * sizeof(double);
  const size_t budget = free_b > ((size_t)2 << 30) ? free_b - ((size_t)1 << 30) : free_b / 2;
  while (max_warps > 1 && max_warps * per_warp > budget / 2) max_warps /= 2;
  size_t fields_cap = (budget - max_warps * per_warp - n * 64 - traj_cap * 40) / field_bytes;
  if (fields_cap < 1) HPP_FAIL(c, HPP_ERR_NOMEM, "not enough device memory for one heuristic field");
  if (fields_cap > 4096) fields_cap = 4096;
  const int grid = (int)((max_warps + PLAN_WARPS - 1) / PLAN_WARPS);
  const size_t n_warps = (size_t)grid * PLAN_WARPS;

  char *d_pool = nullptr, *d_io = nullptr;
  double *d_fields = nullptr;
  const size_t io_bytes = n * 48 + n * 8 + n * sizeof(hpp_plan) + traj_cap * 40 + 64;
  // buffers come from the context's grow-only scratch: a per-call cudaMalloc/cudaFree would synchronise the whole
  // device and serialise plan_batch calls that other contexts (threads) run concurrently on the same GPU
  cudaError_t e = cudaSuccess;
  hpp_status st = hpp_scratch(c, 1, n_warps * per_warp + 256, (void **)&d_pool);
  if (st == HPP_OK) st = hpp_scratch(c, 0, io_bytes, (void **)&d_io);
  if (st != HPP_OK) st = HPP_ERR_NOMEM;
  double *d_starts = nullptr, *d_goals = nullptr, *d_traj = nullptr;
  int *d_qids = nullptr, *d_fq = nullptr;
  hpp_plan *d_plans = nullptr;
  unsigned long long *d_cursor = nullptr;
  std::vector<int> fq(n);
  std::vector<int32_t> chunk_goals;
  size_t fields_alloc = 0;
  if (st == HPP_OK) {
    d_starts = (double *)d_io;
    d_goals = d_starts + 3 * n;
    d_qids = (int *)(d_goals + 3 * n);
    d_fq = d_qids + n;
    d_plans = (hpp_plan *)(d_fq + n);
    d_cursor = (unsigned long long *)(d_plans + n);
    d_traj = (double *)(d_cursor + 2);
    pnode *d_nodes = (pnode *)d_pool;
    double *d_keys = (double *)(d_nodes + n_warps * pc.max_nodes);
    int *d_idx = (int *)(d_keys + n_warps * pc.max_nodes);
    int *d_table = d_idx + n_warps * pc.max_nodes;
    cudaMemcpyAsync(d_starts, starts, n * 24, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(d_goals, goals, n * 24, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(d_qids, order.data(), n * 4, cudaMemcpyHostToDevice, c->stream);
    cudaMemsetAsync(d_cursor, 0, 16, c->stream);
    // chunks of distinct goal cells
    size_t q0 = 0;
    while (q0 < n && st == HPP_OK) {
      chunk_goals.clear();
      size_t q1 = q0;
      long last = -1;
      while (q1 < n) {
        const long cl = cell[order[q1]];
        if (cl != last) {
          if (chunk_goals.size() / 2 >= fields_cap) break;
          chunk_goals.push_back((int32_t)(cl % W));
          chunk_goals.push_back((int32_t)(cl / W));
          last = cl;
        }
        fq[q1] = (int)(chunk_goals.size() / 2 - 1);
        q1++;
      }
      const size_t nf = chunk_goals.size() / 2;
      if (nf > fields_alloc) {
        if (hpp_scratch(c, 3, nf * field_bytes, (void **)&d_fields) != HPP_OK) {
          st = HPP_ERR_NOMEM;
          d_fields = nullptr;
          break;
        }
        fields_alloc = nf;
      }
      st = hpp_heuristic_batch_dev(m, variant, chunk_goals.data(), nf, d_fields, nullptr);
      if (st != HPP_OK) break;
      cudaMemcpyAsync(d_fq + q0, fq.data() + q0, (q1 - q0) * 4, cudaMemcpyHostToDevice, c->stream);
      if (variant == HPP_VARIANT_RS)
        plan_kernel<HPP_VARIANT_RS><<<grid, PLAN_WARPS * 32, 0, c->stream>>>(
            pc, m->d_distance, m->d_occ_bits, m->bits_pitch, d_starts, d_goals, d_qids + q0, (int)(q1 - q0), d_fq + q0,
            d_fields, d_nodes, d_keys, d_idx, d_table, d_plans, d_traj, traj_cap, d_cursor);
      else
        plan_kernel<HPP_VARIANT_DUBINS><<<grid, PLAN_WARPS * 32, 0, c->stream>>>(
            pc, m->d_distance, m->d_occ_bits, m->bits_pitch, d_starts, d_goals, d_qids + q0, (int)(q1 - q0), d_fq + q0,
            d_fields, d_nodes, d_keys, d_idx, d_table, d_plans, d_traj, traj_cap, d_cursor);
      c->launches++;
      e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);  // fq / chunk_goals are reused by the next chunk
      if (e != cudaSuccess) {
        snprintf(c->err, sizeof c->err, "plan kernel failed: %s", cudaGetErrorString(e));
        st = HPP_ERR_CUDA;
      }
      q0 = q1;
    }
  }
  if (st == HPP_OK) {
    unsigned long long used = 0;
    cudaMemcpyAsync(plans, d_plans, n * sizeof(hpp_plan), cudaMemcpyDeviceToHost, c->stream);
    cudaMemcpyAsync(c->h_flags + 48, d_cursor, 8, cudaMemcpyDeviceToHost, c->stream);
    e = cudaStreamSynchronize(c->stream);
    memcpy(&used, c->h_flags + 48, 8);
    const size_t rows = used < traj_cap ? (size_t)used : traj_cap;
    if (e == cudaSuccess && rows) {
      cudaMemcpyAsync(traj, d_traj, rows * 40, cudaMemcpyDeviceToHost, c->stream);
      e = cudaStreamSynchronize(c->stream);
    }
    if (e != cudaSuccess) {
      snprintf(c->err, sizeof c->err, "plan_batch copy-out failed: %s", cudaGetErrorString(e));
      st = HPP_ERR_CUDA;
    } else {
      if (traj_used) *traj_used = (size_t)used;
      if (used > traj_cap) {
        snprintf(c->err, sizeof c->err, "trajectories need %llu rows, caller supplied %zu", used, traj_cap);
        st = HPP_ERR_CAPACITY;
      }
    }
  }
  cudaStreamSynchronize(c->stream);
  return st;
}
