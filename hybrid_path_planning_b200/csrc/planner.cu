// planner.cu -- whole Hybrid A* queries on the device: HybridAstar::run (RS/src/HybridAstar.cpp:553-729,
// DB/src/HybridAstar.cpp:315-496) for a batch of independent start/goal pairs on one map.
//
// Two kernels, one CTA per query (which then takes the next query from a counter): no host round trip per A* iteration.
//   plan_pipe_kernel  Reeds-Shepp planner.  Warp 0 runs the reference's loop statement by statement -- pop, openSet
//                     lookup, closed test, goal test, expansion (lane = motion command), insert / replace in command
//                     order, heap pushes -- ASSUMING that analytic shots fail, and hands every popped node to the worker
//                     warps; each worker evaluates one node's whole shot (46 closed forms, ordered de-duplication,
//                     footprint checks of every candidate path, cost order when one is free); results are retired in
//                     pop order, so the first success ends the query exactly where the sequential loop would have.
//                     The open list itself is kept by the last warp of the CTA: it replays "the pushes of the previous
//                     expansion, then this pop" while the controller expands, and hands back the root the pop left
//                     behind, from which (and from the new keys) the controller knows its next node.
//   plan_kernel       Dubins planner, phase by phase: the shot is a single path (warp 0) and the expansion's SAT checks
//                     (one warp per state) run beside it.
// Per query the CTA keeps, in its own slice of global memory,
//   * a node pool (state, Cost_path, steering angle, direction, parent, command index),
//   * the open list as a binary heap of (f, node) that replays libstdc++'s std::push_heap / std::pop_heap for
//     the reference's cost-only comparator, so equal-f nodes pop in the reference's order,
//   * one open-addressing hash table keyed by the State doubles that plays both openSet and closedSet
//     (value = node index, low bit = "closed").
// The 3500-iteration cap, the 1-state failure trajectory and the back-track order are kept.  Child trajectories are not
// stored: they are re-rolled from (parent state, command) at back-track time.  The heuristic field of every distinct
// (map, goal cell) comes from K3 (hpp_heuristic_batch_dev) beforehand.
#include <algorithm>
#include <chrono>
#include <vector>

#include "expand.cuh"
#include "shots.cuh"

// warps per query: 4 for batches (three CTAs per SM: more queries in flight), 8 for small batches (a lone query's
// iteration is ~11 % shorter with its paths and formulas spread over eight warps; one CTA per SM)
#define PLAN_WARPS_BATCH 4
#define PLAN_WARPS_LONE 8
#define PLAN_ITER_CAP 3500
#ifndef PLAN_PATH_GROUP
#define PLAN_PATH_GROUP 8  // lanes per candidate path in the shot's footprint checks (32 = one path per warp step)
#endif
#define PI_ASTAR 3.14159 /* RS/include/HybridAstar.h:24 */

struct pnode {
  double x, y, h, cost, steer;
  int parent;
  signed char dir;
  signed char cmd;  // -1 = start, -2 = shot node
  short pad;
};

// per-map device pointers of a launch (all maps of one launch have the same geometry and vehicle)
struct plan_map_desc {
  const float *dist;      // grid_map "distance" layer
  const uint32_t *bits;   // bit-packed occupancy rows
  int pitch, pad_;
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

// std::pop_heap by a whole warp: the sift-down is a chain of dependent loads (one heap level per memory round trip
// when one thread walks it), so the 30 descendants of the hole over the next four levels are fetched by the lanes
// together and the walk takes four levels per round trip.  Same comparisons, same moves, same result as gheap_pop.
__device__ __forceinline__ int gheap_pop_warp(double *key, int *idx, int &n) {
  const unsigned FULLM = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int top = 0;
  if (lane == 0) top = idx[0];
  top = __shfl_sync(FULLM, top, 0);
  if (n > 1) {
    const double vk = key[n - 1];
    const int vp = idx[n - 1];
    const int len = n - 1;
    int hole = 0, second = 0;
    double k_above = 0.0;  // key now stored at the hole's parent (= the key that was at the hole)
    int lr = 0;
    for (int t = lane + 2; t > 1; t >>= 1) lr++;          // lane l holds relative level lr (1..4) ...
    const int lj = lane + 2 - (1 << lr);                   // ... position lj of the subtree under the hole
    while (second < (len - 1) / 2) {
      const int base = hole;
      const long node = (((long)base + 1) << lr) - 1 + lj;
      const bool have = lane < 30 && node < len;
      const double kv = have ? key[node] : 0.0;
      const int iv = have ? idx[node] : 0;
      for (int r = 1; r <= 4 && second < (len - 1) / 2; r++) {
        second = 2 * (second + 1);
        const int first = ((base + 1) << r) - 1;           // leftmost node of relative level r
        const int src = (1 << r) - 2 + (second - first);   // lane holding `second`
        const double kr = __shfl_sync(FULLM, kv, src), kl = __shfl_sync(FULLM, kv, src - 1);
        const int ir = __shfl_sync(FULLM, iv, src), il = __shfl_sync(FULLM, iv, src - 1);
        double km = kr;
        int im = ir;
        if (kr > kl) {
          second--;
          km = kl;
          im = il;
        }
        if (lane == 0) {
          key[hole] = km;
          idx[hole] = im;
        }
        k_above = km;
        hole = second;
      }
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
      second = 2 * (second + 1);
      const double km = key[second - 1];
      const int im = idx[second - 1];
      if (lane == 0) {
        key[hole] = km;
        idx[hole] = im;
      }
      k_above = km;
      hole = second - 1;
    }
    __syncwarp();
    // std::__push_heap of the former last element from the hole; its parent's key is known without a load
    if (lane == 0) {
      if (hole > 0 && k_above > vk) {
        int parent = (hole - 1) / 2;
        while (hole > 0 && key[parent] > vk) {
          key[hole] = key[parent];
          idx[hole] = idx[parent];
          hole = parent;
          parent = (hole - 1) / 2;
        }
      }
      key[hole] = vk;
      idx[hole] = vp;
    }
    __syncwarp();
  }
  n--;
  return top;
}

// std::push_heap by a whole warp.  A new f value is usually close to the minimum, so it climbs most of the heap's
// levels; lane t fetches the t-th ancestor of the new slot, a ballot finds where the climb stops (the first ancestor
// whose key is not greater), and the lanes shift the entries down together: one memory round trip per push.
__device__ __forceinline__ void gheap_push_warp(double *key, int *idx, int &n, double k, int p) {
  const unsigned FULLM = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int node = ((n + 1) >> lane) - 1;  // lane 0: the new slot; -1 beyond the root
  const bool anc = lane >= 1 && node >= 0;
  double a = 0.0;
  int ai = 0;
  if (anc) {
    a = key[node];
    ai = idx[node];
  }
  const unsigned gt = __ballot_sync(FULLM, anc && a > k);
  const int moves = __ffs(~(gt >> 1)) - 1;  // leading run of ancestors that move down
  const double a_up = __shfl_down_sync(FULLM, a, 1);
  const int ai_up = __shfl_down_sync(FULLM, ai, 1);
  if (lane < moves) {
    key[node] = a_up;
    idx[node] = ai_up;
  } else if (lane == moves) {
    key[node] = k;
    idx[node] = p;
  }
  n++;
  __syncwarp();
}

// The pushes of one expansion (up to 26, no pop in between) as ONE memory round trip.  gheap_push_warp costs a round
// trip per push (its ancestors are loaded after the previous push has stored), and an expansion's pushes were 5.6 k of
// the 19 k cycles of an iteration on the parking maps.  The ancestors of the new slots n+1 .. n+m overlap almost
// entirely: level L (L-th ancestors) holds the consecutive slots (n+1)>>L .. (n+m)>>L, at most m / 2^L + 2 of them, ~60
// entries in all.  They are loaded into shared memory together, the m sift-ups run there one after the other with the
// same ballot as gheap_push_warp (so the result is the sequential one, std::push_heap after std::push_heap), and
// everything is stored back together.  A batch that crosses a power of two (the new slots are on two depths, so one
// slot would sit on two levels of the cache), a heap below 64 entries or a single push take gheap_push_warp.
#define HEAP_CACHE_LEVELS 19
#define HEAP_CACHE_WIDTH 28
struct heap_cache {
  double key[HEAP_CACHE_LEVELS][HEAP_CACHE_WIDTH];
  int idx[HEAP_CACHE_LEVELS][HEAP_CACHE_WIDTH];
};
// fkey: the key of the lane's entry (lanes of vm, pushed in lane order); payloads first_id, first_id + 1, ...
__device__ __forceinline__ void gheap_push_batch(double *key, int *idx, int &n, heap_cache &hc, double fkey, unsigned vm,
                                                 int first_id) {
  const unsigned FULLM = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int m = __popc(vm);
  const int s1 = n + 1, s2 = n + m;  // 1-based slots of the new entries: the parent of slot s is s >> 1
  const int D = 31 - __clz(s1);      // depth of the new slots = number of ancestor levels
  if (m < 2 || s1 < 64 || 31 - __clz(s2) != D || D >= HEAP_CACHE_LEVELS || m > HEAP_CACHE_WIDTH - 2) {
    int j = 0;
    for (unsigned mset = vm; mset; j++) {
      const int b = __ffs(mset) - 1;
      mset &= mset - 1;
      gheap_push_warp(key, idx, n, __shfl_sync(FULLM, fkey, b), first_id + j);
    }
    return;
  }
  // every lane takes at most one entry of levels 1..4 (at most 14 + 8 + 5 + 3 slots) and one of levels 5..D (at most two
  // slots each): all loads are issued before the first is used, one round trip
  int La = 0, da = 0, Lb = 5 + (lane >> 1), db = lane & 1;
  {
    int rest = lane;
    for (int L = 1; L <= 4 && La == 0; L++) {
      const int cnt = (s2 >> L) - (s1 >> L) + 1;
      if (rest < cnt) {
        La = L;
        da = rest;
      }
      rest -= cnt;
    }
  }
  const bool va = La != 0, vb = Lb <= D && db < (s2 >> Lb) - (s1 >> Lb) + 1;
  const int na_ = va ? (s1 >> La) + da - 1 : 0, nb_ = vb ? (s1 >> Lb) + db - 1 : 0;  // 0-based heap positions
  double ka_ = 0.0, kb_ = 0.0;
  int ia_ = 0, ib_ = 0;
  if (va) {
    ka_ = key[na_];
    ia_ = idx[na_];
  }
  if (vb) {
    kb_ = key[nb_];
    ib_ = idx[nb_];
  }
  if (va) {
    hc.key[La][da] = ka_;
    hc.idx[La][da] = ia_;
  }
  if (vb) {
    hc.key[Lb][db] = kb_;
    hc.idx[Lb][db] = ib_;
  }
  __syncwarp();
  const bool lvl = lane <= D;  // lane t works on level t (0 = the new slot itself)
  int j = 0;
  for (unsigned mset = vm; mset; j++) {
    const int b = __ffs(mset) - 1;
    mset &= mset - 1;
    const double k = __shfl_sync(FULLM, fkey, b);
    const int s = s1 + j;
    const int pos = lvl ? (s >> lane) - (s1 >> lane) : 0;
    double a = 0.0;
    int ai = 0;
    if (lane >= 1 && lvl) {
      a = hc.key[lane][pos];
      ai = hc.idx[lane][pos];
    }
    const unsigned gt = __ballot_sync(FULLM, lane >= 1 && lvl && a > k);
    const int moves = __ffs(~(gt >> 1)) - 1;  // leading run of ancestors that move down
    const double a_up = __shfl_down_sync(FULLM, a, 1);
    const int ai_up = __shfl_down_sync(FULLM, ai, 1);
    if (lane < moves) {
      hc.key[lane][pos] = a_up;
      hc.idx[lane][pos] = ai_up;
    } else if (lane == moves) {
      hc.key[lane][pos] = k;
      hc.idx[lane][pos] = first_id + j;
    }
    __syncwarp();
  }
  if (va) {
    key[na_] = hc.key[La][da];
    idx[na_] = hc.idx[La][da];
  }
  if (vb) {
    key[nb_] = hc.key[Lb][db];
    idx[nb_] = hc.idx[Lb][db];
  }
  if (lane < m) {  // the new slots themselves
    key[n + lane] = hc.key[0][lane];
    idx[n + lane] = hc.idx[0][lane];
  }
  n += m;
  __syncwarp();
}

// what the threads of a query's CTA tell each other (two copies, alternating per iteration, so that one barrier per
// hand-over is enough)
struct plan_mail {
  int hn, tv, slot;         // after the pop
  int n_new, status;        // after the insertions (hn is rewritten too)
  int found, samples;       // Dubins shot (evaluated by warp 0)
  double cost;
};

struct plan_smem {
  plan_mail mail[2];
  int query;
  unsigned hit;                         // expansion: commands with a colliding state
  double child[HPP_MAX_COMMANDS][3];    // expansion: last state of every command's rollout
};

template <int VARIANT, int PLAN_WARPS>
__global__ void __launch_bounds__(PLAN_WARPS * 32) plan_kernel(plan_consts pc, const plan_map_desc *__restrict__ descs,
                                                            const int *__restrict__ map_of_query,
                                                            const double *__restrict__ starts,
                                                            const double *__restrict__ goals,
                                                            const int *__restrict__ query_ids, int n_queries,
                                                            const int *__restrict__ field_of_query,
                                                            const double *__restrict__ fields,
                                                            pnode *__restrict__ node_pool, double *__restrict__ key_pool,
                                                            int *__restrict__ idx_pool, int *__restrict__ table_pool,
                                                            hpp_plan *__restrict__ plans, double *__restrict__ traj,
                                                            unsigned long long traj_cap,
                                                            unsigned long long *__restrict__ cursor,
                                                            unsigned long long *__restrict__ prof) {
  static_assert(PLAN_WARPS >= 2, "warp 0 and at least one helper warp");
  static_assert(VARIANT == HPP_VARIANT_DUBINS, "the Reeds-Shepp planner is plan_pipe_kernel");
  constexpr int PLAN_THREADS = PLAN_WARPS * 32;
  __shared__ plan_smem sm;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const hpp_geom &g = pc.g;
  const int W = g.W, H = g.H;
  const long map_size = (long)W * H;
  pnode *nodes = node_pool + (size_t)blockIdx.x * pc.max_nodes;
  double *hkey = key_pool + (size_t)blockIdx.x * pc.max_nodes;
  int *hidx = idx_pool + (size_t)blockIdx.x * pc.max_nodes;
  int *table = table_pool + (size_t)blockIdx.x * (pc.table_mask + 1);
  const int ncmd = pc.ec.n_cmd, n_poses = pc.ec.n_cmd * pc.ec.sim;

#ifdef HPP_PLAN_PROF
  unsigned long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tk = 0;
#define PLAN_MARK(i) { const long long now_ = clock64(); pf[i] += (unsigned long long)(now_ - tk); tk = now_; }
#else
#define PLAN_MARK(i)
#endif
  // queries are handed out one at a time (cursor[1]): the few long ones (the 3500-iteration failures) then start
  // early on whichever CTA is free instead of queueing behind a fixed share of the batch
  int *next_query = (int *)(cursor + 1);
  for (;;) {
    __syncthreads();  // the previous query is finished by every warp
    if (tid == 0) sm.query = atomicAdd(next_query, 1);
    __syncthreads();
    const int qi = sm.query;
    if (qi >= n_queries) break;
    const int q = query_ids[qi];
    const double *hmap = fields + (size_t)field_of_query[qi] * map_size;
    // the query's map (maps of one launch share their geometry and vehicle, hpp_plan_batch_multi)
    const plan_map_desc md = descs[map_of_query ? map_of_query[q] : 0];
    const uint32_t *__restrict__ bits = md.bits;
    const int pitch = md.pitch;
    const double sx = starts[3 * q], sy = starts[3 * q + 1], sh = starts[3 * q + 2];
    const double gx = goals[3 * q], gy = goals[3 * q + 1], gh = goals[3 * q + 2];
    const int s_cx = (int)(hpp_div_res(g, sx - g.ox)), s_cy = (int)(hpp_div_res(g, sy - g.oy));
    const int g_cx = (int)(hpp_div_res(g, gx - g.ox)), g_cy = (int)(hpp_div_res(g, gy - g.oy));
    for (int i = tid; i <= pc.table_mask; i += PLAN_THREADS) table[i] = -1;
    __syncthreads();
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
      if (tid == 0) {
        pnode nd;
        nd.x = sx; nd.y = sy; nd.h = sh; nd.cost = 0.0; nd.steer = 0.0; nd.parent = -1; nd.dir = 1; nd.cmd = -1; nd.pad = 0;
        nodes[0] = nd;
        int slot;
        table_find(table, nodes, pc.table_mask, sx, sy, sh, slot);
        table[slot] = 0;  // node 0, open
        gheap_push(hkey, hidx, hn, 0.0 + 30.0 * hmap[start_index], 0);
      }
      n_nodes = 1;
      hn = 1;
    }

    while (status == 0 && hn > 0) {
      iterations++;
      if (iterations > PLAN_ITER_CAP) break;
      plan_mail &mail = sm.mail[iterations & 1];
#ifdef HPP_PLAN_PROF
      tk = clock64();
#endif
      // ---- pop (warp 0)
      if (wid == 0) {
        const int top = gheap_pop_warp(hkey, hidx, hn);
        if (lane == 0) {
          int slot = 0;
          const pnode &tn = nodes[top];
          const int tv = table_find(table, nodes, pc.table_mask, tn.x, tn.y, tn.h, slot);
          if (tv >= 0 && !(tv & 1)) table[slot] = tv | 1;  // openSet.erase + closedSet[currentState] = currentNode
          mail.hn = hn;
          mail.tv = tv;
          sm.hit = 0u;
        }
      }
      __syncthreads();
      hn = mail.hn;
      const int tv = mail.tv;
      if (tv < 0 || (tv & 1)) continue;  // already in the closed set
      const int cur = tv >> 1;           // openSet[currentState] (may be a cheaper replacement of the popped node)
      const pnode cn = nodes[cur];
      PLAN_MARK(0)

      // ---- analytic shot on warp 0, the expansion's footprint checks on the other warps meanwhile
      {
        if (wid == 0) {
          DB = db_plan(cn.x, cn.y, cn.h, gx, gy, gh, pc.maxc);
          sat_visitor vis;
          vis.g = &g; vis.rx = pc.rx; vis.ry = pc.ry; vis.bits = bits; vis.pitch = pitch; vis.hit_zero_idx = 0x7fffffff;
          const int n_keep = warp_sample_path<true>(DB.len, DB.typ, DB.nmode, pc.maxc, 0.1, cn.x, cn.y, cn.h, vis);
          const bool collides = (n_keep < 0) || (vis.hit_zero_idx < n_keep);
          const int found = (n_keep > 0 && !collides);
          double cost = cn.cost;
          cost += fabs(DB.best_cost);
          for (int i = 0; i < 3; i++)
            if (DB.typ[i] != 'S') cost += pc.max_steer * 15.0;
          if (lane == 0) {
            mail.found = found;
            mail.samples = n_keep;
            mail.cost = cost;
          }
        }
        if (wid != 0 || PLAN_WARPS == 1) {
          // GetnextNeighbours' SAT checks (DB/src/HybridAstar.cpp:23-79), one pose per warp at a time
          const int w0 = PLAN_WARPS == 1 ? 0 : wid - 1, ws = PLAN_WARPS == 1 ? 1 : PLAN_WARPS - 1;
          for (int p = w0; p < n_poses; p += ws) {
            const int c = p / pc.ec.sim, i = p % pc.ec.sim;
            if ((sm.hit >> c) & 1u) continue;  // an earlier state of this command already collides
            double x = cn.x, y = cn.y, h = cn.h;
            for (int k = 0; k <= i; k++) vehicle_step(x, y, h, pc.ec.dir[c], pc.ec.step, pc.ec.dtheta[c]);
            const int hit = sat_collide_warp(g, pc.rx, pc.ry, bits, pitch, x, y, h);
            if (lane == 0) {
              if (hit) atomicOr(&sm.hit, 1u << c);
              if (i == pc.ec.sim - 1) {
                sm.child[c][0] = x; sm.child[c][1] = y; sm.child[c][2] = h;
              }
            }
          }
        }
        __syncthreads();
        SR.found = mail.found;
        if (SR.found) {
          shot_samples = mail.samples;
          shot_cost = mail.cost;
        }
      }
      PLAN_MARK(2)
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

      PLAN_MARK(3)

      // ---- insert / replace in command order (warp 0: lane = command)
      if (wid == 0) {
        double cxs = 0, cys = 0, chs = 0, ccost = 0;
        int valid = 0;
        if (lane < ncmd) {
          valid = !((sm.hit >> lane) & 1u);
          cxs = sm.child[lane][0]; cys = sm.child[lane][1]; chs = sm.child[lane][2];
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
            {  // (the Dubins planner does not range-check the cell index, DB/src/HybridAstar.cpp:442-444)
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
        int st = 0;
        if (n_nodes + n_new > pc.max_nodes) {
          st = 1;
        } else {
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
          PLAN_MARK(4)
          // heap pushes in command order (the whole warp per push: see gheap_push_warp)
          for (unsigned mset = vm; mset;) {
            const int b = __ffs(mset) - 1;
            mset &= mset - 1;
            const double k = __shfl_sync(0xffffffffu, fkey, b);
            const int my = n_nodes + __popc(vm & ((1u << b) - 1u));
            gheap_push_warp(hkey, hidx, hn, k, my);
          }
        }
        if (lane == 0) {
          mail.hn = hn;
          mail.n_new = n_new;
          mail.status = st;
        }
      }
      __syncthreads();
      hn = mail.hn;
      n_nodes += mail.n_new;
      status = mail.status;
      PLAN_MARK(5)
    }
#ifdef HPP_PLAN_PROF
    if (prof && tid == 0 && iterations > 3000) {
      for (int i = 0; i < 7; i++) atomicAdd(prof + i, pf[i]);
      atomicAdd(prof + 7, (unsigned long long)iterations);
    }
    for (int i = 0; i < 8; i++) pf[i] = 0;
#endif
    if (wid != 0) continue;  // the back-track is warp 0's

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
          warp_sample_path<true>(DB.len, DB.typ, DB.nmode, pc.maxc, 0.1, cn.x, cn.y, cn.h, sv);
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

// ---------------------------------------------------------------------------------------------- pipelined RS planner
// plan_kernel runs the phases of an iteration one after the other, and three quarters of an iteration of a long query is
// the analytic shot (46 closed forms, ordered de-duplication, footprint checks of every candidate path).  But the shot of
// a node is a pure function of the node's state, the goal and the map -- it does not look at the open or closed set --
// and as long as shots FAIL the search does not look at them either.  So here the search runs ahead of its shots:
//   warp 0 (the controller) runs HybridAstar::run's loop statement by statement -- pop, closed test, goal test,
//       expansion, insert / replace, heap pushes -- assuming that every shot fails, and hands every popped node to the
//       workers through a ring in shared memory;
//   warps 1 .. 3 roll out the expansion's 26 x 4 poses together with the controller (one thread per pose, a named
//       barrier before and after), because the expansion is on the controller's chain;
//   warps 4 .. (the workers) each take the next node of the ring and evaluate its whole shot alone;
//   the controller retires the results IN POP ORDER: the first success ends the query at that node's iteration --
//       exactly where the sequential loop would have ended, whatever the controller did beyond it (its later pops only
//       appended nodes; the successful node's ancestors are untouched) -- and a loop that ends by the goal test, an empty
//       open list, the iteration cap or an exhausted pool first waits for the shots of all earlier pops.
// Results (iterations, via_shot, cost, trajectory) are those of plan_kernel; the chain of a long query shrinks from
// pop + shot + expansion + insert + push to pop + expansion + insert + push.
#define PIPE_RING 32
struct pipe_job {
  double x, y, h, cost;
  int cur, iter;
};
struct pipe_res {
  int state;  // 0 pending, 1 shot failed, 2 shot succeeded (payload below)
  int n_samples, nseg, pad_;
  double cost;
  double len[5];
  unsigned char typ[8];
};
struct pipe_worker {
  rs_warp_smem S;
  int verdict[RS_NCAND];
  int next_path, any_free;
};
#define PIPE_EXP_WARPS 3  // warps 1..3 roll out the expansion with the controller (4 x 32 threads >= 26 commands x 4 states)
template <int NW>
struct pipe_smem {
  int query;
  volatile int head;   // jobs published by the controller
  int claim;           // jobs claimed by the workers
  volatile int quit;
  volatile int exp_quit;               // expansion helpers: the query is over
  unsigned exp_hit;                    // expansion: commands with a colliding state
  double exp_from[3];                  // expansion: the node being expanded
  double exp_child[HPP_MAX_COMMANDS][3];  // expansion: last state of every command's rollout
  double exp_sc[HPP_MAX_COMMANDS][HPP_MAX_SIM][2];  // expansion: (m dt cos h_i, m dt sin h_i) of every command's step i
  pipe_job job[PIPE_RING];
  pipe_res res[PIPE_RING];
  pipe_worker w[NW - 2 - PIPE_EXP_WARPS];
  heap_cache hcache;                   // heap warp: the ancestors of an expansion's pushes (gheap_push_batch)
  // controller -> heap warp: command c (parity c & 1) = "push these keys, then pop"; heap warp -> controller: the root
  // the pop left behind
  int hm_m[2], hm_first[2], hm_top[2];  // (hm_top: the node the controller took for this pop -- checked by the heap warp)
  double hm_keys[2][HPP_MAX_COMMANDS];
  double hm_root_key;
  int hm_root_idx;                     // -1: the heap is empty
  volatile int hm_quit;
};
// named barrier of the controller and its expansion helpers (barrier 0 is __syncthreads)
__device__ __forceinline__ void pipe_exp_barrier() {
  asm volatile("bar.sync 1, %0;" ::"n"((1 + PIPE_EXP_WARPS) * 32) : "memory");
}
// producer / consumer barriers of the controller and the heap warp (64 threads): 2 = "a command is posted" (the
// controller arrives, the heap warp waits), 3 = "its root is posted" (the heap warp arrives, the controller waits)
__device__ __forceinline__ void pipe_heap_post(int id) {
  __threadfence_block();
  if (id == 2) asm volatile("bar.arrive 2, 64;" ::: "memory");
  else asm volatile("bar.arrive 3, 64;" ::: "memory");
}
__device__ __forceinline__ void pipe_heap_wait(int id) {
  if (id == 2) asm volatile("bar.sync 2, 64;" ::: "memory");
  else asm volatile("bar.sync 3, 64;" ::: "memory");
}

template <int NW, int MIN_CTAS>
__global__ void __launch_bounds__(NW * 32, MIN_CTAS) plan_pipe_kernel(plan_consts pc, const plan_map_desc *__restrict__ descs,
                                                            const int *__restrict__ map_of_query,
                                                            const double *__restrict__ starts,
                                                            const double *__restrict__ goals,
                                                            const int *__restrict__ query_ids, int n_queries,
                                                            const int *__restrict__ field_of_query,
                                                            const double *__restrict__ fields,
                                                            pnode *__restrict__ node_pool, double *__restrict__ key_pool,
                                                            int *__restrict__ idx_pool, int *__restrict__ table_pool,
                                                            hpp_plan *__restrict__ plans, double *__restrict__ traj,
                                                            unsigned long long traj_cap,
                                                            unsigned long long *__restrict__ cursor, int iter_budget,
                                                            const int *__restrict__ qi_list, const int *__restrict__ n_list,
                                                            int *__restrict__ deferred, int *__restrict__ n_deferred,
                                                            int check_order,
                                                            unsigned long long *__restrict__ prof) {
  constexpr int THREADS = NW * 32;
#ifdef HPP_PLAN_PROF
  unsigned long long pf[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define PIPE_T0 const long long t0_ = clock64();
#define PIPE_ADD(i) pf[i] += (unsigned long long)(clock64() - t0_);
#else
#define PIPE_T0
#define PIPE_ADD(i)
#endif
  extern __shared__ __align__(16) unsigned char pipe_raw[];
  pipe_smem<NW> &sm = *reinterpret_cast<pipe_smem<NW> *>(pipe_raw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const unsigned FULLM = 0xffffffffu;
  const hpp_geom &g = pc.g;
  const int W = g.W;
  const long map_size = (long)W * g.H;
  pnode *nodes = node_pool + (size_t)blockIdx.x * pc.max_nodes;
  double *hkey = key_pool + (size_t)blockIdx.x * pc.max_nodes;
  int *hidx = idx_pool + (size_t)blockIdx.x * pc.max_nodes;
  int *table = table_pool + (size_t)blockIdx.x * (pc.table_mask + 1);
  const int ncmd = pc.ec.n_cmd;
  int *next_query = (int *)(cursor + 1);
  for (;;) {
    __syncthreads();  // the previous query is finished by every warp
    if (tid == 0) {
      sm.query = atomicAdd(next_query, 1);
      sm.head = 0;
      sm.claim = 0;
      sm.quit = 0;
      sm.exp_quit = 0;
      sm.hm_quit = 0;
    }
    __syncthreads();
    // second pass of a two-pass batch: the queries the first pass set aside (qi_list, *n_list of them)
    if (sm.query >= (n_list ? *n_list : n_queries)) break;
    const int qi = qi_list ? qi_list[sm.query] : sm.query;
    const int q = query_ids[qi];
    const double *hmap = fields + (size_t)field_of_query[qi] * map_size;
    const plan_map_desc md = descs[map_of_query ? map_of_query[q] : 0];
    const float *__restrict__ dist = md.dist;
    const double gx = goals[3 * q], gy = goals[3 * q + 1], gh = goals[3 * q + 2];
    for (int i = tid; i <= pc.table_mask; i += THREADS) table[i] = -1;
    __syncthreads();

    // one thread per (command, state of its rollout) of the node in sm.exp_from (GetnextNeighbours,
    // RS/src/HybridAstar.cpp:25-79); run by the controller and its helpers between two named barriers
    auto rollouts = [&]() {
      // tid < 128 here; pathLength 4 (the default) gives 104 poses, pathLength 8 gives 208: a second round.
      // CarData's Euler step (vehicle_step) is x += m dt cos h, y += m dt sin h, h += m dt dtheta: the headings of a
      // command's steps do not depend on the positions, so thread (c, i) takes the sine and cosine of step i alone
      // (one libm call instead of i + 1 in a row) and the positions are the same sums, in the same order, afterwards.
      const int np = ncmd * pc.ec.sim;
      for (int p = tid; p < np; p += (1 + PIPE_EXP_WARPS) * 32) {
        const int c = p / pc.ec.sim, i = p % pc.ec.sim;
        const double m = pc.ec.dir[c], dt = pc.ec.step;
        double h = sm.exp_from[2];
        for (int k = 0; k < i; k++) h = h + (m * dt * pc.ec.dtheta[c]);
        double sn, cs;
        hpp_sincos(h, &sn, &cs);
        sm.exp_sc[c][i][0] = m * dt * cs;
        sm.exp_sc[c][i][1] = m * dt * sn;
      }
      pipe_exp_barrier();
      for (int p = tid; p < np; p += (1 + PIPE_EXP_WARPS) * 32) {
        const int c = p / pc.ec.sim, i = p % pc.ec.sim;
        const double m = pc.ec.dir[c], dt = pc.ec.step;
        double x = sm.exp_from[0], y = sm.exp_from[1], h = sm.exp_from[2];
        for (int k = 0; k <= i; k++) {
          x = x + sm.exp_sc[c][k][0];
          y = y + sm.exp_sc[c][k][1];
          h = h + (m * dt * pc.ec.dtheta[c]);
        }
        if (circles_collide_ilp(g, pc.cc, dist, x, y, h)) atomicOr(&sm.exp_hit, 1u << c);
        if (i == pc.ec.sim - 1) {
          sm.exp_child[c][0] = x; sm.exp_child[c][1] = y; sm.exp_child[c][2] = h;
        }
      }
    };
    if (wid >= 1 && wid <= PIPE_EXP_WARPS) {
      // ================================================================ expansion helper
      for (;;) {
        pipe_exp_barrier();  // the controller has published a node (or the end of the query)
        if (sm.exp_quit) break;
        rollouts();
        pipe_exp_barrier();  // hit bits and child states are complete
      }
      continue;
    }
    if (wid == NW - 1) {
      // ================================================================ heap warp: the open list's priority queue
      // std::priority_queue's arrays are touched by this warp alone once the search runs.  A command is the pushes of
      // the previous expansion (in command order) followed by the pop of the node the controller is already working
      // on; the answer is the root that pop leaves behind, which is all the controller needs to know its next node
      // (see the controller's loop).  The sift-down and the sift-ups overlap the controller's expansion and insertion.
      int n = 1;  // the start node, pushed by the controller before its first command
      for (int seq = 0;; seq++) {
        pipe_heap_wait(2);
        if (sm.hm_quit) break;
        const int par = seq & 1;
        const int m = sm.hm_m[par];
        if (m > 0) {
          const double fk = lane < m ? sm.hm_keys[par][lane] : 0.0;
          gheap_push_batch(hkey, hidx, n, sm.hcache, fk, m >= 32 ? 0xffffffffu : ((1u << m) - 1u), sm.hm_first[par]);
        }
        // The node the controller went on with must be the one std::pop_heap takes now; the host turns a mismatch into
        // an error (HPP_PLAN_CHECK=0 skips the check; measured cost: none, the heap warp has slack).
        if (check_order && lane == 0 && hidx[0] != sm.hm_top[par]) atomicAdd(cursor + 2, 1ull);
        gheap_pop_warp(hkey, hidx, n);
        if (lane == 0) {
          sm.hm_root_idx = n > 0 ? hidx[0] : -1;
          sm.hm_root_key = n > 0 ? hkey[0] : 0.0;
        }
        pipe_heap_post(3);
      }
      continue;
    }
    if (wid != 0) {
      // ================================================================ worker: one whole shot per job
      pipe_worker &me = sm.w[wid - 1 - PIPE_EXP_WARPS];
      for (;;) {
        int c = 0;
        if (lane == 0) {
          c = atomicAdd(&sm.claim, 1);
          while (c >= sm.head && !sm.quit) __nanosleep(100);
          if (sm.quit) c = -1;  // the query is over (every shot that mattered has been retired)
        }
        c = __shfl_sync(FULLM, c, 0);
        if (c < 0) break;
        __threadfence_block();
        const pipe_job jb = sm.job[c % PIPE_RING];
        {
          PIPE_T0
          rs_generate(me.S, jb.x, jb.y, jb.h, gx, gy, gh, pc.maxc);
          PIPE_ADD(0)
        }
        const int n_acc = me.S.n_acc;
        rs_shot_result SR;
        SR.found = 0; SR.chosen = -1; SR.n_samples = 0; SR.cand = -1; SR.cost = 0.0; SR.n_paths = n_acc;
        if (n_acc > 0) {
          int hit0 = 0;
          // sample 0 of every candidate path is the start pose itself (x, y unchanged, pi_2_pi(yaw)): looked at once
          if (lane == 0) {
            hit0 = circles_collide_ilp(g, pc.cc, dist, jb.x, jb.y, pi_2_pi(0.0 + jb.h));
            me.next_path = 0;
            me.any_free = 0;
          }
          hit0 = __shfl_sync(FULLM, hit0, 0);
          __syncwarp();
          {
            PIPE_T0
            rs_check_paths_grouped<PLAN_PATH_GROUP>(me.S, n_acc, &me.next_path, me.verdict, &me.any_free, g, pc.cc, dist,
                                                    jb.x, jb.y, jb.h, pc.maxc, hit0);
            PIPE_ADD(1)
          }
          __syncwarp();
#ifdef HPP_PLAN_PROF
          pf[2] += 1;
          pf[3] += (unsigned long long)n_acc;
#endif
          if (me.any_free) {
            rs_order_paths(me.S, jb.cost, pc.maxc, pc.max_steer);
            double best = INFINITY;
            for (int o = 0; o < n_acc; o++)
              if (rs_fold_path(me.S, o, me.verdict[me.S.order[o]], SR, best)) break;
          }
        }
        if (lane == 0) {
          pipe_res &r = sm.res[c % PIPE_RING];
          if (SR.found) {
            const int k = SR.cand;
            r.n_samples = SR.n_samples;
            r.nseg = me.S.nseg[k];
            r.cost = SR.cost;
            for (int i = 0; i < 5; i++) {
              r.len[i] = me.S.len[k][i];
              r.typ[i] = me.S.typ[k][i];
            }
          }
          __threadfence_block();
          *(volatile int *)&r.state = SR.found ? 2 : 1;
        }
        __syncwarp();
      }
#ifdef HPP_PLAN_PROF
      if (prof && lane == 0 && wid == 1 + PIPE_EXP_WARPS) {
        for (int i = 0; i < 4; i++) atomicAdd(prof + i, pf[i]);
      }
      for (int i = 0; i < 8; i++) pf[i] = 0;
#endif
      continue;  // next query (the controller does the back-track)
    }

    // ================================================================== controller (warp 0)
    const double sx = starts[3 * q], sy = starts[3 * q + 1], sh = starts[3 * q + 2];
    const int s_cx = (int)(hpp_div_res(g, sx - g.ox)), s_cy = (int)(hpp_div_res(g, sy - g.oy));
    const int g_cx = (int)(hpp_div_res(g, gx - g.ox)), g_cy = (int)(hpp_div_res(g, gy - g.oy));
    int n_nodes = 0, hn = 0, iterations = 0, status = 0;
    int goal_node = -1, via_shot = 0, shot_samples = 0;
    double shot_cost = 0.0;
    int head = 0, retired = 0, win = -1;  // win: ring index of the first successful shot
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
        table[slot] = 0;
        gheap_push(hkey, hidx, hn, 0.0 + 30.0 * hmap[start_index], 0);
      }
      n_nodes = 1;
      hn = 1;
      __syncwarp();
    }
    // retires finished shots in pop order; returns true when one of them succeeded (win = its job index)
    auto retire = [&]() {
      for (;;) {
        int st = 0;
        if (lane == 0 && retired < head) st = *(volatile int *)&sm.res[retired % PIPE_RING].state;
        st = __shfl_sync(FULLM, st, 0);
        if (st == 0) return false;
        if (st == 2) {
          win = retired;
          return true;
        }
        retired++;
      }
    };
    int end_node = -1;  // node at which the loop itself ended by the goal test
    bool set_aside = false;
#ifdef HPP_PLAN_PROF
    const long long tloop_ = clock64();
#endif
    // The open list lives with the heap warp; the controller keeps what it needs to go on without waiting for it:
    // `count` entries, and the payload of the top.  After "push k_1 .. k_m, pop" the next top follows from the root R the
    // pop left behind and the keys of THIS expansion's pushes alone: std::push_heap moves a new key up past every
    // ancestor that is strictly greater, and every ancestor is >= the root, so a key below the root's becomes the root
    // and any other key leaves it alone -- the top after the pushes is R, unless some key is strictly smaller, then the
    // first of the smallest keys (an equal later one does not displace it).  The heap warp replays the pushes and the
    // pop in the reference's order meanwhile, so the arrays -- and every tie among equal keys -- are std::priority_queue's.
    int count = hn, top_idx = 0, pend_m = 0, pend_first = 0, par = 0;
    while (status == 0 && count > 0) {
      if (retire()) break;
      iterations++;
      if (iterations > PLAN_ITER_CAP) break;
      if (iterations > iter_budget) {  // a long query: set aside for the second pass (unless an earlier shot succeeds)
        set_aside = true;
        break;
      }
      // ---- pop: the heap warp gets the previous expansion's pushes and this pop; the node is known already
#ifdef HPP_PLAN_PROF
      const long long tp_ = clock64();
#endif
      if (lane == 0) {
        sm.hm_m[par] = pend_m;
        sm.hm_first[par] = pend_first;
        sm.hm_top[par] = top_idx;
      }
      pipe_heap_post(2);
      par ^= 1;
      pend_m = 0;
      count--;
      const int top = top_idx;
      int tv = 0;
      if (lane == 0) {
        int slot = 0;
        const pnode &tn = nodes[top];
        tv = table_find(table, nodes, pc.table_mask, tn.x, tn.y, tn.h, slot);
        if (tv >= 0 && !(tv & 1)) table[slot] = tv | 1;  // openSet.erase + closedSet[currentState] = currentNode
      }
      tv = __shfl_sync(FULLM, tv, 0);
#ifdef HPP_PLAN_PROF
      pf[8] += (unsigned long long)(clock64() - tp_);
#endif
      bool leave = false;
      unsigned vm = 0u;
      double fkey = 0.0;
      int n_new = 0;
      do {
        if (tv < 0 || (tv & 1)) break;  // already in the closed set
        const int cur = tv >> 1;
        const pnode cn = nodes[cur];
        // ---- the analytic shot of this node goes to the workers
        {
          PIPE_T0
          while (head - retired >= PIPE_RING) {
            if (retire()) break;
            __nanosleep(100);
          }
          PIPE_ADD(5)
        }
        if (win >= 0) {
          leave = true;
          break;
        }
        if (lane == 0) {
          pipe_job &jb = sm.job[head % PIPE_RING];
          jb.x = cn.x; jb.y = cn.y; jb.h = cn.h; jb.cost = cn.cost; jb.cur = cur; jb.iter = iterations;
          *(volatile int *)&sm.res[head % PIPE_RING].state = 0;
          __threadfence_block();
          sm.head = head + 1;
        }
        head++;
        __syncwarp();
        // ---- goal test (RS/src/HybridAstar.cpp:641-655)
        {
          double hd = fabs(cn.h - gh);
          if (hd > PI_ASTAR) hd = fabs(hd - 2 * PI_ASTAR);
          const int c_cx = (int)(hpp_div_res(g, cn.x - g.ox)), c_cy = (int)(hpp_div_res(g, cn.y - g.oy));
          if (c_cx == g_cx && c_cy == g_cy && hd <= 0.05) {
            end_node = cur;
            leave = true;
            break;
          }
        }
        // ---- expansion: the controller and its helpers, one thread per (command, state)
        if (lane == 0) {
          sm.exp_from[0] = cn.x; sm.exp_from[1] = cn.y; sm.exp_from[2] = cn.h;
          sm.exp_hit = 0u;
        }
#ifdef HPP_PLAN_PROF
        const long long te_ = clock64();
#endif
        pipe_exp_barrier();
        rollouts();
        pipe_exp_barrier();
#ifdef HPP_PLAN_PROF
        const long long ti_ = clock64();
        pf[9] += (unsigned long long)(ti_ - te_);
#endif
        double cxs = 0, cys = 0, chs = 0, ccost = 0;
        int valid = 0, ins_slot = -1, replace = 0;
        if (lane < ncmd) {
          valid = !((sm.exp_hit >> lane) & 1u);
          cxs = sm.exp_child[lane][0]; cys = sm.exp_child[lane][1]; chs = sm.exp_child[lane][2];
        }
        // ---- insert / replace in command order
        if (valid) {
          ccost = path_cost(pc.ec, lane, cn.cost, cn.steer, cn.dir);
          // (the heuristic of the child's cell is fetched before the table is searched: one round trip less on the chain)
          const int ccx = (int)(hpp_div_res(g, cxs - g.ox)), ccy = (int)(hpp_div_res(g, cys - g.oy));
          const long nidx = (long)ccy * W + ccx;
          const bool in_map = !(nidx < 0 || nidx >= map_size);
          const double hval = in_map ? hmap[nidx] : 0.0;
          int sl;
          const int v = table_find(table, nodes, pc.table_mask, cxs, cys, chs, sl);
          if (v >= 0 && (v & 1)) valid = 0;  // in the closed set
          else {
            if (!in_map) valid = 0;  // :674-677
            else {
              fkey = ccost + 30.0 * hval;
              if (v >= 0) {
                if (ccost < nodes[v >> 1].cost) replace = 1;  // strictly cheaper than the open node
                else valid = 0;
              }
              ins_slot = sl;
            }
          }
        }
        vm = __ballot_sync(FULLM, valid);
        n_new = __popc(vm);
        if (n_nodes + n_new > pc.max_nodes) {
          status = 1;
          vm = 0u;
          n_new = 0;
          leave = true;
          break;
        }
        if (valid) {
          const int rank = __popc(vm & ((1u << lane) - 1u));
          const int my = n_nodes + rank;
          pnode nd;
          nd.x = cxs; nd.y = cys; nd.h = chs; nd.cost = ccost; nd.steer = pc.ec.steer[lane]; nd.parent = cur;
          nd.dir = (signed char)(int)pc.ec.dir[lane]; nd.cmd = (signed char)lane; nd.pad = 0;
          nodes[my] = nd;
          sm.hm_keys[par][rank] = fkey;  // the pushes of the next command, in command order
          __threadfence_block();
          if (replace) table[ins_slot] = my << 1;
          else {
            int sl = ins_slot;
            while (atomicCAS(&table[sl], -1, my << 1) != -1) sl = (sl + 1) & pc.table_mask;
          }
        }
        __syncwarp();
#ifdef HPP_PLAN_PROF
        pf[10] += (unsigned long long)(clock64() - ti_);
#endif
      } while (false);
      // ---- the root this pop left behind, and with it the next top
#ifdef HPP_PLAN_PROF
      const long long tq_ = clock64();
#endif
      pipe_heap_wait(3);
#ifdef HPP_PLAN_PROF
      pf[11] += (unsigned long long)(clock64() - tq_);
#endif
      if (leave) break;
      const int r_idx = sm.hm_root_idx;
      const double r_key = sm.hm_root_key;
      top_idx = r_idx;
      if (n_new > 0) {
        double kmin = (vm >> lane) & 1u ? fkey : INFINITY;
        for (int o = 16; o > 0; o >>= 1) kmin = fmin(kmin, __shfl_xor_sync(FULLM, kmin, o));
        if (r_idx < 0 || kmin < r_key) {
          const unsigned eq = __ballot_sync(FULLM, ((vm >> lane) & 1u) && fkey == kmin);
          const int b = __ffs(eq) - 1;
          top_idx = n_nodes + __popc(vm & ((1u << b) - 1u));
        }
      }
      pend_m = n_new;
      pend_first = n_nodes;
      n_nodes += n_new;
      count += n_new;
    }
#ifdef HPP_PLAN_PROF
    pf[4] += (unsigned long long)(clock64() - tloop_);
#endif
    if (lane == 0) {
      sm.exp_quit = 1;
      sm.hm_quit = 1;
    }
    pipe_exp_barrier();  // releases the expansion helpers
    pipe_heap_post(2);   // and the heap warp
    // ---- the loop is over: shots of earlier pops come first
    {
      PIPE_T0
      while (win < 0 && retired < head) {
        if (retire()) break;
        __nanosleep(100);
      }
      PIPE_ADD(6)
    }
#ifdef HPP_PLAN_PROF
    if (prof && lane == 0 && iterations > 3000) {
      atomicAdd(prof + 4, pf[4]);
      atomicAdd(prof + 5, pf[5]);
      atomicAdd(prof + 6, pf[6]);
      atomicAdd(prof + 7, (unsigned long long)iterations);
      for (int i = 8; i < 12; i++) atomicAdd(prof + i, pf[i]);
    }
    for (int i = 0; i < 12; i++) pf[i] = 0;
#endif
    if (lane == 0) sm.quit = 1;
    if (set_aside && win < 0) {  // nothing is written for this query now: the second pass runs it from its start
      if (lane == 0) deferred[atomicAdd(n_deferred, 1)] = qi;
      continue;
    }
    pipe_res wr;
    wr.n_samples = 0; wr.nseg = 0; wr.cost = 0.0;
    if (win >= 0) {
      __threadfence_block();
      const pipe_job jb = sm.job[win % PIPE_RING];
      wr = sm.res[win % PIPE_RING];
      goal_node = jb.cur;
      iterations = jb.iter;
      via_shot = 1;
      status = 0;
      shot_samples = wr.n_samples;
      shot_cost = wr.cost;
    } else if (end_node >= 0) {
      goal_node = end_node;
    }

    // ---- back-track (RS/src/HybridAstar.cpp:690-710)
    int total = 1;  // failure: the goal state alone (hazard H10)
    if (goal_node >= 0) {
      total = via_shot ? shot_samples : 0;
      for (int b = goal_node; b >= 0; b = nodes[b].parent) total += (nodes[b].cmd == -1) ? 1 : pc.ec.sim;
    }
    unsigned long long off = 0;
    if (lane == 0) off = atomicAdd(cursor, (unsigned long long)total);
    off = __shfl_sync(FULLM, off, 0);
    if (off + (unsigned long long)total <= traj_cap) {
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
          warp_sample_path<false>(wr.len, wr.typ, wr.nseg, pc.maxc, pc.maxc, cn.x, cn.y, cn.h, sv);
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

// ---- unit hook: the warp-collective heap against the one-thread replay on the same operations (one warp)
__global__ void heap_ops_kernel(const double *__restrict__ keys, const int *__restrict__ ops, int n_ops,
                                double *ka, int *ia, double *kb, int *ib, int *pops, long long *first_mismatch) {
  const int lane = threadIdx.x & 31;
  __shared__ heap_cache hc;
  int na = 0, nb = 0;
  long long bad = -1;
  for (int i = 0; i < n_ops; i++) {
    const int op = ops[i];
    int pa = -1, pb = -1;
    if (op >= 0) {
      // a run of pushes (up to 26, an expansion's worth): one by one on heap A, as one gheap_push_batch on heap B
      int run = 1;
      while (run < 26 && i + run < n_ops && ops[i + run] >= 0) run++;
      if (lane == 0)
        for (int r = 0; r < run; r++) gheap_push(ka, ia, na, keys[ops[i + r]], i + r);
      na = __shfl_sync(0xffffffffu, na, 0);
      __syncwarp();
      if (run == 1) {
        gheap_push_warp(kb, ib, nb, keys[op], i);
      } else {
        const double fk = lane < run ? keys[ops[i + lane]] : 0.0;
        gheap_push_batch(kb, ib, nb, hc, fk, run == 32 ? 0xffffffffu : ((1u << run) - 1u), i);
      }
      if (pops && lane == 0)
        for (int r = 0; r < run; r++) pops[i + r] = -1;
      i += run - 1;
    } else if (na > 0) {
      if (lane == 0) pa = gheap_pop(ka, ia, na);
      na = __shfl_sync(0xffffffffu, na, 0);
      pa = __shfl_sync(0xffffffffu, pa, 0);
      __syncwarp();
      pb = gheap_pop_warp(kb, ib, nb);
    }
    __syncwarp();
    if (pops && lane == 0) pops[i] = pb;
    // the two heaps must be the same arrays after every operation
    bool diff = pa != pb || na != nb;
    for (int j = lane; j < na && !diff; j += 32) diff = ka[j] != kb[j] || ia[j] != ib[j];
    if (__any_sync(0xffffffffu, diff) && bad < 0) bad = i;
    __syncwarp();
  }
  if (lane == 0) *first_mismatch = bad;
}

extern "C" hpp_status hpp_test_heap_ops(hpp_ctx *c, const double *keys, const int32_t *ops, size_t n_ops,
                                        int32_t *pops_out, int64_t *first_mismatch) {
  if (!c || !keys || !ops || !first_mismatch || n_ops == 0 || n_ops > (1u << 24)) return HPP_ERR_ARG;
  HPP_CUDA(c, cudaSetDevice(c->device));
  void *buf;
  const size_t n = n_ops;
  hpp_status st = hpp_scratch(c, 1, n * (8 + 4 + 8 + 4 + 8 + 4 + 4) + 64, &buf);
  if (st != HPP_OK) return st;
  double *d_keys = (double *)buf, *ka = d_keys + n, *kb = ka + n;
  int *d_ops = (int *)(kb + n), *ia = d_ops + n, *ib = ia + n, *d_pops = ib + n;
  long long *d_bad = (long long *)(((uintptr_t)(d_pops + n) + 7) & ~(uintptr_t)7);
  HPP_CUDA(c, cudaMemcpyAsync(d_keys, keys, n * 8, cudaMemcpyHostToDevice, c->stream));
  HPP_CUDA(c, cudaMemcpyAsync(d_ops, ops, n * 4, cudaMemcpyHostToDevice, c->stream));
  heap_ops_kernel<<<1, 32, 0, c->stream>>>(d_keys, d_ops, (int)n, ka, ia, kb, ib, d_pops, d_bad);
  HPP_LAUNCHED(c);
  long long bad = 0;
  if (pops_out) HPP_CUDA(c, cudaMemcpyAsync(pops_out, d_pops, n * 4, cudaMemcpyDeviceToHost, c->stream));
  HPP_CUDA(c, cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, c->stream));
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  *first_mismatch = bad;
  return HPP_OK;
}

// Whole queries on one or several maps of one context.  Queries are grouped by (map, goal cell): one heuristic field
// per distinct pair; maps that share geometry and vehicle parameters are planned by ONE plan_kernel launch (the kernel
// takes the per-map layer pointers from a descriptor table), maps that differ are planned group after group.
static hpp_status plan_batch_impl(hpp_map *const *maps, size_t n_maps, const uint32_t *map_of_query, int variant,
                                  const double *starts, const double *goals, size_t n, hpp_plan *plans, double *traj,
                                  size_t traj_cap, size_t *traj_used) {
  if (!maps || n_maps == 0 || variant < 0 || variant > 1 || (n && (!starts || !goals || !plans)) || (traj_cap && !traj))
    return HPP_ERR_ARG;
  for (size_t k = 0; k < n_maps; k++)
    if (!maps[k] || maps[k]->ctx != maps[0]->ctx) return HPP_ERR_ARG;
  if (n_maps > 1 && !map_of_query) return HPP_ERR_ARG;
  if (traj_used) *traj_used = 0;
  if (n == 0) return HPP_OK;
  if (n > 0x7fffffff || n_maps > 0x7fffffff) return HPP_ERR_ARG;
  hpp_ctx *c = maps[0]->ctx;
  const bool timing = getenv("HPP_PLAN_TIMING") != nullptr;
  auto t_now = [] { return std::chrono::steady_clock::now(); };
  auto t_us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::micro>(b - a).count();
  };
  const auto t0 = t_now();
  HPP_CUDA(c, cudaSetDevice(c->device));

  // maps with the same geometry and vehicle form a group (one launch per group)
  std::vector<int> grp_of_map(n_maps, -1), grp_first;
  for (size_t k = 0; k < n_maps; k++) {
    for (size_t gi = 0; gi < grp_first.size() && grp_of_map[k] < 0; gi++) {
      const hpp_map *a = maps[grp_first[gi]], *b = maps[k];
      if (a->g.W == b->g.W && a->g.H == b->g.H && a->g.res_f == b->g.res_f && a->g.ox == b->g.ox && a->g.oy == b->g.oy &&
          memcmp(&a->car.p, &b->car.p, sizeof a->car.p) == 0)
        grp_of_map[k] = (int)gi;
    }
    if (grp_of_map[k] < 0) {
      grp_of_map[k] = (int)grp_first.size();
      grp_first.push_back((int)k);
    }
  }

  // one heuristic field per distinct (map, goal cell): sort the queries by (group, map, cell)
  std::vector<long> cell(n);
  std::vector<int> mq(n);
  for (size_t q = 0; q < n; q++) {
    const uint32_t k = map_of_query ? map_of_query[q] : 0u;
    if (k >= n_maps) HPP_FAIL(c, HPP_ERR_ARG, "query %zu names map %u of %zu", q, k, n_maps);
    const hpp_map *m = maps[k];
    const int gx = (int)((goals[3 * q] - m->g.ox) / m->g.res), gy = (int)((goals[3 * q + 1] - m->g.oy) / m->g.res);
    if (gx < 0 || gx >= m->g.W || gy < 0 || gy >= m->g.H)
      HPP_FAIL(c, HPP_ERR_ARG, "goal of query %zu lies outside the map (cell %d,%d)", q, gx, gy);
    cell[q] = (long)gy * m->g.W + gx;
    mq[q] = (int)k;
  }
  std::vector<int> order(n);
  for (size_t q = 0; q < n; q++) order[q] = (int)q;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    if (grp_of_map[mq[a]] != grp_of_map[mq[b]]) return grp_of_map[mq[a]] < grp_of_map[mq[b]];
    if (mq[a] != mq[b]) return mq[a] < mq[b];
    return cell[a] < cell[b];
  });

  // device buffers shared by all groups: inputs, outputs, map descriptors
  size_t free_b = c->plan_free_bytes, total_b = 0;
  if (free_b == 0 || n > 64) {
    HPP_CUDA(c, cudaMemGetInfo(&free_b, &total_b));
    c->plan_free_bytes = free_b;
  }
  const auto t1 = t_now();
  char *d_io = nullptr;
  const size_t io_bytes = n * 48 + n * 16 + n * sizeof(hpp_plan) + n_maps * sizeof(plan_map_desc) + traj_cap * 40 + 128;
  hpp_status st = hpp_scratch(c, 0, io_bytes, (void **)&d_io);
  if (st != HPP_OK) {
    c->plan_free_bytes = 0;  // the cached figure was stale: ask again next time
    return HPP_ERR_NOMEM;
  }
  double *d_starts = (double *)d_io, *d_goals = d_starts + 3 * n;
  int *d_qids = (int *)(d_goals + 3 * n), *d_fq = d_qids + n, *d_mq = d_fq + n;
  hpp_plan *d_plans = (hpp_plan *)(d_mq + n + (n & 1));
  unsigned long long *d_cursor = (unsigned long long *)(d_plans + n);
  plan_map_desc *d_descs = (plan_map_desc *)(d_cursor + 4);  // cursor: trajectory rows, query counter, open-list check
  int *d_defer = (int *)(d_descs + n_maps), *d_ndefer = d_defer + n;  // queries set aside for the second pass
  double *d_traj = (double *)(d_ndefer + 2 + (n & 1));
  std::vector<plan_map_desc> descs(n_maps);
  for (size_t k = 0; k < n_maps; k++) {
    descs[k].dist = maps[k]->d_distance;
    descs[k].bits = maps[k]->d_occ_bits;
    descs[k].pitch = maps[k]->bits_pitch;
    descs[k].pad_ = 0;
  }
  cudaMemcpyAsync(d_starts, starts, n * 24, cudaMemcpyHostToDevice, c->stream);
  cudaMemcpyAsync(d_goals, goals, n * 24, cudaMemcpyHostToDevice, c->stream);
  cudaMemcpyAsync(d_qids, order.data(), n * 4, cudaMemcpyHostToDevice, c->stream);
  cudaMemcpyAsync(d_mq, mq.data(), n * 4, cudaMemcpyHostToDevice, c->stream);
  cudaMemcpyAsync(d_descs, descs.data(), n_maps * sizeof(plan_map_desc), cudaMemcpyHostToDevice, c->stream);
  cudaMemsetAsync(d_cursor, 0, 32, c->stream);
  unsigned long long *d_prof = nullptr;
#ifdef HPP_PLAN_PROF
  if (hpp_scratch(c, 5, 128, (void **)&d_prof) == HPP_OK) cudaMemsetAsync(d_prof, 0, 128, c->stream);
#endif
  std::vector<int> fq(n);
  std::vector<int32_t> chunk_goals;
  cudaError_t e = cudaSuccess;

  size_t q0 = 0;
  while (q0 < n && st == HPP_OK) {
    // ---- one group of maps: queries [q0, g1) of the sorted order
    const int grp = grp_of_map[mq[order[q0]]];
    size_t g1 = q0;
    while (g1 < n && grp_of_map[mq[order[g1]]] == grp) g1++;
    hpp_map *m0 = maps[grp_first[grp]];
    const int W = m0->g.W;
    const size_t ng = g1 - q0;
    plan_consts pc;
    memset(&pc, 0, sizeof pc);
    pc.g = m0->g;
    hpp_make_expand_consts(m0, variant, &pc.ec);
    pc.cc = hpp_make_circ_consts(m0);
    for (int k = 0; k < 5; k++) {
      pc.rx[k] = m0->car.rx[k];
      pc.ry[k] = m0->car.ry[k];
    }
    pc.maxc = m0->car.maxc;
    pc.max_steer = m0->car.p.max_steer;
    pc.max_nodes = PLAN_ITER_CAP * pc.ec.n_cmd + 2;
    int slots = 1;
    while (slots < 2 * pc.max_nodes) slots <<= 1;
    pc.table_mask = slots - 1;
    const size_t per_warp = (size_t)pc.max_nodes * (sizeof(pnode) + 12) + (size_t)slots * 4;
    // kernel shape: eight warps per query while the batch leaves SMs free anyway, four otherwise (HPP_PLAN_WARPS forces one)
    int plan_warps = ng <= (size_t)c->sm_count ? PLAN_WARPS_LONE : PLAN_WARPS_BATCH;
    if (const char *ev = getenv("HPP_PLAN_WARPS")) {
      const int v = atoi(ev);
      if (v == PLAN_WARPS_LONE || v == PLAN_WARPS_BATCH) plan_warps = v;
    }
    const void *kern = plan_warps == PLAN_WARPS_LONE ? (const void *)plan_kernel<HPP_VARIANT_DUBINS, PLAN_WARPS_LONE>
                                                     : (const void *)plan_kernel<HPP_VARIANT_DUBINS, PLAN_WARPS_BATCH>;
    // Reeds-Shepp: the pipelined kernel (search on warp 0, the open list on the last warp, one whole shot per worker warp)
    size_t plan_smem_bytes = 0;
    // Two passes for big Reeds-Shepp batches (HPP_PLAN_BUDGET iterations in the first, default 64; 0 = one pass): 95 % of
    // the cfg5 queries end within 56 iterations and 2 % run into the 3 500-iteration cap, which makes the batch time the
    // time of the SM that drew the most cap hitters from the queue.  The first pass sets a query aside when it passes the
    // budget; the second pass runs exactly those, from their start -- every long query starts at once on an SM of its
    // own instead of behind whatever its CTA drew before.  Same results: a query's
    // outcome does not depend on what ran beside it.
    static const int pass_budget = getenv("HPP_PLAN_BUDGET") ? atoi(getenv("HPP_PLAN_BUDGET")) : 64;
    bool two_pass = false;
    const bool pipe = variant == HPP_VARIANT_RS;
    if (pipe) {
      // One CTA per SM: controller, three expansion helpers, the shot workers, the heap warp.  A pipelined query keeps
      // an SM's instruction fetch busy on its own -- its warps are at different places of 140 KB of code -- and two of
      // them on one SM ran at half speed each, so there is no two-per-SM build any more.  Ten warps (five workers): a
      // shot costs a worker ~47 k cycles and the search now asks for one every ~10 k (HPP_PLAN_PIPE_NW=9: four workers).
      static const int nw_env = getenv("HPP_PLAN_PIPE_NW") ? atoi(getenv("HPP_PLAN_PIPE_NW")) : 10;
      if (nw_env == 9) {
        plan_warps = 9;
        kern = (const void *)plan_pipe_kernel<9, 1>;
        plan_smem_bytes = sizeof(pipe_smem<9>);
      } else {
        plan_warps = 10;
        kern = (const void *)plan_pipe_kernel<10, 1>;
        plan_smem_bytes = sizeof(pipe_smem<10>);
      }
      two_pass = ng > (size_t)32 * c->sm_count && pass_budget > 0;
      if (plan_smem_bytes > 48 * 1024)
        HPP_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan_smem_bytes));
    }
    // cached per kernel shape (the RS slots hold the pipelined kernel's two builds; HPP_PLAN_PIPE_NW is read once)
    const int occ_slot = pipe ? (plan_warps == 9 ? 0 : 1)
                              : variant * 2 + (plan_warps == PLAN_WARPS_LONE ? 1 : 0);
    int occ_blocks = c->plan_occ_blocks[occ_slot];
    if (occ_blocks == 0) {
      HPP_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_blocks, kern, plan_warps * 32, plan_smem_bytes));
      c->plan_occ_blocks[occ_slot] = occ_blocks;
    }
    if (occ_blocks < 1) occ_blocks = 1;
    size_t max_warps = (size_t)c->sm_count * occ_blocks;  // CTAs (= queries in flight): one resident wave, persistent
    if (max_warps > ng) max_warps = ng;
    const size_t field_bytes = m0->n * sizeof(double);
    const size_t budget = free_b > ((size_t)2 << 30) ? free_b - ((size_t)1 << 30) : free_b / 2;
    while (max_warps > 1 && max_warps * per_warp > budget / 2) max_warps /= 2;
    const size_t fixed_bytes = max_warps * per_warp + io_bytes;
    if (fixed_bytes + field_bytes > budget)
      HPP_FAIL(c, HPP_ERR_NOMEM, "plan_batch: pools (%zu B) + query/trajectory buffers (%zu B) + one heuristic field (%zu B) "
               "exceed the device memory budget (%zu B): lower traj_cap or the batch size", max_warps * per_warp, io_bytes,
               field_bytes, budget);
    size_t fields_cap = (budget - fixed_bytes) / field_bytes;
    if (fields_cap > 65536) fields_cap = 65536;  // (a chunk ends with a synchronisation: as few chunks as memory allows)
    const int grid = (int)max_warps;
    const size_t n_warps = (size_t)grid;  // one node pool / heap / table per CTA
    char *d_pool = nullptr;
    if (hpp_scratch(c, 1, n_warps * per_warp + 256, (void **)&d_pool) != HPP_OK) {
      c->plan_free_bytes = 0;
      return HPP_ERR_NOMEM;
    }
    pnode *d_nodes = (pnode *)d_pool;
    double *d_keys = (double *)(d_nodes + n_warps * pc.max_nodes);
    int *d_idx = (int *)(d_keys + n_warps * pc.max_nodes);
    int *d_table = d_idx + n_warps * pc.max_nodes;
    double *d_fields = nullptr;
    size_t fields_alloc = 0;

    // ---- chunks of distinct (map, goal cell) pairs of this group
    while (q0 < g1 && st == HPP_OK) {
      chunk_goals.clear();
      std::vector<std::pair<int, size_t>> runs;  // (map, first field of the map's run) inside the chunk
      size_t q1 = q0;
      long last_cell = -1;
      int last_map = -1;
      while (q1 < g1) {
        const long cl = cell[order[q1]];
        const int mk = mq[order[q1]];
        if (cl != last_cell || mk != last_map) {
          if (chunk_goals.size() / 2 >= fields_cap) break;
          if (mk != last_map) runs.push_back({mk, chunk_goals.size() / 2});
          chunk_goals.push_back((int32_t)(cl % W));
          chunk_goals.push_back((int32_t)(cl / W));
          last_cell = cl;
          last_map = mk;
        }
        fq[q1] = (int)(chunk_goals.size() / 2 - 1);
        q1++;
      }
      const size_t nf = chunk_goals.size() / 2;
      if (nf > fields_alloc) {
        if (hpp_scratch(c, 3, nf * field_bytes, (void **)&d_fields) != HPP_OK) {
          st = HPP_ERR_NOMEM;
          c->plan_free_bytes = 0;
          break;
        }
        fields_alloc = nf;
      }
      const auto t2 = t_now();
      for (size_t r = 0; r < runs.size() && st == HPP_OK; r++) {
        const size_t f0 = runs[r].second, f1 = r + 1 < runs.size() ? runs[r + 1].second : nf;
        st = hpp_heuristic_batch_dev(maps[runs[r].first], variant, chunk_goals.data() + 2 * f0, f1 - f0,
                                     d_fields + f0 * m0->n, nullptr);
      }
      if (st != HPP_OK) break;
      const auto t3 = t_now();
      if (timing) fprintf(stderr, "[plan timing] setup %.0f us, alloc+h2d %.0f us, heuristic (returns after the kernel) %.0f us\n",
                          t_us(t0, t1), t_us(t1, t2), t_us(t2, t3));
      cudaMemcpyAsync(d_fq + q0, fq.data() + q0, (q1 - q0) * 4, cudaMemcpyHostToDevice, c->stream);
      cudaMemsetAsync(d_cursor + 1, 0, 8, c->stream);  // the kernel's query counter
      {
        const plan_map_desc *a_descs = d_descs;
        const int *a_mq = n_maps > 1 ? d_mq : nullptr;
        const double *a_starts = d_starts, *a_goals = d_goals, *a_fields = d_fields;
        const int *a_qids = d_qids + q0, *a_fq = d_fq + q0;
        int a_nq = (int)(q1 - q0);
        unsigned long long a_cap = traj_cap;
        if (!pipe) {
          void *params[] = {&pc,       &a_descs, &a_mq,   &a_starts, &a_goals, &a_qids,  &a_nq,   &a_fq,
                            &a_fields, &d_nodes, &d_keys, &d_idx,    &d_table, &d_plans, &d_traj, &a_cap,
                            &d_cursor, &d_prof};
          cudaLaunchKernel(kern, dim3(grid), dim3(plan_warps * 32), params, plan_smem_bytes, c->stream);
        } else {
          static const int check_env = getenv("HPP_PLAN_CHECK") ? atoi(getenv("HPP_PLAN_CHECK")) : 1;
          int a_check = check_env;
          int a_budget = two_pass ? pass_budget : PLAN_ITER_CAP + 1;
          const int *a_list = nullptr, *a_nlist = nullptr;
          if (two_pass) cudaMemsetAsync(d_ndefer, 0, 4, c->stream);
          void *params[] = {&pc,       &a_descs, &a_mq,   &a_starts, &a_goals, &a_qids,  &a_nq,     &a_fq,    &a_fields,
                            &d_nodes,  &d_keys,  &d_idx,  &d_table,  &d_plans, &d_traj,  &a_cap,    &d_cursor, &a_budget,
                            &a_list,   &a_nlist, &d_defer, &d_ndefer, &a_check, &d_prof};
          cudaLaunchKernel(kern, dim3(grid), dim3(plan_warps * 32), params, plan_smem_bytes, c->stream);
          if (two_pass) {  // the queries set aside, one CTA per SM (the count stays on the device)
            c->launches++;
            cudaMemsetAsync(d_cursor + 1, 0, 8, c->stream);
            a_budget = PLAN_ITER_CAP + 1;
            a_list = d_defer;
            a_nlist = d_ndefer;
            const int grid2 = grid < c->sm_count ? grid : c->sm_count;
            cudaLaunchKernel(kern, dim3(grid2), dim3(plan_warps * 32), params, plan_smem_bytes, c->stream);
          }
        }
      }
      c->launches++;
      e = cudaGetLastError();
      const auto t4 = t_now();
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);  // fq / chunk_goals are reused by the next chunk
      if (timing) fprintf(stderr, "[plan timing] plan kernel + sync %.0f us\n", t_us(t4, t_now()));
      if (e != cudaSuccess) {
        snprintf(c->err, sizeof c->err, "plan kernel failed: %s", cudaGetErrorString(e));
        st = HPP_ERR_CUDA;
      }
      q0 = q1;
    }
  }
#ifdef HPP_PLAN_PROF
  if (st == HPP_OK) {
    unsigned long long hp[16];
    void *dp = nullptr;
    if (hpp_scratch(c, 5, 128, &dp) == HPP_OK && cudaMemcpy(hp, dp, 128, cudaMemcpyDeviceToHost) == cudaSuccess && hp[7])
      if (variant == HPP_VARIANT_RS)
        fprintf(stderr, "[pipe prof] worker 1 of every CTA: %llu shots, %.1f paths/shot, generate %.0f + path checks %.0f cycles per "
                "shot; controllers of long queries: %llu iterations, loop %.0f cycles/iteration of which ring-full waits %.0f, "
                "drain %.0f cycles/query-iteration; pop %.0f expansion %.0f insert %.0f pushes %.0f\n", hp[2], hp[2] ? (double)hp[3] / hp[2] : 0.0, hp[2] ? (double)hp[0] / hp[2] : 0.0,
                hp[2] ? (double)hp[1] / hp[2] : 0.0, hp[7], (double)hp[4] / hp[7], (double)hp[5] / hp[7], (double)hp[6] / hp[7],
                (double)hp[8] / hp[7], (double)hp[9] / hp[7], (double)hp[10] / hp[7], (double)hp[11] / hp[7]);
      else
      fprintf(stderr, "[plan prof] %llu iterations of long queries; cycles/iteration: pop %.0f formulas %.0f dedup+order %.0f "
              "path checks %.0f expand %.0f insert %.0f push %.0f\n", hp[7], (double)hp[0] / hp[7], (double)hp[6] / hp[7],
              (double)hp[1] / hp[7], (double)hp[2] / hp[7], (double)hp[3] / hp[7], (double)hp[4] / hp[7],
              (double)hp[5] / hp[7]);
  }
#endif
  if (st == HPP_OK) {
    unsigned long long used = 0;
    cudaMemcpyAsync(plans, d_plans, n * sizeof(hpp_plan), cudaMemcpyDeviceToHost, c->stream);
    cudaMemcpyAsync(c->h_flags + 48, d_cursor, 24, cudaMemcpyDeviceToHost, c->stream);
    e = cudaStreamSynchronize(c->stream);
    memcpy(&used, c->h_flags + 48, 8);
    unsigned long long order_bad = 0;
    memcpy(&order_bad, c->h_flags + 52, 8);
    if (e == cudaSuccess && order_bad)  // never seen; a wrong answer must not pass silently
      HPP_FAIL(c, HPP_ERR_CUDA, "plan_batch: internal check failed -- the open list's top differed from the node the "
               "search went on with in %llu pops", order_bad);
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
  if (timing) fprintf(stderr, "[plan timing] total %.0f us\n", t_us(t0, t_now()));
  return st;
}

extern "C" hpp_status hpp_plan_batch(hpp_map *m, int variant, const double *starts, const double *goals, size_t n,
                                     hpp_plan *plans, double *traj, size_t traj_cap, size_t *traj_used) {
  if (!m) return HPP_ERR_ARG;
  hpp_map *one[1] = {m};
  return plan_batch_impl(one, 1, nullptr, variant, starts, goals, n, plans, traj, traj_cap, traj_used);
}

extern "C" hpp_status hpp_plan_batch_multi(hpp_map *const *maps, size_t n_maps, const uint32_t *map_of_query, int variant,
                                           const double *starts, const double *goals, size_t n, hpp_plan *plans,
                                           double *traj, size_t traj_cap, size_t *traj_used) {
  return plan_batch_impl(maps, n_maps, map_of_query, variant, starts, goals, n, plans, traj, traj_cap, traj_used);
}
