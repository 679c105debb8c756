// shots.cu -- K7 / K8 batch kernels: analytic-expansion shots, one warp per start node (see shots.cuh).
#include "shots.cuh"

#define SHOT_WARPS 4

struct rect_consts {
  float rx[5], ry[5];
};

static circ_consts make_cc(const hpp_map *m) { return hpp_make_circ_consts(m); }
static rect_consts make_rc(const hpp_map *m) {
  rect_consts rc;
  for (int k = 0; k < 5; k++) {
    rc.rx[k] = m->car.rx[k];
    rc.ry[k] = m->car.ry[k];
  }
  return rc;
}

__global__ void __launch_bounds__(SHOT_WARPS * 32) rs_shot_kernel(hpp_geom g, circ_consts cc,
                                                                  const float *__restrict__ dist, double maxc,
                                                                  double max_steer,
                                                                  const hpp_node *__restrict__ starts, size_t n,
                                                                  const double *__restrict__ goals, size_t goal_stride,
                                                                  hpp_shot *__restrict__ shots,
                                                                  double *__restrict__ samples,
                                                                  unsigned long long samples_cap,
                                                                  unsigned long long *__restrict__ cursor) {
  __shared__ rs_warp_smem S_all[SHOT_WARPS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  rs_warp_smem &S = S_all[wid];
  const size_t warps = (size_t)gridDim.x * SHOT_WARPS;
  for (size_t q = (size_t)blockIdx.x * SHOT_WARPS + wid; q < n; q += warps) {
    const hpp_node st = starts[q];
    const double *gp = goals + q * goal_stride;
    const double gx = gp[0], gy = gp[1], gh = gp[2];
    rs_generate(S, st.x, st.y, st.heading, gx, gy, gh, maxc);
    rs_shot_result R = rs_select(S, g, cc, dist, st.x, st.y, st.heading, st.cost, maxc, max_steer);
    unsigned long long off = 0;
    if (R.found) {
      if (lane == 0) off = atomicAdd(cursor, (unsigned long long)R.n_samples);
      off = __shfl_sync(0xffffffffu, off, 0);
      if (off + R.n_samples <= samples_cap) {
        store_visitor sv;
        sv.out = samples + off * 5;
        sv.g = &g;
        sv.cap = R.n_samples;
        const int k = R.cand;
        warp_sample_path<false>(S.len[k], S.typ[k], S.nseg[k], maxc, maxc, st.x, st.y, st.heading, sv);
      }
    }
    if (lane == 0) {
      hpp_shot o;
      o.found = R.found;
      o.n_paths = R.n_paths;
      o.chosen = R.chosen;
      o.n_samples = R.n_samples;
      o.cost = R.cost;
      o.sample_offset = off;
      shots[q] = o;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(SHOT_WARPS * 32) dubins_shot_kernel(hpp_geom g, rect_consts rc,
                                                                      const uint32_t *__restrict__ bits, int pitch,
                                                                      double curv, double max_steer,
                                                                      const hpp_node *__restrict__ starts, size_t n,
                                                                      const double *__restrict__ goals,
                                                                      size_t goal_stride, hpp_shot *__restrict__ shots,
                                                                      double *__restrict__ samples,
                                                                      unsigned long long samples_cap,
                                                                      unsigned long long *__restrict__ cursor) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const size_t warps = (size_t)gridDim.x * SHOT_WARPS;
  for (size_t q = (size_t)blockIdx.x * SHOT_WARPS + wid; q < n; q += warps) {
    const hpp_node st = starts[q];
    const double *gp = goals + q * goal_stride;
    db_best B = db_plan(st.x, st.y, st.heading, gp[0], gp[1], gp[2], curv);
    hpp_shot o;
    o.found = 0; o.n_paths = 0; o.chosen = -1; o.n_samples = 0; o.cost = 0.0; o.sample_offset = 0;
    sat_visitor vis;
    vis.g = &g; vis.rx = rc.rx; vis.ry = rc.ry; vis.bits = bits; vis.pitch = pitch; vis.hit_zero_idx = 0x7fffffff;
    // calc_dubins_path(..., step_size = 0.1) (DB/include/dubins_path.h:57): the step is NOT scaled by the curvature
    int n_keep = warp_sample_path<true>(B.len, B.typ, B.nmode, curv, 0.1, st.x, st.y, st.heading, vis);
    bool collides = (n_keep < 0) || (vis.hit_zero_idx < n_keep);
    if (n_keep != 0) o.n_paths = 1;
    if (n_keep > 0 && !collides) {
      // dubins_Path_Cost (DB/src/HybridAstar.cpp:167-184): the NORMALISED length is added
      double cost = st.cost;
      cost += fabs(B.best_cost);
      for (int i = 0; i < 3; i++)
        if (B.typ[i] != 'S') cost += max_steer * 15.0;
      unsigned long long off = 0;
      if (lane == 0) off = atomicAdd(cursor, (unsigned long long)n_keep);
      off = __shfl_sync(0xffffffffu, off, 0);
      if (off + n_keep <= samples_cap) {
        store_visitor sv;
        sv.out = samples + off * 5;
        sv.g = &g;
        sv.cap = n_keep;
        warp_sample_path<true>(B.len, B.typ, B.nmode, curv, 0.1, st.x, st.y, st.heading, sv);
      }
      o.found = 1; o.chosen = 0; o.n_samples = n_keep; o.cost = cost; o.sample_offset = off;
    }
    if (lane == 0) shots[q] = o;
    __syncwarp();
  }
}

typedef enum { SHOT_RS, SHOT_DB } shot_kind;

static hpp_status shot_host(hpp_map *m, shot_kind kind, const hpp_node *starts, size_t n, const double *goals,
                            size_t goal_stride, hpp_shot *shots, double *samples, size_t samples_cap,
                            size_t *samples_used) {
  if (!m || (n && (!starts || !goals || !shots)) || (samples_cap && !samples)) return HPP_ERR_ARG;
  if (goal_stride != 0 && goal_stride != 3) return HPP_ERR_ARG;
  if (samples_used) *samples_used = 0;
  if (n == 0) return HPP_OK;
  hpp_ctx *c = m->ctx;
  HPP_CUDA(c, cudaSetDevice(c->device));
  const size_t n_goals = goal_stride ? n : 1;
  const size_t in_bytes = n * sizeof(hpp_node) + n_goals * 24;
  const size_t out_bytes = n * sizeof(hpp_shot) + samples_cap * 40 + 16;
  void *d_in, *d_out;
  hpp_status s = hpp_scratch(c, 0, in_bytes, &d_in);
  if (s != HPP_OK) return s;
  s = hpp_scratch(c, 1, out_bytes, &d_out);
  if (s != HPP_OK) return s;
  hpp_node *d_starts = (hpp_node *)d_in;
  double *d_goals = (double *)((char *)d_in + n * sizeof(hpp_node));
  unsigned long long *d_cursor = (unsigned long long *)d_out;
  hpp_shot *d_shots = (hpp_shot *)((char *)d_out + 16);
  double *d_samples = (double *)((char *)d_out + 16 + n * sizeof(hpp_shot));
  HPP_CUDA(c, cudaMemcpyAsync(d_starts, starts, n * sizeof(hpp_node), cudaMemcpyHostToDevice, c->stream));
  HPP_CUDA(c, cudaMemcpyAsync(d_goals, goals, n_goals * 24, cudaMemcpyHostToDevice, c->stream));
  HPP_CUDA(c, cudaMemsetAsync(d_cursor, 0, 16, c->stream));
  long long want = hpp_div_up((long long)n, SHOT_WARPS);
  int grid = (int)(want < (long long)c->sm_count * 8 ? want : (long long)c->sm_count * 8);
  if (kind == SHOT_RS)
    rs_shot_kernel<<<grid, SHOT_WARPS * 32, 0, c->stream>>>(m->g, make_cc(m), m->d_distance, m->car.maxc,
                                                            m->car.p.max_steer, d_starts, n, d_goals, goal_stride,
                                                            d_shots, d_samples, samples_cap, d_cursor);
  else
    dubins_shot_kernel<<<grid, SHOT_WARPS * 32, 0, c->stream>>>(m->g, make_rc(m), m->d_occ_bits, m->bits_pitch,
                                                                m->car.maxc, m->car.p.max_steer, d_starts, n, d_goals,
                                                                goal_stride, d_shots, d_samples, samples_cap, d_cursor);
  HPP_LAUNCHED(c);
  unsigned long long used = 0;
  HPP_CUDA(c, cudaMemcpyAsync(shots, d_shots, n * sizeof(hpp_shot), cudaMemcpyDeviceToHost, c->stream));
  HPP_CUDA(c, cudaMemcpyAsync(c->h_flags + 48, d_cursor, 8, cudaMemcpyDeviceToHost, c->stream));
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(&used, c->h_flags + 48, 8);
  const size_t copy_rows = used < samples_cap ? (size_t)used : samples_cap;
  if (copy_rows) {
    HPP_CUDA(c, cudaMemcpyAsync(samples, d_samples, copy_rows * 40, cudaMemcpyDeviceToHost, c->stream));
    HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  if (samples_used) *samples_used = (size_t)used;
  if (used > samples_cap)
    HPP_FAIL(c, HPP_ERR_CAPACITY, "shot trajectories need %llu sample rows, caller supplied %zu", used, samples_cap);
  return HPP_OK;
}

extern "C" hpp_status hpp_rs_shot(hpp_map *m, const hpp_node *starts, size_t n, const double *goals_xyh,
                                  size_t goal_stride, hpp_shot *shots, double *samples, size_t samples_cap,
                                  size_t *samples_used) {
  return shot_host(m, SHOT_RS, starts, n, goals_xyh, goal_stride, shots, samples, samples_cap, samples_used);
}
extern "C" hpp_status hpp_dubins_shot(hpp_map *m, const hpp_node *starts, size_t n, const double *goals_xyh,
                                      size_t goal_stride, hpp_shot *shots, double *samples, size_t samples_cap,
                                      size_t *samples_used) {
  return shot_host(m, SHOT_DB, starts, n, goals_xyh, goal_stride, shots, samples, samples_cap, samples_used);
}

// ------------------------------------------------------------------ unit-level parity hooks (a7 / a11)
struct all_paths_hdr {
  int n_paths;
  int n_seg[RS_NCAND];
  int n_samples[RS_NCAND];
  double lengths[RS_NCAND][5];
  double total[RS_NCAND];
  char types[RS_NCAND][5];
  int seg_samples[RS_NCAND][5];  // per segment: interior samples | (those written with pd > 0) << 16
  unsigned long long samples_needed;
};

__global__ void __launch_bounds__(32) rs_all_paths_kernel(double sx, double sy, double syaw, double gx, double gy,
                                                          double gyaw, double maxc, double step_size,
                                                          all_paths_hdr *__restrict__ hdr,
                                                          double *__restrict__ samples, unsigned long long cap) {
  __shared__ rs_warp_smem S;
  const int lane = threadIdx.x & 31;
  rs_generate(S, sx, sy, syaw, gx, gy, gyaw, maxc);
  unsigned long long base = 0;
  for (int e = 0; e < S.n_acc; e++) {
    const int k = S.acc[e];
    store3_visitor sv;
    sv.out = samples;
    sv.cap = cap;
    sv.base = base;
    if (lane == 0)
      for (int i = 0; i < 5; i++) hdr->seg_samples[e][i] = 0;
    __syncwarp();
    int n_keep = warp_sample_path<false>(S.len[k], S.typ[k], S.nseg[k], maxc, step_size * maxc, sx, sy, syaw, sv, true,
                                         hdr->seg_samples[e]);
    if (lane == 0) {
      hdr->n_seg[e] = S.nseg[k];
      hdr->n_samples[e] = n_keep;
      for (int i = 0; i < 5; i++) {
        hdr->lengths[e][i] = (i < S.nseg[k]) ? S.len[k][i] / maxc : 0.0;  // :87-90
        hdr->types[e][i] = (char)S.typ[k][i];
      }
      hdr->total[e] = S.L[k] / maxc;  // :92
    }
    base += (unsigned long long)n_keep;
  }
  if (lane == 0) {
    hdr->n_paths = S.n_acc;
    hdr->samples_needed = base;
  }
}

extern "C" hpp_status hpp_rs_all_paths(hpp_ctx *c, const double start[3], const double goal[3], double maxc,
                                       double step_size, int32_t max_paths, int32_t *n_paths, int32_t *n_seg,
                                       double *lengths, char *types, double *total_len, int32_t *n_samples,
                                       double *samples, size_t samples_cap, int32_t *directions) {
  if (!c || !start || !goal || !n_paths || !n_seg || !lengths || !types || !total_len || !n_samples || max_paths < 1)
    return HPP_ERR_ARG;
  HPP_CUDA(c, cudaSetDevice(c->device));
  void *d_out, *h_hdr;
  hpp_status s = hpp_scratch(c, 1, sizeof(all_paths_hdr) + 256 + samples_cap * 24, &d_out);
  if (s != HPP_OK) return s;
  s = hpp_pinned(c, 1, sizeof(all_paths_hdr), &h_hdr);
  if (s != HPP_OK) return s;
  all_paths_hdr *d_hdr = (all_paths_hdr *)d_out;
  double *d_samples = (double *)((char *)d_out + ((sizeof(all_paths_hdr) + 255) & ~(size_t)255));
  rs_all_paths_kernel<<<1, 32, 0, c->stream>>>(start[0], start[1], start[2], goal[0], goal[1], goal[2], maxc, step_size,
                                               d_hdr, d_samples, samples_cap);
  HPP_LAUNCHED(c);
  HPP_CUDA(c, cudaMemcpyAsync(h_hdr, d_hdr, sizeof(all_paths_hdr), cudaMemcpyDeviceToHost, c->stream));
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  const all_paths_hdr *h = (const all_paths_hdr *)h_hdr;
  *n_paths = h->n_paths;
  if (h->n_paths > max_paths) HPP_FAIL(c, HPP_ERR_CAPACITY, "%d paths, room for %d", h->n_paths, max_paths);
  for (int e = 0; e < h->n_paths; e++) {
    n_seg[e] = h->n_seg[e];
    n_samples[e] = h->n_samples[e];
    total_len[e] = h->total[e];
    for (int i = 0; i < 5; i++) {
      lengths[e * 5 + i] = h->lengths[e][i];
      types[e * 5 + i] = h->types[e][i];
    }
  }
  if (h->samples_needed > samples_cap)
    HPP_FAIL(c, HPP_ERR_CAPACITY, "%llu samples, room for %zu", h->samples_needed, samples_cap);
  if (h->samples_needed && samples) {
    HPP_CUDA(c, cudaMemcpyAsync(samples, d_samples, h->samples_needed * 24, cudaMemcpyDeviceToHost, c->stream));
    HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  if (directions) {
    // PATH::directions: slot 0 carries the sign of the first segment's length (RS/src/reeds_shepp_path.cpp:697); every
    // other slot the sign of the length argument of the interpolate() call that wrote it (:791) -- the running pd for
    // interior samples (negative for the first samples of a segment that continues in its predecessor's direction,
    // :716-719), the segment length for the end sample of the last segment; trimmed like the samples.  pd moves
    // monotonically by +-step inside a segment, so its positive samples are one contiguous run.
    size_t o = 0;
    for (int e = 0; e < h->n_paths; e++) {
      const int keep = h->n_samples[e];
      int j = 0;
      if (keep > 0) directions[o + j++] = h->lengths[e][0] > 0.0 ? 1 : -1;
      for (int i = 0; i < h->n_seg[e]; i++) {
        const bool fwd = h->lengths[e][i] > 0.0;  // step d > 0: pd rises, non-positive samples first
        const int cnt = h->seg_samples[e][i] & 0xffff, npos = h->seg_samples[e][i] >> 16;
        for (int q = 0; q < cnt && j < keep; q++) {
          const bool positive = fwd ? q >= cnt - npos : q < npos;
          directions[o + j++] = positive ? 1 : -1;
        }
        if (i == h->n_seg[e] - 1 && j < keep) directions[o + j++] = fwd ? 1 : -1;
      }
      o += (size_t)keep;
    }
  }
  return HPP_OK;
}

struct dubins_hdr {
  double length;
  char mode[4];
  int n_samples;
};

__global__ void __launch_bounds__(32) dubins_path_kernel(double sx, double sy, double syaw, double gx, double gy,
                                                         double gyaw, double curv, double step_size,
                                                         dubins_hdr *__restrict__ hdr, double *__restrict__ samples,
                                                         unsigned long long cap) {
  const int lane = threadIdx.x & 31;
  db_best B = db_plan(sx, sy, syaw, gx, gy, gyaw, curv);
  store3_visitor sv;
  sv.out = samples;
  sv.cap = cap;
  sv.base = 0;
  int n_keep = warp_sample_path<true>(B.len, B.typ, B.nmode, curv, step_size, sx, sy, syaw, sv);
  if (lane == 0) {
    hdr->length = B.best_cost;
    for (int i = 0; i < 3; i++) hdr->mode[i] = (char)B.typ[i];
    hdr->mode[3] = 0;
    hdr->n_samples = n_keep;
  }
}

extern "C" hpp_status hpp_dubins_path(hpp_ctx *c, const double start[3], const double goal[3], double curv,
                                      double step_size, double *length, char mode[3], int32_t *n_samples,
                                      double *samples, size_t samples_cap) {
  if (!c || !start || !goal || !length || !mode || !n_samples) return HPP_ERR_ARG;
  HPP_CUDA(c, cudaSetDevice(c->device));
  void *d_out, *h_hdr;
  hpp_status s = hpp_scratch(c, 1, 256 + samples_cap * 24, &d_out);
  if (s != HPP_OK) return s;
  s = hpp_pinned(c, 1, sizeof(dubins_hdr), &h_hdr);
  if (s != HPP_OK) return s;
  dubins_hdr *d_hdr = (dubins_hdr *)d_out;
  double *d_samples = (double *)((char *)d_out + 256);
  dubins_path_kernel<<<1, 32, 0, c->stream>>>(start[0], start[1], start[2], goal[0], goal[1], goal[2], curv, step_size,
                                              d_hdr, d_samples, samples_cap);
  HPP_LAUNCHED(c);
  HPP_CUDA(c, cudaMemcpyAsync(h_hdr, d_hdr, sizeof(dubins_hdr), cudaMemcpyDeviceToHost, c->stream));
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  const dubins_hdr *h = (const dubins_hdr *)h_hdr;
  *length = h->length;
  mode[0] = h->mode[0];
  mode[1] = h->mode[1];
  mode[2] = h->mode[2];
  *n_samples = h->n_samples;
  if ((size_t)h->n_samples > samples_cap)
    HPP_FAIL(c, HPP_ERR_CAPACITY, "%d samples, room for %zu", h->n_samples, samples_cap);
  if (h->n_samples && samples) {
    HPP_CUDA(c, cudaMemcpyAsync(samples, d_samples, (size_t)h->n_samples * 24, cudaMemcpyDeviceToHost, c->stream));
    HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return HPP_OK;
}
