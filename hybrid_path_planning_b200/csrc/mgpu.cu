// mgpu.cu -- the multi-GPU layer of the C ABI: one process per GPU, NCCL over NVLink.
//
// Only independent units are partitioned, exactly as the north_star names them:
//   * batches of planning queries (hpp_plan_batch_sharded) and pose batches (hpp_collide_circles_sharded): the unit
//     list is split contiguously over the ranks, every rank computes its slice with the single-GPU entry points, and
//     ONE collective (ncclAllGather of the plan headers / verdict bytes) returns the whole result to every rank;
//   * row bands of a large map for the distance layer (hpp_distance_layer_sharded): pass 1 of the EDT (row distances g)
//     is local to a band; pass 2 needs the columns across bands, so the bands exchange halo rows OF THE PASS-1 RESULT g
//     with their two neighbours (ncclSend / ncclRecv, one group) right after the row kernel, on the same stream.  A band
//     result is exact when no distance in it exceeds the distance to the first row beyond its halo; otherwise the
//     ranks agree (ncclAllReduce of one flag) to double the halo and repeat pass 2, and once the halo spans whole
//     neighbouring bands the g rows are all-gathered.
// A single heuristic field does not shard (replicas only): shard by goal / query instead.
//
// NCCL is reached through dlopen("libnccl.so.2") -- the library has no link-time dependency on it, a process that
// already loaded NCCL (e.g. through torch) shares that copy, and single-GPU users never touch it.  The communicator
// is created from a 128-byte unique id that rank 0 makes (hpp_comm_unique_id) and the caller ships to the other
// ranks by whatever means it has (MPI, a file, torch.distributed in bench.py: plumbing).
#include <dlfcn.h>
#include <nccl.h>  // types and enums only; every call goes through the table below

#include <vector>

#include "hpp_internal.cuh"

namespace {

struct nccl_api {
  void *handle;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char *(*GetErrorString)(ncclResult_t);
  char err[256];
};

nccl_api *nccl() {
  static nccl_api api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      snprintf(api.err, sizeof api.err, "dlopen(libnccl.so.2) failed: %s", dlerror());
      return &api;
    }
#define HPP_NCCL_SYM(field, sym)                                                   \
  *(void **)(&api.field) = dlsym(api.handle, sym);                                 \
  if (!api.field) {                                                                \
    snprintf(api.err, sizeof api.err, "libnccl has no symbol %s", sym);            \
    api.handle = nullptr;                                                          \
    return &api;                                                                   \
  }
    HPP_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    HPP_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    HPP_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    HPP_NCCL_SYM(Send, "ncclSend")
    HPP_NCCL_SYM(Recv, "ncclRecv")
    HPP_NCCL_SYM(AllGather, "ncclAllGather")
    HPP_NCCL_SYM(AllReduce, "ncclAllReduce")
    HPP_NCCL_SYM(GroupStart, "ncclGroupStart")
    HPP_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    HPP_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef HPP_NCCL_SYM
  }
  return &api;
}

}  // namespace

struct hpp_comm {
  hpp_ctx *ctx;
  ncclComm_t comm;
  int rank, world;
};

#define HPP_NCCL(c, call)                                                                              \
  do {                                                                                                 \
    ncclResult_t r_ = (call);                                                                          \
    if (r_ != ncclSuccess) {                                                                           \
      snprintf((c)->err, sizeof((c)->err), "%s:%d: %s -> %s", __FILE__, __LINE__, #call,               \
               nccl()->GetErrorString(r_));                                                            \
      return HPP_ERR_CUDA;                                                                             \
    }                                                                                                  \
  } while (0)

extern "C" void hpp_shard_range(size_t n, int world, int rank, size_t *lo, size_t *hi) {
  // contiguous split: sizes differ by at most one, the first n % world ranks are the larger ones
  const size_t base = n / (size_t)world, rem = n % (size_t)world;
  const size_t r = (size_t)rank;
  const size_t l = r * base + (r < rem ? r : rem);
  if (lo) *lo = l;
  if (hi) *hi = l + base + (r < rem ? 1 : 0);
}

extern "C" hpp_status hpp_comm_unique_id(uint8_t id[HPP_COMM_ID_BYTES]) {
  static_assert(sizeof(ncclUniqueId) == HPP_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  nccl_api *a = nccl();
  if (!id) return HPP_ERR_ARG;
  if (!a->handle) return HPP_ERR_UNSUPPORTED;
  ncclUniqueId u;
  if (a->GetUniqueId(&u) != ncclSuccess) return HPP_ERR_CUDA;
  memcpy(id, &u, sizeof u);
  return HPP_OK;
}

extern "C" hpp_status hpp_comm_create(hpp_ctx *c, const uint8_t id[HPP_COMM_ID_BYTES], int rank, int world,
                                      hpp_comm **out) {
  if (!c || !id || !out || world < 1 || rank < 0 || rank >= world) return HPP_ERR_ARG;
  nccl_api *a = nccl();
  if (!a->handle) HPP_FAIL(c, HPP_ERR_UNSUPPORTED, "NCCL is not available: %s", a->err);
  HPP_CUDA(c, cudaSetDevice(c->device));
  ncclUniqueId u;
  memcpy(&u, id, sizeof u);
  ncclComm_t comm;
  HPP_NCCL(c, a->CommInitRank(&comm, world, u, rank));
  hpp_comm *k = new hpp_comm;
  k->ctx = c;
  hpp_ctx_retain(c);  // released by hpp_comm_destroy
  k->comm = comm;
  k->rank = rank;
  k->world = world;
  *out = k;
  return HPP_OK;
}

extern "C" void hpp_comm_destroy(hpp_comm *k) {
  if (!k) return;
  cudaSetDevice(k->ctx->device);
  cudaStreamSynchronize(k->ctx->stream);
  nccl()->CommDestroy(k->comm);
  hpp_ctx *c = k->ctx;
  delete k;
  hpp_ctx_release(c);
}

extern "C" int hpp_comm_rank(const hpp_comm *k) { return k ? k->rank : -1; }
extern "C" int hpp_comm_size(const hpp_comm *k) { return k ? k->world : 0; }

// every rank contributes `count` units of `unit` bytes at units [lo, lo + count) of a list of n_total units held in
// HOST memory (`all`) and receives the units of all other ranks: one ncclAllGather over slots padded to the longest
// slice (a rank with an empty slice still takes part)
static hpp_status gather_slices_host(hpp_comm *k, void *all, size_t n_total, size_t unit) {
  hpp_ctx *c = k->ctx;
  nccl_api *a = nccl();
  if (n_total == 0) return HPP_OK;
  const size_t longest = (n_total + (size_t)k->world - 1) / (size_t)k->world;
  const size_t slot = (longest * unit + 15) & ~(size_t)15;
  void *d = nullptr;
  hpp_status s = hpp_scratch(c, 0, slot * (size_t)k->world, &d);
  if (s != HPP_OK) return s;
  size_t lo, hi;
  hpp_shard_range(n_total, k->world, k->rank, &lo, &hi);
  char *mine = (char *)d + slot * (size_t)k->rank;
  if (hi > lo) HPP_CUDA(c, cudaMemcpyAsync(mine, (const char *)all + lo * unit, (hi - lo) * unit, cudaMemcpyHostToDevice, c->stream));
  HPP_NCCL(c, a->AllGather(mine, d, slot, ncclUint8, k->comm, c->stream));  // in place: sendbuff = recvbuff + rank * slot
  for (int r = 0; r < k->world; r++) {
    if (r == k->rank) continue;
    size_t l, h;
    hpp_shard_range(n_total, k->world, r, &l, &h);
    if (h > l)
      HPP_CUDA(c, cudaMemcpyAsync((char *)all + l * unit, (const char *)d + slot * (size_t)r, (h - l) * unit,
                                  cudaMemcpyDeviceToHost, c->stream));
  }
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  return HPP_OK;
}

extern "C" hpp_status hpp_plan_batch_sharded(hpp_comm *k, hpp_map *const *maps, size_t n_maps, const uint32_t *map_of_query,
                                             int variant, const double *starts, const double *goals, size_t n_total,
                                             hpp_plan *plans_all, double *traj_local, size_t traj_cap,
                                             size_t *traj_used, size_t *lo_out, size_t *hi_out) {
  if (!k || !maps || n_maps == 0 || !maps[0] || maps[0]->ctx != k->ctx) return HPP_ERR_ARG;
  if (n_total && (!starts || !goals || !plans_all)) return HPP_ERR_ARG;
  size_t lo, hi;
  hpp_shard_range(n_total, k->world, k->rank, &lo, &hi);
  if (lo_out) *lo_out = lo;
  if (hi_out) *hi_out = hi;
  if (traj_used) *traj_used = 0;
  // queries [lo, hi) on this rank: no inter-GPU traffic during the search
  hpp_status s = hpp_plan_batch_multi(maps, n_maps, map_of_query ? map_of_query + lo : nullptr, variant, starts + 3 * lo,
                                      goals + 3 * lo, hi - lo, plans_all + lo, traj_local, traj_cap, traj_used);
  if (s != HPP_OK && s != HPP_ERR_CAPACITY) return s;
  // result gather: the plan headers of all ranks (trajectories stay on the rank that produced them)
  hpp_status g = gather_slices_host(k, plans_all, n_total, sizeof(hpp_plan));
  return g != HPP_OK ? g : s;
}

extern "C" hpp_status hpp_collide_circles_sharded(hpp_comm *k, hpp_map *m, const double *xyh, size_t n_total,
                                                  uint8_t *verdict_all) {
  if (!k || !m || m->ctx != k->ctx || (n_total && (!xyh || !verdict_all))) return HPP_ERR_ARG;
  size_t lo, hi;
  hpp_shard_range(n_total, k->world, k->rank, &lo, &hi);
  hpp_status s = hpp_collide_circles(m, xyh + 3 * lo, hi - lo, verdict_all + lo);
  if (s != HPP_OK) return s;
  return gather_slices_host(k, verdict_all, n_total, 1);
}

// all-gather of row bands of different heights (shard_range split of total_rows) into one contiguous buffer:
// grouped send / recv, every rank sends its band to every other rank
static hpp_status gather_bands_dev(hpp_comm *k, const void *band, void *full, size_t total_rows, size_t row_bytes) {
  hpp_ctx *c = k->ctx;
  nccl_api *a = nccl();
  size_t lo, hi;
  hpp_shard_range(total_rows, k->world, k->rank, &lo, &hi);
  if (hi > lo && (const char *)band != (const char *)full + lo * row_bytes)
    HPP_CUDA(c, cudaMemcpyAsync((char *)full + lo * row_bytes, band, (hi - lo) * row_bytes, cudaMemcpyDeviceToDevice,
                                c->stream));
  if (k->world == 1) return HPP_OK;
  HPP_NCCL(c, a->GroupStart());
  for (int r = 0; r < k->world; r++) {
    if (r == k->rank) continue;
    size_t l, h;
    hpp_shard_range(total_rows, k->world, r, &l, &h);
    if (hi > lo) HPP_NCCL(c, a->Send(band, (hi - lo) * row_bytes, ncclUint8, r, k->comm, c->stream));
    if (h > l) HPP_NCCL(c, a->Recv((char *)full + l * row_bytes, (h - l) * row_bytes, ncclUint8, r, k->comm, c->stream));
  }
  HPP_NCCL(c, a->GroupEnd());
  return HPP_OK;
}

__global__ void __launch_bounds__(256) flip_layer_kernel(const float *__restrict__ edt, float *__restrict__ layer, size_t n) {
  // grid_map layout of the reference: distance[k] = edt[n - 1 - k]  (RS/src/Grid_map.cpp:67, cv::flip(-1))
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    layer[n - 1 - i] = edt[i];
}

extern "C" hpp_status hpp_distance_layer_sharded(hpp_comm *k, const int8_t *occ_band_dev, uint32_t W, uint32_t total_rows,
                                                 float resolution, uint32_t halo_rows, float *distance_band_dev,
                                                 float *distance_full_dev, hpp_band_report *report) {
  if (!k || !occ_band_dev || !distance_band_dev || W == 0 || total_rows == 0) return HPP_ERR_ARG;
  hpp_ctx *c = k->ctx;
  nccl_api *a = nccl();
  if (W > 46340 || total_rows > 46340) HPP_FAIL(c, HPP_ERR_UNSUPPORTED, "EDT supports maps up to 46340 cells per side");
  if ((size_t)k->world > total_rows) HPP_FAIL(c, HPP_ERR_ARG, "more ranks (%d) than map rows (%u)", k->world, total_rows);
  HPP_CUDA(c, cudaSetDevice(c->device));
  size_t lo, hi;
  hpp_shard_range(total_rows, k->world, k->rank, &lo, &hi);
  const int rows = (int)(hi - lo);
  const int min_rows = (int)(total_rows / (uint32_t)k->world);  // the shortest band: the same number on every rank
  const bool up = k->rank > 0, dn = k->rank < k->world - 1;
  int halo = k->world == 1 ? 0 : (int)(halo_rows == 0 ? 64 : halo_rows);
  if (halo > min_rows) halo = min_rows;  // rank-independent clamp: every exchange moves the same number of rows
  // scratch: g of the band with room for the widest halo on both sides (slot 2); the flag word lives in d_flags[44]
  const size_t gcap = ((size_t)rows + 2 * (size_t)min_rows) * W;
  uint16_t *g_ext = nullptr;
  hpp_status s = hpp_scratch(c, 2, gcap * sizeof(uint16_t) + 256, (void **)&g_ext);
  if (s != HPP_OK) return s;
  uint16_t *g_band = g_ext + (size_t)(up ? min_rows : 0) * W;  // fixed place: halos grow outwards around it
  uint32_t *d_flag = c->d_flags + 44;
  hpp_band_report rep;
  memset(&rep, 0, sizeof rep);
  // pass 1 on the band
  s = hpp_edt_band_rows(c, occ_band_dev, (int)W, rows, g_band);
  if (s != HPP_OK) return s;
  bool done = false;
  while (!done) {
    rep.rounds++;
    rep.halo_rows = (uint32_t)halo;
    // halo exchange of g rows with the two neighbouring bands, fused behind the row kernel on the same stream
    if (k->world > 1 && halo > 0) {
      const size_t bytes = (size_t)halo * W * sizeof(uint16_t);
      HPP_NCCL(c, a->GroupStart());
      if (up) {
        HPP_NCCL(c, a->Send(g_band, bytes, ncclUint8, k->rank - 1, k->comm, c->stream));
        HPP_NCCL(c, a->Recv(g_band - (size_t)halo * W, bytes, ncclUint8, k->rank - 1, k->comm, c->stream));
      }
      if (dn) {
        HPP_NCCL(c, a->Send(g_band + (size_t)(rows - halo) * W, bytes, ncclUint8, k->rank + 1, k->comm, c->stream));
        HPP_NCCL(c, a->Recv(g_band + (size_t)rows * W, bytes, ncclUint8, k->rank + 1, k->comm, c->stream));
      }
      HPP_NCCL(c, a->GroupEnd());
      rep.halo_bytes_sent += (uint64_t)bytes * ((up ? 1 : 0) + (dn ? 1 : 0));
    }
    const int h_up = up ? halo : 0, h_dn = dn ? halo : 0;
    // rows of the map that the buffer does not hold: above iff the halo does not reach row 0, below likewise
    const int more_above = up && (size_t)h_up < lo, more_below = dn && hi + (size_t)h_dn < total_rows;
    HPP_CUDA(c, cudaMemsetAsync(d_flag, 0, sizeof(uint32_t), c->stream));
    s = hpp_edt_band_cols(c, g_band - (size_t)h_up * W, (int)W, rows + h_up + h_dn, h_up, rows, more_above, more_below,
                          resolution, distance_band_dev, d_flag);
    if (s != HPP_OK) return s;
    if (k->world > 1) HPP_NCCL(c, a->AllReduce(d_flag, d_flag, 1, ncclUint32, ncclMax, k->comm, c->stream));
    HPP_CUDA(c, cudaMemcpyAsync(c->h_flags + 44, d_flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    HPP_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->h_flags[44] == 0) {
      done = true;
    } else if (halo < min_rows) {
      halo = halo * 2 > min_rows ? min_rows : halo * 2;  // some distance reaches past the halo: widen and redo pass 2
    } else {
      // the halo already spans whole neighbouring bands: gather all g rows and run pass 2 on the full columns
      uint16_t *g_full = nullptr;
      s = hpp_scratch(c, 4, (size_t)total_rows * W * sizeof(uint16_t) + 256, (void **)&g_full);
      if (s != HPP_OK) return s;
      s = gather_bands_dev(k, g_band, g_full, total_rows, (size_t)W * sizeof(uint16_t));
      if (s != HPP_OK) return s;
      s = hpp_edt_band_cols(c, g_full, (int)W, (int)total_rows, (int)lo, rows, 0, 0, resolution, distance_band_dev, d_flag);
      if (s != HPP_OK) return s;
      rep.gathered_all = 1;
      rep.halo_bytes_sent += (uint64_t)rows * W * sizeof(uint16_t) * (uint64_t)(k->world - 1);
      done = true;
    }
  }
  if (distance_full_dev) {
    // result gather: every rank receives the whole layer, in the reference's (flipped) grid_map layout
    const size_t n = (size_t)total_rows * W;
    float *edt = nullptr;
    s = hpp_scratch(c, 3, n * sizeof(float) + 256, (void **)&edt);
    if (s != HPP_OK) return s;
    s = gather_bands_dev(k, distance_band_dev, edt, total_rows, (size_t)W * sizeof(float));
    if (s != HPP_OK) return s;
    flip_layer_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(edt, distance_full_dev, n);
    HPP_LAUNCHED(c);
  }
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  if (report) *report = rep;
  return HPP_OK;
}
