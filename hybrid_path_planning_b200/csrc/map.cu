// map.cu -- context + device-resident map state (replaces the Grid_map constructor,
// RS/src/Grid_map.cpp:3-81, DB/src/Grid_map.cpp:3-21, and Grid_map::setcarData, RS/src/Grid_map.cpp:217).
#include <math.h>
#include <stdlib.h>

#include "hpp_internal.cuh"

extern "C" const char *hpp_version(void) { return "hpp_b200 0.1 (sm_100a)"; }

extern "C" hpp_status hpp_ctx_create(int device, hpp_ctx **out) {
  if (!out) return HPP_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return HPP_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return HPP_ERR_CUDA;
  if (prop.major != 10) return HPP_ERR_CUDA;  // sm_100a cubins only; no fallback path exists
  hpp_ctx *c = (hpp_ctx *)calloc(1, sizeof(hpp_ctx));
  if (!c) return HPP_ERR_NOMEM;
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMallocHost((void **)&c->h_flags, 64 * sizeof(uint32_t)) != cudaSuccess ||
      cudaMalloc((void **)&c->d_flags, 64 * sizeof(uint32_t)) != cudaSuccess) {
    free(c);
    return HPP_ERR_CUDA;
  }
  c->refs = 1;
  *out = c;
  return HPP_OK;
}

void hpp_ctx_retain(hpp_ctx *c) { __atomic_add_fetch(&c->refs, 1, __ATOMIC_RELAXED); }

void hpp_ctx_release(hpp_ctx *c) {
  if (__atomic_sub_fetch(&c->refs, 1, __ATOMIC_ACQ_REL) > 0) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (int i = 0; i < 8; i++)
    if (c->d_scratch[i]) cudaFree(c->d_scratch[i]);
  for (int i = 0; i < 2; i++)
    if (c->h_pinned[i]) cudaFreeHost(c->h_pinned[i]);
  cudaFreeHost(c->h_flags);
  cudaFree(c->d_flags);
  if (c->copy_stream) {
    cudaStreamDestroy(c->copy_stream);
    cudaEventDestroy(c->copy_done[0]);
    cudaEventDestroy(c->copy_done[1]);
  }
  cudaStreamDestroy(c->stream);
  free(c);
}

extern "C" hpp_status hpp_ctx_destroy(hpp_ctx *c) {
  if (!c) return HPP_ERR_ARG;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  hpp_ctx_release(c);  // the resources go now, or with the last map / communicator still alive
  return HPP_OK;
}

extern "C" hpp_status hpp_ctx_sync(hpp_ctx *c) {
  if (!c) return HPP_ERR_ARG;
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  return HPP_OK;
}
extern "C" void *hpp_ctx_stream(hpp_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" const char *hpp_last_error(hpp_ctx *c) { return c ? c->err : "null ctx"; }
extern "C" uint64_t hpp_ctx_launch_count(hpp_ctx *c) { return c ? c->launches : 0; }

hpp_status hpp_scratch(hpp_ctx *c, int slot, size_t bytes, void **out) {
  if (c->d_scratch_bytes[slot] < bytes) {
    if (c->d_scratch[slot]) {
      HPP_CUDA(c, cudaStreamSynchronize(c->stream));
      cudaFree(c->d_scratch[slot]);
      c->d_scratch[slot] = nullptr;
      c->d_scratch_bytes[slot] = 0;
    }
    HPP_CUDA(c, cudaMalloc(&c->d_scratch[slot], bytes));
    c->d_scratch_bytes[slot] = bytes;
  }
  *out = c->d_scratch[slot];
  return HPP_OK;
}

hpp_status hpp_pinned(hpp_ctx *c, int slot, size_t bytes, void **out) {
  if (c->h_pinned_bytes[slot] < bytes) {
    if (c->h_pinned[slot]) {
      HPP_CUDA(c, cudaStreamSynchronize(c->stream));
      cudaFreeHost(c->h_pinned[slot]);
      c->h_pinned[slot] = nullptr;
      c->h_pinned_bytes[slot] = 0;
    }
    HPP_CUDA(c, cudaMallocHost(&c->h_pinned[slot], bytes));
    c->h_pinned_bytes[slot] = bytes;
  }
  *out = c->h_pinned[slot];
  return HPP_OK;
}

// ------------------------------------------------------------------ per-cell tables
// One pass over the occupancy: class byte (blocked / near-obstacle / 3x3 collision count) used by both
// heuristics, and the bit-packed "data > 0" rows used by the SAT window scan.
__global__ void __launch_bounds__(256) build_cls_kernel(const int8_t *__restrict__ occ, int W, int H,
                                                        uint8_t *__restrict__ cls, uint32_t *__restrict__ bits,
                                                        uint32_t *__restrict__ near_bits, int pitch) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  bool in = x < W;
  bool blocked = false, near = false;
  if (in) {
    auto col = [&](int xx, int yy) -> int {
      if (xx < 0 || xx >= W || yy < 0 || yy >= H) return 1;  // isInCollision: out of bounds = collision
      return occ[(size_t)yy * W + xx] > 0;
    };
    auto inb_col = [&](int xx, int yy) -> int {  // RS near-obstacle test requires in-bounds (hazard note A.3)
      if (xx < 0 || xx >= W || yy < 0 || yy >= H) return 0;
      return occ[(size_t)yy * W + xx] > 0;
    };
    int cnt = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
      for (int dx = -1; dx <= 1; dx++) cnt += col(x + dx, y + dy);
    int near4 = inb_col(x - 1, y) | inb_col(x, y - 1) | inb_col(x + 1, y) | inb_col(x, y + 1);
    blocked = occ[(size_t)y * W + x] > 0;
    near = near4 != 0;
    cls[(size_t)y * W + x] = (uint8_t)((blocked ? CLS_BLOCKED : 0) | (near4 ? CLS_NEAR4 : 0) | cnt);
  }
  unsigned m = __ballot_sync(0xffffffffu, in && blocked);
  unsigned mn = __ballot_sync(0xffffffffu, in && near);
  if ((threadIdx.x & 31) == 0 && (x >> 5) < pitch) {
    bits[(size_t)y * pitch + (x >> 5)] = m;
    near_bits[(size_t)y * pitch + (x >> 5)] = mn;
  }
}

// Same tables, four consecutive cells per thread (needs W % 4 == 0): three rows x (one aligned 32-bit word + the two
// bytes beside it) instead of 13 bounds-checked byte loads per cell; the class bytes leave as one 32-bit store and
// the bit rows are assembled from four ballots (lane l holds cells 4l..4l+3 of a 128-cell span = 4 words).
__device__ __forceinline__ uint32_t spread8(uint32_t b) {  // bit i of the low byte -> bit 4i
  b &= 0xffu;
  b = (b | (b << 12)) & 0x000F000Fu;
  b = (b | (b << 6)) & 0x03030303u;
  b = (b | (b << 3)) & 0x11111111u;
  return b;
}

__global__ void __launch_bounds__(256) build_cls4_kernel(const int8_t *__restrict__ occ, int W, int H,
                                                         uint8_t *__restrict__ cls, uint32_t *__restrict__ bits,
                                                         uint32_t *__restrict__ near_bits, int pitch) {
  const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int y = blockIdx.y;
  const bool in = x0 < W;
  uint32_t blocked4 = 0, near4m = 0;
  if (in) {
    // per row: bit c (c = 0..5) <-> column x0 - 1 + c.  cb: isInCollision (out of bounds = 1), ib: in bounds and data > 0
    uint32_t cb[3], ib[3];
#pragma unroll
    for (int dy = -1; dy <= 1; dy++) {
      const int yy = y + dy;
      uint32_t c = 0x3fu, i = 0;
      if (yy >= 0 && yy < H) {
        const int8_t *row = occ + (size_t)yy * W;
        const uint32_t w = *(const uint32_t *)(row + x0);
        i = ((int8_t)(w & 0xff) > 0 ? 2u : 0u) | ((int8_t)((w >> 8) & 0xff) > 0 ? 4u : 0u) |
            ((int8_t)((w >> 16) & 0xff) > 0 ? 8u : 0u) | ((int8_t)(w >> 24) > 0 ? 16u : 0u);
        c = i;
        if (x0 > 0) {
          if (row[x0 - 1] > 0) i |= 1u, c |= 1u;
        } else {
          c |= 1u;
        }
        if (x0 + 4 < W) {
          if (row[x0 + 4] > 0) i |= 32u, c |= 32u;
        } else {
          c |= 32u;
        }
      }
      cb[dy + 1] = c;
      ib[dy + 1] = i;
    }
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t cnt = __popc((cb[0] >> k) & 7u) + __popc((cb[1] >> k) & 7u) + __popc((cb[2] >> k) & 7u);
      const uint32_t nr = ((ib[1] >> k) | (ib[1] >> (k + 2)) | (ib[0] >> (k + 1)) | (ib[2] >> (k + 1))) & 1u;
      const uint32_t bl = (ib[1] >> (k + 1)) & 1u;
      out |= ((bl ? CLS_BLOCKED : 0u) | (nr ? CLS_NEAR4 : 0u) | cnt) << (8 * k);
      blocked4 |= bl << k;
      near4m |= nr << k;
    }
    *(uint32_t *)(cls + (size_t)y * W + x0) = out;
  }
  uint32_t wb = 0, wn = 0;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint32_t bb = __ballot_sync(0xffffffffu, (blocked4 >> k) & 1u);
    const uint32_t bn = __ballot_sync(0xffffffffu, (near4m >> k) & 1u);
    if (lane < 4) {
      wb |= spread8(bb >> (8 * lane)) << k;
      wn |= spread8(bn >> (8 * lane)) << k;
    }
  }
  const int word = ((x0 - 4 * lane) >> 5) + lane;  // first word of the warp's 128-cell span + my word
  if (lane < 4 && word < pitch) {
    bits[(size_t)y * pitch + word] = wb;
    near_bits[(size_t)y * pitch + word] = wn;
  }
}

hpp_status hpp_build_cls(hpp_map *m) {
  hpp_ctx *c = m->ctx;
  if ((m->g.W & 3) == 0) {
    dim3 grid(hpp_div_up(m->g.W, 4 * 256), m->g.H);
    build_cls4_kernel<<<grid, 256, 0, c->stream>>>(m->d_occ, m->g.W, m->g.H, m->d_cls, m->d_occ_bits, m->d_near_bits,
                                                   m->bits_pitch);
  } else {
    dim3 grid(hpp_div_up(m->g.W, 256), m->g.H);
    build_cls_kernel<<<grid, 256, 0, c->stream>>>(m->d_occ, m->g.W, m->g.H, m->d_cls, m->d_occ_bits, m->d_near_bits,
                                                  m->bits_pitch);
  }
  HPP_LAUNCHED(c);
  return HPP_OK;
}

// ------------------------------------------------------------------ vehicle constants (host, glibc)
static void setup_car(hpp_car_dev *car, const hpp_params *p) {
  car->p = *p;
  // RS/src/CarData.cpp:84-94
  double shift = p->width / 4.0;
  car->cr = sqrt(2 * pow(shift, 2));
  car->cx[0] = -p->axle_to_back + shift; car->cy[0] = p->width / 2.0;
  car->cx[1] = -p->axle_to_back + shift; car->cy[1] = -p->width / 2.0;
  car->cx[2] = p->axle_to_front - shift; car->cy[2] = p->width / 2.0;
  car->cx[3] = p->axle_to_front - shift; car->cy[3] = -p->width / 2.0;
  // DB/src/CarData.cpp:15-37: Point32 (float32) corners, closed polygon
  car->rx[0] = (float)p->axle_to_front;   car->ry[0] = (float)(p->width / 2.0);
  car->rx[1] = (float)p->axle_to_front;   car->ry[1] = (float)(-p->width / 2.0);
  car->rx[2] = (float)(-p->axle_to_back); car->ry[2] = (float)(-p->width / 2.0);
  car->rx[3] = (float)(-p->axle_to_back); car->ry[3] = (float)(p->width / 2.0);
  car->rx[4] = car->rx[0];                car->ry[4] = car->ry[0];
  // RS/src/HybridAstar.cpp:14-23 / DB/src/HybridAstar.cpp:14-21, steer_step = maxSteerAngle / 6 (CarData.h:21-22)
  double steer_step = p->max_steer / 6.0;
  for (int v = 0; v < 2; v++) {
    int n = 0;
    for (double i = p->max_steer; i >= -p->max_steer && n + 2 <= HPP_MAX_COMMANDS; i -= steer_step) {
      car->cmd_steer[v][n] = i;
      car->cmd_dir[v][n++] = 1.0;
      if (v == HPP_VARIANT_RS) {
        car->cmd_steer[v][n] = i;
        car->cmd_dir[v][n++] = -1.0;
      }
    }
    car->n_cmd[v] = n;
  }
  car->maxc = tan(p->max_steer) / p->wheel_base;
}

extern "C" int32_t hpp_num_commands(const hpp_params *params, int variant) {
  if (!params || variant < 0 || variant > 1) return 0;
  hpp_car_dev car;
  setup_car(&car, params);
  return car.n_cmd[variant];
}

static hpp_status map_alloc(hpp_ctx *c, uint32_t W, uint32_t H, float resolution, double ox, double oy,
                            const hpp_params *params, hpp_map **out) {
  if (!c || !out || !params || W == 0 || H == 0) return HPP_ERR_ARG;
  if (W >= 65535u || H >= 65535u) HPP_FAIL(c, HPP_ERR_UNSUPPORTED, "map side %u x %u >= 65535 not supported", W, H);
  if (params->path_length > HPP_MAX_SIM) HPP_FAIL(c, HPP_ERR_UNSUPPORTED, "pathLength > %d", HPP_MAX_SIM);
  HPP_CUDA(c, cudaSetDevice(c->device));
  hpp_map *m = (hpp_map *)calloc(1, sizeof(hpp_map));
  if (!m) return HPP_ERR_NOMEM;
  m->ctx = c;
  hpp_ctx_retain(c);  // released by hpp_map_destroy
  m->n = (size_t)W * H;
  m->g.W = (int)W;
  m->g.H = (int)H;
  m->g.res_f = resolution;
  m->g.res = (double)resolution;
  m->g.rcp_res = 1.0 / m->g.res;
  m->g.ox = ox;
  m->g.oy = oy;
  {  // RS/src/Grid_map.cpp:15-25 + grid_map_core setGeometry (size = round(length/res), length = size*res)
    double length_x = W * m->g.res, length_y = H * m->g.res;
    m->g.posx = ox + length_x / 2.0;
    m->g.posy = oy + length_y / 2.0;
    m->g.lenx = (double)(int)round(length_x / m->g.res) * m->g.res;
    m->g.leny = (double)(int)round(length_y / m->g.res) * m->g.res;
  }
  setup_car(&m->car, params);
  m->bits_pitch = hpp_div_up(W, 32);
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_occ, m->n);
  if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_cls, m->n);
  if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_occ_bits, (size_t)m->bits_pitch * H * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_near_bits, (size_t)m->bits_pitch * H * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_g, m->n * sizeof(uint16_t));
  if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_obstacle, m->n * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_distance, m->n * sizeof(float));
  m->d2_tiles_x = hpp_div_up(W, HPP_D2_TILE_W);
  if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_d2u8, (size_t)m->d2_tiles_x * hpp_div_up(H, HPP_D2_TILE_H) * 32);
  if (e != cudaSuccess) {
    snprintf(c->err, sizeof(c->err), "map allocation failed: %s", cudaGetErrorString(e));
    hpp_map_destroy(m);
    return HPP_ERR_NOMEM;
  }
  *out = m;
  return HPP_OK;
}

static hpp_status map_build(hpp_map *m) {
  hpp_status s = hpp_build_cls(m);
  if (s != HPP_OK) return s;
  return hpp_edt_build(m);
}

extern "C" hpp_status hpp_map_create(hpp_ctx *c, const int8_t *occ, uint32_t W, uint32_t H, float resolution,
                                     double ox, double oy, const hpp_params *params, hpp_map **out) {
  if (!occ) return HPP_ERR_ARG;
  hpp_status s = map_alloc(c, W, H, resolution, ox, oy, params, out);
  if (s != HPP_OK) return s;
  s = hpp_map_update(*out, occ);
  if (s != HPP_OK) {
    hpp_map_destroy(*out);
    *out = nullptr;
  }
  return s;
}

extern "C" hpp_status hpp_map_create_dev(hpp_ctx *c, const int8_t *occ_dev, uint32_t W, uint32_t H, float resolution,
                                         double ox, double oy, const hpp_params *params, hpp_map **out) {
  if (!occ_dev) return HPP_ERR_ARG;
  hpp_status s = map_alloc(c, W, H, resolution, ox, oy, params, out);
  if (s != HPP_OK) return s;
  s = hpp_map_update_dev(*out, occ_dev);
  if (s != HPP_OK) {
    hpp_map_destroy(*out);
    *out = nullptr;
  }
  return s;
}

extern "C" hpp_status hpp_map_update(hpp_map *m, const int8_t *occ) {
  if (!m || !occ) return HPP_ERR_ARG;
  hpp_ctx *c = m->ctx;
  HPP_CUDA(c, cudaSetDevice(c->device));
  HPP_CUDA(c, cudaMemcpyAsync(m->d_occ, occ, m->n, cudaMemcpyHostToDevice, c->stream));
  hpp_status s = map_build(m);
  if (s != HPP_OK) return s;
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  return HPP_OK;
}

extern "C" hpp_status hpp_map_update_dev(hpp_map *m, const int8_t *occ_dev) {
  if (!m || !occ_dev) return HPP_ERR_ARG;
  hpp_ctx *c = m->ctx;
  HPP_CUDA(c, cudaSetDevice(c->device));
  HPP_CUDA(c, cudaMemcpyAsync(m->d_occ, occ_dev, m->n, cudaMemcpyDeviceToDevice, c->stream));
  return map_build(m);
}

extern "C" hpp_status hpp_map_set_params(hpp_map *m, const hpp_params *params) {
  if (!m || !params) return HPP_ERR_ARG;
  if (params->path_length > HPP_MAX_SIM) HPP_FAIL(m->ctx, HPP_ERR_UNSUPPORTED, "pathLength > %d", HPP_MAX_SIM);
  setup_car(&m->car, params);
  return HPP_OK;
}

extern "C" hpp_status hpp_map_destroy(hpp_map *m) {
  if (!m) return HPP_ERR_ARG;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  cudaFree(m->d_occ);
  cudaFree(m->d_cls);
  cudaFree(m->d_occ_bits);
  cudaFree(m->d_near_bits);
  cudaFree(m->d_g);
  cudaFree(m->d_obstacle);
  cudaFree(m->d_distance);
  cudaFree(m->d_d2u8);
  hpp_ctx *c = m->ctx;
  free(m);
  hpp_ctx_release(c);
  return HPP_OK;
}

extern "C" hpp_status hpp_map_layers(hpp_map *m, float *obstacle, float *distance) {
  if (!m) return HPP_ERR_ARG;
  hpp_ctx *c = m->ctx;
  HPP_CUDA(c, cudaSetDevice(c->device));
  if (obstacle)
    HPP_CUDA(c, cudaMemcpyAsync(obstacle, m->d_obstacle, m->n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (distance)
    HPP_CUDA(c, cudaMemcpyAsync(distance, m->d_distance, m->n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  return HPP_OK;
}

extern "C" hpp_status hpp_map_layers_dev(hpp_map *m, const float **obstacle_dev, const float **distance_dev) {
  if (!m) return HPP_ERR_ARG;
  if (obstacle_dev) *obstacle_dev = m->d_obstacle;
  if (distance_dev) *distance_dev = m->d_distance;
  return HPP_OK;
}

// ------------------------------------------------------------------ cell predicates (a2)
// Grid_map::toCellID (:186-192, truncation toward zero), isPointInBounds (:200-203), isInCollision (:205-213)
__global__ void __launch_bounds__(256) cells_kernel(const double *__restrict__ xy, size_t n, hpp_geom g,
                                                    const int8_t *__restrict__ occ, int32_t *__restrict__ cell,
                                                    uint8_t *__restrict__ flags) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cx = (int)hpp_div_res(g, xy[2 * i] - g.ox);
  int cy = (int)hpp_div_res(g, xy[2 * i + 1] - g.oy);
  bool inb = cx >= 0 && cx < g.W && cy >= 0 && cy < g.H;
  bool col = !inb || occ[(size_t)cy * g.W + cx] > 0;
  cell[2 * i] = cx;
  cell[2 * i + 1] = cy;
  flags[i] = (uint8_t)((inb ? 1 : 0) | (col ? 2 : 0));
}

extern "C" hpp_status hpp_map_cells(hpp_map *m, const double *xy, size_t n, int32_t *cell_xy, uint8_t *flags) {
  if (!m || !xy || !cell_xy || !flags) return HPP_ERR_ARG;
  if (n == 0) return HPP_OK;
  hpp_ctx *c = m->ctx;
  HPP_CUDA(c, cudaSetDevice(c->device));
  void *d_in, *d_out;
  hpp_status s = hpp_scratch(c, 0, n * 16, &d_in);
  if (s != HPP_OK) return s;
  s = hpp_scratch(c, 1, n * 9, &d_out);
  if (s != HPP_OK) return s;
  int32_t *d_cell = (int32_t *)d_out;
  uint8_t *d_fl = (uint8_t *)d_out + n * 8;
  HPP_CUDA(c, cudaMemcpyAsync(d_in, xy, n * 16, cudaMemcpyHostToDevice, c->stream));
  cells_kernel<<<hpp_div_up((long long)n, 256), 256, 0, c->stream>>>((const double *)d_in, n, m->g, m->d_occ, d_cell, d_fl);
  HPP_LAUNCHED(c);
  HPP_CUDA(c, cudaMemcpyAsync(cell_xy, d_cell, n * 8, cudaMemcpyDeviceToHost, c->stream));
  HPP_CUDA(c, cudaMemcpyAsync(flags, d_fl, n, cudaMemcpyDeviceToHost, c->stream));
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  return HPP_OK;
}
