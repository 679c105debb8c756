// expand.cu -- K6: batched successor expansion (see expand.cuh for the reference citations).
//   RS     : one warp per parent, lane = motion command (13 steer angles x {forward, reverse} = 26 lanes);
//            every lane rolls its command out for simulationLength Euler steps and runs the circle-footprint
//            check on each state.
//   Dubins : one warp per (parent, command); the lanes share the SAT window scan of each rolled-out state.
#include "expand.cuh"

struct circ_rect {
  circ_consts cc;
  float rx[5], ry[5];
};

__global__ void __launch_bounds__(256) expand_rs_kernel(hpp_geom g, expand_consts ec, circ_consts cc,
                                                        const float *__restrict__ dist,
                                                        const hpp_node *__restrict__ parents, size_t n,
                                                        hpp_child *__restrict__ children) {
  const int lane = threadIdx.x & 31;
  const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t p = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < n; p += warps) {
    if (lane >= ec.n_cmd) continue;
    const hpp_node nd = parents[p];
    hpp_child *ch = children + p * ec.n_cmd + lane;
    double x = nd.x, y = nd.y, h = nd.heading;
    int hit = 0;
    for (int i = 0; i < ec.sim; i++) {
      vehicle_step(x, y, h, ec.dir[lane], ec.step, ec.dtheta[lane]);
      ch->traj[i][0] = x;
      ch->traj[i][1] = y;
      ch->traj[i][2] = h;
      ch->traj[i][3] = (double)(int)(hpp_div_res(g, x - g.ox));  // toCellID (RS/src/HybridAstar.cpp:38-40)
      ch->traj[i][4] = (double)(int)(hpp_div_res(g, y - g.oy));
      if (!hit) hit = circles_collide(g, cc, dist, x, y, h);
    }
    ch->cost = path_cost(ec, lane, nd.cost, nd.steer, nd.direction);
    ch->steer = ec.steer[lane];
    ch->direction = (int)ec.dir[lane];
    ch->valid = !hit;
  }
}

struct rect_consts {
  float rx[5], ry[5];
};

__global__ void __launch_bounds__(256) expand_db_kernel(hpp_geom g, expand_consts ec, rect_consts rc,
                                                        const uint32_t *__restrict__ bits, int pitch,
                                                        const hpp_node *__restrict__ parents, size_t n,
                                                        hpp_child *__restrict__ children) {
  const int lane = threadIdx.x & 31;
  const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const size_t total = n * ec.n_cmd;
  for (size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += warps) {
    const size_t p = w / ec.n_cmd;
    const int cmd = (int)(w - p * ec.n_cmd);
    const hpp_node nd = parents[p];
    hpp_child *ch = children + w;
    double x = nd.x, y = nd.y, h = nd.heading;
    int hit = 0;
    for (int i = 0; i < ec.sim; i++) {
      vehicle_step(x, y, h, ec.dir[cmd], ec.step, ec.dtheta[cmd]);
      if (lane == 0) {
        ch->traj[i][0] = x;
        ch->traj[i][1] = y;
        ch->traj[i][2] = h;
        ch->traj[i][3] = (double)(int)(hpp_div_res(g, x - g.ox));
        ch->traj[i][4] = (double)(int)(hpp_div_res(g, y - g.oy));
      }
      if (!hit) hit = sat_collide_warp(g, rc.rx, rc.ry, bits, pitch, x, y, h);
    }
    if (lane == 0) {
      ch->cost = path_cost(ec, cmd, nd.cost, nd.steer, nd.direction);
      ch->steer = ec.steer[cmd];
      ch->direction = (int)ec.dir[cmd];
      ch->valid = !hit;
    }
  }
}

void hpp_make_expand_consts(const hpp_map *m, int variant, expand_consts *ec);  // shots.cu shares it

void hpp_make_expand_consts(const hpp_map *m, int variant, expand_consts *ec) {
  memset(ec, 0, sizeof *ec);
  ec->n_cmd = m->car.n_cmd[variant];
  ec->sim_len = m->car.p.path_length;
  ec->sim = 0;
  while (ec->sim < m->car.p.path_length && ec->sim < HPP_MAX_SIM) ec->sim++;  // for (int i = 0; i < simulationLength; ++i)
  ec->step = m->car.p.step_car;
  for (int k = 0; k < ec->n_cmd; k++) {
    ec->steer[k] = m->car.cmd_steer[variant][k];
    ec->dir[k] = m->car.cmd_dir[variant][k];
    ec->dtheta[k] = tan(ec->steer[k]) / m->car.p.wheel_base;  // host libm: same value as the reference's
  }
  ec->dir_change = (variant == HPP_VARIANT_RS) ? 250.0 : 200.0;
}

extern "C" hpp_status hpp_expand_dev(hpp_map *m, int variant, const hpp_node *parents_dev, size_t n,
                                     hpp_child *children_dev) {
  if (!m || variant < 0 || variant > 1 || (n && (!parents_dev || !children_dev))) return HPP_ERR_ARG;
  if (n == 0) return HPP_OK;
  hpp_ctx *c = m->ctx;
  HPP_CUDA(c, cudaSetDevice(c->device));
  expand_consts ec;
  hpp_make_expand_consts(m, variant, &ec);
  if (variant == HPP_VARIANT_RS) {
    const circ_consts cc = hpp_make_circ_consts(m);
    long long want = hpp_div_up((long long)n, 8);
    int grid = (int)(want < (long long)c->sm_count * 8 ? want : (long long)c->sm_count * 8);
    expand_rs_kernel<<<grid, 256, 0, c->stream>>>(m->g, ec, cc, m->d_distance, parents_dev, n, children_dev);
  } else {
    rect_consts rc;
    for (int k = 0; k < 5; k++) {
      rc.rx[k] = m->car.rx[k];
      rc.ry[k] = m->car.ry[k];
    }
    long long want = hpp_div_up((long long)n * ec.n_cmd, 8);
    int grid = (int)(want < (long long)c->sm_count * 8 ? want : (long long)c->sm_count * 8);
    expand_db_kernel<<<grid, 256, 0, c->stream>>>(m->g, ec, rc, m->d_occ_bits, m->bits_pitch, parents_dev, n,
                                                   children_dev);
  }
  HPP_LAUNCHED(c);
  return HPP_OK;
}

extern "C" hpp_status hpp_expand(hpp_map *m, int variant, const hpp_node *parents, size_t n, hpp_child *children) {
  if (!m || variant < 0 || variant > 1 || (n && (!parents || !children))) return HPP_ERR_ARG;
  if (n == 0) return HPP_OK;
  hpp_ctx *c = m->ctx;
  HPP_CUDA(c, cudaSetDevice(c->device));
  const int nc = m->car.n_cmd[variant];
  void *d_in, *d_out;
  hpp_status s = hpp_scratch(c, 0, n * sizeof(hpp_node), &d_in);
  if (s != HPP_OK) return s;
  s = hpp_scratch(c, 1, n * nc * sizeof(hpp_child), &d_out);
  if (s != HPP_OK) return s;
  HPP_CUDA(c, cudaMemcpyAsync(d_in, parents, n * sizeof(hpp_node), cudaMemcpyHostToDevice, c->stream));
  s = hpp_expand_dev(m, variant, (const hpp_node *)d_in, n, (hpp_child *)d_out);
  if (s != HPP_OK) return s;
  HPP_CUDA(c, cudaMemcpyAsync(children, d_out, n * nc * sizeof(hpp_child), cudaMemcpyDeviceToHost, c->stream));
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  return HPP_OK;
}
