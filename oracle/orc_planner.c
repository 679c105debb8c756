/*
 * oracle/orc_planner.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 * Restates a5/a10 (motionCommands, GetnextNeighbours, PathCost) and the host loop HybridAstar::run
 * (RS/src/HybridAstar.cpp:14-107, :553-729; DB/src/HybridAstar.cpp:14-105, :315-496).
 */
#include "orc.h"
#include "orc_stdheap.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI_A 3.14159 /* RS/include/HybridAstar.h:24 -- the A* goal test uses this, not M_PI */

/* RS/src/HybridAstar.cpp:14-23, DB/src/HybridAstar.cpp:14-21; steer_step = maxSteerAngle / 6
 * (RS/include/CarData.h:21-22).  The loop variable accumulates in fp64 exactly as the reference's. */
int orc_motion_commands(const orc_car *car, int variant, double *steer, double *dir) {
  double steer_ratio = 6;
  double steer_step = car->max_steer / steer_ratio;
  double i;
  int n = 0;
  for (i = car->max_steer; i >= -car->max_steer; i -= steer_step) {
    if (n + 2 > 64) break;
    steer[n] = i;
    dir[n++] = 1.0;
    if (variant == ORC_VARIANT_RS) {
      steer[n] = i;
      dir[n++] = -1.0;
    }
  }
  return n;
}

/* RS/src/HybridAstar.cpp:25-107 / DB/src/HybridAstar.cpp:23-105 with CarData::getVehicleStep
 * (RS/src/CarData.cpp:63-74). */
int orc_expand(const orc_map *m, const orc_car *car, int variant, double sim_len, double step,
               const orc_state *parent, double parent_cost, double parent_steer, int parent_dir,
               orc_child *children) {
  const double reverse = 10.0, steerAngle = 15.0, steerAngleChange = 20.0;
  const double directionChange = (variant == ORC_VARIANT_RS) ? 250.0 : 200.0;
  double steer[64], dir[64];
  int ncmd = orc_motion_commands(car, variant, steer, dir), c, i;
  for (c = 0; c < ncmd; c++) {
    orc_child *ch = &children[c];
    orc_state cur = *parent;
    int has_collision = 0;
    memset(ch, 0, sizeof *ch);
    for (i = 0; i < sim_len && i < 8; ++i) {
      double dx = cos(cur.heading), dy = sin(cur.heading);
      double dtheta = tan(steer[c]) / car->wheel_base;
      orc_state nx;
      nx.x = cur.x + (dir[c] * step * dx);
      nx.y = cur.y + (dir[c] * step * dy);
      nx.heading = cur.heading + (dir[c] * step * dtheta);
      nx.gridx = orc_cell_x(m, nx.x);
      nx.gridy = orc_cell_y(m, nx.y);
      ch->traj[ch->n_traj++] = nx;
      cur = nx;
    }
    for (i = 0; i < ch->n_traj; i++) {
      const orc_state *s = &ch->traj[i];
      int hit = (variant == ORC_VARIANT_RS) ? orc_collide_circles(m, car, s->x, s->y, s->heading)
                                            : orc_collide_sat(m, car, s->x, s->y, s->heading);
      if (hit) {
        has_collision = 1;
        break;
      }
    }
    ch->steer = steer[c];
    ch->direction = (int)dir[c];
    if (!has_collision) {
      double cost = parent_cost;
      if (dir[c] == 1) cost += sim_len;
      else cost += sim_len * reverse;
      if (parent_dir != dir[c]) cost += directionChange;
      cost += fabs(steer[c]) * steerAngle;
      cost += fabs(steer[c] - parent_steer) * steerAngleChange;
      ch->cost = cost;
      ch->valid = 1;
    }
  }
  return ncmd;
}

/* ---- State-keyed hash map (stand-in for std::unordered_map<State, ..., StateHash>) ----
 * The reference's key equality has a 1e-9 tolerance (State.h:24-29) but its hash is on the raw doubles
 * (State.h:58-65), so two keys only meet when they land in the same bucket; the oracle uses exact
 * value equality (documented deviation). */
typedef struct {
  double x, y, h;
  int node; /* -1 = empty slot, -2 = tombstone */
} slot;
typedef struct {
  slot *s;
  size_t cap, used;
} smap;

static uint64_t dbits(double d) {
  uint64_t u;
  if (d == 0.0) d = 0.0; /* -0.0 == 0.0 */
  memcpy(&u, &d, 8);
  return u;
}
static size_t shash(double x, double y, double h) {
  uint64_t a = dbits(x) * 0x9E3779B97F4A7C15ull;
  a ^= (dbits(y) + 0x7F4A7C15ull + (a << 6) + (a >> 2));
  a ^= (dbits(h) * 0xC2B2AE3D27D4EB4Full + (a << 6) + (a >> 2));
  a ^= a >> 29;
  return (size_t)a;
}
static void smap_init(smap *m, size_t cap) {
  size_t i;
  m->cap = cap;
  m->used = 0;
  m->s = (slot *)malloc(sizeof(slot) * cap);
  for (i = 0; i < cap; i++) m->s[i].node = -1;
}
static slot *smap_find(smap *m, double x, double y, double h) {
  size_t i = shash(x, y, h) & (m->cap - 1);
  while (m->s[i].node != -1) {
    if (m->s[i].node != -2 && m->s[i].x == x && m->s[i].y == y && m->s[i].h == h) return &m->s[i];
    i = (i + 1) & (m->cap - 1);
  }
  return NULL;
}
static void smap_put(smap *m, double x, double y, double h, int node);
static void smap_grow(smap *m) {
  smap n;
  size_t i;
  smap_init(&n, m->cap * 2);
  for (i = 0; i < m->cap; i++)
    if (m->s[i].node >= 0) smap_put(&n, m->s[i].x, m->s[i].y, m->s[i].h, m->s[i].node);
  free(m->s);
  *m = n;
}
static void smap_put(smap *m, double x, double y, double h, int node) {
  slot *f = smap_find(m, x, y, h);
  size_t i;
  if (f) {
    f->node = node;
    return;
  }
  if ((m->used + 1) * 2 > m->cap) smap_grow(m);
  i = shash(x, y, h) & (m->cap - 1);
  while (m->s[i].node >= 0) i = (i + 1) & (m->cap - 1);
  if (m->s[i].node == -1) m->used++;
  m->s[i].x = x;
  m->s[i].y = y;
  m->s[i].h = h;
  m->s[i].node = node;
}
static void smap_erase(smap *m, double x, double y, double h) {
  slot *f = smap_find(m, x, y, h);
  if (f) f->node = -2;
}

typedef struct {
  orc_state state;
  orc_state *traj;
  int n_traj;
  double cost, steer;
  int direction, parent;
} pnode;

void orc_run(const orc_map *m, const orc_car *car, int variant, double sim_len, double step, double sx,
             double sy, double sh, double gx, double gy, double gh, int heuristic_quirks, orc_plan *out) {
  const double hybridCost = 30.0, HEADING_TOLERANCE = 0.05;
  size_t map_size = (size_t)m->W * m->H;
  double *hmap = (double *)malloc(sizeof(double) * map_size);
  orc_state start, goal;
  pnode *nodes = NULL;
  int nn = 0, ncap = 0, goal_node, iterations = 0, start_index, via_shot = 0;
  orc_heap_item *heap = NULL;
  int hn = 0, hcap = 0;
  smap open_set, closed_set;
  orc_child children[64];

#define NEW_NODE(idx)                                                   \
  do {                                                                  \
    if (nn == ncap) {                                                   \
      ncap = ncap ? ncap * 2 : 1024;                                    \
      nodes = (pnode *)realloc(nodes, sizeof(pnode) * (size_t)ncap);    \
    }                                                                   \
    (idx) = nn++;                                                       \
  } while (0)
#define HEAP_PUSH(k, p)                                                              \
  do {                                                                               \
    orc_heap_item it_;                                                               \
    if (hn + 1 >= hcap) {                                                            \
      hcap = hcap ? hcap * 2 : 4096;                                                 \
      heap = (orc_heap_item *)realloc(heap, sizeof(orc_heap_item) * (size_t)hcap);   \
    }                                                                                \
    it_.key = (k);                                                                   \
    it_.payload = (p);                                                               \
    orc_stdheap_push(heap, &hn, it_);                                                \
  } while (0)

  memset(out, 0, sizeof *out);
  memset(&start, 0, sizeof start);
  memset(&goal, 0, sizeof goal);
  start.x = sx; start.y = sy; start.heading = sh;
  goal.x = gx; goal.y = gy; goal.heading = gh;
  start.gridx = orc_cell_x(m, sx);
  start.gridy = orc_cell_y(m, sy);
  start_index = (int)start.gridy * m->W + (int)start.gridx;
  goal.gridx = orc_cell_x(m, gx);
  goal.gridy = orc_cell_y(m, gy);

  if (variant == ORC_VARIANT_RS)
    orc_heuristic_rs(m->occ, m->W, m->H, (int)goal.gridx, (int)goal.gridy, (int)start.gridx, (int)start.gridy,
                     heuristic_quirks, hmap, NULL);
  else
    orc_heuristic_dubins(m->occ, m->W, m->H, (int)goal.gridx, (int)goal.gridy, hmap);

  smap_init(&open_set, 1024);
  smap_init(&closed_set, 1024);
  {
    int s_idx, g_idx;
    NEW_NODE(s_idx);
    nodes[s_idx].state = start;
    nodes[s_idx].traj = (orc_state *)malloc(sizeof(orc_state));
    nodes[s_idx].traj[0] = start;
    nodes[s_idx].n_traj = 1;
    nodes[s_idx].cost = 0;
    nodes[s_idx].steer = 0;
    nodes[s_idx].direction = 1;
    nodes[s_idx].parent = -1;
    NEW_NODE(g_idx);
    nodes[g_idx] = nodes[s_idx];
    nodes[g_idx].state = goal;
    nodes[g_idx].traj = (orc_state *)malloc(sizeof(orc_state));
    nodes[g_idx].traj[0] = goal;
    goal_node = g_idx;
    smap_put(&open_set, start.x, start.y, start.heading, s_idx);
    HEAP_PUSH(nodes[s_idx].cost + hybridCost * hmap[start_index], s_idx);
  }

  while (hn > 0) {
    orc_heap_item top;
    orc_state cs;
    slot *os;
    int cur, nchild, c;
    orc_shot shot;
    double heading_difference;
    iterations++;
    if (iterations > 3500) break;
    top = orc_stdheap_pop(heap, &hn);
    cs = nodes[top.payload].state;
    os = smap_find(&open_set, cs.x, cs.y, cs.heading);
    if (smap_find(&closed_set, cs.x, cs.y, cs.heading)) continue;
    if (!os) continue; /* cannot happen: a state leaves the open set only by entering the closed set */
    cur = os->node;
    smap_erase(&open_set, cs.x, cs.y, cs.heading);
    smap_put(&closed_set, cs.x, cs.y, cs.heading, cur);

    if (variant == ORC_VARIANT_RS) orc_rs_shot(m, car, &nodes[cur].state, nodes[cur].cost, &goal, &shot);
    else orc_dubins_shot(m, car, &nodes[cur].state, nodes[cur].cost, &goal, &shot);
    if (shot.found) {
      int d_idx;
      NEW_NODE(d_idx);
      nodes[d_idx].state = shot.traj[shot.n - 1];
      nodes[d_idx].traj = shot.traj;
      nodes[d_idx].n_traj = shot.n;
      nodes[d_idx].cost = shot.cost;
      nodes[d_idx].steer = 0;
      nodes[d_idx].direction = 1;
      nodes[d_idx].parent = cur;
      goal_node = d_idx;
      via_shot = 1;
      break;
    }
    orc_shot_free(&shot);

    heading_difference = fabs(cs.heading - goal.heading);
    if (heading_difference > PI_A) heading_difference = fabs(heading_difference - 2 * PI_A);
    if (cs.gridx == goal.gridx && cs.gridy == goal.gridy && heading_difference <= HEADING_TOLERANCE) {
      goal_node = cur;
      break;
    }

    nchild = orc_expand(m, car, variant, sim_len, step, &nodes[cur].state, nodes[cur].cost, nodes[cur].steer,
                        nodes[cur].direction, children);
    for (c = 0; c < nchild; c++) {
      orc_child *ch = &children[c];
      orc_state ns;
      int neighbor_index, idx;
      double neighborCost;
      slot *o;
      if (!ch->valid) continue;
      ns = ch->traj[ch->n_traj - 1];
      if (smap_find(&closed_set, ns.x, ns.y, ns.heading)) continue;
      neighbor_index = (int)ns.gridy * m->W + (int)ns.gridx;
      if (variant == ORC_VARIANT_RS && (neighbor_index < 0 || neighbor_index >= (int)map_size)) continue;
      neighborCost = ch->cost + hybridCost * hmap[neighbor_index];
      o = smap_find(&open_set, ns.x, ns.y, ns.heading);
      if (!o || ch->cost < nodes[o->node].cost) {
        NEW_NODE(idx);
        nodes[idx].state = ns;
        nodes[idx].traj = (orc_state *)malloc(sizeof(orc_state) * (size_t)ch->n_traj);
        memcpy(nodes[idx].traj, ch->traj, sizeof(orc_state) * (size_t)ch->n_traj);
        nodes[idx].n_traj = ch->n_traj;
        nodes[idx].cost = ch->cost;
        nodes[idx].steer = ch->steer;
        nodes[idx].direction = ch->direction;
        nodes[idx].parent = cur;
        smap_put(&open_set, ns.x, ns.y, ns.heading, idx);
        HEAP_PUSH(neighborCost, idx);
      }
    }
  }

  /* backtrack (RS/src/HybridAstar.cpp:695-710) */
  {
    int b, total = 0, k = 0, i;
    for (b = goal_node; b >= 0; b = nodes[b].parent) total += nodes[b].n_traj;
    out->traj = (orc_state *)malloc(sizeof(orc_state) * (size_t)(total + 1));
    out->n = total;
    /* concatenating reversed trajectories goal->start and reversing the whole equals start->goal order */
    for (b = goal_node; b >= 0; b = nodes[b].parent)
      for (i = nodes[b].n_traj - 1; i >= 0; i--) out->traj[total - 1 - (k++)] = nodes[b].traj[i];
    out->iterations = iterations;
    out->via_shot = via_shot;
    out->cost = nodes[goal_node].cost;
  }
  {
    int i;
    for (i = 0; i < nn; i++) free(nodes[i].traj);
  }
  free(nodes);
  free(heap);
  free(open_set.s);
  free(closed_set.s);
  free(hmap);
#undef NEW_NODE
#undef HEAP_PUSH
}

void orc_plan_free(orc_plan *p) {
  free(p->traj);
  p->traj = NULL;
  p->n = 0;
}
