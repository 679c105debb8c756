/*
 * oracle/orc_heur.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 * Restates the two holonomic-with-obstacles heuristics: a1 (RS) and a8 (Dubins).
 */
#include "orc.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int *v;
  int n, cap;
} ivec;

static void ivec_push(ivec *b, int x) {
  if (b->n == b->cap) {
    b->cap = b->cap ? b->cap * 2 : 16;
    b->v = (int *)realloc(b->v, sizeof(int) * (size_t)b->cap);
  }
  b->v[b->n++] = x;
}

static int occ_collision(const int8_t *occ, int W, int H, int x, int y) {
  if (x < 0 || x >= W || y < 0 || y >= H) return 1; /* RS/src/Grid_map.cpp:207-210 */
  return occ[(size_t)y * W + x] > 0;                /* :212 */
}

/* RS edge weight into cell (x,y) from direction (dx,dy): RS/src/HybridAstar.cpp:386-390, :500-524 */
static double rs_move_cost(const int8_t *occ, int W, int H, int x, int y, int dx, int dy) {
  static const int ndx[4] = {-1, 0, 1, 0};
  static const int ndy[4] = {0, -1, 0, 1};
  double movement_cost = (dx == 0 || dy == 0) ? 1.0 : 1.414;
  int i;
  for (i = 0; i < 4; i++) {
    int nx = x + ndx[i], ny = y + ndy[i];
    if (nx >= 0 && nx < W && ny >= 0 && ny < H && occ_collision(occ, W, H, nx, ny)) {
      movement_cost += 5.0;
      break;
    }
  }
  return movement_cost;
}

static const int DIRS[8][2] = {{-1, 0}, {-1, 1}, {0, 1}, {1, 1}, {1, 0}, {1, -1}, {0, -1}, {-1, -1}};

/*
 * RS/src/HybridAstar.cpp:353-551.  emulate_quirks=1 follows the reference statement by statement:
 * bucket queue indexed by (int)cost clamped to MAX_COST-1 (:395, :534), FIFO inside a bucket, closed
 * set, and the progress check every 70000 settled cells whose `break` leaves the for-loop over the
 * swapped-out bucket and so drops its remaining cells (:452-472).  The function-static cache (:407-416,
 * hazard H3) is a property of the caller's process, not of the arithmetic, and is not emulated.
 * emulate_quirks=0 computes the least fixed point d[v] = min_u fl(d[u] + w(u->v)) by the same bucket
 * sweep without the progress check and with an unbounded last bucket handled by a final
 * label-correcting loop.
 */
void orc_heuristic_rs(const int8_t *occ, int W, int H, int goal_x, int goal_y, int start_x, int start_y,
                      int emulate_quirks, double *cost, orc_quirks *q) {
  size_t map_size = (size_t)W * H, k;
  unsigned char *closed = (unsigned char *)calloc(map_size, 1);
  int MAX_COST = (W + H > 5000) ? W + H : 5000;
  ivec *buckets = (ivec *)calloc((size_t)MAX_COST, sizeof(ivec));
  int goal_index = goal_y * W + goal_x, current_bucket = 0, dir, b;
  long processed = 0;
  orc_quirks qq;
  memset(&qq, 0, sizeof qq);
  for (k = 0; k < map_size; k++) cost[k] = INFINITY;
  cost[goal_index] = 0.0;
  ivec_push(&buckets[0], goal_index);

  for (;;) {
    ivec cur;
    int ci;
    while (current_bucket < MAX_COST && buckets[current_bucket].n == 0) current_bucket++;
    if (current_bucket >= MAX_COST) break;
    cur = buckets[current_bucket];
    memset(&buckets[current_bucket], 0, sizeof(ivec));
    for (ci = 0; ci < cur.n; ci++) {
      int current_index = cur.v[ci], cx, cy;
      double current_cost;
      if (closed[current_index]) continue;
      current_cost = cost[current_index];
      closed[current_index] = 1;
      processed++;
      if (emulate_quirks && processed % 70000 == 0) {
        int sufficient = 1;
        if (start_x >= 0 && start_x < W && start_y >= 0 && start_y < H) {
          int si = start_y * W + start_x;
          if (!closed[si] && cost[si] == INFINITY) sufficient = 0;
        }
        if (sufficient || processed > map_size * 0.75) {
          if (ci + 1 < cur.n) qq.h1_fired = 1;
          /* the current cell was closed but never expanded; that alone already deviates */
          qq.h1_fired = 1;
          break;
        }
      }
      cx = current_index % W;
      cy = current_index / W;
      for (dir = 0; dir < 8; dir++) {
        int x = cx + DIRS[dir][0], y = cy + DIRS[dir][1], ni, bi;
        double new_cost;
        if (x < 0 || x >= W || y < 0 || y >= H) continue;
        ni = y * W + x;
        if (closed[ni] || occ_collision(occ, W, H, x, y)) continue;
        new_cost = current_cost + rs_move_cost(occ, W, H, x, y, DIRS[dir][0], DIRS[dir][1]);
        if (new_cost < cost[ni]) {
          cost[ni] = new_cost;
          bi = (int)new_cost;
          if (bi >= MAX_COST - 1) {
            bi = MAX_COST - 1;
            qq.h2_clamped = 1;
          }
          ivec_push(&buckets[bi], ni);
          if (bi < current_bucket) current_bucket = bi;
        }
      }
    }
    free(cur.v);
  }

  if (!emulate_quirks && qq.h2_clamped) {
    /* exact mode beyond the clamp: finish with plain label-correcting sweeps (order-independent) */
    int changed = 1;
    while (changed) {
      int x, y;
      changed = 0;
      for (y = 0; y < H; y++)
        for (x = 0; x < W; x++) {
          size_t u = (size_t)y * W + x;
          if (cost[u] == INFINITY) continue;
          for (dir = 0; dir < 8; dir++) {
            int vx = x + DIRS[dir][0], vy = y + DIRS[dir][1];
            double nc;
            if (vx < 0 || vx >= W || vy < 0 || vy >= H || occ_collision(occ, W, H, vx, vy)) continue;
            nc = cost[u] + rs_move_cost(occ, W, H, vx, vy, DIRS[dir][0], DIRS[dir][1]);
            if (nc < cost[(size_t)vy * W + vx]) {
              cost[(size_t)vy * W + vx] = nc;
              changed = 1;
            }
          }
        }
    }
  }
  for (k = 0; k < map_size; k++)
    if (cost[k] != INFINITY && cost[k] > qq.max_cost) qq.max_cost = cost[k];
  qq.processed = processed;
  if (q) *q = qq;
  for (b = 0; b < MAX_COST; b++) free(buckets[b].v);
  free(buckets);
  free(closed);
}

/*
 * DB/src/HybridAstar.cpp:198-313.  Edge u->v: w = fl(hypot(dx,dy) + penalty(u)), penalty = 10 per cell of
 * the 3x3 block around the SOURCE u that isInCollision (out of bounds included) (:200-220).  All cells
 * start at 1000 (Node.h:35), the goal at 0 (:233).  The reference runs a lazy priority queue of shared
 * nodes whose keys mutate while queued and has no closed set; it terminates at the fixed point
 * cost[v] = min(1000, min_u fl(cost[u] + w)) -- strict '>' at :289 -- which is what is computed here with a
 * binary heap on (cost, index) and lazy deletion.
 */
typedef struct {
  double c;
  int i;
} hent;

static void heap_push(hent **h, int *n, int *cap, hent e) {
  int i;
  if (*n == *cap) {
    *cap = *cap ? *cap * 2 : 1024;
    *h = (hent *)realloc(*h, sizeof(hent) * (size_t)*cap);
  }
  i = (*n)++;
  while (i > 0) {
    int p = (i - 1) / 2;
    if ((*h)[p].c <= e.c) break;
    (*h)[i] = (*h)[p];
    i = p;
  }
  (*h)[i] = e;
}

static hent heap_pop(hent *h, int *n) {
  hent top = h[0], last = h[--(*n)];
  int i = 0;
  for (;;) {
    int l = 2 * i + 1, r = l + 1, m = l;
    if (l >= *n) break;
    if (r < *n && h[r].c < h[l].c) m = r;
    if (h[m].c >= last.c) break;
    h[i] = h[m];
    i = m;
  }
  h[i] = last;
  return top;
}

void orc_heuristic_dubins(const int8_t *occ, int W, int H, int goal_x, int goal_y, double *cost) {
  size_t map_size = (size_t)W * H, k;
  hent *heap = NULL;
  int hn = 0, hcap = 0, dir;
  hent e;
  for (k = 0; k < map_size; k++) cost[k] = 1000;
  cost[(size_t)goal_y * W + goal_x] = 0;
  e.c = 0;
  e.i = goal_y * W + goal_x;
  heap_push(&heap, &hn, &hcap, e);
  while (hn > 0) {
    hent cur = heap_pop(heap, &hn);
    int cx = cur.i % W, cy = cur.i / W, dx, dy;
    double current_cost, obstacle_penalty = 0.0;
    if (cur.c > cost[cur.i]) continue; /* stale entry */
    current_cost = cost[cur.i];
    for (dx = -1; dx <= 1; ++dx)
      for (dy = -1; dy <= 1; ++dy)
        if (occ_collision(occ, W, H, cx + dx, cy + dy)) obstacle_penalty += 10.0;
    for (dir = 0; dir < 8; dir++) {
      int x = cx + DIRS[dir][0], y = cy + DIRS[dir][1];
      double new_cost;
      if (occ_collision(occ, W, H, x, y)) continue;
      new_cost = current_cost + (hypot((double)DIRS[dir][0], (double)DIRS[dir][1]) + obstacle_penalty);
      if (cost[(size_t)y * W + x] > new_cost) {
        cost[(size_t)y * W + x] = new_cost;
        e.c = new_cost;
        e.i = y * W + x;
        heap_push(&heap, &hn, &hcap, e);
      }
    }
  }
  free(heap);
}
