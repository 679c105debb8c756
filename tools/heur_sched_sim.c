// heur_sched_sim.c -- design tool (not product, not oracle): simulates the tile / cost-window schedule of the
// K3a heuristic kernel on the CPU to count rounds, tile visits, in-tile sweeps on the critical path and cell
// evaluations for a given tile side T and window width DELTA.  Input: raw int8 occupancy (side x side).
//   gcc -O2 -o /tmp/heur_sched_sim tools/heur_sched_sim.c -lm && /tmp/heur_sched_sim map.raw 4096 2048 2048 128 128
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int W, H, T, TX, TY;
static int8_t *occ;
static uint8_t *fre, *nearb;
static double *cost, *snap;  // snap = values at the start of the round (what concurrent CTAs see in the halo)

int main(int argc, char **argv) {
  if (argc < 7) return 1;
  W = H = atoi(argv[2]);
  int gx = atoi(argv[3]), gy = atoi(argv[4]);
  T = atoi(argv[5]);
  double DELTA = atof(argv[6]);
  int max_sweeps = argc > 7 ? atoi(argv[7]) : 1 << 30;  // cap of sweeps per visit (then the tile stays active)
  size_t n = (size_t)W * H;
  occ = malloc(n);
  FILE *f = fopen(argv[1], "rb");
  if (!f || fread(occ, 1, n, f) != n) return 2;
  fclose(f);
  fre = malloc(n); nearb = malloc(n); cost = malloc(n * 8); snap = malloc(n * 8);
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      size_t i = (size_t)y * W + x;
      fre[i] = !(occ[i] > 0);
      int nb = 0;
      if (x > 0 && occ[i - 1] > 0) nb = 1;
      if (x < W - 1 && occ[i + 1] > 0) nb = 1;
      if (y > 0 && occ[i - W] > 0) nb = 1;
      if (y < H - 1 && occ[i + W] > 0) nb = 1;
      nearb[i] = nb;
      cost[i] = INFINITY; snap[i] = INFINITY;
    }
  TX = (W + T - 1) / T; TY = (H + T - 1) / T;
  int NT = TX * TY;
  // per-cell state bits: need (1), deferred (2)
  uint8_t *st = calloc(n, 1);
  double *pmin = malloc(NT * 8);
  uint8_t *act = calloc(NT, 1), *actn = calloc(NT, 1);
  for (int t = 0; t < NT; t++) pmin[t] = INFINITY;
  cost[(size_t)gy * W + gx] = 0.0; snap[(size_t)gy * W + gx] = 0.0;
  // the goal's change is "deferred" with cost 0: released at the first visit
  st[(size_t)gy * W + gx] = 2;
  pmin[(gy / T) * TX + gx / T] = 0.0;
  double bound = 0.0;
  long rounds = 0, visits = 0, crit_sweeps = 0, evals = 0, windows = 0, max_active = 0, sum_sweeps = 0;
  int *list = malloc(NT * sizeof(int));
  // snapshot-free: process tiles sequentially in place (halo reads see earlier tiles' writes of this round; the
  // real kernel is chaotic in the same way)
  int *fl = malloc((size_t)T * T * sizeof(int)), *fl2 = malloc((size_t)T * T * sizeof(int));
  for (;;) {
    int na = 0;
    for (int t = 0; t < NT; t++)
      if (act[t] || pmin[t] < bound) list[na++] = t;
    if (na == 0) {
      double g = INFINITY;
      for (int t = 0; t < NT; t++) if (pmin[t] < g) g = pmin[t];
      if (!(g < INFINITY)) break;
      bound = g + DELTA;
      windows++;
      continue;
    }
    rounds++;
    if (na > max_active) max_active = na;
    long round_max = 0;
    for (int k = 0; k < na; k++) {
      int t = list[k];
      int tx = t % TX, ty = t / TX, x0 = tx * T, y0 = ty * T;
      int x1 = x0 + T < W ? x0 + T : W, y1 = y0 + T < H ? y0 + T : H;
      visits++;
      // initial need: ring sides flagged by neighbours + dilation of released deferred cells
      int nf = 0;
      uint8_t a = act[t];
      act[t] = 0;
      for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
          size_t i = (size_t)y * W + x;
          int ring = 0;
          if ((a & 1) && y == y0) ring = 1;
          if ((a & 2) && y == y1 - 1) ring = 1;
          if ((a & 4) && x == x0) ring = 1;
          if ((a & 8) && x == x1 - 1) ring = 1;
          if (ring && fre[i]) st[i] |= 1;
        }
      unsigned out = 0;
      // release deferred
      for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
          size_t i = (size_t)y * W + x;
          if ((st[i] & 2) && cost[i] < bound) {
            st[i] &= ~2;
            for (int dy = -1; dy <= 1; dy++)
              for (int dx = -1; dx <= 1; dx++) {
                if (!dx && !dy) continue;
                int xx = x + dx, yy = y + dy;
                if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
                if (xx < x0 || xx >= x1 || yy < y0 || yy >= y1) {
                  out |= (yy < y0 ? 1 : 0) | (yy >= y1 ? 2 : 0) | (xx < x0 ? 4 : 0) | (xx >= x1 ? 8 : 0);
                  continue;
                }
                size_t j = (size_t)yy * W + xx;
                if (fre[j]) st[j] |= 1;
              }
          }
        }
      for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++)
          if (st[(size_t)y * W + x] & 1) fl[nf++] = y * W + x;
      long sweeps = 0;
      while (nf > 0 && sweeps < max_sweeps) {
        sweeps++;
        int nf2 = 0;
        // evaluate (pull)
        static int chg[1 << 16];
        int nc = 0;
        for (int q = 0; q < nf; q++) {
          int i = fl[q];
          st[i] &= ~1;
          int x = i % W, y = i / W;
          double mc = INFINITY, md = INFINITY;
          for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
              if (!dx && !dy) continue;
              int xx = x + dx, yy = y + dy;
              if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
              int inside = xx >= x0 && xx < x1 && yy >= y0 && yy < y1;
              double c = inside ? cost[(size_t)yy * W + xx] : snap[(size_t)yy * W + xx];
              if (dx && dy) { if (c < md) md = c; } else { if (c < mc) mc = c; }
            }
          double wc = 1.0, wd = 1.414;
          if (nearb[i]) { wc += 5.0; wd += 5.0; }
          double b = fmin(mc + wc, md + wd);
          evals++;
          if (b < cost[i]) { cost[i] = b; chg[nc++] = i; }
        }
        for (int q = 0; q < nc; q++) {
          int i = chg[q];
          if (!(cost[i] < bound)) { st[i] |= 2; continue; }
          st[i] &= ~2;
          int x = i % W, y = i / W;
          for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
              if (!dx && !dy) continue;
              int xx = x + dx, yy = y + dy;
              if (xx < 0 || xx >= W || yy < 0 || yy >= H) continue;
              if (xx < x0 || xx >= x1 || yy < y0 || yy >= y1) {
                out |= (yy < y0 ? 1 : 0) | (yy >= y1 ? 2 : 0) | (xx < x0 ? 4 : 0) | (xx >= x1 ? 8 : 0);
                continue;
              }
              size_t j = (size_t)yy * W + xx;
              if (fre[j] && !(st[j] & 1)) { st[j] |= 1; fl2[nf2++] = (int)j; }
            }
        }
        int *tmp = fl; fl = fl2; fl2 = tmp;
        nf = nf2;
      }
      if (nf > 0) actn[t] |= 16;  // sweep cap hit: keep active (need bits stay set)
      sum_sweeps += sweeps;
      if (sweeps > round_max) round_max = sweeps;
      // pending min over deferred
      double pm = INFINITY;
      for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
          size_t i = (size_t)y * W + x;
          if ((st[i] & 2) && cost[i] < pm) pm = cost[i];
        }
      pmin[t] = pm;
      // wake neighbours: out bits combine (N=1,S=2,W=4,E=8)
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          if (!dx && !dy) continue;
          int need = (dy < 0 ? 1 : 0) | (dy > 0 ? 2 : 0) | (dx < 0 ? 4 : 0) | (dx > 0 ? 8 : 0);
          if ((out & need) != need) continue;   // (approximation for corners: both bits set)
          int nx = tx + dx, ny = ty + dy;
          if (nx < 0 || nx >= TX || ny < 0 || ny >= TY) continue;
          // neighbour evaluates the side(s) facing us
          int side = (dy < 0 ? 2 : 0) | (dy > 0 ? 1 : 0) | (dx < 0 ? 8 : 0) | (dx > 0 ? 4 : 0);
          actn[ny * TX + nx] |= side;
        }
    }
    crit_sweeps += round_max;
    for (int k = 0; k < na; k++) {
      int t = list[k]; int x0 = (t % TX) * T, y0 = (t / TX) * T;
      int x1 = x0 + T < W ? x0 + T : W, y1 = y0 + T < H ? y0 + T : H;
      for (int y = y0; y < y1; y++) memcpy(snap + (size_t)y * W + x0, cost + (size_t)y * W + x0, (size_t)(x1 - x0) * 8);
    }
    for (int t = 0; t < NT; t++) { act[t] |= actn[t]; actn[t] = 0; }
  }
  // checksum
  double s = 0; long reach = 0; double mx = 0;
  for (size_t i = 0; i < n; i++) if (cost[i] < INFINITY) { s += cost[i]; reach++; if (cost[i] > mx) mx = cost[i]; }
  printf("T=%d DELTA=%g cap=%d: rounds=%ld windows=%ld visits=%ld (%.2f/tile) crit_sweeps=%ld sum_sweeps=%ld evals=%ld (%.2f/cell) max_active=%ld  reach=%ld max=%.3f sum=%.6f\n",
         T, DELTA, max_sweeps, rounds, windows, visits, (double)visits / NT, crit_sweeps, sum_sweeps, evals, (double)evals / n, max_active, reach, mx, s);
  return 0;
}
