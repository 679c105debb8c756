/*
 * oracle/orc_dubins.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 * Restates a11: dubins_path (DB/src/dubins_path.cpp:13-374) and HybridAstar::dubins_Path /
 * dubins_Path_Cost (DB/src/HybridAstar.cpp:108-184).
 */
#include "orc.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* :13-24 */
static double pi_2_pi(double theta) {
  while (theta > M_PI) theta -= 2.0 * M_PI;
  while (theta < -M_PI) theta += 2.0 * M_PI;
  return theta;
}
/* :26-29 */
static double mod2pi(double theta) { return theta - 2.0 * M_PI * floor(theta / M_PI / 2.0); }

typedef struct {
  double t, p, q;
  const char *mode;
} word;

#define TRIG                                                                                            \
  double sin_a = sin(alpha), sin_b = sin(beta), cos_a = cos(alpha), cos_b = cos(beta),                  \
         cos_a_b = cos(alpha - beta)

/* :33-57 */
static word w_LSL(double alpha, double beta, double dist) {
  word w = {NAN, NAN, NAN, "LSL"};
  TRIG;
  double p = 2 + dist * dist - 2 * cos_a_b + 2 * dist * (sin_a - sin_b), den;
  if (p < 0) return w;
  p = sqrt(p);
  den = dist + sin_a - sin_b;
  w.t = mod2pi(-alpha + atan2(cos_b - cos_a, den));
  w.p = p;
  w.q = mod2pi(beta - atan2(cos_b - cos_a, den));
  return w;
}
/* :59-83 */
static word w_RSR(double alpha, double beta, double dist) {
  word w = {NAN, NAN, NAN, "RSR"};
  TRIG;
  double p = 2 + dist * dist - 2 * cos_a_b + 2 * dist * (sin_b - sin_a), den;
  if (p < 0) return w;
  p = sqrt(p);
  den = dist - sin_a + sin_b;
  w.t = mod2pi(alpha - atan2(cos_a - cos_b, den));
  w.p = p;
  w.q = mod2pi(-beta + atan2(cos_a - cos_b, den));
  return w;
}
/* :85-109 */
static word w_LSR(double alpha, double beta, double dist) {
  word w = {NAN, NAN, NAN, "LSR"};
  TRIG;
  double p = -2 + dist * dist + 2 * cos_a_b + 2 * dist * (sin_a + sin_b), rec;
  if (p < 0) return w;
  p = sqrt(p);
  rec = atan2(-cos_a - cos_b, dist + sin_a + sin_b) - atan2(-2.0, p);
  w.t = mod2pi(-alpha + rec);
  w.p = p;
  w.q = mod2pi(-mod2pi(beta) + rec);
  return w;
}
/* :111-135 */
static word w_RSL(double alpha, double beta, double dist) {
  word w = {NAN, NAN, NAN, "RSL"};
  TRIG;
  double p = -2 + dist * dist + 2 * cos_a_b - 2 * dist * (sin_a + sin_b), rec;
  if (p < 0) return w;
  p = sqrt(p);
  rec = atan2(cos_a + cos_b, dist - sin_a - sin_b) - atan2(2.0, p);
  w.t = mod2pi(alpha - rec);
  w.p = p;
  w.q = mod2pi(beta - rec);
  return w;
}
/* :137-157 */
static word w_RLR(double alpha, double beta, double dist) {
  word w = {NAN, NAN, NAN, "RLR"};
  TRIG;
  double rec = (6.0 - dist * dist + 2.0 * cos_a_b + 2.0 * dist * (sin_a - sin_b)) / 8.0;
  if (fabs(rec) > 1.0) return w;
  w.p = mod2pi(2 * M_PI - acos(rec));
  w.t = mod2pi(alpha - atan2(cos_a - cos_b, dist - sin_a + sin_b) + mod2pi(w.p / 2.0));
  w.q = mod2pi(alpha - beta - w.t + mod2pi(w.p));
  return w;
}
/* :159-179 */
static word w_LRL(double alpha, double beta, double dist) {
  word w = {NAN, NAN, NAN, "LRL"};
  TRIG;
  double rec = (6.0 - dist * dist + 2.0 * cos_a_b + 2.0 * dist * (sin_b - sin_a)) / 8.0;
  if (fabs(rec) > 1.0) return w;
  w.p = mod2pi(2 * M_PI - acos(rec));
  w.t = mod2pi(-alpha - atan2(cos_a - cos_b, dist + sin_a - sin_b) + w.p / 2.0);
  w.q = mod2pi(mod2pi(beta) - alpha - w.t + mod2pi(w.p));
  return w;
}

/* :181-230 */
static void interpolate(int ind, double l, char m, double maxc, double ox, double oy, double oyaw, double *px,
                        double *py, double *pyaw) {
  if (m == 'S') {
    px[ind] = ox + l / maxc * cos(oyaw);
    py[ind] = oy + l / maxc * sin(oyaw);
    pyaw[ind] = oyaw;
  } else {
    double ldx = sin(l) / maxc, ldy = 0.0, gdx, gdy;
    if (m == 'L') ldy = (1.0 - cos(l)) / maxc;
    else if (m == 'R') ldy = (1.0 - cos(l)) / (-maxc);
    gdx = cos(-oyaw) * ldx + sin(-oyaw) * ldy;
    gdy = -sin(-oyaw) * ldx + cos(-oyaw) * ldy;
    px[ind] = ox + gdx;
    py[ind] = oy + gdy;
    if (m == 'L') pyaw[ind] = oyaw + l;
    else if (m == 'R') pyaw[ind] = oyaw - l;
  }
}

/* :232-294; returns sample count (0 when <= 1 point survives the trim) */
static int generate_local_course(double L, const double *lengths, const char *mode, int nmode, double maxc,
                                 double step_size, double **opx, double **opy, double **opyaw) {
  int point_num = (int)(L / step_size) + 3 + 3;
  double *px = (double *)calloc((size_t)point_num + 8, sizeof(double));
  double *py = (double *)calloc((size_t)point_num + 8, sizeof(double));
  double *pyaw = (double *)calloc((size_t)point_num + 8, sizeof(double));
  int ind = 1, i, n;
  double d = (lengths[0] > 0.0) ? step_size : -step_size, ll = 0.0;
  for (i = 0; i < nmode; ++i) {
    double l = lengths[i], ox, oy, oyaw, pd;
    char m = mode[i];
    d = (l > 0.0) ? step_size : -step_size;
    ox = px[ind];
    oy = py[ind];
    oyaw = pyaw[ind];
    ind -= 1;
    pd = (i >= 1 && (lengths[i - 1] * lengths[i]) > 0) ? -d - ll : d - ll;
    while (fabs(pd) <= fabs(l)) {
      ind += 1;
      interpolate(ind, pd, m, maxc, ox, oy, oyaw, px, py, pyaw);
      pd += d;
    }
    ll = l - pd - d;
    ind += 1;
    interpolate(ind, l, m, maxc, ox, oy, oyaw, px, py, pyaw);
  }
  n = point_num;
  while (n > 1 && px[n - 1] == 0.0) n--;
  if (n <= 1) n = 0;
  *opx = px;
  *opy = py;
  *opyaw = pyaw;
  return n;
}

/* :296-344 (planning_from_origin) + :346-374 (calc_dubins_path) */
void orc_dubins_path(double sx, double sy, double syaw, double gx, double gy, double gyaw, double curv,
                     double step_size, orc_dubins *out) {
  double cos_yaw, sin_yaw, le_x, le_y, le_yaw, D, d, theta, alpha, beta, best_cost = INFINITY;
  double bt = 0.0, bp = 0.0, bq = 0.0, lengths[3], *lx, *ly, *lyaw;
  const char *best_mode = "";
  word w[6];
  int i, n;
  memset(out, 0, sizeof *out);
  gx -= sx;
  gy -= sy;
  cos_yaw = cos(syaw);
  sin_yaw = sin(syaw);
  le_x = gx * cos_yaw + gy * sin_yaw;
  le_y = -gx * sin_yaw + gy * cos_yaw;
  le_yaw = gyaw - syaw;
  D = hypot(le_x, le_y);
  d = D * curv;
  theta = mod2pi(atan2(le_y, le_x));
  alpha = mod2pi(-theta);
  beta = mod2pi(le_yaw - theta);
  w[0] = w_LSL(alpha, beta, d);
  w[1] = w_RSR(alpha, beta, d);
  w[2] = w_LSR(alpha, beta, d);
  w[3] = w_RSL(alpha, beta, d);
  w[4] = w_RLR(alpha, beta, d);
  w[5] = w_LRL(alpha, beta, d);
  for (i = 0; i < 6; i++) {
    double cost;
    if (isnan(w[i].t)) continue;
    cost = fabs(w[i].t) + fabs(w[i].p) + fabs(w[i].q);
    if (best_cost > cost) {
      bt = w[i].t;
      bp = w[i].p;
      bq = w[i].q;
      best_mode = w[i].mode;
      best_cost = cost;
    }
  }
  lengths[0] = bt;
  lengths[1] = bp;
  lengths[2] = bq;
  n = generate_local_course(0.0 + bt + bp + bq, lengths, best_mode, (int)strlen(best_mode), curv, step_size, &lx,
                            &ly, &lyaw);
  out->length = best_cost;
  out->t = bt;
  out->p = bp;
  out->q = bq;
  memcpy(out->mode, best_mode, strlen(best_mode));
  out->n = n;
  out->x = (double *)malloc(sizeof(double) * (size_t)(n + 1));
  out->y = (double *)malloc(sizeof(double) * (size_t)(n + 1));
  out->yaw = (double *)malloc(sizeof(double) * (size_t)(n + 1));
  for (i = 0; i < n; ++i) {
    out->x[i] = lx[i] * cos_yaw - ly[i] * sin_yaw + sx;
    out->y[i] = lx[i] * sin_yaw + ly[i] * cos_yaw + sy;
    out->yaw[i] = pi_2_pi(lyaw[i] + syaw);
  }
  free(lx);
  free(ly);
  free(lyaw);
}

void orc_dubins_free(orc_dubins *d) {
  free(d->x);
  free(d->y);
  free(d->yaw);
  d->x = d->y = d->yaw = NULL;
  d->n = 0;
}

/* DB/src/HybridAstar.cpp:108-165 (dubins_Path), :167-184 (dubins_Path_Cost) */
void orc_dubins_shot(const orc_map *m, const orc_car *car, const orc_state *cur, double cur_cost,
                     const orc_state *goal, orc_shot *out) {
  double radius = tan(car->max_steer) / car->wheel_base;
  orc_dubins p;
  int i, collide = 0;
  memset(out, 0, sizeof *out);
  out->chosen = -1;
  orc_dubins_path(cur->x, cur->y, cur->heading, goal->x, goal->y, goal->heading, radius, 0.1, &p);
  out->n_paths = p.n > 0;
  if (p.n == 0) {
    orc_dubins_free(&p);
    return;
  }
  for (i = 0; i < p.n && !collide; i++) collide = orc_collide_sat(m, car, p.x[i], p.y[i], p.yaw[i]);
  if (!collide) {
    double cost = cur_cost;
    cost += fabs(p.length);
    for (i = 0; i < 3; i++)
      if (p.mode[i] != 'S') cost += car->max_steer * 15.0;
    out->found = 1;
    out->cost = cost;
    out->chosen = 0;
    out->n = p.n;
    out->traj = (orc_state *)malloc(sizeof(orc_state) * (size_t)(p.n + 1));
    for (i = 0; i < p.n; i++) {
      out->traj[i].x = p.x[i];
      out->traj[i].y = p.y[i];
      out->traj[i].heading = p.yaw[i];
      out->traj[i].gridx = orc_cell_x(m, p.x[i]);
      out->traj[i].gridy = orc_cell_y(m, p.y[i]);
    }
  }
  orc_dubins_free(&p);
}
