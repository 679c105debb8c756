/*
 * oracle/orc_rs.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 * Restates a7 (reeds_shepp_path::calc_all_paths and helpers, RS/src/reeds_shepp_path.cpp:40-860)
 * and a6 (HybridAstar::reeds_shepp_Path*, reeds_Path_Cost, RS/src/HybridAstar.cpp:110-313).
 */
#include "orc.h"
#include "orc_stdheap.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI M_PI /* RS/include/reeds_shepp_path.h:15 */
#define MAX_LENGTH 1000.0

typedef struct {
  int n;
  int nseg[ORC_RS_MAX_PATHS];
  double len[ORC_RS_MAX_PATHS][5];
  char typ[ORC_RS_MAX_PATHS][5];
  double L[ORC_RS_MAX_PATHS];
} cand_set;

/* :846-860 */
static double M(double theta) {
  double phi = fmod(theta, 2.0 * PI);
  if (phi < -PI) phi += 2.0 * PI;
  if (phi > PI) phi -= 2.0 * PI;
  return phi;
}
/* :837-843 */
static void R(double x, double y, double *r, double *theta) {
  *r = hypot(x, y);
  *theta = atan2(y, x);
}
/* :821-834 */
static double pi_2_pi(double theta) {
  while (theta > PI) theta -= 2.0 * PI;
  while (theta < -PI) theta += 2.0 * PI;
  return theta;
}

/* :98-141 (set_path): ordered, signed de-duplication; L >= MAX_LENGTH dropped.  The reference asserts
 * L >= 0.01 (:135, aborts when start ~ goal); the oracle follows the NDEBUG behaviour and keeps it. */
static void set_path(cand_set *ps, int nseg, const double *lengths, const char *types) {
  int e, i;
  double L = 0.0;
  for (e = 0; e < ps->n; e++) {
    if (ps->nseg[e] == nseg && memcmp(ps->typ[e], types, (size_t)nseg) == 0) {
      double sum_diff = 0.0;
      for (i = 0; i < nseg; i++) sum_diff += ps->len[e][i] - lengths[i];
      if (sum_diff <= 0.01) return;
    }
  }
  for (i = 0; i < nseg; i++) L = L + fabs(lengths[i]);
  if (L >= MAX_LENGTH) return;
  if (ps->n >= ORC_RS_MAX_PATHS) return;
  e = ps->n++;
  ps->nseg[e] = nseg;
  memcpy(ps->len[e], lengths, sizeof(double) * (size_t)nseg);
  memcpy(ps->typ[e], types, (size_t)nseg);
  ps->L[e] = L;
}

static void add3(cand_set *ps, int flag, double a, double b, double c, const char *t) {
  double l[3];
  if (!flag) return;
  l[0] = a; l[1] = b; l[2] = c;
  set_path(ps, 3, l, t);
}
static void add4(cand_set *ps, int flag, double a, double b, double c, double d, const char *t) {
  double l[4];
  if (!flag) return;
  l[0] = a; l[1] = b; l[2] = c; l[3] = d;
  set_path(ps, 4, l, t);
}
static void add5(cand_set *ps, int flag, double a, double b, double c, double d, double e, const char *t) {
  double l[5];
  if (!flag) return;
  l[0] = a; l[1] = b; l[2] = c; l[3] = d; l[4] = e;
  set_path(ps, 5, l, t);
}

/* :217-239 */
static int SLS(double x, double y, double phi, double *t, double *u, double *v) {
  phi = M(phi);
  if (y > 0.0 && 0.0 < phi && phi < PI * 0.99) {
    double xd = -y / tan(phi) + x;
    *t = xd - tan(phi / 2.0);
    *u = phi;
    *v = sqrt((x - xd) * (x - xd) + y * y) - tan(phi / 2.0);
    return 1;
  } else if (y < 0.0 && 0.0 < phi && phi < PI * 0.99) {
    double xd = -y / tan(phi) + x;
    *t = xd - tan(phi / 2.0);
    *u = phi;
    *v = -sqrt((x - xd) * (x - xd) + y * y) - tan(phi / 2.0);
    return 1;
  }
  return 0;
}
/* :143-157 */
static int LSL(double x, double y, double phi, double *t, double *u, double *v) {
  double uu, tt;
  R(x - sin(phi), y - 1.0 + cos(phi), &uu, &tt);
  if (tt >= 0.0) {
    double vv = M(phi - tt);
    if (vv >= 0.0) {
      *t = tt; *u = uu; *v = vv;
      return 1;
    }
  }
  return 0;
}
/* :159-179 */
static int LSR(double x, double y, double phi, double *t, double *u, double *v) {
  double u1, t1;
  R(x + sin(phi), y - 1.0 - cos(phi), &u1, &t1);
  u1 = u1 * u1;
  if (u1 >= 4.0) {
    double uu = sqrt(u1 - 4.0);
    double theta = atan2(2.0, uu);
    double tt = M(t1 + theta);
    double vv = M(tt - phi);
    if (tt >= 0.0 && vv >= 0.0) {
      *t = tt; *u = uu; *v = vv;
      return 1;
    }
  }
  return 0;
}
/* :181-199 */
static int LRL(double x, double y, double phi, double *t, double *u, double *v) {
  double u1, t1;
  R(x - sin(phi), y - 1.0 + cos(phi), &u1, &t1);
  if (u1 <= 4.0) {
    double uu = -2.0 * asin(0.25 * u1);
    double tt = M(t1 + 0.5 * uu + PI);
    double vv = M(phi - tt + uu);
    if (tt >= 0.0 && uu <= 0.0) {
      *t = tt; *u = uu; *v = vv;
      return 1;
    }
  }
  return 0;
}
/* :359-372 */
static void calc_tauOmega(double u, double v, double xi, double eta, double phi, double *tau, double *omega) {
  double delta = M(u - v);
  double A = sin(u) - sin(delta);
  double B = cos(u) - cos(delta) - 1.0;
  double t1 = atan2(eta * A - xi * B, xi * A + eta * B);
  double t2 = 2.0 * (cos(delta) - cos(v) - cos(u)) + 3.0;
  *tau = t2 < 0 ? M(t1 + PI) : M(t1);
  *omega = M(*tau - u + v - phi);
}
/* :374-393 */
static int LRLRn(double x, double y, double phi, double *t, double *u, double *v) {
  double xi = x + sin(phi);
  double eta = y - 1.0 - cos(phi);
  double rho = 0.25 * (2.0 + sqrt(xi * xi + eta * eta));
  if (rho <= 1.0) {
    double uu = acos(rho), tt, vv;
    calc_tauOmega(uu, -uu, xi, eta, phi, &tt, &vv);
    if (tt >= 0.0 && vv <= 0.0) {
      *t = tt; *u = uu; *v = vv;
      return 1;
    }
  }
  return 0;
}
/* :395-417 */
static int LRLRp(double x, double y, double phi, double *t, double *u, double *v) {
  double xi = x + sin(phi);
  double eta = y - 1.0 - cos(phi);
  double rho = (20.0 - xi * xi - eta * eta) / 16.0;
  if (rho >= 0.0 && rho <= 1.0) {
    double uu = -acos(rho);
    if (uu >= -0.5 * PI) {
      double tt, vv;
      calc_tauOmega(uu, uu, xi, eta, phi, &tt, &vv);
      if (tt >= 0.0 && vv >= 0.0) {
        *t = tt; *u = uu; *v = vv;
        return 1;
      }
    }
  }
  return 0;
}
/* :475-494 */
static int LRSR(double x, double y, double phi, double *t, double *u, double *v) {
  double xi = x + sin(phi);
  double eta = y - 1.0 - cos(phi);
  double rho, theta;
  R(-eta, xi, &rho, &theta);
  if (rho >= 2.0) {
    double tt = theta;
    double uu = 2.0 - rho;
    double vv = M(tt + 0.5 * PI - phi);
    if (tt >= 0.0 && uu <= 0.0 && vv <= 0.0) {
      *t = tt; *u = uu; *v = vv;
      return 1;
    }
  }
  return 0;
}
/* :496-516 */
static int LRSL(double x, double y, double phi, double *t, double *u, double *v) {
  double xi = x - sin(phi);
  double eta = y - 1.0 + cos(phi);
  double rho, theta;
  R(xi, eta, &rho, &theta);
  if (rho >= 2.0) {
    double r = sqrt(rho * rho - 4.0);
    double uu = 2.0 - r;
    double tt = M(theta + atan2(r, -2.0));
    double vv = M(phi - 0.5 * PI - tt);
    if (tt >= 0.0 && uu <= 0.0 && vv <= 0.0) {
      *t = tt; *u = uu; *v = vv;
      return 1;
    }
  }
  return 0;
}
/* :626-650 */
static int LRSLR(double x, double y, double phi, double *t, double *u, double *v) {
  double xi = x + sin(phi);
  double eta = y - 1.0 - cos(phi);
  double rho, theta;
  R(xi, eta, &rho, &theta);
  if (rho >= 2.0) {
    double uu = 4.0 - sqrt(rho * rho - 4.0);
    if (uu <= 0.0) {
      double tt = M(atan2((4.0 - uu) * xi - 2.0 * eta, -2.0 * xi + (uu - 4.0) * eta));
      double vv = M(tt - phi);
      if (tt >= 0.0 && vv >= 0.0) {
        *t = tt; *u = uu; *v = vv;
        return 1;
      }
    }
  }
  return 0;
}

/* :795-817 (generate_path) with the six family drivers :201-215, :241-297, :299-357, :419-473,
 * :518-624, :652-682 inlined in their fixed order: 2 + 8 + 8 + 8 + 16 + 4 = 46 formula calls. */
static void generate_path(const double q0[3], const double q1[3], double maxc, cand_set *ps) {
  double dx = q1[0] - q0[0], dy = q1[1] - q0[1], dth = q1[2] - q0[2];
  double c = cos(q0[2]), s = sin(q0[2]);
  double x = (c * dx + s * dy) * maxc;
  double y = (-s * dx + c * dy) * maxc;
  double phi = dth, t, u, v, xb, yb;
  int f;
  ps->n = 0;
  /* SCS */
  f = SLS(x, y, phi, &t, &u, &v);    add3(ps, f, t, u, v, "SLS");
  f = SLS(x, -y, -phi, &t, &u, &v);  add3(ps, f, t, u, v, "SRS");
  /* CSC */
  f = LSL(x, y, phi, &t, &u, &v);    add3(ps, f, t, u, v, "LSL");
  f = LSL(-x, y, -phi, &t, &u, &v);  add3(ps, f, -t, -u, -v, "LSL");
  f = LSL(x, -y, -phi, &t, &u, &v);  add3(ps, f, t, u, v, "RSR");
  f = LSL(-x, -y, phi, &t, &u, &v);  add3(ps, f, -t, -u, -v, "RSR");
  f = LSR(x, y, phi, &t, &u, &v);    add3(ps, f, t, u, v, "LSR");
  f = LSR(-x, y, -phi, &t, &u, &v);  add3(ps, f, -t, -u, -v, "LSR");
  f = LSR(x, -y, -phi, &t, &u, &v);  add3(ps, f, t, u, v, "RSL");
  f = LSR(-x, -y, phi, &t, &u, &v);  add3(ps, f, -t, -u, -v, "RSL");
  /* CCC */
  f = LRL(x, y, phi, &t, &u, &v);    add3(ps, f, t, u, v, "LRL");
  f = LRL(-x, y, -phi, &t, &u, &v);  add3(ps, f, -t, -u, -v, "LRL");
  f = LRL(x, -y, -phi, &t, &u, &v);  add3(ps, f, t, u, v, "RLR");
  f = LRL(-x, -y, phi, &t, &u, &v);  add3(ps, f, -t, -u, -v, "RLR");
  xb = x * cos(phi) + y * sin(phi);
  yb = x * sin(phi) - y * cos(phi);
  f = LRL(xb, yb, phi, &t, &u, &v);    add3(ps, f, v, u, t, "LRL");
  f = LRL(-xb, yb, -phi, &t, &u, &v);  add3(ps, f, -v, -u, -t, "LRL");
  f = LRL(xb, -yb, -phi, &t, &u, &v);  add3(ps, f, v, u, t, "RLR");
  f = LRL(-xb, -yb, phi, &t, &u, &v);  add3(ps, f, -v, -u, -t, "RLR");
  /* CCCC */
  f = LRLRn(x, y, phi, &t, &u, &v);    add4(ps, f, t, u, -u, v, "LRLR");
  f = LRLRn(-x, y, -phi, &t, &u, &v);  add4(ps, f, -t, -u, u, -v, "LRLR");
  f = LRLRn(x, -y, -phi, &t, &u, &v);  add4(ps, f, t, u, -u, v, "RLRL");
  f = LRLRn(-x, -y, phi, &t, &u, &v);  add4(ps, f, -t, -u, u, -v, "RLRL");
  f = LRLRp(x, y, phi, &t, &u, &v);    add4(ps, f, t, u, u, v, "LRLR");
  f = LRLRp(-x, y, -phi, &t, &u, &v);  add4(ps, f, -t, -u, -u, -v, "LRLR");
  f = LRLRp(x, -y, -phi, &t, &u, &v);  add4(ps, f, t, u, u, v, "RLRL");
  f = LRLRp(-x, -y, phi, &t, &u, &v);  add4(ps, f, -t, -u, -u, -v, "RLRL");
  /* CCSC */
  f = LRSL(x, y, phi, &t, &u, &v);    add4(ps, f, t, -0.5 * PI, u, v, "LRSL");
  f = LRSL(-x, y, -phi, &t, &u, &v);  add4(ps, f, -t, 0.5 * PI, -u, -v, "LRSL");
  f = LRSL(x, -y, -phi, &t, &u, &v);  add4(ps, f, t, -0.5 * PI, u, v, "RLSR");
  f = LRSL(-x, -y, phi, &t, &u, &v);  add4(ps, f, -t, 0.5 * PI, -u, -v, "RLSR");
  f = LRSR(x, y, phi, &t, &u, &v);    add4(ps, f, t, -0.5 * PI, u, v, "LRSR");
  f = LRSR(-x, y, -phi, &t, &u, &v);  add4(ps, f, -t, 0.5 * PI, -u, -v, "LRSR");
  f = LRSR(x, -y, -phi, &t, &u, &v);  add4(ps, f, t, -0.5 * PI, u, v, "RLSL");
  f = LRSR(-x, -y, phi, &t, &u, &v);  add4(ps, f, -t, 0.5 * PI, -u, -v, "RLSL");
  xb = x * cos(phi) + y * sin(phi);
  yb = x * sin(phi) - y * cos(phi);
  f = LRSL(xb, yb, phi, &t, &u, &v);    add4(ps, f, v, u, -0.5 * PI, t, "LSRL");
  f = LRSL(-xb, yb, -phi, &t, &u, &v);  add4(ps, f, -v, -u, 0.5 * PI, -t, "LSRL");
  f = LRSL(xb, -yb, -phi, &t, &u, &v);  add4(ps, f, v, u, -0.5 * PI, t, "RSLR");
  f = LRSL(-xb, -yb, phi, &t, &u, &v);  add4(ps, f, -v, -u, 0.5 * PI, -t, "RSLR");
  f = LRSR(xb, yb, phi, &t, &u, &v);    add4(ps, f, v, u, -0.5 * PI, t, "RSRL");
  f = LRSR(-xb, yb, -phi, &t, &u, &v);  add4(ps, f, -v, -u, 0.5 * PI, -t, "RSRL");
  f = LRSR(xb, -yb, -phi, &t, &u, &v);  add4(ps, f, v, u, -0.5 * PI, t, "LSLR");
  f = LRSR(-xb, -yb, phi, &t, &u, &v);  add4(ps, f, -v, -u, 0.5 * PI, -t, "LSLR");
  /* CCSCC */
  f = LRSLR(x, y, phi, &t, &u, &v);    add5(ps, f, t, -0.5 * PI, u, -0.5 * PI, v, "LRSLR");
  f = LRSLR(-x, y, -phi, &t, &u, &v);  add5(ps, f, -t, 0.5 * PI, -u, 0.5 * PI, -v, "LRSLR");
  f = LRSLR(x, -y, -phi, &t, &u, &v);  add5(ps, f, t, -0.5 * PI, u, -0.5 * PI, v, "RLSRL");
  f = LRSLR(-x, -y, phi, &t, &u, &v);  add5(ps, f, -t, 0.5 * PI, -u, 0.5 * PI, -v, "RLSRL");
}

/* :750-793 */
static void interpolate(int ind, double l, char m, double maxc, double ox, double oy, double oyaw, double *px,
                        double *py, double *pyaw, int *directions) {
  if (m == 'S') {
    px[ind] = ox + l / maxc * cos(oyaw);
    py[ind] = oy + l / maxc * sin(oyaw);
    pyaw[ind] = oyaw;
  } else {
    double ldx = sin(l) / maxc;
    double ldy = 0.0;
    double gdx, gdy;
    if (m == 'L') ldy = (1.0 - cos(l)) / maxc;
    else if (m == 'R') ldy = (1.0 - cos(l)) / (-maxc);
    gdx = cos(-oyaw) * ldx + sin(-oyaw) * ldy;
    gdy = -sin(-oyaw) * ldx + cos(-oyaw) * ldy;
    px[ind] = ox + gdx;
    py[ind] = oy + gdy;
  }
  if (m == 'L') pyaw[ind] = oyaw + l;
  else if (m == 'R') pyaw[ind] = oyaw - l;
  directions[ind] = (l > 0.0) ? 1 : -1;
}

/* :684-748; returns the number of samples kept after the trailing "px == 0.0" trim */
static int generate_local_course(double L, const double *lengths, const char *mode, int nseg, double maxc,
                                 double step_size, double **opx, double **opy, double **opyaw, int **odir) {
  int point_num = (int)(L / step_size) + nseg + 3;
  double *px = (double *)calloc((size_t)point_num + 8, sizeof(double));
  double *py = (double *)calloc((size_t)point_num + 8, sizeof(double));
  double *pyaw = (double *)calloc((size_t)point_num + 8, sizeof(double));
  int *directions = (int *)calloc((size_t)point_num + 8, sizeof(int));
  int ind = 1, i, n;
  double d, pd, ll = 0.0;
  directions[0] = (lengths[0] > 0.0) ? 1 : -1;
  d = (lengths[0] > 0.0) ? step_size : -step_size;
  pd = d;
  for (i = 0; i < nseg; ++i) {
    double l = lengths[i], ox, oy, oyaw;
    char m = mode[i];
    d = (l > 0.0) ? step_size : -step_size;
    ox = px[ind];
    oy = py[ind];
    oyaw = pyaw[ind];
    --ind;
    if (i >= 1 && (lengths[i - 1] * lengths[i]) > 0) pd = -d - ll;
    else pd = d - ll;
    while (fabs(pd) <= fabs(l)) {
      ++ind;
      interpolate(ind, pd, m, maxc, ox, oy, oyaw, px, py, pyaw, directions);
      pd += d;
    }
    ll = l - pd - d;
    ++ind;
    interpolate(ind, l, m, maxc, ox, oy, oyaw, px, py, pyaw, directions);
  }
  n = point_num;
  while (n > 0 && px[n - 1] == 0.0) n--;
  *opx = px;
  *opy = py;
  *opyaw = pyaw;
  *odir = directions;
  return n;
}

/* :40-96 */
void orc_rs_all_paths(double sx, double sy, double syaw, double gx, double gy, double gyaw, double maxc,
                      double step_size, orc_rs_paths *out) {
  double q0[3], q1[3];
  cand_set ps;
  int k, i;
  q0[0] = sx; q0[1] = sy; q0[2] = syaw;
  q1[0] = gx; q1[1] = gy; q1[2] = gyaw;
  generate_path(q0, q1, maxc, &ps);
  out->n = ps.n;
  for (k = 0; k < ps.n; k++) {
    orc_rs_path *p = &out->p[k];
    double *x, *y, *yaw;
    int *dir, n;
    p->n_seg = ps.nseg[k];
    memcpy(p->types, ps.typ[k], 5);
    n = generate_local_course(ps.L[k], ps.len[k], ps.typ[k], ps.nseg[k], maxc, step_size * maxc, &x, &y, &yaw,
                              &dir);
    p->n = n;
    p->x = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    p->y = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    p->yaw = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    p->dir = dir;
    for (i = 0; i < n; i++) {
      p->x[i] = cos(-q0[2]) * x[i] + sin(-q0[2]) * y[i] + q0[0];
      p->y[i] = -sin(-q0[2]) * x[i] + cos(-q0[2]) * y[i] + q0[1];
      p->yaw[i] = pi_2_pi(yaw[i] + q0[2]);
    }
    for (i = 0; i < p->n_seg; i++) p->lengths[i] = ps.len[k][i] / maxc;
    p->L = ps.L[k] / maxc;
    free(x);
    free(y);
    free(yaw);
  }
}

void orc_rs_paths_free(orc_rs_paths *p) {
  int k;
  for (k = 0; k < p->n; k++) {
    free(p->p[k].x);
    free(p->p[k].y);
    free(p->p[k].yaw);
    free(p->p[k].dir);
  }
  p->n = 0;
}

/* RS/src/HybridAstar.cpp:276-313 */
static double reeds_path_cost(const orc_car *car, double parent_cost, const orc_rs_path *p) {
  const double reverse = 10.0, directionChange = 250.0, steerAngle = 15.0;
  double cost = parent_cost;
  int i;
  for (i = 0; i < p->n_seg; i++) {
    if (p->lengths[i] >= 0) cost += 1.0;
    else cost += fabs(p->lengths[i]) * reverse;
  }
  for (i = 0; i + 1 < p->n_seg; ++i)
    if (p->lengths[i] * p->lengths[i + 1] < 0) cost += directionChange;
  for (i = 0; i < p->n_seg; i++)
    if (p->types[i] != 'S') cost += car->max_steer * steerAngle;
  return cost;
}

static int path_collides(const orc_map *m, const orc_car *car, const orc_rs_path *p) {
  int i;
  for (i = 0; i < p->n; i++)
    if (orc_collide_circles(m, car, p->x[i], p->y[i], p->yaw[i])) return 1;
  return 0;
}

static void fill_shot(const orc_map *m, const orc_rs_path *p, double cost, int chosen, orc_shot *out) {
  int i;
  out->found = 1;
  out->cost = cost;
  out->chosen = chosen;
  out->n = p->n;
  out->traj = (orc_state *)malloc(sizeof(orc_state) * (size_t)(p->n + 1));
  for (i = 0; i < p->n; i++) {
    out->traj[i].x = p->x[i];
    out->traj[i].y = p->y[i];
    out->traj[i].heading = p->yaw[i];
    out->traj[i].gridx = orc_cell_x(m, p->x[i]);
    out->traj[i].gridy = orc_cell_y(m, p->y[i]);
  }
}

/* RS/src/HybridAstar.cpp:110-143 (dispatch), :145-210 (iterative), :212-274 (priority queue) */
void orc_rs_shot(const orc_map *m, const orc_car *car, const orc_state *cur, double cur_cost,
                 const orc_state *goal, orc_shot *out) {
  double radius = tan(car->max_steer) / car->wheel_base;
  double max_c = 1;
  orc_rs_paths paths;
  int k;
  memset(out, 0, sizeof *out);
  out->chosen = -1;
  orc_rs_all_paths(cur->x, cur->y, cur->heading, goal->x, goal->y, goal->heading, radius, max_c, &paths);
  out->n_paths = paths.n;
  if (paths.n == 0) return;
  if (paths.n > 5) {
    orc_heap_item heap[ORC_RS_MAX_PATHS];
    int hn = 0;
    for (k = 0; k < paths.n; k++) {
      orc_heap_item it;
      it.key = reeds_path_cost(car, cur_cost, &paths.p[k]);
      it.payload = k;
      orc_stdheap_push(heap, &hn, it);
    }
    while (hn > 0) {
      orc_heap_item it = orc_stdheap_pop(heap, &hn);
      if (!path_collides(m, car, &paths.p[it.payload])) {
        fill_shot(m, &paths.p[it.payload], it.key, it.payload, out);
        break;
      }
    }
  } else {
    double min_cost = INFINITY;
    int best = -1;
    for (k = 0; k < paths.n; k++) {
      if (!path_collides(m, car, &paths.p[k])) {
        double cost = reeds_path_cost(car, cur_cost, &paths.p[k]);
        if (cost < min_cost) {
          min_cost = cost;
          best = k;
        }
      }
    }
    if (best >= 0) fill_shot(m, &paths.p[best], min_cost, best, out);
  }
  orc_rs_paths_free(&paths);
}

void orc_shot_free(orc_shot *s) {
  free(s->traj);
  s->traj = NULL;
  s->n = 0;
}
