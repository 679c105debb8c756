/*
 * oracle/orc_map.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 * Restates: cell predicates (a2), obstacle/distance layers (a3), map_inflate (a12),
 * circle-footprint collision (a4) and SAT-footprint collision (a9).
 */
#include "orc.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ a2 */
/* RS/src/Grid_map.cpp:186-192 (toCellID): static_cast<int>((x - originX) / resolution), truncation */
int orc_cell_x(const orc_map *m, double x) { return (int)((x - m->ox) / m->res); }
int orc_cell_y(const orc_map *m, double y) { return (int)((y - m->oy) / m->res); }

/* RS/src/Grid_map.cpp:200-203 */
int orc_in_bounds(const orc_map *m, int x, int y) { return x >= 0 && x < m->W && y >= 0 && y < m->H; }

/* RS/src/Grid_map.cpp:205-213: out of bounds counts as collision; occupied <=> data > 0 */
int orc_in_collision(const orc_map *m, int x, int y) {
  if (!orc_in_bounds(m, x, y)) return 1;
  return m->occ[(size_t)y * m->W + x] > 0;
}

/* ------------------------------------------------------------------ a3 */
/*
 * cv::distanceTransform(src, dst, DIST_L2, DIST_MASK_PRECISE) as published in OpenCV's
 * imgproc/src/distransform.cpp (trueDistTrans: Felzenszwalb-Huttenlocher, column pass in
 * integers through a squares table, row pass with the lower envelope of parabolas).
 * OpenCV is an un-vendored dependency of the reference (find_package(OpenCV), no version pin;
 * ROS Humble => 4.5.4); this restatement is pinned against cv2 4.13 with IPP disabled
 * (tests/golden/edt_*.npz, tests/test_cpu_oracle_golden.py, and -- where cv2 is importable -- the
 * 8192-wide cfg4 map in tests/test_cpu_oracle_vs_ref.py).
 * The parabola intersections are evaluated in DOUBLE here: a float32 evaluation (q*q + f[q] with q > 4096 is
 * not representable in 24 bits) puts 11 % of the cells of the 8192^2 cfg4 map on the wrong parabola, while
 * cv2 4.13 returns the exact transform there (= scipy's exact EDT, 0 of 67 M cells differ); with doubles every
 * intersection of integer parabolas is ordered exactly, so this pass returns sqrtf of the exact squared
 * distance like cv2 4.13 does.  What OpenCV 4.5.4 itself returns on images wider than 4096 px is unpinned.
 * src is an m-row x n-col 8-bit image (row-major, step n); non-zero = free.
 */
static void true_dist_trans(const uint8_t *src, int m, int n, float *dst) {
  /* sqr_tab entries for "no zero pixel in this column".  Any value above every finite candidate gives the
   * same result whenever the image has at least one zero pixel; 2^32 reproduces what cv2 4.13 returns for an
   * image with none (the constant 65536). */
  const float inf_col = 4294967296.0f;
  const double inf = 1e15;              /* envelope sentinels of the row pass */
  float *sqr_tab = (float *)malloc(sizeof(float) * (size_t)(2 * (m > n ? m : n) + 2));
  int *d = (int *)malloc(sizeof(int) * (size_t)m);
  int i, j, k;
  for (i = 0; i < m; i++) sqr_tab[i] = (float)(i * i);
  for (i = m; i < 2 * m; i++) sqr_tab[i] = inf_col;
  /* stage 1: 1-D distance along each column (index i runs over rows) */
  for (k = 0; k < n; k++) {
    int dist = m - 1;
    for (j = m - 1; j >= 0; j--) {
      dist = (src[(size_t)j * n + k] == 0) ? 0 : dist + 1;
      d[j] = dist;
    }
    dist = m - 1;
    for (j = 0; j < m; j++) {
      dist = (dist + 1 < d[j]) ? dist + 1 : d[j]; /* dist + 1 - sat(dist + 1 - d[j]) */
      d[j] = dist;
      dst[(size_t)j * n + k] = sqr_tab[dist];
    }
  }
  free(d);
  /* stage 2: lower envelope of parabolas along each row */
  {
    float *f = (float *)malloc(sizeof(float) * (size_t)(n + 1));
    double *z = (double *)malloc(sizeof(double) * (size_t)(n + 2));
    int *v = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    sqr_tab[0] = 0.f;
    for (i = 1; i < n; i++) sqr_tab[i] = (float)(i * i); /* used below only for |q - p|, where it is exact */
    for (i = 0; i < m; i++) {
      float *row = dst + (size_t)i * n;
      int p, q;
      v[0] = 0;
      z[0] = -inf;
      z[1] = inf;
      f[0] = row[0];
      for (q = 1, k = 0; q < n; q++) {
        float fq = row[q];
        f[q] = fq;
        for (;; k--) {
          double s;
          p = v[k];
          s = (((double)fq + (double)q * q) - ((double)row[p] + (double)p * p)) / (2.0 * (q - p));
          if (s > z[k]) {
            k++;
            v[k] = q;
            z[k] = s;
            z[k + 1] = inf;
            break;
          }
        }
      }
      for (q = 0, k = 0; q < n; q++) {
        while (z[k + 1] < q) k++;
        p = v[k];
        row[q] = sqrtf(sqr_tab[abs(q - p)] + f[p]);
      }
    }
    free(f);
    free(z);
    free(v);
  }
  free(sqr_tab);
}

/*
 * RS/src/Grid_map.cpp:22-74.  grid_map size is (W,H); Eigen layers are W x H column-major, so
 * element (x,y) lives at [x + y*W].  obstacle(x,y) = 0 if data > free_thres_(20) or data < 0 else 1
 * (:35-51).  binary = obstacle.cast<uchar>; cv::eigen2cv of a column-major Eigen matrix yields a
 * W-row x H-col cv::Mat with Mat(x,y) = binary(x,y) (:56-60).  distanceTransform (:64), flip(-1)
 * (:67), cv2eigen (:71), then "distance * resolution" (:74) = float matrix times the scalar cast to
 * float (Eigen scalar promotion).  Net: distance[k] = edt[n-1-k] * (float)res  (hazard H7).
 */
void orc_edt_layers(const int8_t *occ, int W, int H, float resolution, float *obstacle, float *distance) {
  size_t n = (size_t)W * H, k;
  uint8_t *bin = (uint8_t *)malloc(n);  /* Mat (W rows, H cols): bin[x*H + y] */
  float *edt = (float *)malloc(sizeof(float) * n);
  int x, y;
  for (y = 0; y < H; y++)
    for (x = 0; x < W; x++) {
      int8_t v = occ[(size_t)y * W + x];
      int obst = ((double)v > 20.0) || (v < 0);
      if (obstacle) obstacle[(size_t)x + (size_t)y * W] = obst ? 0.0f : 1.0f;
      bin[(size_t)x * H + y] = obst ? 0 : 1;
    }
  true_dist_trans(bin, W, H, edt);
  /* flip(-1) of the W x H Mat: flipped(x,y) = edt(W-1-x, H-1-y); column-major Eigen index x + y*W */
  for (y = 0; y < H; y++)
    for (x = 0; x < W; x++) {
      k = (size_t)x + (size_t)y * W;
      distance[k] = edt[(size_t)(W - 1 - x) * H + (H - 1 - y)] * resolution;
    }
  free(bin);
  free(edt);
}

/* exact integer squared EDT by brute force over the column distances (test helper, O(W*H*H)) */
void orc_edt_sq_exact(const int8_t *occ, int W, int H, int64_t *d2) {
  const int64_t BIG = (int64_t)1 << 40;
  int64_t *g = (int64_t *)malloc(sizeof(int64_t) * (size_t)W * H);
  int x, y, yy;
  for (y = 0; y < H; y++) {
    int64_t dist = BIG;
    for (x = 0; x < W; x++) {
      int8_t v = occ[(size_t)y * W + x];
      int obst = ((double)v > 20.0) || (v < 0);
      dist = obst ? 0 : (dist >= BIG ? BIG : dist + 1);
      g[(size_t)y * W + x] = dist;
    }
    dist = BIG;
    for (x = W - 1; x >= 0; x--) {
      int8_t v = occ[(size_t)y * W + x];
      int obst = ((double)v > 20.0) || (v < 0);
      dist = obst ? 0 : (dist >= BIG ? BIG : dist + 1);
      if (dist < g[(size_t)y * W + x]) g[(size_t)y * W + x] = dist;
    }
  }
  for (y = 0; y < H; y++)
    for (x = 0; x < W; x++) {
      int64_t best = INT64_MAX;
      for (yy = 0; yy < H; yy++) {
        int64_t gx = g[(size_t)yy * W + x];
        if (gx >= BIG) continue;
        {
          int64_t c = gx * gx + (int64_t)(yy - y) * (yy - y);
          if (c < best) best = c;
        }
      }
      d2[(size_t)y * W + x] = best;
    }
  free(g);
}

/* ------------------------------------------------------------------ a12 */
/* cv::threshold(..., THRESH_OTSU) threshold selection as published in OpenCV imgproc/src/thresh.cpp
 * (getThreshVal_Otsu_8u); called at MI/src/map_inflate_node.cpp:127. */
int orc_otsu_threshold(const uint8_t *gray, size_t n) {
  int h[256];
  int i;
  size_t k;
  double mu = 0, scale = 1. / (double)n;
  double mu1 = 0, q1 = 0, max_sigma = 0, max_val = 0;
  memset(h, 0, sizeof h);
  for (k = 0; k < n; k++) h[gray[k]]++;
  for (i = 0; i < 256; i++) mu += i * (double)h[i];
  mu *= scale;
  for (i = 0; i < 256; i++) {
    double p_i, q2, mu2, sigma;
    p_i = h[i] * scale;
    mu1 *= q1;
    q1 += p_i;
    q2 = 1. - q1;
    if ((q1 < q2 ? q1 : q2) < FLT_EPSILON || (q1 > q2 ? q1 : q2) > 1. - FLT_EPSILON) continue;
    mu1 = (mu1 + i * p_i) / q1;
    mu2 = (mu - q1 * mu1) / q2;
    sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
    if (sigma > max_sigma) {
      max_sigma = sigma;
      max_val = i;
    }
  }
  return (int)max_val;
}

/* cv::getStructuringElement(MORPH_ELLIPSE, (2r+1)x(2r+1)) row half-widths, as published in OpenCV
 * imgproc/src/morph.dispatch.cpp: dx = cvRound(c*sqrt((r*r - dy*dy)*inv_r2)), round-half-even.
 * Called at MI/src/map_inflate_node.cpp:134 (hazard H8). */
void orc_ellipse_halfwidths(int r, int *dx) {
  int i;
  double inv_r2 = r ? 1. / ((double)r * r) : 0;
  for (i = 0; i < 2 * r + 1; i++) {
    int dy = i - r;
    dx[i] = (int)nearbyint(r * sqrt((r * r - dy * dy) * inv_r2)); /* default FE_TONEAREST = cvRound */
  }
}

/* MI/src/map_inflate_node.cpp:90-98 (to gray, vertical flip), :118-138 (InflateMap), :151-157 (back). */
void orc_inflate(const int8_t *occ, int W, int H, double resolution, double inflation_radius, int8_t *out) {
  size_t n = (size_t)W * H, k;
  uint8_t *gray = (uint8_t *)malloc(n), *inv = (uint8_t *)malloc(n), *dil = (uint8_t *)malloc(n);
  int i, j, r, thr, *dx;
  for (i = 0; i < H; i++)
    for (j = 0; j < W; j++) {
      int occ_prob = occ[(size_t)i * W + j];
      occ_prob = (occ_prob < 0) ? 100 : occ_prob;
      gray[(size_t)(H - i - 1) * W + j] = (uint8_t)(int)(255 - round(occ_prob * 255.0 / 100.0));
    }
  r = (int)round(inflation_radius / resolution);
  if (r < 1) r = 1;
  thr = orc_otsu_threshold(gray, n);
  for (k = 0; k < n; k++) inv[k] = (uint8_t) ~(gray[k] > thr ? 255 : 0); /* threshold then bitwise_not */
  dx = (int *)malloc(sizeof(int) * (size_t)(2 * r + 1));
  orc_ellipse_halfwidths(r, dx);
  for (i = 0; i < H; i++)
    for (j = 0; j < W; j++) {
      uint8_t best = 0;
      int di, dj;
      for (di = -r; di <= r && !best; di++) {
        int ii = i + di, w = dx[di + r];
        if (ii < 0 || ii >= H) continue; /* dilate border: pixels outside never win the max */
        for (dj = -w; dj <= w; dj++) {
          int jj = j + dj;
          if (jj < 0 || jj >= W) continue;
          if (inv[(size_t)ii * W + jj]) {
            best = 255;
            break;
          }
        }
      }
      dil[(size_t)i * W + j] = (uint8_t)~best;
    }
  for (i = 0; i < H; i++)
    for (j = 0; j < W; j++) {
      uint8_t g = dil[(size_t)(H - i - 1) * W + j];
      out[(size_t)i * W + j] = (int8_t)(100 - round(g * 100.0 / 255.0));
    }
  free(dx);
  free(gray);
  free(inv);
  free(dil);
}

/* ------------------------------------------------------------------ a4 */
/* RS/src/CarData.cpp:76-98 */
void orc_car_circles(const orc_car *car, double cx[4], double cy[4], double *r) {
  double shift = car->width / 4.0;
  *r = sqrt(2 * pow(shift, 2));
  cx[0] = -car->axle_to_back + shift; cy[0] = car->width / 2.0;   /* rl */
  cx[1] = -car->axle_to_back + shift; cy[1] = -car->width / 2.0;  /* rr */
  cx[2] = car->axle_to_front - shift; cy[2] = car->width / 2.0;   /* fl */
  cx[3] = car->axle_to_front - shift; cy[3] = -car->width / 2.0;  /* fr */
}

/*
 * grid_map::GridMap geometry as set at RS/src/Grid_map.cpp:15-25 and the grid_map_core
 * (un-vendored dependency, ros-humble-grid-map 2.0.x) functions used by
 * RS/src/Grid_map.cpp:222-237: isInside -> checkIfPositionWithinMap, getIndex ->
 * getIndexFromPosition, getPosition -> getPositionFromIndex, atPosition(INTER_LINEAR) ->
 * atPositionLinearInterpolated with INTER_NEAREST fallback.  Restated from the published
 * grid_map_core sources (GridMapMath.cpp, GridMap.cpp) as recalled: PARITY UNPINNED.
 */
typedef struct {
  double lenx, leny, posx, posy, res;
  int W, H;
} gm_geom;

static void gm_setup(const orc_map *m, gm_geom *g) {
  double length_x = m->W * m->res, length_y = m->H * m->res; /* :15-16 */
  g->res = m->res;
  g->W = m->W;
  g->H = m->H;
  g->posx = m->ox + length_x / 2.0; /* :19-20 */
  g->posy = m->oy + length_y / 2.0;
  /* setGeometry: size = round(length/res); length_ = size * res */
  g->lenx = (double)(int)round(length_x / m->res) * m->res;
  g->leny = (double)(int)round(length_y / m->res) * m->res;
}

static int gm_is_inside(const gm_geom *g, double px, double py) {
  double tx = -((px - g->posx) - 0.5 * g->lenx);
  double ty = -((py - g->posy) - 0.5 * g->leny);
  return tx >= 0.0 && ty >= 0.0 && tx < g->lenx && ty < g->leny;
}

static void gm_index(const gm_geom *g, double px, double py, int *i, int *j) {
  double vx = ((px - 0.5 * g->lenx) - g->posx) / g->res;
  double vy = ((py - 0.5 * g->leny) - g->posy) / g->res;
  *i = (int)(-vx);
  *j = (int)(-vy);
}

static int gm_index_in_range(const gm_geom *g, int i, int j) { return i >= 0 && j >= 0 && i < g->W && j < g->H; }

/* returns 0 and leaves *x,*y untouched when the index is out of range (as getPositionFromIndex does) */
static int gm_position(const gm_geom *g, int i, int j, double *x, double *y) {
  if (!gm_index_in_range(g, i, j)) return 0;
  *x = (g->posx + (0.5 * g->lenx - 0.5 * g->res)) + g->res * (double)(-i);
  *y = (g->posy + (0.5 * g->leny - 0.5 * g->res)) + g->res * (double)(-j);
  return 1;
}

static int gm_at_linear(const gm_geom *g, const float *layer, double px, double py, float *value) {
  int idx[4][2], shift[4], t, dirx;
  double ptx = 0, pty = 0, rx, ry, fx, fy;
  float f[4];
  size_t n = (size_t)g->W * g->H;
  gm_index(g, px, py, &idx[0][0], &idx[0][1]);
  gm_position(g, idx[0][0], idx[0][1], &ptx, &pty);
  if (px >= ptx) { idx[1][0] = idx[0][0] - 1; idx[1][1] = idx[0][1]; dirx = 1; }
  else           { idx[1][0] = idx[0][0] + 1; idx[1][1] = idx[0][1]; dirx = 0; }
  if (py >= pty) {
    idx[2][0] = idx[0][0]; idx[2][1] = idx[0][1] - 1;
    if (dirx) { shift[0] = 0; shift[1] = 1; shift[2] = 2; shift[3] = 3; }
    else      { shift[0] = 1; shift[1] = 0; shift[2] = 3; shift[3] = 2; }
  } else {
    idx[2][0] = idx[0][0]; idx[2][1] = idx[0][1] + 1;
    if (dirx) { shift[0] = 2; shift[1] = 3; shift[2] = 0; shift[3] = 1; }
    else      { shift[0] = 3; shift[1] = 2; shift[2] = 1; shift[3] = 0; }
  }
  idx[3][0] = idx[1][0];
  idx[3][1] = idx[2][1];
  for (t = 0; t < 4; t++) {
    /* getLinearIndexFromIndex (column-major): j*W + i as int, compared as size_t.  The library's
     * test is "lin > bufferSize" (so lin == bufferSize would read one past the end: UB); the oracle
     * and the product both treat lin >= bufferSize as "not interpolable". */
    long lin = (long)idx[shift[t]][1] * g->W + idx[shift[t]][0];
    if (lin < 0 || (size_t)lin >= n) return 0;
    f[t] = layer[lin];
  }
  gm_position(g, idx[shift[0]][0], idx[shift[0]][1], &ptx, &pty);
  rx = (px - ptx) / g->res;
  ry = (py - pty) / g->res;
  fx = 1.0 - rx;
  fy = 1.0 - ry;
  *value = (float)(f[0] * fx * fy + f[1] * rx * fy + f[2] * fx * ry + f[3] * rx * ry);
  return 1;
}

/* RS/src/Grid_map.cpp:239-284 with CarData::getCircles/local2Global (RS/src/CarData.cpp:100-124) */
int orc_collide_circles(const orc_map *m, const orc_car *car, double x, double y, double heading) {
  gm_geom g;
  double cx[4], cy[4], r;
  int c;
  gm_setup(m, &g);
  orc_car_circles(car, cx, cy, &r);
  for (c = 0; c < 4; c++) {
    double px = cx[c] * cos(heading) - cy[c] * sin(heading) + x;
    double py = cx[c] * sin(heading) + cy[c] * cos(heading) + y;
    float value;
    double clearance;
    if (!gm_is_inside(&g, px, py)) return 1;
    if (!gm_at_linear(&g, m->distance, px, py, &value)) {
      int i, j;
      gm_index(&g, px, py, &i, &j); /* INTER_NEAREST fallback */
      value = m->distance[(size_t)j * g.W + i];
    }
    clearance = (double)value;
    if (clearance < r) return 1;
  }
  return 0;
}

/* RS/src/Grid_map.cpp:286-317 isSingleStateCollisionFreeImproved with CarData::getBoundingCircle
 * (RS/src/CarData.cpp:78-81, :112-124); returns the method's value as the reference returns it */
int orc_collide_circles_improved(const orc_map *m, const orc_car *car, double x, double y, double heading) {
  gm_geom g;
  double bx = (car->axle_to_front - car->axle_to_back) / 2.0;
  double br = sqrt(pow((car->axle_to_front + car->axle_to_back) / 2.0, 2) + pow(car->width / 2.0, 2));
  double px = bx * cos(heading) - 0.0 * sin(heading) + x;
  double py = bx * sin(heading) + 0.0 * cos(heading) + y;
  float value;
  gm_setup(m, &g);
  if (!gm_is_inside(&g, px, py)) return 0;
  if (!gm_at_linear(&g, m->distance, px, py, &value)) {
    int i, j;
    gm_index(&g, px, py, &i, &j);
    value = m->distance[(size_t)j * g.W + i];
  }
  if ((double)value < br) return orc_collide_circles(m, car, x, y, heading);
  return 1;
}

/* the same check over n poses (xyh: n x 3), one call per pose exactly as above: the loop the CPU baselines time */
void orc_collide_circles_batch(const orc_map *m, const orc_car *car, const double *xyh, size_t n, uint8_t *verdict) {
  size_t i;
  for (i = 0; i < n; i++) verdict[i] = (uint8_t)orc_collide_circles(m, car, xyh[3 * i], xyh[3 * i + 1], xyh[3 * i + 2]);
}

/* ------------------------------------------------------------------ a9 */
typedef struct { float x, y; } pt32;

/* DB/src/sat_collision_checker.cc:80-103 (project) */
static void sat_project(const pt32 *p, double ax, double ay, double *mn, double *mx) {
  int i;
  for (i = 0; i < 5; i++) {
    double d = p[i].x * ax + p[i].y * ay;
    if (i == 0 || d < *mn) *mn = d;
    if (i == 0 || d > *mx) *mx = d;
  }
}

/* DB/src/sat_collision_checker.cc:131-161 (check_collision), :113-120 (is_overlapping) */
static int sat_check(const pt32 *rect, const pt32 *poly) {
  double ax[6], ay[6];
  int e, a;
  for (e = 0; e < 3; e++) { /* first 3 of the 4 edges of each closed 5-point polygon */
    float ex = rect[e + 1].x - rect[e].x, ey = rect[e + 1].y - rect[e].y; /* float32 subtraction (:35-36) */
    ax[e] = (double)ey;
    ay[e] = -(double)ex;
    ex = poly[e + 1].x - poly[e].x;
    ey = poly[e + 1].y - poly[e].y;
    ax[3 + e] = (double)ey;
    ay[3 + e] = -(double)ex;
  }
  for (a = 0; a < 6; a++) {
    double min1, max1, min2, max2;
    int overlap;
    sat_project(rect, ax[a], ay[a], &min1, &max1);
    sat_project(poly, ax[a], ay[a], &min2, &max2);
    if (min1 < min2) overlap = (min2 - max1) < 0;
    else overlap = (min1 - max2) < 0;
    if (!overlap) return 0;
  }
  return 1;
}

/* DB/src/Grid_map.cpp:50-95 (checkCollision), :97-123 (createObstaclePolygon),
 * DB/src/CarData.cpp:15-59 (vehicle rectangle, float32 vertices), DB/src/HybridAstar.cpp:52. */
int orc_collide_sat(const orc_map *m, const orc_car *car, double x, double y, double heading) {
  pt32 base[5], rect[5], poly[5];
  double offset = 5.0, min_x, max_x, min_y, max_y, res = m->res;
  int min_i, max_i, min_j, max_j, i, j, k, gx, gy;
  base[0].x = (float)car->axle_to_front;  base[0].y = (float)(car->width / 2.0);
  base[1].x = (float)car->axle_to_front;  base[1].y = (float)(-car->width / 2.0);
  base[2].x = (float)(-car->axle_to_back); base[2].y = (float)(-car->width / 2.0);
  base[3].x = (float)(-car->axle_to_back); base[3].y = (float)(car->width / 2.0);
  base[4] = base[0];
  for (k = 0; k < 5; k++) {
    double px = x + base[k].x * cos(heading) - base[k].y * sin(heading);
    double py = y + base[k].x * sin(heading) + base[k].y * cos(heading);
    rect[k].x = (float)px;
    rect[k].y = (float)py;
  }
  min_x = fmax(m->ox, x - offset);
  max_x = fmin(m->ox + m->W * res, x + offset);
  min_y = fmax(m->oy, y - offset);
  max_y = fmin(m->oy + m->H * res, y + offset);
  min_i = (int)((min_y - m->oy) / res); if (min_i < 0) min_i = 0;
  max_i = (int)((max_y - m->oy) / res); if (max_i > m->H - 1) max_i = m->H - 1;
  min_j = (int)((min_x - m->ox) / res); if (min_j < 0) min_j = 0;
  max_j = (int)((max_x - m->ox) / res); if (max_j > m->W - 1) max_j = m->W - 1;
  /* the reference stores these in uint32_t; a negative max (pose > 5 m outside the map) would wrap and
   * read out of bounds (UB).  The oracle treats that window as empty. */
  for (i = min_i; i <= max_i; i++)
    for (j = min_j; j <= max_j; j++) {
      if (m->occ[(size_t)i * m->W + j] > 0) {
        double px = m->ox + j * res, py = m->oy + i * res, half = res / 2.0;
        px += res / 2;
        py += res / 2;
        poly[0].x = (float)(px - half); poly[0].y = (float)(py - half);
        poly[1].x = (float)(px + half); poly[1].y = (float)(py - half);
        poly[2].x = (float)(px + half); poly[2].y = (float)(py + half);
        poly[3].x = (float)(px - half); poly[3].y = (float)(py + half);
        poly[4] = poly[0];
        if (sat_check(rect, poly)) return 1;
      }
    }
  gx = orc_cell_x(m, x);
  gy = orc_cell_y(m, y);
  return !orc_in_bounds(m, gx, gy);
}

void orc_collide_sat_batch(const orc_map *m, const orc_car *car, const double *xyh, size_t n, uint8_t *verdict) {
  size_t i;
  for (i = 0; i < n; i++) verdict[i] = (uint8_t)orc_collide_sat(m, car, xyh[3 * i], xyh[3 * i + 1], xyh[3 * i + 2]);
}

/* ------------------------------------------------------------------ f1 */
/* GP/src/grid_publisher.cpp:46-50: GridMapCvConverter::addLayerFromImage<unsigned char, 1>(img_src, "obstacle", map,
 * OCCUPY = 255, FREE = 0, 1):  mapValue = lower + (upper - lower) * ((float)pixel / 255), index (i, j) = pixel (i, j).
 * :78-86: GridMapRosConverter::toOccupancyGrid(map, "obstacle", 0, 255, grid): value = (mapValue - dataMin) /
 * (dataMax - dataMin), clamped to [0, 1], * 100, float -> int8; data[nCells - 1 - (j * rows + i)]; width = rows. */
void orc_image_to_occ(const uint8_t *img, int rows, int cols, int8_t *occ) {
  const size_t n = (size_t)rows * cols;
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) {
      const float lower = 255.0f, upper = 0.0f;
      const float map_value = lower + (upper - lower) * ((float)img[(size_t)i * cols + j] / 255.0f);
      float value = (map_value - 0.0f) / (255.0f - 0.0f);
      value = fminf(fmaxf(0.0f, value), 1.0f);
      value = 0.0f + value * 100.0f;
      occ[n - 1 - ((size_t)j * rows + i)] = (int8_t)value;
    }
}
