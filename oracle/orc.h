/*
 * oracle/orc.h -- CPU restatement ("oracle") of the hot path of
 * armando-genis/hybrid_path_planning.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it, and only as the
 * checker.  The product (libhpp_b200.so) never links or calls anything in oracle/.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference root; RS = reeds_shepp_hybrid_astar, DB = dubins_hybrid_star,
 * MI = map_inflate).  Plain C99, no FMA contraction (-ffp-contract=off), glibc libm.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - a1/a5/a6/a7 (RS) and a8/a9/a10/a11 (Dubins) are pinned against the reference's own
 *     translation units compiled into oracle/_ref (tests/test_cpu_oracle_vs_ref.py, and the
 *     committed tests/golden/ vectors generated from them).
 *   - a3 (EDT layers) and a12 (inflate) are pinned against cv2 (IPP off) goldens.
 *   - a4 (grid_map::atPosition bilinear + isInside) restates grid_map_core from its
 *     published source as recalled; grid_map_core is not vendored in the reference
 *     => "parity unpinned" for that third-party piece.
 */
#ifndef ORC_H
#define ORC_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  double max_steer;      /* maxSteerAngle */
  double wheel_base;     /* wheelBase     */
  double axle_to_front;  /* axleToFront   */
  double axle_to_back;   /* axleToBack    */
  double width;          /* width         */
} orc_car;

/* Occupancy map as the planners hold it (RS/src/Grid_map.cpp:3-12). */
typedef struct {
  const int8_t *occ; /* row-major data[y*W+x] */
  int W, H;
  double res; /* (double)(float)info.resolution  (hazard H6) */
  double ox, oy;
  const float *distance; /* grid_map "distance" layer, column-major [x + y*W] (may be NULL) */
} orc_map;

/* ---- a2: cell predicates (RS/src/Grid_map.cpp:186-213) ---- */
int orc_cell_x(const orc_map *m, double x);
int orc_cell_y(const orc_map *m, double y);
int orc_in_bounds(const orc_map *m, int x, int y);
int orc_in_collision(const orc_map *m, int x, int y);

/* ---- a3: obstacle + distance layers (RS/src/Grid_map.cpp:22-74) ---- */
void orc_edt_layers(const int8_t *occ, int W, int H, float resolution, float *obstacle, float *distance);
/* exact squared EDT (integer) of the same binary image; used to cross-check the float path */
void orc_edt_sq_exact(const int8_t *occ, int W, int H, int64_t *d2);

/* ---- a12: map_inflate (MI/src/map_inflate_node.cpp:65-161) ---- */
void orc_inflate(const int8_t *occ, int W, int H, double resolution, double inflation_radius, int8_t *out);
/* f1: image -> OccupancyGrid as grid_map_publisher does it (GP/src/grid_publisher.cpp:32-50,78-86; grid_map_cv /
 * grid_map_ros arithmetic restated as recalled: parity unpinned).  img rows x cols, occ [y*rows + x], W = rows, H = cols */
void orc_image_to_occ(const uint8_t *img, int rows, int cols, int8_t *occ);
int orc_otsu_threshold(const uint8_t *gray, size_t n);
void orc_ellipse_halfwidths(int r, int *dx /* 2r+1 entries, -1 = empty row */);

/* ---- a4: circle footprint collision (RS/src/Grid_map.cpp:222-284, RS/src/CarData.cpp:76-124) ---- */
/* returns 1 = collision or out of map (reference polarity, hazard H5) */
int orc_collide_circles(const orc_map *m, const orc_car *car, double x, double y, double heading);
void orc_car_circles(const orc_car *car, double cx[4], double cy[4], double *r);
int orc_collide_circles_improved(const orc_map *m, const orc_car *car, double x, double y, double heading);
void orc_collide_circles_batch(const orc_map *m, const orc_car *car, const double *xyh, size_t n, uint8_t *verdict);

/* ---- a9: SAT footprint collision (DB/src/Grid_map.cpp:50-123, DB/src/sat_collision_checker.cc:12-161) ---- */
/* returns checkCollision(state, poly) || !isPointInBounds(gridx, gridy)  (DB/src/HybridAstar.cpp:52) */
int orc_collide_sat(const orc_map *m, const orc_car *car, double x, double y, double heading);
void orc_collide_sat_batch(const orc_map *m, const orc_car *car, const double *xyh, size_t n, uint8_t *verdict);

/* ---- a1: RS heuristic (RS/src/HybridAstar.cpp:353-551) ---- */
typedef struct {
  int h1_fired;          /* progress-check break dropped part of a bucket (hazard H1) */
  int h2_clamped;        /* some cost reached bucket MAX_COST-1 (hazard H2) */
  int64_t processed;     /* settled cells */
  double max_cost;       /* largest finite cost */
} orc_quirks;
/* emulate_quirks=1: the reference's exact sequential behaviour incl. H1/H2.
 * emulate_quirks=0: the order-independent fp64 fixed point (SURVEY hazard H4). */
void orc_heuristic_rs(const int8_t *occ, int W, int H, int goal_x, int goal_y, int start_x, int start_y,
                      int emulate_quirks, double *cost, orc_quirks *q);

/* ---- a8: Dubins heuristic (DB/src/HybridAstar.cpp:198-313) ---- */
void orc_heuristic_dubins(const int8_t *occ, int W, int H, int goal_x, int goal_y, double *cost);

/* ---- a5/a10: successor expansion ---- */
typedef struct {
  double x, y, heading;  /* State */
  double gridx, gridy;
} orc_state;

typedef struct {
  orc_state traj[8];     /* simulationLength states (<= 8) */
  int n_traj;
  double cost, steer;
  int direction;
  int valid;             /* 0 = trajectory rejected by the collision check */
} orc_child;

#define ORC_VARIANT_RS 0
#define ORC_VARIANT_DUBINS 1

/* steer commands: RS/src/HybridAstar.cpp:14-23 (26) / DB/src/HybridAstar.cpp:14-21 (13) */
int orc_motion_commands(const orc_car *car, int variant, double *steer, double *dir /* >= 64 entries */);
/* GetnextNeighbours + PathCost; children has one entry per motion command, in command order */
int orc_expand(const orc_map *m, const orc_car *car, int variant, double sim_len, double step,
               const orc_state *parent, double parent_cost, double parent_steer, int parent_dir,
               orc_child *children);

/* ---- a7: Reeds-Shepp all paths (RS/src/reeds_shepp_path.cpp:40-860) ---- */
#define ORC_RS_MAX_PATHS 48
typedef struct {
  int n_seg;
  double lengths[5];   /* metres (already divided by maxc), signed */
  char types[5];       /* 'S','L','R' */
  double L;            /* metres */
  int n;               /* samples */
  double *x, *y, *yaw; /* malloc'ed */
  int *dir;
} orc_rs_path;
typedef struct {
  int n;
  orc_rs_path p[ORC_RS_MAX_PATHS];
} orc_rs_paths;
void orc_rs_all_paths(double sx, double sy, double syaw, double gx, double gy, double gyaw, double maxc,
                      double step_size, orc_rs_paths *out);
void orc_rs_paths_free(orc_rs_paths *p);

/* ---- a6: RS analytic shot (RS/src/HybridAstar.cpp:110-313) ---- */
typedef struct {
  int found;
  double cost;
  int n;               /* trajectory samples */
  orc_state *traj;     /* malloc'ed */
  int n_paths;         /* candidates generated */
  int chosen;          /* index of the chosen candidate in generation order, -1 if none */
} orc_shot;
void orc_rs_shot(const orc_map *m, const orc_car *car, const orc_state *cur, double cur_cost,
                 const orc_state *goal, orc_shot *out);
void orc_shot_free(orc_shot *s);

/* ---- a11: Dubins path + shot (DB/src/dubins_path.cpp:13-374, DB/src/HybridAstar.cpp:108-184) ---- */
typedef struct {
  double length;       /* normalised best cost |t|+|p|+|q| */
  char mode[3];        /* empty path: n == 0 */
  double t, p, q;
  int n;
  double *x, *y, *yaw; /* malloc'ed */
} orc_dubins;
void orc_dubins_path(double sx, double sy, double syaw, double gx, double gy, double gyaw, double curv,
                     double step_size, orc_dubins *out);
void orc_dubins_free(orc_dubins *d);
void orc_dubins_shot(const orc_map *m, const orc_car *car, const orc_state *cur, double cur_cost,
                     const orc_state *goal, orc_shot *out);

/* ---- host loop: HybridAstar::run (RS/src/HybridAstar.cpp:553-729, DB/src/HybridAstar.cpp:315-496) ---- */
typedef struct {
  int n;               /* states in the returned trajectory (1 = failure, hazard H10) */
  orc_state *traj;     /* malloc'ed */
  int iterations;
  int via_shot;        /* terminated by an analytic shot */
  double cost;         /* Cost_path of the terminating node (0 on failure) */
} orc_plan;
void orc_run(const orc_map *m, const orc_car *car, int variant, double sim_len, double step,
             double sx, double sy, double sh, double gx, double gy, double gh, int heuristic_quirks,
             orc_plan *out);
void orc_plan_free(orc_plan *p);

#ifdef __cplusplus
}
#endif
#endif
