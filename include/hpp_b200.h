/*
 * hpp_b200.h -- C ABI of libhpp_b200.so: the B200-native (sm_100a) data-parallel core of the two
 * Hybrid A* planners of armando-genis/hybrid_path_planning.
 *
 * The reference has no FFI layer: its drop-in boundary is the public methods of the C++ classes
 * Grid_map / HybridAstar / CarData / reeds_shepp_path / dubins_path and map_inflate_node::InflateMap.
 * Each entry point below names the reference method it replaces (paths relative to the reference
 * root; RS = reeds_shepp_hybrid_astar, DB = dubins_hybrid_star, MI = map_inflate).  The host C++
 * shells in hybrid_path_planning_b200/host/ keep those class/method signatures and call this ABI.
 *
 * Conventions
 *  - plain pointers and sizes only; the caller owns every host buffer; the library owns device
 *    memory behind the opaque hpp_ctx / hpp_map handles; no callbacks into the host.
 *  - every function returns an hpp_status (0 = ok); hpp_last_error(ctx) gives the text.  No exceptions.
 *  - one hpp_ctx per (thread, GPU); it owns one CUDA stream.  Functions with host pointers are
 *    synchronous (H2D + kernels + D2H inside the call).  "_dev" variants take device pointers, only
 *    enqueue work on the ctx stream and return; call hpp_ctx_sync() before reading results.
 *  - there is NO CPU fallback: every function fails with HPP_ERR_CUDA when no sm_100 device is usable.
 *  - occupancy is the nav_msgs/OccupancyGrid payload: int8 row-major data[y*W + x]
 *    (RS/src/Grid_map.cpp:194-198); resolution is the message's float32 (hazard H6).
 *  - collision verdicts keep the reference's polarity: 1 = collision or outside the map
 *    (RS/src/Grid_map.cpp:259-283 returns true on collision).
 */
#ifndef HPP_B200_H
#define HPP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  HPP_OK = 0,
  HPP_ERR_ARG = 1,       /* null pointer, zero size, goal outside the map, ... */
  HPP_ERR_CUDA = 2,      /* CUDA runtime failure or no usable device */
  HPP_ERR_NOMEM = 3,     /* host or device allocation failed */
  HPP_ERR_CAPACITY = 4,  /* an output buffer supplied by the caller was too small (results truncated) */
  HPP_ERR_UNSUPPORTED = 5 /* input outside the supported envelope (e.g. W >= 65535) */
} hpp_status;

#define HPP_VARIANT_RS 0     /* reeds_shepp_hybrid_astar */
#define HPP_VARIANT_DUBINS 1 /* dubins_hybrid_star */

typedef struct hpp_ctx hpp_ctx;
typedef struct hpp_map hpp_map;

/* ROS 2 node parameters of both planners (RS/src/main_path_planning.cpp:5-21, GL/config/reeds_config.yaml). */
typedef struct {
  double max_steer;     /* maxSteerAngle */
  double wheel_base;    /* wheelBase */
  double axle_to_front; /* axleToFront */
  double axle_to_back;  /* axleToBack */
  double width;         /* width */
  double path_length;   /* pathLength  -> HybridAstar::simulationLength */
  double step_car;      /* step_car */
} hpp_params;

/* ---------------------------------------------------------------- context */
hpp_status hpp_ctx_create(int device, hpp_ctx **out);
/* Maps and communicators keep their context alive: hpp_ctx_destroy gives up the caller's reference, and the stream and
 * scratch memory are released with the last map / communicator of the context (so a garbage-collected binding may
 * finalise its handles in any order).  The ctx pointer itself must not be passed to any call after hpp_ctx_destroy. */
hpp_status hpp_ctx_destroy(hpp_ctx *ctx);
hpp_status hpp_ctx_sync(hpp_ctx *ctx);
void *hpp_ctx_stream(hpp_ctx *ctx); /* cudaStream_t the ctx enqueues on */
const char *hpp_last_error(hpp_ctx *ctx);
/* number of kernel launches this ctx has enqueued so far (bench.py's gpu_launches) */
uint64_t hpp_ctx_launch_count(hpp_ctx *ctx);
const char *hpp_version(void);

/* ---------------------------------------------------------------- map: Grid_map
 * Replaces Grid_map::Grid_map(const nav_msgs::msg::OccupancyGrid&) (RS/include/Grid_map.h:85,
 * RS/src/Grid_map.cpp:3-81; DB/src/Grid_map.cpp:3-21) and Grid_map::setcarData (RS/include/Grid_map.h:99).
 * Uploads the occupancy, builds the grid_map "obstacle" and "distance" layers on the device (exact
 * Euclidean distance transform) plus the per-cell tables the other kernels use.
 * origin_x/origin_y = info.origin.position.{x,y}. */
hpp_status hpp_map_create(hpp_ctx *ctx, const int8_t *occ, uint32_t W, uint32_t H, float resolution,
                          double origin_x, double origin_y, const hpp_params *params, hpp_map **out);
/* same, occupancy already in device memory (copied; the caller keeps ownership of occ_dev) */
hpp_status hpp_map_create_dev(hpp_ctx *ctx, const int8_t *occ_dev, uint32_t W, uint32_t H, float resolution,
                              double origin_x, double origin_y, const hpp_params *params, hpp_map **out);
hpp_status hpp_map_destroy(hpp_map *map);
/* Grid_map::setcarData (RS/include/Grid_map.h:99, RS/src/Grid_map.cpp:217-220): replace the vehicle / planner
 * parameters of an existing map (footprint circles, rectangle, motion commands); no device work */
hpp_status hpp_map_set_params(hpp_map *map, const hpp_params *params);
/* re-run the layer build on new occupancy of the same geometry (the reference rebuilds its Grid_map on
 * every 1 Hz map message, RS/src/main_path_planning.cpp:63-69) */
hpp_status hpp_map_update(hpp_map *map, const int8_t *occ);
hpp_status hpp_map_update_dev(hpp_map *map, const int8_t *occ_dev);
/* copy the layers out in grid_map's memory layout: Eigen::MatrixXf (W rows x H cols, column-major), i.e.
 * obstacle[x + y*W] (1 = free, 0 = obstacle, RS/src/Grid_map.cpp:35-51) and
 * distance[k] = edt[n-1-k] * (float)resolution (RS/src/Grid_map.cpp:56-74).  Either pointer may be NULL. */
hpp_status hpp_map_layers(hpp_map *map, float *obstacle, float *distance);
/* device views of the same layers (valid until the map is destroyed or updated) */
hpp_status hpp_map_layers_dev(hpp_map *map, const float **obstacle_dev, const float **distance_dev);

/* Grid_map::toCellID / toCellIndex / isPointInBounds / isInCollision (RS/src/Grid_map.cpp:186-213) for
 * n world points: cell_xy[2i], cell_xy[2i+1] = truncated cell ids; flags[i] bit0 = in bounds,
 * bit1 = isInCollision.  xy is n x 2 doubles. */
hpp_status hpp_map_cells(hpp_map *map, const double *xy, size_t n, int32_t *cell_xy, uint8_t *flags);

/* ---------------------------------------------------------------- map_inflate
 * Replaces map_inflate_node::MapCallback/InflateMap/CreateInflatedOccupancyGrid
 * (MI/src/map_inflate_node.cpp:65-161): gray conversion, Otsu binarisation, dilation with OpenCV's
 * MORPH_ELLIPSE element of radius max(1, round(inflation_radius / resolution)), back-conversion. */
hpp_status hpp_inflate(hpp_ctx *ctx, const int8_t *occ, uint32_t W, uint32_t H, double resolution,
                       double inflation_radius, int8_t *out);
hpp_status hpp_inflate_dev(hpp_ctx *ctx, const int8_t *occ_dev, uint32_t W, uint32_t H, double resolution,
                           double inflation_radius, int8_t *out_dev);

/* ---------------------------------------------------------------- image -> OccupancyGrid
 * Replaces the map source of the demo stack: grid_map_publisher's image -> grid_map "obstacle" layer ->
 * nav_msgs/OccupancyGrid conversion (GP/src/grid_publisher.cpp:32-50,78-86: GridMapCvConverter::addLayerFromImage
 * <unsigned char,1>(img, "obstacle", map, 255, 0, 1) then GridMapRosConverter::toOccupancyGrid(map, "obstacle", 0, 255)).
 * img: rows x cols 8-bit gray, row-major.  The grid gets width = rows, height = cols, and cell (x, y) comes from pixel
 * (rows-1-x, cols-1-y); value = (int8)(clamp(((255 - 255 * (v / 255.f)) - 0) / 255, 0, 1) * 100) in float32:
 * black -> 100, white -> 0.  occ: cols x rows int8, row-major [y*rows + x]. */
hpp_status hpp_image_to_occupancy(hpp_ctx *ctx, const uint8_t *img, uint32_t rows, uint32_t cols, int8_t *occ);
hpp_status hpp_image_to_occupancy_dev(hpp_ctx *ctx, const uint8_t *img_dev, uint32_t rows, uint32_t cols,
                                      int8_t *occ_dev);

/* ---------------------------------------------------------------- footprint collision
 * hpp_collide_circles replaces Grid_map::isSingleStateCollisionFree (RS/include/Grid_map.h:102,
 * RS/src/Grid_map.cpp:239-284) with CarData::getCircles (RS/src/CarData.cpp:100-110).
 * hpp_collide_sat replaces the Dubins planner predicate
 *   checkCollision(state, getVehicleGeometry_state(state)) || !isPointInBounds(gridx, gridy)
 * (DB/include/Grid_map.h:60, DB/src/Grid_map.cpp:50-95, DB/src/HybridAstar.cpp:52).
 * xyh is n x 3 doubles (x, y, heading); verdict[i] = 1 collision/outside, 0 free. */
hpp_status hpp_collide_circles(hpp_map *map, const double *xyh, size_t n, uint8_t *verdict);
hpp_status hpp_collide_circles_dev(hpp_map *map, const double *xyh_dev, size_t n, uint8_t *verdict_dev);
hpp_status hpp_collide_sat(hpp_map *map, const double *xyh, size_t n, uint8_t *verdict);
hpp_status hpp_collide_sat_dev(hpp_map *map, const double *xyh_dev, size_t n, uint8_t *verdict_dev);
/* Grid_map::isSingleStateCollisionFreeImproved (RS/include/Grid_map.h:103, RS/src/Grid_map.cpp:286-317): the bounding
 * circle of CarData::getBoundingCircle (RS/src/CarData.cpp:78-81, :112-116) as a pre-filter of the four-circle check.
 * result[i] is the method's return value AS THE REFERENCE RETURNS IT, which is not the convention of
 * isSingleStateCollisionFree: 1 when the bounding circle is clear of obstacles, 0 when its centre lies outside the map,
 * and isSingleStateCollisionFree's value (1 = collision) when the bounding circle touches an obstacle.  The reference
 * never calls the method; it is provided so that code written against the class finds it. */
hpp_status hpp_collide_improved(hpp_map *map, const double *xyh, size_t n, uint8_t *result);
hpp_status hpp_collide_improved_dev(hpp_map *map, const double *xyh_dev, size_t n, uint8_t *result_dev);

/* ---------------------------------------------------------------- holonomic-with-obstacles heuristic
 * Replaces HybridAstar::holonomicCostsWithObstacles_planning (RS/include/HybridAstar.h:62,
 * RS/src/HybridAstar.cpp:353-551; DB/include/HybridAstar.h:51, DB/src/HybridAstar.cpp:198-313).
 * cost is W*H doubles, row-major [y*W + x] like the reference's return value (RS: +inf where
 * unreachable; Dubins: 1000).  The device computes the order-independent fp64 fixed point, which is
 * bit-identical to the reference wherever its sequential quirks do not fire; hpp_quirk_report says
 * whether they could have (SURVEY hazards H1/H2). */
typedef struct {
  uint64_t reachable;    /* cells with a finite (< 1000 for Dubins) cost, goal included */
  double max_cost;       /* largest finite cost */
  int32_t h1_progress_break; /* 0 = cannot fire, 1 = fires for certain, 2 = may fire (start reachable) */
  int32_t h2_bucket_clamp;   /* 1 = some cost >= max(5000, W+H) - 1: reference's last bucket is FIFO */
  uint32_t rounds;       /* scheduling rounds of the persistent kernel (0 for the asynchronous kernel) */
  uint32_t tile_visits;  /* tile visits (a 32x32 tile staged, relaxed under the cost window, written back) */
} hpp_quirk_report;

hpp_status hpp_heuristic_rs(hpp_map *map, int32_t goal_x, int32_t goal_y, int32_t start_x, int32_t start_y,
                            double *cost, hpp_quirk_report *report /* may be NULL */);
hpp_status hpp_heuristic_dubins(hpp_map *map, int32_t goal_x, int32_t goal_y, double *cost,
                                hpp_quirk_report *report /* may be NULL */);
/* batched, device-resident: n_fields goal cells (goal_xy: n_fields x 2 int32, HOST memory), one W*H field
 * each written to cost_dev + f*W*H.  All fields are relaxed by ONE persistent launch (cost-window tile scheduling
 * on the device, no host polling); the call returns after that launch has completed, so a device fault is
 * reported here.  Many fields per call is the efficient form: every field gets its own group of CTAs. */
hpp_status hpp_heuristic_batch_dev(hpp_map *map, int variant, const int32_t *goal_xy, size_t n_fields,
                                   double *cost_dev, hpp_quirk_report *reports /* n_fields or NULL */);

/* same with a host output buffer: cost = n_fields x W*H doubles (pinned memory recommended); the multi-query
 * driver's form of holonomicCostsWithObstacles_planning (one field per distinct goal cell of a batch of queries) */
hpp_status hpp_heuristic_batch(hpp_map *map, int variant, const int32_t *goal_xy, size_t n_fields, double *cost,
                               hpp_quirk_report *reports /* n_fields or NULL */);

/* ---------------------------------------------------------------- successor expansion
 * Replaces HybridAstar::GetnextNeighbours + PathCost (RS/include/HybridAstar.h:52-53,
 * RS/src/HybridAstar.cpp:25-107; DB/src/HybridAstar.cpp:23-105) for n parents at once. */
typedef struct {
  double x, y, heading; /* State */
  double cost;          /* Node::Cost_path */
  double steer;         /* Node::steeringAngle */
  int32_t direction;    /* Node::direction */
  int32_t pad_;
} hpp_node;

#define HPP_MAX_COMMANDS 26
#define HPP_MAX_SIM 8
typedef struct {
  double traj[HPP_MAX_SIM][5]; /* x, y, heading, gridx, gridy per Euler step (State) */
  double cost;
  double steer;
  int32_t direction;
  int32_t valid;               /* 0 = a state of the rollout collides: the reference drops the child */
} hpp_child;

/* number of motion commands of a variant (26 / 13 with the default steer_ratio 6) */
int32_t hpp_num_commands(const hpp_params *params, int variant);
/* children: n x hpp_num_commands entries, command-major per parent, in the reference's command order */
hpp_status hpp_expand(hpp_map *map, int variant, const hpp_node *parents, size_t n, hpp_child *children);
hpp_status hpp_expand_dev(hpp_map *map, int variant, const hpp_node *parents_dev, size_t n,
                          hpp_child *children_dev);

/* ---------------------------------------------------------------- analytic expansion ("shots")
 * hpp_rs_shot replaces HybridAstar::reeds_shepp_Path{,_iterative,_priority_queue} + reeds_Path_Cost
 * (RS/include/HybridAstar.h:56-59, RS/src/HybridAstar.cpp:110-313) on top of
 * reeds_shepp_path::calc_all_paths (RS/include/reeds_shepp_path.h:51-53, RS/src/reeds_shepp_path.cpp:40-96).
 * hpp_dubins_shot replaces HybridAstar::dubins_Path + dubins_Path_Cost (DB/include/HybridAstar.h:47,
 * DB/src/HybridAstar.cpp:108-184) on top of dubins_path::calc_dubins_path (DB/include/dubins_path.h:57).
 * One shot per start node towards its goal; goals is n x 3 doubles (or 1 x 3 when goal_stride == 0). */
typedef struct {
  int32_t found;         /* 0 = the reference would return nullptr */
  int32_t n_paths;       /* candidate paths generated (RS: after de-duplication) */
  int32_t chosen;        /* index of the chosen candidate in generation order, -1 if none */
  int32_t n_samples;     /* states of the chosen trajectory */
  double cost;           /* Cost_path of the resulting node */
  uint64_t sample_offset; /* first state of the trajectory in samples[] (states of 5 doubles) */
} hpp_shot;

hpp_status hpp_rs_shot(hpp_map *map, const hpp_node *starts, size_t n, const double *goals_xyh,
                       size_t goal_stride, hpp_shot *shots, double *samples /* cap x 5 */, size_t samples_cap,
                       size_t *samples_used);
hpp_status hpp_dubins_shot(hpp_map *map, const hpp_node *starts, size_t n, const double *goals_xyh,
                           size_t goal_stride, hpp_shot *shots, double *samples, size_t samples_cap,
                           size_t *samples_used);

/* unit-level parity hook for reeds_shepp_path::calc_all_paths (a7): all de-duplicated candidates of ONE
 * pose pair.  lengths: max_paths x 5 (metres, signed), types: max_paths x 5 chars, n_seg/n_samples per
 * path; samples: concatenated x,y,yaw triples. */
hpp_status hpp_rs_all_paths(hpp_ctx *ctx, const double start_xyh[3], const double goal_xyh[3], double maxc,
                            double step_size, int32_t max_paths, int32_t *n_paths, int32_t *n_seg,
                            double *lengths, char *types, double *total_len, int32_t *n_samples,
                            double *samples_xyyaw, size_t samples_cap,
                            int32_t *directions /* may be NULL: PATH::directions, one per sample (+1 / -1) */);
/* idem for dubins_path::calc_dubins_path (a11): best word + samples */
hpp_status hpp_dubins_path(hpp_ctx *ctx, const double start_xyh[3], const double goal_xyh[3], double curv,
                           double step_size, double *length, char mode[3], int32_t *n_samples,
                           double *samples_xyyaw, size_t samples_cap);

/* unit-level hook for the planner's open list: replays n_ops operations (op[i] >= 0: push key[op[i]] with payload
 * i; op[i] < 0: pop) twice on the device -- with the one-thread libstdc++ heap replay and with the warp-collective
 * pop / push the planner kernel uses -- and reports the first operation after which the two heaps or the popped
 * payloads differ (-1: identical throughout).  pops_out (n_ops ints, may be NULL): payload of every pop, -1 elsewhere. */
hpp_status hpp_test_heap_ops(hpp_ctx *ctx, const double *keys, const int32_t *ops, size_t n_ops, int32_t *pops_out,
                             int64_t *first_mismatch);

/* ---------------------------------------------------------------- whole queries
 * Replaces HybridAstar::run (RS/include/HybridAstar.h:47, RS/src/HybridAstar.cpp:553-729;
 * DB/src/HybridAstar.cpp:315-496) for a batch of independent start/goal queries on one map: heuristic
 * fields, open lists, shots and expansions all stay on the device; pop order, tie-breaking, the
 * 3500-iteration cap and the failure convention (a 1-state trajectory = the goal, hazard H10) follow the
 * reference.  starts/goals: n x 3 doubles. */
typedef struct {
  int32_t n_states;      /* trajectory length; 1 = failure */
  int32_t iterations;
  int32_t via_shot;      /* ended by an analytic shot */
  int32_t status;        /* 0 ok, 1 = node pool exhausted (query aborted as failure), 2 = the start pose's cell index
                          * lies outside [0, W*H): the reference indexes its heuristic vector out of range there
                          * (RS/src/HybridAstar.cpp:597, undefined behaviour); reported as a failure (1-state trajectory) */
  double cost;           /* Cost_path of the terminating node (0 on failure) */
  uint64_t traj_offset;  /* first state in traj[] (states of 5 doubles: x,y,heading,gridx,gridy) */
} hpp_plan;

hpp_status hpp_plan_batch(hpp_map *map, int variant, const double *starts_xyh, const double *goals_xyh,
                          size_t n, hpp_plan *plans, double *traj /* cap x 5 */, size_t traj_cap,
                          size_t *traj_used);

/* The same for queries on SEVERAL maps of one context (a fleet planning on different lots, cfg5): query q runs on
 * maps[map_of_query[q]].  One heuristic field per distinct (map, goal cell); maps that share geometry (W, H,
 * resolution, origin) and vehicle parameters are planned by ONE kernel launch with no host threads, maps that differ
 * are planned group after group.  Results are identical to one hpp_plan_batch call per map. */
hpp_status hpp_plan_batch_multi(hpp_map *const *maps, size_t n_maps, const uint32_t *map_of_query /* n */, int variant,
                                const double *starts_xyh, const double *goals_xyh, size_t n, hpp_plan *plans,
                                double *traj /* cap x 5 */, size_t traj_cap, size_t *traj_used);

/* ---------------------------------------------------------------- the OpenCV calls of map_inflate_node::InflateMap
 * MI/src/map_inflate_node.cpp:115-139 works on a gray cv::Mat: cv::threshold(0, 255, THRESH_BINARY | THRESH_OTSU) (:127),
 * cv::bitwise_not (:131), cv::dilate with getStructuringElement(MORPH_ELLIPSE, (2r+1)^2, (r,r)) (:134-135),
 * cv::bitwise_not (:138).  These two entry points replace the two image operations; the header
 * hybrid_path_planning_b200/host/dropin/mi/opencv2/opencv.hpp maps the node's cv:: calls onto them, so the node source
 * compiles and runs unmodified.  8-bit single-channel images, row-major, host buffers. */
hpp_status hpp_cv_threshold_otsu(hpp_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, uint8_t *dst /* 0 | 255 */,
                                 double *threshold_out /* may be NULL: the Otsu threshold, cv::threshold's return value */);
/* src is a binary image (0 / non-zero, what cv::threshold returns); set pixels of dst are 255.  r = element radius.
 * src_complemented != 0: the bytes of src are the COMPLEMENT of the image to dilate (the cv::bitwise_not in front of
 * the dilation folded into the call: 255 = unset pixel). */
hpp_status hpp_cv_dilate_ellipse(hpp_ctx *ctx, const uint8_t *src, uint32_t W, uint32_t H, int32_t r,
                                 int32_t src_complemented, uint8_t *dst);

/* ---------------------------------------------------------------- multi-GPU (one process per GPU, NCCL over NVLink)
 * only independent units are partitioned -- batches of queries, pose batches, and row bands of a large
 * map for the distance layer; a single heuristic field does not shard (replicas only).  NCCL is loaded at run time
 * (dlopen of libnccl.so.2; HPP_ERR_UNSUPPORTED when it is absent) and used only for the halo exchange of the row-band
 * distance transform and for the result gathers.  The reference has no multi-process path: these entry points are what
 * a multi-GPU deployment of HybridAstar::run (RS/src/HybridAstar.cpp:553-729), Grid_map::isSingleStateCollisionFree
 * (RS/src/Grid_map.cpp:222-284) and the Grid_map constructor's layers (RS/src/Grid_map.cpp:22-74) calls instead of
 * the single-GPU ones; every rank makes the same call with the same arguments. */
typedef struct hpp_comm hpp_comm;
#define HPP_COMM_ID_BYTES 128
/* rank 0 makes the id; the caller ships the bytes to the other ranks (MPI, a file, torch.distributed ...) */
hpp_status hpp_comm_unique_id(uint8_t id[HPP_COMM_ID_BYTES]);
hpp_status hpp_comm_create(hpp_ctx *ctx, const uint8_t id[HPP_COMM_ID_BYTES], int rank, int world, hpp_comm **out);
void hpp_comm_destroy(hpp_comm *comm);
int hpp_comm_rank(const hpp_comm *comm);
int hpp_comm_size(const hpp_comm *comm);
/* contiguous split of n units: rank r owns [lo, hi); sizes differ by at most one (the first n % world are larger) */
void hpp_shard_range(size_t n, int world, int rank, size_t *lo, size_t *hi);

/* cfg5: queries [lo, hi) of the list are planned on this rank (maps replicated: every rank holds all maps); the plan
 * headers of ALL n_total queries are returned to every rank by one ncclAllGather; trajectories stay on the rank that
 * produced them (traj_local; hpp_plan.traj_offset of a foreign query refers to its own rank's buffer).  Arguments as
 * hpp_plan_batch_multi, all arrays of n_total entries. */
hpp_status hpp_plan_batch_sharded(hpp_comm *comm, hpp_map *const *maps, size_t n_maps, const uint32_t *map_of_query,
                                  int variant, const double *starts_xyh, const double *goals_xyh, size_t n_total,
                                  hpp_plan *plans_all, double *traj_local, size_t traj_cap, size_t *traj_used,
                                  size_t *lo, size_t *hi);
/* cfg4: poses [lo, hi) checked on this rank (distance layer replicated), all n_total verdicts gathered to every rank */
hpp_status hpp_collide_circles_sharded(hpp_comm *comm, hpp_map *map, const double *xyh, size_t n_total,
                                       uint8_t *verdict_all);

/* large-map distance layer in row bands: this rank holds rows [lo, hi) (hpp_shard_range of total_rows) of the
 * occupancy in device memory.  Pass 1 (row distances) is local; the bands then exchange halo_rows rows (0 = 64) of
 * the pass-1 result with their neighbours (ncclSend/ncclRecv) and run pass 2; when a distance reaches past the halo
 * the ranks agree to double it, and finally to gather all rows.  distance_band_dev: (hi-lo) x W floats, metres,
 * row-major in MAP orientation (not flipped).  distance_full_dev (may be NULL): the whole layer W*total_rows in the
 * reference's grid_map layout (distance[k] = edt[n-1-k]*res, RS/src/Grid_map.cpp:67-74), gathered to every rank. */
typedef struct {
  uint32_t rounds;          /* pass-2 rounds (1 = the first halo was enough) */
  uint32_t halo_rows;       /* halo of the last round */
  uint32_t gathered_all;    /* 1 = fell back to gathering all rows */
  uint32_t pad_;
  uint64_t halo_bytes_sent; /* bytes this rank sent in halo exchanges (and the fall-back gather) */
} hpp_band_report;
hpp_status hpp_distance_layer_sharded(hpp_comm *comm, const int8_t *occ_band_dev, uint32_t W, uint32_t total_rows,
                                      float resolution, uint32_t halo_rows, float *distance_band_dev,
                                      float *distance_full_dev, hpp_band_report *report /* may be NULL */);

#ifdef __cplusplus
}
#endif
#endif /* HPP_B200_H */
