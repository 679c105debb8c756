// oracle/ref_driver_rs.cpp -- TEST INFRASTRUCTURE, not product code.
// extern "C" driver around the reference's own RS planner translation units
// (reeds_shepp_hybrid_astar/src/{HybridAstar,CarData,reeds_shepp_path}.cpp, compiled unmodified from
// /root/reference by oracle/build_ref.sh against oracle/ref_shim and oracle/ref_standin/Grid_map.h).
#include "HybridAstar.h"
#include "ref_driver_common.hpp"

#define API extern "C" __attribute__((visibility("default")))

static CarData make_car(const double *car)
{
    CarData c(car[0], car[1], car[2], car[3], car[4]);
    c.createVehicleGeometry();
    return c;
}

// hazard H3: the heuristic caches its result in function statics keyed by the goal cell only.
// A call on a throw-away 2x2 map with a different goal cell invalidates that cache.
static void poison_heuristic_cache(int goal_gx, int goal_gy)
{
    int8_t occ[4] = {0, 0, 0, 0};
    auto g = make_grid(occ, 2, 2, 1.0f, 0.0, 0.0);
    Grid_map gm(g);
    double car[5] = {0.6, 2.75, 1.1, 1.2, 1.9};
    CarData c = make_car(car);
    gm.setcarData(c);
    HybridAstar ha(gm, c, 4.0, 1.0);
    int px = (goal_gx == 0 && goal_gy == 0) ? 1 : 0;
    auto n = make_node(0, 0, 0, px, px, 0, 0, 1);
    (void)ha.holonomicCostsWithObstacles_planning(n);
}

API void ref_rs_heuristic(const int8_t *occ, int W, int H, float res, double ox, double oy, int goal_gx, int goal_gy, double *cost)
{
    MuteCout mute;
    poison_heuristic_cache(goal_gx, goal_gy);
    auto g = make_grid(occ, W, H, res, ox, oy);
    // a minimal Grid_map is enough for the heuristic, but the stand-in ctor also builds the layers;
    // that is fine for the map sizes the tests use.
    Grid_map gm(g);
    double car[5] = {0.6, 2.75, 1.1, 1.2, 1.9};
    CarData c = make_car(car);
    gm.setcarData(c);
    HybridAstar ha(gm, c, 4.0, 1.0);
    auto n = make_node(0, 0, 0, goal_gx, goal_gy, 0, 0, 1);
    std::vector<double> out = ha.holonomicCostsWithObstacles_planning(n);
    std::memcpy(cost, out.data(), sizeof(double) * out.size());
}

// returns number of paths; per path: nseg, lengths[5], types[5], L, n samples; samples appended to xs/ys/yaws
API int ref_rs_all_paths(double sx, double sy, double syaw, double gx, double gy, double gyaw, double maxc, double step,
                         int *nseg, double *lengths, char *types, double *L, int *nsamp,
                         double *xs, double *ys, double *yaws, int *dirs, int max_samples)
{
    reeds_shepp_path rs;
    std::vector<PATH> paths = rs.calc_all_paths(sx, sy, syaw, gx, gy, gyaw, maxc, step);
    int off = 0;
    for (size_t k = 0; k < paths.size(); ++k)
    {
        const PATH &p = paths[k];
        nseg[k] = (int)p.lengths.size();
        for (size_t i = 0; i < p.lengths.size(); ++i)
        {
            lengths[k * 5 + i] = p.lengths[i];
            types[k * 5 + i] = p.ctypes[i][0];
        }
        L[k] = p.L;
        nsamp[k] = (int)p.x.size();
        for (size_t i = 0; i < p.x.size(); ++i)
        {
            if (off >= max_samples)
                return -1;
            xs[off] = p.x[i];
            ys[off] = p.y[i];
            yaws[off] = p.yaw[i];
            dirs[off] = p.directions[i];
            ++off;
        }
    }
    return (int)paths.size();
}

struct RsCtx
{
    Grid_map gm;
    CarData car;
    HybridAstar ha;
    RsCtx(const nav_msgs::msg::OccupancyGrid &g, const double *c, double sim_len, double step)
        : gm(g), car(make_car(c)), ha((gm.setcarData(car), gm), car, sim_len, step) {}
};

API void *ref_rs_create(const int8_t *occ, int W, int H, float res, double ox, double oy, const double *car, double sim_len, double step)
{
    MuteCout mute;
    auto g = make_grid(occ, W, H, res, ox, oy);
    return new RsCtx(g, car, sim_len, step);
}
API void ref_rs_destroy(void *ctx) { delete (RsCtx *)ctx; }

// Same, with the obstacle / distance layers supplied by the caller (grid_map layout, W*H floats each) instead of being
// computed by the stand-in constructor: the reference arm of bench.py builds them once per map message with
// cv2.distanceTransform -- the call RS/src/Grid_map.cpp:64 makes -- and hands them to every planner instance.
API void *ref_rs_create_layers(const int8_t *occ, int W, int H, float res, double ox, double oy, const double *car,
                               double sim_len, double step, const float *obstacle, const float *distance)
{
    MuteCout mute;
    auto g = make_grid(occ, W, H, res, ox, oy);
    ref_standin_ext_obstacle = obstacle;
    ref_standin_ext_distance = distance;
    RsCtx *c = new RsCtx(g, car, sim_len, step);
    ref_standin_ext_obstacle = ref_standin_ext_distance = nullptr;
    return c;
}

// holonomicCostsWithObstacles_planning of an existing planner instance (no second Grid_map): one field per call
API void ref_rs_heuristic_ctx(void *ctx, int goal_gx, int goal_gy, double *cost)
{
    MuteCout mute;
    RsCtx *c = (RsCtx *)ctx;
    poison_heuristic_cache(goal_gx, goal_gy);
    auto n = make_node(0, 0, 0, goal_gx, goal_gy, 0, 0, 1);
    std::vector<double> out = c->ha.holonomicCostsWithObstacles_planning(n);
    if (cost)
        std::memcpy(cost, out.data(), sizeof(double) * out.size());
}

// isSingleStateCollisionFree over n poses (verdict 1 = collision / outside, the C ABI's polarity)
API void ref_rs_collide_batch(void *ctx, const double *xyh, size_t n, uint8_t *verdict)
{
    RsCtx *c = (RsCtx *)ctx;
    for (size_t i = 0; i < n; ++i)
        verdict[i] = c->gm.isSingleStateCollisionFree(State(xyh[3 * i], xyh[3 * i + 1], xyh[3 * i + 2])) ? 1 : 0;
}

API void ref_rs_distance_layer(void *ctx, float *out)
{
    RsCtx *c = (RsCtx *)ctx;
    std::memcpy(out, c->gm.distanceLayer().data(), sizeof(float) * c->gm.distanceLayer().size());
}

API int ref_rs_collide(void *ctx, double x, double y, double h)
{
    RsCtx *c = (RsCtx *)ctx;
    return c->gm.isSingleStateCollisionFree(State(x, y, h)) ? 1 : 0;
}

// children: per child 4 states x (x,y,h,gridx,gridy) + cost + steer + dir; returns count of valid children
API int ref_rs_expand(void *ctx, double x, double y, double h, double cost, double steer, int dir,
                      double *traj /* n x sim x 5 */, double *ccost, double *csteer, int *cdir, int sim)
{
    MuteCout mute;
    RsCtx *c = (RsCtx *)ctx;
    auto cell = c->gm.toCellID(State(x, y, h));
    auto n = make_node(x, y, h, std::get<0>(cell), std::get<1>(cell), cost, steer, dir);
    auto kids = c->ha.GetnextNeighbours(n);
    for (size_t k = 0; k < kids.size(); ++k)
    {
        for (int i = 0; i < sim && i < (int)kids[k]->Trajectory.size(); ++i)
        {
            const State &s = kids[k]->Trajectory[i];
            double *t = traj + (k * sim + i) * 5;
            t[0] = s.x; t[1] = s.y; t[2] = s.heading; t[3] = s.gridx; t[4] = s.gridy;
        }
        ccost[k] = kids[k]->Cost_path;
        csteer[k] = kids[k]->steeringAngle;
        cdir[k] = kids[k]->direction;
    }
    return (int)kids.size();
}

// returns number of trajectory states (0 = no shot); traj: n x 5
API int ref_rs_shot(void *ctx, double x, double y, double h, double cost, double gx, double gy, double gh,
                    double *out_cost, double *traj, int max_states)
{
    MuteCout mute;
    RsCtx *c = (RsCtx *)ctx;
    auto cc = c->gm.toCellID(State(x, y, h));
    auto gc = c->gm.toCellID(State(gx, gy, gh));
    auto cur = make_node(x, y, h, std::get<0>(cc), std::get<1>(cc), cost, 0, 1);
    auto goal = make_node(gx, gy, gh, std::get<0>(gc), std::get<1>(gc), 0, 0, 1);
    auto d = c->ha.reeds_shepp_Path(cur, goal);
    if (!d)
        return 0;
    *out_cost = d->Cost_path;
    int n = (int)d->Trajectory.size();
    for (int i = 0; i < n && i < max_states; ++i)
    {
        const State &s = d->Trajectory[i];
        traj[i * 5 + 0] = s.x; traj[i * 5 + 1] = s.y; traj[i * 5 + 2] = s.heading;
        traj[i * 5 + 3] = s.gridx; traj[i * 5 + 4] = s.gridy;
    }
    return n;
}

// full query; returns number of states; traj n x 5
API int ref_rs_run(void *ctx, double sx, double sy, double sh, double gx, double gy, double gh, double *traj, int max_states)
{
    MuteCout mute;
    RsCtx *c = (RsCtx *)ctx;
    auto gc = c->gm.toCellID(State(gx, gy, gh));
    poison_heuristic_cache(std::get<0>(gc), std::get<1>(gc));
    std::vector<State> t = c->ha.run(State(sx, sy, sh), State(gx, gy, gh));
    int n = (int)t.size();
    for (int i = 0; i < n && i < max_states; ++i)
    {
        traj[i * 5 + 0] = t[i].x; traj[i * 5 + 1] = t[i].y; traj[i * 5 + 2] = t[i].heading;
        traj[i * 5 + 3] = t[i].gridx; traj[i * 5 + 4] = t[i].gridy;
    }
    return n;
}
