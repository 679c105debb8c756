// oracle/ref_driver_db.cpp -- TEST INFRASTRUCTURE, not product code.
// extern "C" driver around the reference's own Dubins planner translation units
// (dubins_hybrid_star/src/{HybridAstar,Grid_map,CarData,dubins_path}.cpp + sat_collision_checker.cc,
// compiled unmodified from /root/reference by oracle/build_ref.sh against oracle/ref_shim).
#include "HybridAstar.h"
#include "ref_driver_common.hpp"

#define API extern "C" __attribute__((visibility("default")))

static CarData make_car(const double *car)
{
    CarData c(car[0], car[1], car[2], car[3], car[4]);
    c.createVehicleGeometry();
    return c;
}

struct DbCtx
{
    Grid_map gm;
    CarData car;
    HybridAstar ha;
    DbCtx(const nav_msgs::msg::OccupancyGrid &g, const double *c, double sim_len, double step)
        : gm(g), car(make_car(c)), ha(gm, car, sim_len, step) {}
};

API void *ref_db_create(const int8_t *occ, int W, int H, float res, double ox, double oy, const double *car, double sim_len, double step)
{
    MuteCout mute;
    auto g = make_grid(occ, W, H, res, ox, oy);
    return new DbCtx(g, car, sim_len, step);
}
API void ref_db_destroy(void *ctx) { delete (DbCtx *)ctx; }

API void ref_db_heuristic(void *ctx, int goal_gx, int goal_gy, double *cost)
{
    MuteCout mute;
    DbCtx *c = (DbCtx *)ctx;
    auto n = make_node(0, 0, 0, goal_gx, goal_gy, 0, 0, 1);
    auto out = c->ha.holonomicCostsWithObstacles_planning(n);
    for (size_t i = 0; i < out.size(); ++i)
        cost[i] = out[i]->cost;
}

// planner predicate of DB/src/HybridAstar.cpp:52
API int ref_db_collide(void *ctx, double x, double y, double h)
{
    DbCtx *c = (DbCtx *)ctx;
    State s(x, y, h);
    auto cell = c->gm.toCellID(s);
    s.gridx = std::get<0>(cell);
    s.gridy = std::get<1>(cell);
    auto poly = c->car.getVehicleGeometry_state(s);
    return (c->gm.checkCollision(s, poly) || !c->gm.isPointInBounds(s.gridx, s.gridy)) ? 1 : 0;
}

// the same predicate over n poses: the loop the CPU baseline of the SAT footprint check times
API void ref_db_collide_batch(void *ctx, const double *xyh, size_t n, uint8_t *verdict)
{
    for (size_t i = 0; i < n; ++i)
        verdict[i] = (uint8_t)ref_db_collide(ctx, xyh[3 * i], xyh[3 * i + 1], xyh[3 * i + 2]);
}

API int ref_db_expand(void *ctx, double x, double y, double h, double cost, double steer, int dir,
                      double *traj, double *ccost, double *csteer, int *cdir, int sim)
{
    MuteCout mute;
    DbCtx *c = (DbCtx *)ctx;
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

// mode: 3 chars; returns number of samples
API int ref_db_dubins_path(double sx, double sy, double syaw, double gx, double gy, double gyaw, double curv,
                           double *length, char *mode, double *xs, double *ys, double *yaws, int max_samples)
{
    dubins_path dp;
    std::streambuf *olderr = std::cerr.rdbuf(nullptr);
    Path p = dp.calc_dubins_path(sx, sy, syaw, gx, gy, gyaw, curv);
    std::cerr.rdbuf(olderr);
    *length = p.length;
    for (size_t i = 0; i < 3; ++i)
        mode[i] = i < p.mode.size() ? p.mode[i][0] : 0;
    int n = (int)p.x.size();
    for (int i = 0; i < n && i < max_samples; ++i)
    {
        xs[i] = p.x[i]; ys[i] = p.y[i]; yaws[i] = p.yaw[i];
    }
    return n;
}

API int ref_db_shot(void *ctx, double x, double y, double h, double cost, double gx, double gy, double gh,
                    double *out_cost, double *traj, int max_states)
{
    MuteCout mute;
    std::streambuf *olderr = std::cerr.rdbuf(nullptr);
    DbCtx *c = (DbCtx *)ctx;
    auto cc = c->gm.toCellID(State(x, y, h));
    auto gc = c->gm.toCellID(State(gx, gy, gh));
    auto cur = make_node(x, y, h, std::get<0>(cc), std::get<1>(cc), cost, 0, 1);
    auto goal = make_node(gx, gy, gh, std::get<0>(gc), std::get<1>(gc), 0, 0, 1);
    auto d = c->ha.dubins_Path(cur, goal);
    std::cerr.rdbuf(olderr);
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

API int ref_db_run(void *ctx, double sx, double sy, double sh, double gx, double gy, double gh, double *traj, int max_states)
{
    MuteCout mute;
    std::streambuf *olderr = std::cerr.rdbuf(nullptr);
    DbCtx *c = (DbCtx *)ctx;
    std::vector<State> t = c->ha.run(State(sx, sy, sh), State(gx, gy, gh));
    std::cerr.rdbuf(olderr);
    int n = (int)t.size();
    for (int i = 0; i < n && i < max_states; ++i)
    {
        traj[i * 5 + 0] = t[i].x; traj[i * 5 + 1] = t[i].y; traj[i * 5 + 2] = t[i].heading;
        traj[i * 5 + 3] = t[i].gridx; traj[i * 5 + 4] = t[i].gridy;
    }
    return n;
}
