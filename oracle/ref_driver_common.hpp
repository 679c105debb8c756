// oracle/ref_driver_common.hpp -- TEST INFRASTRUCTURE, not product code.
// Helpers shared by the two drivers that expose the compiled reference planners through extern "C".
#pragma once
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <vector>

static inline nav_msgs::msg::OccupancyGrid make_grid(const int8_t *occ, int W, int H, float res, double ox, double oy)
{
    nav_msgs::msg::OccupancyGrid g;
    g.header.frame_id = "map";
    g.info.resolution = res;
    g.info.width = (uint32_t)W;
    g.info.height = (uint32_t)H;
    g.info.origin.position.x = ox;
    g.info.origin.position.y = oy;
    g.data.assign(occ, occ + (size_t)W * H);
    return g;
}

struct MuteCout
{
    std::streambuf *old;
    MuteCout() : old(std::cout.rdbuf(nullptr)) {}
    ~MuteCout() { std::cout.rdbuf(old); }
};

static inline std::shared_ptr<planner::Node> make_node(double x, double y, double h, int gx, int gy, double cost, double steer, int dir)
{
    State s(x, y, h);
    s.gridx = gx;
    s.gridy = gy;
    std::vector<State> traj = {s};
    return std::make_shared<planner::Node>(s, traj, cost, steer, dir, std::weak_ptr<planner::Node>());
}
