// oracle/ref_standin/Grid_map.h -- TEST INFRASTRUCTURE, not product code.
//
// Stand-in for reeds_shepp_hybrid_astar/include/Grid_map.h, force-included when the reference's
// RS translation units are compiled into oracle/_ref (same include guard, so the real header --
// which needs grid_map_core, OpenCV and Eigen, none of which exist in this image -- is skipped).
// It offers the public methods of RS/include/Grid_map.h:84-110 that RS/src/HybridAstar.cpp calls.
// The cell predicates restate RS/src/Grid_map.cpp:186-213 inline; the layers (a3) and the
// circle-footprint check (a4) delegate to the C oracle (orc_edt_layers / orc_collide_circles), because
// RS/src/Grid_map.cpp itself cannot be compiled here.  Consequently oracle/_ref pins a1/a5/a6/a7 and
// run() of the RS planner, but NOT a3/a4 (those are pinned against cv2 / left "parity unpinned").
#ifndef GRID_MAP_H
#define GRID_MAP_H

#include <nav_msgs/msg/occupancy_grid.hpp>
#include <geometry_msgs/msg/polygon.hpp>
#include <geometry_msgs/msg/point32.hpp>
#include <iostream>
#include <tuple>
#include <vector>
#include "CarData.h"
#include "State.h"
#include "../orc.h"

// Layers computed elsewhere (bench.py's reference arm builds them with cv2.distanceTransform, the library call the real
// RS/src/Grid_map.cpp:64 makes): when set, the constructor copies them instead of running the oracle's EDT.
inline const float *ref_standin_ext_obstacle = nullptr;
inline const float *ref_standin_ext_distance = nullptr;

class Grid_map
{
private:
    nav_msgs::msg::OccupancyGrid map_data_;
    double resolution;
    double originX;
    double originY;
    unsigned int width;
    unsigned int height;
    CarData car_data_;
    std::vector<float> obstacle_layer_;
    std::vector<float> distance_layer_;

public:
    Grid_map(const nav_msgs::msg::OccupancyGrid &map_data)
    {
        map_data_ = map_data;
        resolution = map_data.info.resolution;
        originX = map_data.info.origin.position.x;
        originY = map_data.info.origin.position.y;
        width = map_data.info.width;
        height = map_data.info.height;
        obstacle_layer_.resize((size_t)width * height);
        distance_layer_.resize((size_t)width * height);
        if (ref_standin_ext_obstacle && ref_standin_ext_distance)
        {
            std::copy(ref_standin_ext_obstacle, ref_standin_ext_obstacle + obstacle_layer_.size(), obstacle_layer_.begin());
            std::copy(ref_standin_ext_distance, ref_standin_ext_distance + distance_layer_.size(), distance_layer_.begin());
        }
        else
            orc_edt_layers(map_data_.data.data(), (int)width, (int)height, map_data.info.resolution,
                           obstacle_layer_.data(), distance_layer_.data());
    }
    ~Grid_map() {}

    double getResolution() { return resolution; }
    double getOriginX() { return originX; }
    double getOriginY() { return originY; }
    unsigned int getWidth() { return width; }
    unsigned int getHeight() { return height; }

    void setcarData(CarData car_data) { car_data_ = car_data; }

    bool isSingleStateCollisionFree(const State &current)
    {
        orc_map m{map_data_.data.data(), (int)width, (int)height, resolution, originX, originY, distance_layer_.data()};
        orc_car c{car_data_.maxSteerAngle, car_data_.wheelBase, car_data_.axleToFront, car_data_.axleToBack, car_data_.width};
        return orc_collide_circles(&m, &c, current.x, current.y, current.heading) != 0;
    }

    std::tuple<int, int> toCellID(State state_)
    {
        int cell_x = static_cast<int>((state_.x - originX) / resolution);
        int cell_y = static_cast<int>((state_.y - originY) / resolution);
        return std::make_tuple(cell_x, cell_y);
    }
    int toCellIndex(int x, int y) { return y * width + x; }
    bool isPointInBounds(int x, int y) { return x >= 0 && x < (int)width && y >= 0 && y < (int)height; }
    bool isInCollision(int x, int y)
    {
        if (!isPointInBounds(x, y))
            return true;
        return map_data_.data[toCellIndex(x, y)] > 0;
    }
    const std::vector<float> &distanceLayer() const { return distance_layer_; }
};

#endif // GRID_MAP_H
