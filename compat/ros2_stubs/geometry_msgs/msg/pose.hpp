// Minimal stand-in for geometry_msgs/msg/Pose, written for this repo (NOT ROS code).
#pragma once
#include "geometry_msgs/msg/quaternion.hpp"
namespace geometry_msgs {
namespace msg {
struct Point { double x = 0.0, y = 0.0, z = 0.0; };
struct Pose {
    Point position;
    Quaternion orientation;
};
}  // namespace msg
}  // namespace geometry_msgs
