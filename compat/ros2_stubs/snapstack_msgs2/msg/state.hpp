// Minimal stand-in for snapstack_msgs2/msg/State, written for this repo.  Only the fields the reference node reads
// (TrajectoryGenerator.cpp:613-619: pos, quat); the trajectory samplers include the header and use nothing from it.
#pragma once
#include "geometry_msgs/msg/quaternion.hpp"
#include "geometry_msgs/msg/vector3.hpp"
namespace snapstack_msgs2 {
namespace msg {
struct State {
    geometry_msgs::msg::Vector3 pos;
    geometry_msgs::msg::Vector3 vel;
    geometry_msgs::msg::Quaternion quat;
};
}  // namespace msg
}  // namespace snapstack_msgs2
