// Minimal stand-in for the generated header of snapstack_msgs2/msg/Goal.msg, written for this repo.
//
// snapstack_msgs2 is an external, un-vendored ROS2 interface package (reference package.xml:17,
// CMakeLists.txt:32).  Only plain data lives there; the fields below are exactly the ones the
// reference reads or writes (Circle.cpp:105-127, TrajectoryGenerator.cpp:621-634).
#pragma once

#include <cstdint>
#include <string>

#include "builtin_interfaces/msg/time.hpp"

namespace snapstack_msgs2 {
namespace msg {

using GoalStamp = builtin_interfaces::msg::Time;   // header.stamp = this->now() (TrajectoryGenerator.cpp:606)

struct GoalHeader {
    GoalStamp stamp;
    std::string frame_id;
};

struct GoalVector3 {
    double x = 0.0;
    double y = 0.0;
    double z = 0.0;
};

struct Goal {
    GoalHeader header;
    GoalVector3 p;   // position
    GoalVector3 v;   // velocity
    GoalVector3 a;   // acceleration
    GoalVector3 j;   // jerk
    double psi = 0.0;   // yaw
    double dpsi = 0.0;  // yaw rate
    bool power = false;
    uint8_t mode_xy = 0;
    uint8_t mode_z = 0;

    static constexpr uint8_t MODE_POSITION_CONTROL = 0;
    static constexpr uint8_t MODE_VELOCITY_CONTROL = 1;
    static constexpr uint8_t MODE_ACCELERATION_CONTROL = 2;
};

}  // namespace msg
}  // namespace snapstack_msgs2
