// Minimal stand-in for <rclcpp/rclcpp.hpp>, written for this repo (NOT ROS code).
//
// ROS2 (rclcpp) is not installed in the build image, so this header provides only the handful of
// names the reference's trajectory classes touch (Trajectory.hpp:15,35; Circle.cpp:35,92):
//   rclcpp::Clock, Clock::SharedPtr, Clock::now(), rclcpp::Time, (Time - Time).seconds().
// In a real ROS2 workspace this directory is simply left off the include path.
#pragma once

#include <chrono>
#include <memory>

#include "rclcpp/logger.hpp"
#include "rclcpp/logging.hpp"

namespace rclcpp {

class Duration {
public:
    explicit Duration(double s) : s_(s) {}
    double seconds() const { return s_; }
private:
    double s_;
};

class Time {
public:
    Time() : s_(0.0) {}
    explicit Time(double s) : s_(s) {}
    double seconds() const { return s_; }
    Duration operator-(const Time& o) const { return Duration(s_ - o.s_); }
private:
    double s_;
};

class Clock {
public:
    using SharedPtr = std::shared_ptr<Clock>;
    Time now() const {
        using namespace std::chrono;
        return Time(duration<double>(steady_clock::now().time_since_epoch()).count());
    }
};

}  // namespace rclcpp
