// Minimal stand-in for builtin_interfaces/msg/Time, written for this repo (NOT ROS code).
#pragma once

#include <cstdint>

namespace builtin_interfaces {
namespace msg {

struct Time {
    int32_t sec = 0;
    uint32_t nanosec = 0;
};

}  // namespace msg
}  // namespace builtin_interfaces
