// Minimal stand-in for snapstack_msgs2/msg/QuadFlightMode, written for this repo.  The reference node compares
// msg.mode with msg.GO / msg.LAND / msg.KILL; its own comment gives the values (TrajectoryGenerator.cpp:437-440:
// "START -> GO (4), END -> LAND (2), ESTOP -> KILL (6)").
#pragma once
#include <cstdint>
namespace snapstack_msgs2 {
namespace msg {
struct QuadFlightMode {
    uint8_t mode = 0;
    static constexpr uint8_t NOT_FLYING = 0;
    static constexpr uint8_t LAND = 2;
    static constexpr uint8_t GO = 4;
    static constexpr uint8_t KILL = 6;
};
}  // namespace msg
}  // namespace snapstack_msgs2
