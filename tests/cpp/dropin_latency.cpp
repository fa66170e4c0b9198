// dropin_latency.cpp — BASELINE.json configs[0] / BASELINE.md §4: what ONE generateTraj of the default.yaml circle
// costs through the C++ drop-in class (count -> plan -> evaluate on the GPU -> D2H -> repack into
// std::vector<Goal> + index_msgs), i.e. the call TrajectoryGenerator.cpp:71 makes at start-up.  Links only the
// drop-in classes and libtgx.so — no reference source; the reference's own CPU time for the same call is measured
// beside it by bench.py through oracle/_ref (the reference's self-timing hook is Circle.cpp:92).
// Prints one JSON line: median / min / max in ms over `reps` calls after 3 warm-ups.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "tgx_trajectories.hpp"

using snapstack_msgs2::msg::Goal;

int main(int argc, char** argv) {
    const int reps = argc > 1 && std::atoi(argv[1]) > 0 ? std::atoi(argv[1]) : 20;
    // config/default.yaml:5-6,38-44 with traj_type Circle: alt 1.8, r 3.4, c (0,0), v_goals [1,2,2], t_traj 80, accel 0.4
    tgx_dropin::Circle circle(1.8, 3.4, 0.0, 0.0, {1.0, 2.0, 2.0}, 80.0, 0.4, 0.01);
    auto clock = std::make_shared<rclcpp::Clock>();
    std::vector<double> ms;
    size_t n = 0, msgs = 0;
    for (int r = 0; r < reps + 3; ++r) {
        std::vector<Goal> goals;
        std::unordered_map<int, std::string> index_msgs;
        const auto t0 = std::chrono::steady_clock::now();
        circle.generateTraj(goals, index_msgs, clock);
        const auto t1 = std::chrono::steady_clock::now();
        if (r >= 3) ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
        n = goals.size();
        msgs = index_msgs.size();
    }
    std::sort(ms.begin(), ms.end());
    if (n != 25001 || msgs != 7) {
        std::fprintf(stderr, "unexpected result: %zu goals, %zu index_msgs\n", n, msgs);
        return 1;
    }
    std::printf("{\"call\": \"tgx drop-in Circle::generateTraj (default.yaml circle) into std::vector<Goal>\", "
                "\"samples\": %zu, \"reps\": %d, \"median_ms\": %.4f, \"min_ms\": %.4f, \"max_ms\": %.4f}\n",
                n, reps, ms[ms.size() / 2], ms.front(), ms.back());
    return 0;
}
