// node_shim.cpp — in-process harness around the UNMODIFIED reference node (test infrastructure, NOT product code).
//
// Compiles together with the reference's own src/TrajectoryGenerator.cpp (from where it lies under /root/reference)
// behind the fake rclcpp::Node of compat/ros2_stubs, in two flavours (oracle/Makefile, tests/cpp/Makefile):
//   oracle/_ref/libnoderef.so      node + the reference's own eleven trajectory classes            -> the oracle
//   tests/cpp/bin/libnodegpu.so    the SAME node source + this repo's GPU-backed drop-in classes    -> the product,
//                                  found through compat/dropin_include, which shadows only the eleven class headers
// Both export node_run(): construct trajectory_generator::TrajectoryGenerator (TrajectoryGenerator.cpp:44-93) with a
// parameter table (the YAML of a real launch), then tick its 100 Hz timer (pubCB, :525-611), delivering
// /globalflightmode events (modeCB, :427-523) at scripted ticks and a `state` message before every tick that reports
// perfect tracking (pose = the goal published on the previous tick).  Every published Goal is returned.
// No reference source is copied: this file only drives the node through the interfaces it subscribes to.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "trajectory_generator_ros2/TrajectoryGenerator.hpp"

using snapstack_msgs2::msg::Goal;
using snapstack_msgs2::msg::QuadFlightMode;
using snapstack_msgs2::msg::State;

namespace {

// "key=value" lines; value is a number, a comma-separated list of numbers (v_goals) or a word (traj_type).
void load_config(const char* text) {
    auto& env = tgx_stub::node_environment();
    env.overrides.clear();
    env.shutdown_requested = false;
    std::istringstream in(text ? text : "");
    std::string line;
    while (std::getline(in, line)) {
        const size_t eq = line.find('=');
        if (eq == std::string::npos) continue;
        const std::string key = line.substr(0, eq), val = line.substr(eq + 1);
        if (key == "namespace") {
            env.ns = val;
            continue;
        }
        rclcpp::ParameterValue p;
        if (key == "traj_type") {
            p.kind = rclcpp::ParameterValue::STRING;
            p.s = val;
        } else if (key == "v_goals") {
            p.kind = rclcpp::ParameterValue::DOUBLE_ARRAY;
            std::istringstream vs(val);
            std::string tok;
            // the node declares v_goals as std::vector<float> and reads doubles (TrajectoryGenerator.cpp:136, :183):
            // the YAML value reaches the class as the double nearest to the decimal text
            while (std::getline(vs, tok, ',')) if (!tok.empty()) p.v.push_back(std::strtod(tok.c_str(), nullptr));
        } else {
            p.kind = rclcpp::ParameterValue::DOUBLE;
            p.d = std::strtod(val.c_str(), nullptr);
        }
        env.overrides[key] = p;
    }
}

geometry_msgs::msg::Quaternion yaw_quat(double yaw) {
    geometry_msgs::msg::Quaternion q;
    q.x = 0.0;
    q.y = 0.0;
    q.z = std::sin(0.5 * yaw);
    q.w = std::cos(0.5 * yaw);
    return q;
}

}  // namespace

extern "C" {

// One row of `out` per published Goal: {tick, p.xyz, v.xyz, a.xyz, j.xyz, psi, dpsi, power, mode_xy, mode_z} = 18 doubles.
// Events: at tick ev_tick[i] the mode ev_mode[i] (QuadFlightMode: GO 4, LAND 2, KILL 6) is delivered before the timer
// fires.  start = {x, y, z, yaw} of the vehicle on the ground.  Returns the number of rows (<= out_cap are written), or
// -1 if the node refused to start (bad parameters: readParameters() returned false, TrajectoryGenerator.cpp:54-57).
int64_t node_run(const char* config, const int32_t* ev_tick, const uint8_t* ev_mode, int32_t n_events,
                 int64_t n_ticks, const double* start, double* out, int64_t out_cap) {
    load_config(config);
    tgx_stub::LogState saved = tgx_stub::log_state();
    tgx_stub::log_state() = tgx_stub::LogState();
    int64_t rows = -1;
    {
        State st;
        st.pos.x = start[0];
        st.pos.y = start[1];
        st.pos.z = start[2];
        st.quat = yaw_quat(start[3]);
        std::unique_ptr<trajectory_generator::TrajectoryGenerator> node;
        try {
            tgx_stub::log_state().throw_on_error = false;
            node = std::make_unique<trajectory_generator::TrajectoryGenerator>();
        } catch (...) {
            node.reset();
        }
        if (node && !tgx_stub::node_environment().shutdown_requested) {
            auto* pub = node->tgx_stub_publisher<Goal>();
            rows = 0;
            size_t seen = 0;
            // the constructor copied pose_ (all zeros) into goal_ before any state arrived (:84-89); like a real launch,
            // the first state message arrives before the operator presses anything
            for (int64_t t = 0; t < n_ticks; ++t) {
                node->tgx_stub_deliver(st);
                for (int32_t i = 0; i < n_events; ++i)
                    if (ev_tick[i] == t) {
                        QuadFlightMode m;
                        m.mode = ev_mode[i];
                        node->tgx_stub_deliver(m);
                    }
                node->tgx_stub_fire_timer(0);
                for (; seen < pub->sent.size(); ++seen) {
                    const Goal& g = pub->sent[seen];
                    if (rows < out_cap) {
                        double* r = out + rows * 18;
                        r[0] = (double)t;
                        r[1] = g.p.x; r[2] = g.p.y; r[3] = g.p.z;
                        r[4] = g.v.x; r[5] = g.v.y; r[6] = g.v.z;
                        r[7] = g.a.x; r[8] = g.a.y; r[9] = g.a.z;
                        r[10] = g.j.x; r[11] = g.j.y; r[12] = g.j.z;
                        r[13] = g.psi; r[14] = g.dpsi;
                        r[15] = g.power ? 1.0 : 0.0;
                        r[16] = (double)g.mode_xy;
                        r[17] = (double)g.mode_z;
                    }
                    ++rows;
                }
                // perfect tracking: a powered vehicle is where the last published goal told it to be (on the ground,
                // motors off, it stays where it is)
                if (!pub->sent.empty() && pub->sent.back().power) {
                    const Goal& g = pub->sent.back();
                    st.pos.x = g.p.x;
                    st.pos.y = g.p.y;
                    st.pos.z = g.p.z;
                    st.quat = yaw_quat(g.psi);
                }
            }
        }
    }
    tgx_stub::log_state() = saved;
    return rows;
}

}  // extern "C"
