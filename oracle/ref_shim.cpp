// ref_shim.cpp — C-ABI driver around the UNMODIFIED reference trajectory classes (test infrastructure).
//
// Built ONLY by oracle/Makefile into oracle/_ref/libtrajref.so together with the reference's own
// src/trajectories/{Circle,Line,Figure8,Boomerang,Square,Rectangle,Reciprocating,Bounce,M,I,T}.cpp, compiled from where they lie under /root/reference behind the
// stub headers in compat/ros2_stubs.  No reference source is copied into this repository; this file is the
// repo's own glue: it constructs trajectory_generator::{Circle,Line,Figure8} with the constructor arguments
// held in a tgx_params record, calls generateTraj / generateStopTraj / trajectoryInsideBounds and repacks the
// std::vector<snapstack_msgs2::msg::Goal> into the same SoA convention the oracle uses, so that
// tests can compare reference == oracle bit for bit, and bench.py can time the reference's CPU path.
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "trajectory_generator_ros2/trajectories/Boomerang.hpp"
#include "trajectory_generator_ros2/trajectories/Bounce.hpp"
#include "trajectory_generator_ros2/trajectories/Circle.hpp"
#include "trajectory_generator_ros2/trajectories/Figure8.hpp"
#include "trajectory_generator_ros2/trajectories/I.hpp"
#include "trajectory_generator_ros2/trajectories/Line.hpp"
#include "trajectory_generator_ros2/trajectories/M.hpp"
#include "trajectory_generator_ros2/trajectories/Reciprocating.hpp"
#include "trajectory_generator_ros2/trajectories/Rectangle.hpp"
#include "trajectory_generator_ros2/trajectories/Square.hpp"
#include "trajectory_generator_ros2/trajectories/T.hpp"

#include "../include/tgx.h"

using snapstack_msgs2::msg::Goal;
namespace tg = trajectory_generator;

namespace {

bool finite_pos(double x) { return std::isfinite(x) && x > 0.0; }

// Same acceptance rule as the oracle (traj_oracle.c: params_ok): parameters the node would reject, or for
// which the reference loops cannot terminate, are never handed to the reference classes.
bool params_ok(const tgx_params& p) {
    if (!finite_pos(p.dt) || !std::isfinite(p.alt)) return false;
    if (p.type == TGX_CIRCLE || p.type == TGX_FIGURE8) {
        const tgx_orbit_params& o = p.u.orbit;
        // goal speeds beyond the eighth live in the continuation records that follow p (tgx.h: TGX_VGOALS_MORE); the
        // caller guarantees they are there
        if (p.n_vgoals < 0 || p.n_vgoals > TGX_MAX_VGOALS_TOTAL) return false;
        if (!std::isfinite(o.r) || o.r == 0.0 || !finite_pos(o.accel)) return false;
        if (!std::isfinite(o.cx) || !std::isfinite(o.cy) || !std::isfinite(o.t_traj)) return false;
        for (int q = 1; q < TGX_ORBIT_RECORDS(p.n_vgoals); ++q)
            if ((&p)[q].type != TGX_VGOALS_MORE) return false;
        for (int i = 0; i < p.n_vgoals; ++i)
            if (!finite_pos(i < TGX_MAX_VGOALS ? o.v_goals[i] : (&p)[i >> 3].u.orbit.v_goals[i & 7])) return false;
        return true;
    }
    if (p.type == TGX_LINE || p.type == TGX_BOOMERANG) {
        const tgx_line_params& l = p.u.line;
        for (int i = 0; i < 3; ++i)
            if (!std::isfinite(l.A[i]) || !std::isfinite(l.B[i])) return false;
        return finite_pos(l.v_goal) && finite_pos(l.a1) && finite_pos(l.a3);
    }
    if (TGX_IS_POLYLINE(p.type)) {
        const tgx_polyline_params& q = p.u.poly;
        if (!finite_pos(q.v_goal) || !std::isfinite(q.t_traj) || !std::isfinite(q.orientation)) return false;
        for (int i = 0; i < 5; ++i)
            if (!std::isfinite(q.g[i])) return false;
        switch (p.type) {
            case TGX_SQUARE: return finite_pos(q.g[0]) && finite_pos(q.decel);
            case TGX_RECTANGLE: return finite_pos(q.g[0]) && finite_pos(q.g[1]) && finite_pos(q.decel);
            case TGX_RECIPROCATING:
                return std::isfinite(q.g[5]) && finite_pos(q.decel) && (q.g[0] != q.g[3] || q.g[1] != q.g[4]);
            case TGX_BOUNCE: return q.g[2] != q.g[3];
            default: return finite_pos(q.g[2]) && finite_pos(q.g[3]);
        }
    }
    return false;
}

std::unique_ptr<tg::Trajectory> make_traj(const tgx_params& p) {
    if (TGX_IS_POLYLINE(p.type)) {
        const tgx_polyline_params& q = p.u.poly;
        std::vector<double> vg{q.v_goal};
        switch (p.type) {
            case TGX_SQUARE:
                return std::make_unique<tg::Square>(p.alt, q.g[0], q.g[1], q.g[2], q.orientation, vg, q.t_traj, q.decel,
                                                    p.dt);
            case TGX_RECTANGLE:
                return std::make_unique<tg::Rectangle>(p.alt, q.g[0], q.g[1], q.g[2], q.g[3], q.orientation, vg,
                                                       q.t_traj, q.decel, p.dt);
            case TGX_RECIPROCATING:
                // a1 is stored and never used by the class (Reciprocating.hpp)
                return std::make_unique<tg::Reciprocating>(p.alt, Eigen::Vector3d(q.g[0], q.g[1], q.g[2]),
                                                           Eigen::Vector3d(q.g[3], q.g[4], q.g[5]), vg, q.decel,
                                                           q.decel, q.t_traj, p.dt);
            case TGX_BOUNCE:
                return std::make_unique<tg::Bounce>(q.g[0], q.g[1], q.g[2], q.g[3], vg, q.t_traj, q.orientation, p.dt);
            case TGX_M:
                return std::make_unique<tg::M>(q.g[0], q.g[1], q.g[2], q.g[3], p.alt, vg, q.t_traj, q.orientation, p.dt);
            case TGX_I:
                return std::make_unique<tg::I>(q.g[0], q.g[1], q.g[2], q.g[3], p.alt, vg, q.t_traj, q.orientation, p.dt);
            default:
                return std::make_unique<tg::T>(q.g[0], q.g[1], q.g[2], q.g[3], p.alt, vg, q.t_traj, q.orientation, p.dt);
        }
    }
    if (p.type == TGX_BOOMERANG) {
        const tgx_line_params& l = p.u.line;
        std::vector<double> vg{l.v_goal};
        return std::make_unique<tg::Boomerang>(p.alt, Eigen::Vector3d(l.A[0], l.A[1], l.A[2]),
                                               Eigen::Vector3d(l.B[0], l.B[1], l.B[2]), vg, l.a1, l.a3, p.dt);
    }
    if (p.type == TGX_LINE) {
        const tgx_line_params& l = p.u.line;
        std::vector<double> vg{l.v_goal};
        return std::make_unique<tg::Line>(p.alt, Eigen::Vector3d(l.A[0], l.A[1], l.A[2]),
                                          Eigen::Vector3d(l.B[0], l.B[1], l.B[2]), vg, l.a1, l.a3, p.dt);
    }
    const tgx_orbit_params& o = p.u.orbit;
    std::vector<double> vg;
    for (int i = 0; i < p.n_vgoals; ++i)
        vg.push_back(i < TGX_MAX_VGOALS ? o.v_goals[i] : (&p)[i >> 3].u.orbit.v_goals[i & 7]);
    if (p.type == TGX_FIGURE8)
        return std::make_unique<tg::Figure8>(p.alt, o.r, o.cx, o.cy, vg, o.t_traj, o.accel, p.dt);
    return std::make_unique<tg::Circle>(p.alt, o.r, o.cx, o.cy, vg, o.t_traj, o.accel, p.dt);
}

void goal_to_array(const Goal& g, double a[TGX_NCHAN]) {
    a[TGX_PX] = g.p.x; a[TGX_PY] = g.p.y; a[TGX_PZ] = g.p.z;
    a[TGX_VX] = g.v.x; a[TGX_VY] = g.v.y; a[TGX_VZ] = g.v.z;
    a[TGX_AX] = g.a.x; a[TGX_AY] = g.a.y; a[TGX_AZ] = g.a.z;
    a[TGX_JX] = g.j.x; a[TGX_JY] = g.j.y; a[TGX_JZ] = g.j.z;
    a[TGX_PSI] = g.psi; a[TGX_DPSI] = g.dpsi;
}

Goal array_to_goal(const double a[TGX_NCHAN]) {
    Goal g;
    g.p.x = a[TGX_PX]; g.p.y = a[TGX_PY]; g.p.z = a[TGX_PZ];
    g.v.x = a[TGX_VX]; g.v.y = a[TGX_VY]; g.v.z = a[TGX_VZ];
    g.a.x = a[TGX_AX]; g.a.y = a[TGX_AY]; g.a.z = a[TGX_AZ];
    g.j.x = a[TGX_JX]; g.j.y = a[TGX_JY]; g.j.z = a[TGX_JZ];
    g.psi = a[TGX_PSI]; g.dpsi = a[TGX_DPSI];
    return g;
}

void repack(const std::vector<Goal>& goals, double* out, int64_t chan_stride, int64_t cap) {
    if (!out) return;
    const int64_t n = std::min<int64_t>((int64_t)goals.size(), cap);
    double a[TGX_NCHAN];
    for (int64_t k = 0; k < n; ++k) {
        goal_to_array(goals[k], a);
        for (int c = 0; c < TGX_NCHAN; ++c) out[c * chan_stride + k] = a[c];
    }
}

// index_msgs -> "key\tmessage\n" lines sorted by key.
void dump_msgs(const std::unordered_map<int, std::string>& m, char* buf, int64_t cap) {
    if (!buf || cap <= 0) return;
    std::map<int, std::string> sorted(m.begin(), m.end());
    std::string s;
    for (const auto& kv : sorted) s += std::to_string(kv.first) + "\t" + kv.second + "\n";
    std::snprintf(buf, (size_t)cap, "%s", s.c_str());
}

struct LogScope {
    tgx_stub::LogState saved;
    LogScope() : saved(tgx_stub::log_state()) {
        tgx_stub::log_state() = tgx_stub::LogState();
        tgx_stub::log_state().throw_on_error = true;
    }
    ~LogScope() { tgx_stub::log_state() = saved; }
};

int64_t generate_one(const tgx_params& p, std::vector<Goal>& goals, std::unordered_map<int, std::string>& msgs,
                     uint32_t* status) {
    uint32_t st = 0;
    if (!params_ok(p)) {
        if (status) *status = TGX_ST_BAD_PARAM;
        return -1;
    }
    LogScope scope;
    auto clock = std::make_shared<rclcpp::Clock>();
    auto traj = make_traj(p);
    try {
        traj->generateTraj(goals, msgs, clock);
    } catch (const tgx_stub::ErrorLogged&) {
        // the reference would now call exit(1) (Circle.cpp:87, Line.cpp:78, Figure8.cpp:87)
        st |= (p.type == TGX_LINE || p.type == TGX_BOOMERANG) ? TGX_ST_LINE_END_NOT_B : TGX_ST_FINAL_V_NONZERO;
    }
    if (tgx_stub::log_state().n_warn > 0) st |= TGX_ST_VGOALS_NOT_INCREASING;
    if (status) *status = st;
    return (int64_t)goals.size();
}

}  // namespace

extern "C" {

int64_t ref_generate(const tgx_params* p, double* out, int64_t chan_stride, int64_t cap, uint32_t* status,
                     char* msgs_buf, int64_t msgs_cap) {
    std::vector<Goal> goals;
    std::unordered_map<int, std::string> msgs;
    int64_t n = generate_one(*p, goals, msgs, status);
    if (n < 0) return n;
    repack(goals, out, chan_stride, cap);
    dump_msgs(msgs, msgs_buf, msgs_cap);
    return n;
}

int64_t ref_stop(const tgx_params* p, const double* from, double* out, int64_t chan_stride, int64_t cap,
                 uint32_t* status, char* msgs_buf, int64_t msgs_cap) {
    if (!params_ok(*p)) {
        if (status) *status = TGX_ST_BAD_PARAM;
        return -1;
    }
    LogScope scope;
    auto clock = std::make_shared<rclcpp::Clock>();
    auto traj = make_traj(*p);
    std::vector<Goal> goals{array_to_goal(from)};
    std::unordered_map<int, std::string> msgs;
    int pub_index = 0;
    traj->generateStopTraj(goals, msgs, pub_index, clock);
    if (status) *status = 0;
    repack(goals, out, chan_stride, cap);
    dump_msgs(msgs, msgs_buf, msgs_cap);
    return (int64_t)goals.size();
}

// Returns 1 / 0; *n_errors (may be NULL) receives how many RCLCPP_ERROR lines the call logged (Line.cpp:166 logs
// "Line trajectory not feasible" when d2 < 0).
static int inside_bounds_impl(const tgx_params* p, const double box[6], long* n_errors) {
    LogScope scope;
    tgx_stub::log_state().throw_on_error = false;   // Line.cpp:166 logs an error and returns false
    auto traj = make_traj(*p);
    const bool ok = traj->trajectoryInsideBounds(box[0], box[1], box[2], box[3], box[4], box[5]);
    if (n_errors) *n_errors = tgx_stub::log_state().n_error;
    return ok ? 1 : 0;
}

int ref_inside_bounds(const tgx_params* p, const double box[6]) { return inside_bounds_impl(p, box, nullptr); }

// Multi-threaded batch generation with SoA repack (for parity checks at moderate scale).
int ref_generate_batch(const tgx_params* p, int64_t n, double* out, int64_t traj_stride, int64_t chan_stride,
                       int64_t cap, int32_t* counts, uint32_t* status, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> th;
    auto work = [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            std::vector<Goal> goals;
            std::unordered_map<int, std::string> msgs;
            uint32_t st = 0;
            int64_t m = generate_one(p[i], goals, msgs, &st);
            if (m > 0 && out) repack(goals, out + i * traj_stride, chan_stride, cap);
            if (counts) counts[i] = (int32_t)(m < 0 ? 0 : m);
            if (status) status[i] = st;
        }
    };
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, n * t / nthreads, n * (t + 1) / nthreads);
    work(0, n / nthreads);
    for (auto& t : th) t.join();
    return 0;
}

// Feasibility as BASELINE.json config 4 defines it, reduced over the reference's own samples.
int ref_feasibility_batch(const tgx_params* p, int64_t n, const tgx_limits* limits, uint8_t* flags,
                          double* max_v, double* max_a, int32_t* counts, uint32_t* status, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> th;
    auto work = [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            std::vector<Goal> goals;
            std::unordered_map<int, std::string> msgs;
            uint32_t st = 0;
            int64_t m = generate_one(p[i], goals, msgs, &st);
            double mv = 0.0, ma = 0.0;
            for (const Goal& g : goals) {
                double nv = std::sqrt(g.v.x * g.v.x + g.v.y * g.v.y + g.v.z * g.v.z);
                double na = std::sqrt(g.a.x * g.a.x + g.a.y * g.a.y + g.a.z * g.a.z);
                if (nv > mv) mv = nv;
                if (na > ma) ma = na;
            }
            if (limits && limits->check_box && !(st & TGX_ST_BAD_PARAM)) {
                long n_err = 0;
                if (!inside_bounds_impl(&p[i], limits->box, &n_err)) {
                    st |= TGX_ST_OUTSIDE_BOUNDS;
                    if (n_err > 0) st |= TGX_ST_LINE_D2_NEGATIVE;   // the only error trajectoryInsideBounds logs
                }
            }
            if (limits && mv > limits->v_max) st |= TGX_ST_VMAX_EXCEEDED;
            if (limits && ma > limits->a_max) st |= TGX_ST_AMAX_EXCEEDED;
            if (max_v) max_v[i] = mv;
            if (max_a) max_a[i] = ma;
            if (flags) flags[i] = (st == 0) ? 1 : 0;
            if (counts) counts[i] = (int32_t)(m < 0 ? 0 : m);
            if (status) status[i] = st;
        }
    };
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, n * t / nthreads, n * (t + 1) / nthreads);
    work(0, n / nthreads);
    for (auto& t : th) t.join();
    return 0;
}

// Timing leg: the faithful reference path — generateTraj into std::vector<Goal> exactly as the node does
// (TrajectoryGenerator.cpp:71), one trajectory object per task, nthreads host threads over contiguous blocks.
// Returns the total number of samples; *checksum keeps the work observable.
int64_t ref_time_batch(const tgx_params* p, int64_t n, int nthreads, double* checksum) {
    if (nthreads < 1) nthreads = 1;
    std::vector<int64_t> totals((size_t)nthreads, 0);
    std::vector<double> sums((size_t)nthreads, 0.0);
    std::vector<std::thread> th;
    auto work = [&](int t, int64_t lo, int64_t hi) {
        int64_t total = 0;
        double cs = 0.0;
        for (int64_t i = lo; i < hi; ++i) {
            std::vector<Goal> goals;
            std::unordered_map<int, std::string> msgs;
            uint32_t st = 0;
            int64_t m = generate_one(p[i], goals, msgs, &st);
            if (m > 0) {
                total += m;
                cs += goals.back().p.x + goals[goals.size() / 2].psi;
            }
        }
        totals[(size_t)t] = total;
        sums[(size_t)t] = cs;
    };
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t, n * t / nthreads, n * (t + 1) / nthreads);
    work(0, 0, n / nthreads);
    for (auto& t : th) t.join();
    int64_t total = 0;
    double cs = 0.0;
    for (int t = 0; t < nthreads; ++t) { total += totals[(size_t)t]; cs += sums[(size_t)t]; }
    if (checksum) *checksum = cs;
    return total;
}

}  // extern "C"
