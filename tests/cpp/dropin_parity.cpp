// dropin_parity.cpp — C++ parity driver (test infrastructure).
//
// Builds the reference's own Circle / Line / Figure8 (unmodified sources, stub ROS headers) and this repo's
// GPU-backed drop-in classes (compiled into namespace tgx_dropin so both fit in one binary) and drives both through
// the SAME abstract interface, trajectory_generator::Trajectory, exactly as TrajectoryGenerator.cpp does:
//   generateTraj (:71), trajectoryInsideBounds (:419), generateStopTraj (:516), and the public create*Goal helpers.
// Exit code 0 and a final "DROPIN PARITY OK" line mean every comparison passed.
#include <cmath>
#include <cstdio>
#include <memory>
#include <random>
#include <string>
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

#include "tgx_trajectories.hpp"

using snapstack_msgs2::msg::Goal;
namespace ref = trajectory_generator;
namespace gpu = tgx_dropin;

static int g_failures = 0;
static double g_worst_pos = 0.0, g_worst_rel = 0.0;

#define CHECK(cond, ...)                          \
    do {                                          \
        if (!(cond)) {                            \
            ++g_failures;                         \
            std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
            std::printf(__VA_ARGS__);             \
            std::printf("\n");                    \
        }                                         \
    } while (0)

static double norm3(double x, double y, double z) { return std::sqrt(x * x + y * y + z * z); }

struct Peaks { double v = 0, a = 0, j = 0, dpsi = 0; };

static Peaks peaks(const std::vector<Goal>& g) {
    Peaks p;
    for (const Goal& s : g) {
        p.v = std::max(p.v, norm3(s.v.x, s.v.y, s.v.z));
        p.a = std::max(p.a, norm3(s.a.x, s.a.y, s.a.z));
        p.j = std::max(p.j, norm3(s.j.x, s.j.y, s.j.z));
        p.dpsi = std::max(p.dpsi, std::fabs(s.dpsi));
    }
    return p;
}

// Tolerances of the parity gate (tests/parity.py).
static void compareGoals(const char* what, const std::vector<Goal>& got, const std::vector<Goal>& want) {
    CHECK(got.size() == want.size(), "%s: %zu goals, reference has %zu", what, got.size(), want.size());
    if (got.size() != want.size()) return;
    const Peaks pk = peaks(want);
    for (size_t k = 0; k < want.size(); ++k) {
        const Goal &g = got[k], &w = want[k];
        const double dp = std::max(std::fabs(g.p.x - w.p.x), std::max(std::fabs(g.p.y - w.p.y), std::fabs(g.p.z - w.p.z)));
        g_worst_pos = std::max(g_worst_pos, dp);
        CHECK(dp <= 1e-9, "%s[%zu]: position off by %.3e m", what, k, dp);
        auto rel = [&](double dx, double dy, double dz, double mag, double peak) {
            return norm3(dx, dy, dz) / std::max(std::max(mag, 1e-3 * peak), 1e-6);
        };
        const double rv = rel(g.v.x - w.v.x, g.v.y - w.v.y, g.v.z - w.v.z, norm3(w.v.x, w.v.y, w.v.z), pk.v);
        const double ra = rel(g.a.x - w.a.x, g.a.y - w.a.y, g.a.z - w.a.z, norm3(w.a.x, w.a.y, w.a.z), pk.a);
        const double rj = rel(g.j.x - w.j.x, g.j.y - w.j.y, g.j.z - w.j.z, norm3(w.j.x, w.j.y, w.j.z), pk.j);
        const double dpsi = std::fabs(std::remainder(g.psi - w.psi, 2.0 * M_PI)) / std::max(std::fabs(w.psi), 1.0);
        const double rd = std::fabs(g.dpsi - w.dpsi) / std::max(std::max(std::fabs(w.dpsi), 1e-3 * pk.dpsi), 1e-6);
        const double worst = std::max(std::max(rv, ra), std::max(std::max(rj, dpsi), rd));
        g_worst_rel = std::max(g_worst_rel, worst);
        CHECK(worst <= 1e-8, "%s[%zu]: v %.2e a %.2e j %.2e psi %.2e dpsi %.2e (relative)", what, k, rv, ra, rj, dpsi, rd);
        CHECK(g.header.frame_id == w.header.frame_id && g.power == w.power, "%s[%zu]: frame_id / power", what, k);
    }
}

static void compareMsgs(const char* what, const std::unordered_map<int, std::string>& got,
                        const std::unordered_map<int, std::string>& want) {
    CHECK(got.size() == want.size(), "%s: %zu index_msgs, reference has %zu", what, got.size(), want.size());
    for (const auto& kv : want) {
        auto it = got.find(kv.first);
        CHECK(it != got.end(), "%s: index_msgs[%d] missing", what, kv.first);
        if (it != got.end())
            CHECK(it->second == kv.second, "%s: index_msgs[%d] = \"%s\", reference \"%s\"", what, kv.first,
                  it->second.c_str(), kv.second.c_str());
    }
}

static void runPair(const char* what, ref::Trajectory& r, ref::Trajectory& g, int stop_from) {
    auto clock = std::make_shared<rclcpp::Clock>();
    std::vector<Goal> rg, gg;
    std::unordered_map<int, std::string> rm, gm;
    // the node appends to vectors it owns; start both from a non-empty vector to exercise the key offset
    rg.emplace_back(); gg.emplace_back();
    r.generateTraj(rg, rm, clock);
    g.generateTraj(gg, gm, clock);
    compareGoals(what, gg, rg);
    compareMsgs(what, gm, rm);
    std::printf("%-28s N = %zu, %zu index_msgs\n", what, rg.size() - 1, rm.size());

    for (double half : {5.0, 2.0}) {
        const bool rb = r.trajectoryInsideBounds(-half, half, -half, half, 0.0, 3.0);
        const bool gb = g.trajectoryInsideBounds(-half, half, -half, half, 0.0, 3.0);
        CHECK(rb == gb, "%s: trajectoryInsideBounds(+-%.0f) = %d, reference %d", what, half, (int)gb, (int)rb);
    }

    // braking: both get the SAME input vector (the reference's), as the count depends on that sample's bits
    std::vector<Goal> rs = rg, gs = rg;
    std::unordered_map<int, std::string> rsm = rm, gsm = rm;
    int rpi = stop_from, gpi = stop_from;
    r.generateStopTraj(rs, rsm, rpi, clock);
    g.generateStopTraj(gs, gsm, gpi, clock);
    std::string w2 = std::string(what) + " stop";
    CHECK(rpi == 0 && gpi == 0, "%s: pub_index %d, reference %d", w2.c_str(), gpi, rpi);
    compareGoals(w2.c_str(), gs, rs);
    compareMsgs(w2.c_str(), gsm, rsm);
    std::printf("%-28s braking from %d: %zu samples\n", what, stop_from, rs.size());
}

int main() {
    const double dt = 0.01;
    {   // config/default.yaml, the three classes on the path
        ref::Circle r(1.8, 3.4, 0.0, 0.0, {1.0, 2.0, 2.0}, 80.0, 0.4, dt);
        gpu::Circle g(1.8, 3.4, 0.0, 0.0, {1.0, 2.0, 2.0}, 80.0, 0.4, dt);
        runPair("default Circle", r, g, 12500);
        const Goal a = r.createCircleGoal(1.7, 0.4, 2.5), b = g.createCircleGoal(1.7, 0.4, 2.5);
        compareGoals("createCircleGoal", {b}, {a});
    }
    {
        ref::Figure8 r(1.8, 3.4, 0.0, 0.0, {1.0, 2.0, 2.0}, 80.0, 0.4, dt);
        gpu::Figure8 g(1.8, 3.4, 0.0, 0.0, {1.0, 2.0, 2.0}, 80.0, 0.4, dt);
        runPair("default Figure8", r, g, 12500);
        const Goal a = r.createFigure8Goal(1.2, 0.4, 4.0), b = g.createFigure8Goal(1.2, 0.4, 4.0);
        compareGoals("createFigure8Goal", {b}, {a});
    }
    {
        const Eigen::Vector3d A(0.0, -3.0, 1.8), B(0.0, 3.0, 1.8);
        ref::Line r(1.8, A, B, {1.0}, 1.5, 1.0, dt);
        gpu::Line g(1.8, A, B, {1.0}, 1.5, 1.0, dt);
        runPair("default Line", r, g, 342);
        const Goal a = r.createLineGoal(0.3, -1.0, 0.8, 1.5, 0.7), b = g.createLineGoal(0.3, -1.0, 0.8, 1.5, 0.7);
        compareGoals("createLineGoal", {b}, {a});
    }
    {   // Line out and back
        const Eigen::Vector3d A(0.0, -3.0, 1.8), B(0.0, 3.0, 1.8);
        ref::Boomerang r(1.8, A, B, {1.0}, 1.5, 1.0, dt);
        gpu::Boomerang g(1.8, A, B, {1.0}, 1.5, 1.0, dt);
        runPair("default Boomerang", r, g, 1000);
        const Eigen::Vector3d C(-4.25, -3.5, 1.0), D(4.5, 4.25, 1.0);
        ref::Boomerang r2(1.0, C, D, {3.0}, 1.5, 1.0, dt);
        gpu::Boomerang g2(1.0, C, D, {3.0}, 1.5, 1.0, dt);
        runPair("diagonal Boomerang", r2, g2, 500);
    }
    {   // a Line that does not fit its bounds check (d2 < 0) must report false on both sides
        const Eigen::Vector3d A(0.0, -3.0, 1.8), B(0.0, -2.5, 1.8);
        ref::Line r(1.8, A, B, {1.0}, 1.5, 1.0, dt);
        gpu::Line g(1.8, A, B, {1.0}, 1.5, 1.0, dt);
        CHECK(!r.trajectoryInsideBounds(-5, 5, -5, 5, -5, 5) && !g.trajectoryInsideBounds(-5, 5, -5, 5, -5, 5),
              "short line: both sides must reject");
    }
    {   // the constant-speed polyline family, config/default.yaml values (traj_type: T ships as
        // the default, default.yaml:9) and a rotated / off-centre variant of each
        const std::vector<double> vg{1.0, 2.0, 2.0};
        for (int variant = 0; variant < 2; ++variant) {
            const double ori = variant ? 0.5 : 0.0, cx = variant ? 0.4 : 0.0, cy = variant ? -0.3 : 0.0;
            const double T = variant ? 23.7 : 80.0, alt = 1.8;
            const std::string tag = variant ? "rotated " : "default ";
            {
                ref::Square r(alt, 2.0, cx, cy, ori, vg, T, 0.4, dt);
                gpu::Square g(alt, 2.0, cx, cy, ori, vg, T, 0.4, dt);
                runPair((tag + "Square").c_str(), r, g, 1234);
                compareGoals("createSquareGoal", {g.createSquareGoal(1.0, 2.0, 1.5, -0.4, 0.7)},
                             {r.createSquareGoal(1.0, 2.0, 1.5, -0.4, 0.7)});
            }
            {
                ref::Rectangle r(alt, 2.0, 4.0, cx, cy, ori, vg, T, 0.4, dt);
                gpu::Rectangle g(alt, 2.0, 4.0, cx, cy, ori, vg, T, 0.4, dt);
                runPair((tag + "Rectangle").c_str(), r, g, 777);
            }
            {
                const Eigen::Vector3d A(cx, -3.0, alt), B(cy, 3.0, alt);
                ref::Reciprocating r(alt, A, B, {1.0}, 1.5, 1.0, T, dt);
                gpu::Reciprocating g(alt, A, B, {1.0}, 1.5, 1.0, T, dt);
                runPair((tag + "Reciprocating").c_str(), r, g, 300);
            }
            {
                ref::Bounce r(cx, cy, 4.0, 1.0, vg, T, ori, dt);
                gpu::Bounce g(cx, cy, 4.0, 1.0, vg, T, ori, dt);
                runPair((tag + "Bounce").c_str(), r, g, 450);
                compareGoals("createBounceGoal", {g.createBounceGoal(0.1, 0.2, 2.5, -1.0, 0.3)},
                             {r.createBounceGoal(0.1, 0.2, 2.5, -1.0, 0.3)});
            }
            {
                ref::M r(cx, cy, 3.0, 4.0, alt, vg, T, ori, dt);
                gpu::M g(cx, cy, 3.0, 4.0, alt, vg, T, ori, dt);
                runPair((tag + "M").c_str(), r, g, 2000);
            }
            {
                ref::I r(cx, cy, 3.0, 4.0, alt, vg, T, ori, dt);
                gpu::I g(cx, cy, 3.0, 4.0, alt, vg, T, ori, dt);
                runPair((tag + "I").c_str(), r, g, 2100);
            }
            {
                ref::T r(cx, cy, 3.0, 4.0, alt, vg, T, ori, dt);
                gpu::T g(cx, cy, 3.0, 4.0, alt, vg, T, ori, dt);
                runPair((tag + "T").c_str(), r, g, 999);
                compareGoals("createTGoal", {g.createTGoal(-1.0, 0.5, 2.0, 0.0, -2.1)},
                             {r.createTGoal(-1.0, 0.5, 2.0, 0.0, -2.1)});
            }
        }
    }
    {   // v_goals of any length (Circle.cpp:43 loops over a std::vector): 12 and 17 goal speeds go through continuation
        // records (tgx.h: TGX_VGOALS_MORE); an empty vector is the start sample alone; a negative radius is the
        // mirrored circle.  The node checks vel > 0 and accel > 0 only (TrajectoryGenerator.cpp:184-195).
        std::vector<double> v12, v17;
        for (int i = 0; i < 12; ++i) v12.push_back(0.3 + 0.2 * i);
        for (int i = 0; i < 17; ++i) v17.push_back(0.25 + 0.15 * i);
        ref::Circle r12(1.5, 2.0, 0.3, -0.2, v12, 0.7, 1.0, dt);
        gpu::Circle g12(1.5, 2.0, 0.3, -0.2, v12, 0.7, 1.0, dt);
        runPair("12-goal Circle", r12, g12, 700);
        ref::Figure8 r17(1.2, 1.5, 0.0, 0.5, v17, 0.4, 1.5, dt);
        gpu::Figure8 g17(1.2, 1.5, 0.0, 0.5, v17, 0.4, 1.5, dt);
        runPair("17-goal Figure8", r17, g17, 900);
        ref::Circle r0(1.5, 2.0, 0.3, -0.2, {}, 0.7, 1.0, dt);
        gpu::Circle g0(1.5, 2.0, 0.3, -0.2, {}, 0.7, 1.0, dt);
        runPair("empty-v_goals Circle", r0, g0, 0);
        ref::Circle rn(1.5, -2.0, 0.3, -0.2, {1.0, 2.0}, 0.7, 1.0, dt);
        gpu::Circle gn(1.5, -2.0, 0.3, -0.2, {1.0, 2.0}, 0.7, 1.0, dt);
        runPair("negative-radius Circle", rn, gn, 200);
    }
    // random parameters
    std::mt19937_64 rng(20261018);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    for (int i = 0; i < 12; ++i) {
        const double r0 = 0.5 + 4.5 * U(rng), cx = -2 + 4 * U(rng), cy = -2 + 4 * U(rng), alt = 1 + 1.5 * U(rng);
        const double v1 = 0.5 + 2.5 * U(rng), acc = 0.7 + 1.3 * U(rng), t = 1.0 + 4.0 * U(rng);
        std::vector<double> vg = (i % 3 == 0) ? std::vector<double>{0.5 * v1, v1} : std::vector<double>{v1};
        const std::string name = "random #" + std::to_string(i);
        if (i % 3 == 0) {
            ref::Circle r(alt, r0, cx, cy, vg, t, acc, dt);
            gpu::Circle g(alt, r0, cx, cy, vg, t, acc, dt);
            runPair((name + " Circle").c_str(), r, g, 150);
        } else if (i % 3 == 1) {
            ref::Figure8 r(alt, r0, cx, cy, vg, t, acc, dt);
            gpu::Figure8 g(alt, r0, cx, cy, vg, t, acc, dt);
            runPair((name + " Figure8").c_str(), r, g, 150);
        } else {
            const Eigen::Vector3d A(-4 + 3 * U(rng), -4 + 3 * U(rng), alt), B(1 + 3 * U(rng), 1 + 3 * U(rng), alt);
            const double vl = 0.5 + 1.0 * U(rng), a1 = 0.8 + U(rng), a3 = 0.5 + U(rng);
            ref::Line r(alt, A, B, {vl}, a1, a3, dt);
            gpu::Line g(alt, A, B, {vl}, a1, a3, dt);
            runPair((name + " Line").c_str(), r, g, 90);
        }
    }
    std::printf("worst position error %.3e m, worst relative error %.3e\n", g_worst_pos, g_worst_rel);
    if (g_failures) {
        std::printf("DROPIN PARITY FAILED: %d checks\n", g_failures);
        return 1;
    }
    std::printf("DROPIN PARITY OK\n");
    return 0;
}
