// tgx_trajectories.cpp — see tgx_trajectories.hpp.  Host glue only: every sample comes from libtgx (CUDA).
#include "tgx_trajectories.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace TGX_DROPIN_NAMESPACE {

namespace {

using snapstack_msgs2::msg::Goal;

[[noreturn]] void die(const rclcpp::Logger& logger, const char* what, int rc) {
    RCLCPP_ERROR(logger, "tgx: %s failed: %s (%s)", what, tgx_strerror(rc), tgx_last_cuda_error());
    std::exit(1);   // generateTraj "can exit the program" (Trajectory.hpp:32); there is no CPU fallback
}

// std::to_string(double) is "%f" (Circle.cpp:45, 61-62).
std::string f6(double x) { return std::to_string(x); }

std::string phaseText(const std::string& shape, int type, int kind, double value, double value2, bool stop_traj) {
    switch (kind) {
        case TGX_PH_ACCEL_TO: return shape + " traj: accelerating to " + f6(value) + " m/s";
        case TGX_PH_REACHED:
            return shape + " traj: reached " + f6(value) + " m/s, keeping constant v for " + f6(value2) + " s";
        case TGX_PH_DECEL: return shape + " traj: decelerating to 0 m/s";
        case TGX_PH_STOPPED:
            // Figure8::generateTraj announces "Figure 8 traj: stopped" (Figure8.cpp:89); its braking path and the
            // other classes use the plain shape name.
            if (type == TGX_FIGURE8 && !stop_traj) return "Figure 8 traj: stopped";
            return shape + " traj: stopped";
        case TGX_PH_PRESSED_END: return shape + " traj: pressed END, decelerating to 0 m/s";
        default: return shape + " traj: ?";
    }
}

// index_msgs text of sample k (of n) of a polyline-family trajectory whose leg is `leg` (tgx_polyline_legs numbering):
// Square.cpp:61,79,88; Rectangle.cpp:61,79,88; Reciprocating.cpp:47,57; Bounce.cpp:42,50; M.cpp:57,65; I.cpp:65,73;
// T.cpp:63,71.
std::string polylineText(const std::string& shape, int type, int leg, int64_t k, int64_t n) {
    const bool last = k == n - 1;
    switch (type) {
        case TGX_SQUARE:
        case TGX_RECTANGLE:
            if (last) return shape + " traj: completed";
            if (leg < 0) return shape + " traj: starting at corner 0";
            return shape + " traj: moving along side " + std::to_string(leg);
        case TGX_RECIPROCATING:
            if (leg == 1 || leg == 3) return "Reciprocating: yaw flip at endpoint";
            return leg == 0 ? "Reciprocating: forward" : "Reciprocating: reverse";
        case TGX_BOUNCE:
            if (last) return "Bounce: completed";
            return leg == 0 ? "Bounce: ascending" : "Bounce: descending";
        default: {
            const int nseg = type == TGX_M ? 4 : (type == TGX_I ? 5 : 3);
            if (last) return shape + " traj: completed";
            return shape + " traj: segment " + std::to_string(leg % nseg) + (leg < nseg ? " fwd" : " rev");
        }
    }
}

// The reference announces every sample of a polyline trajectory; rebuild its map from the leg structure.
void polylineMessages(const std::string& shape, int type, const tgx_polyline_legs& L, int base,
                      std::unordered_map<int, std::string>& index_msgs) {
    const int64_t n = L.n;
    if (n <= 0) {
        // index_msgs[goals.size() - 1] = "... completed" on an empty vector (Bounce.cpp:50, M.cpp:65): key size()-1
        if (type != TGX_RECIPROCATING) index_msgs[base - 1] = polylineText(shape, type, 0, -1, 0);
        return;
    }
    int64_t k = 0;
    if (L.first_special) index_msgs[base + (int)k++] = polylineText(shape, type, -1, 0, n);
    int leg = 0;
    while (k < n && L.n_legs > 0) {
        for (int c = 0; c < L.count[leg] && k < n; ++c, ++k)
            index_msgs[base + (int)k] = polylineText(shape, type, leg, k, n);
        leg = (leg + 1) % L.n_legs;
    }
    // the yaw-flip goal appended after a leg that t_traj cut short (Reciprocating.cpp:50-57) replaces the last entry
    if (L.last_special) {
        int m = (int)((n - 2 - L.first_special) % L.period), l = 0;
        while (l + 1 < L.n_legs && m >= L.count[l]) m -= L.count[l++];
        index_msgs[base + (int)(n - 1)] = polylineText(shape, type, l + 1, n - 1, n);
    }
}

Goal goalFromPlanes(const double* row, int64_t cap, int64_t k) {
    Goal g;
    g.header.frame_id = "world";                 // Circle.cpp:106
    g.p.x = row[TGX_PX * cap + k];  g.p.y = row[TGX_PY * cap + k];  g.p.z = row[TGX_PZ * cap + k];
    g.v.x = row[TGX_VX * cap + k];  g.v.y = row[TGX_VY * cap + k];  g.v.z = row[TGX_VZ * cap + k];
    g.a.x = row[TGX_AX * cap + k];  g.a.y = row[TGX_AY * cap + k];  g.a.z = row[TGX_AZ * cap + k];
    g.j.x = row[TGX_JX * cap + k];  g.j.y = row[TGX_JY * cap + k];  g.j.z = row[TGX_JZ * cap + k];
    g.psi = row[TGX_PSI * cap + k];
    g.dpsi = row[TGX_DPSI * cap + k];
    g.power = true;                              // Circle.cpp:127
    return g;
}

// The compact wire format (tgx_generate_host_compact): the 10 planes that vary along a trajectory; the z-components
// are the constants the reference writes literally (p.z = alt_, v.z = a.z = j.z = 0: Circle.cpp:109-121).
Goal goalFromCompact(const double* row, int64_t cap, int64_t k, double alt) {
    Goal g;
    g.header.frame_id = "world";                 // Circle.cpp:106
    g.p.x = row[0 * cap + k];  g.p.y = row[1 * cap + k];  g.p.z = alt;
    g.v.x = row[2 * cap + k];  g.v.y = row[3 * cap + k];  g.v.z = 0.0;
    g.a.x = row[4 * cap + k];  g.a.y = row[5 * cap + k];  g.a.z = 0.0;
    g.j.x = row[6 * cap + k];  g.j.y = row[7 * cap + k];  g.j.z = 0.0;
    g.psi = row[8 * cap + k];
    g.dpsi = row[9 * cap + k];
    g.power = true;                              // Circle.cpp:127
    return g;
}

void goalToArray(const Goal& g, double a[TGX_NCHAN]) {
    a[TGX_PX] = g.p.x; a[TGX_PY] = g.p.y; a[TGX_PZ] = g.p.z;
    a[TGX_VX] = g.v.x; a[TGX_VY] = g.v.y; a[TGX_VZ] = g.v.z;
    a[TGX_AX] = g.a.x; a[TGX_AY] = g.a.y; a[TGX_AZ] = g.a.z;
    a[TGX_JX] = g.j.x; a[TGX_JY] = g.j.y; a[TGX_JZ] = g.j.z;
    a[TGX_PSI] = g.psi; a[TGX_DPSI] = g.dpsi;
}

// Page-locked staging buffer, reused across calls.
struct Staging {
    double* p = nullptr;
    int64_t doubles = 0;
    double* reserve(int64_t want) {
        if (want <= doubles) return p;
        if (p) tgx_free_host(p);
        p = static_cast<double*>(tgx_alloc_host(want * (int64_t)sizeof(double)));
        doubles = p ? want : 0;
        return p;
    }
};

Staging& staging() {
    static Staging s;
    return s;
}

// The reference takes a std::vector of any length (Circle.cpp:43): goal speeds beyond the eighth go into continuation
// records (tgx.h: TGX_VGOALS_MORE).  More than TGX_MAX_VGOALS_TOTAL of them is reported by the engine as
// TGX_ST_BAD_PARAM when the trajectory is generated (RCLCPP_ERROR + exit(1), like the reference's own fatal paths).
std::vector<tgx_params> orbitParams(int type, double alt, double r, double cx, double cy,
                                    const std::vector<double>& v_goals, double t_traj, double accel, double dt) {
    const size_t k = v_goals.size();
    const size_t rows = k <= TGX_MAX_VGOALS_TOTAL ? (size_t)TGX_ORBIT_RECORDS((int)k) : 1;
    std::vector<tgx_params> recs(rows);
    std::memset(recs.data(), 0, rows * sizeof(tgx_params));
    tgx_params& p = recs[0];
    p.type = type;
    p.n_vgoals = (int32_t)std::min<size_t>(k, 0x7fffffff);
    p.dt = dt;
    p.alt = alt;
    p.u.orbit.r = r;
    p.u.orbit.cx = cx;
    p.u.orbit.cy = cy;
    p.u.orbit.t_traj = t_traj;
    p.u.orbit.accel = accel;
    for (size_t q = 0; q < rows; ++q) {
        tgx_params& rec = recs[q];
        if (q) {
            rec.type = TGX_VGOALS_MORE;
            rec.dt = dt;
            rec.alt = alt;
        }
        int cnt = 0;
        for (size_t g = q * TGX_MAX_VGOALS; g < k && g < (q + 1) * TGX_MAX_VGOALS; ++g) rec.u.orbit.v_goals[cnt++] = v_goals[g];
        if (q) rec.n_vgoals = cnt;
    }
    return recs;
}

}  // namespace

tgx_engine* sharedEngine() {
    static tgx_engine* engine = nullptr;
    if (!engine) {
        int device = 0;
        if (const char* env = std::getenv("TGX_DEVICE")) device = std::atoi(env);
        const int rc = tgx_create(&engine, device);
        if (rc != TGX_OK) {
            std::fprintf(stderr, "tgx: cannot create the GPU engine on device %d: %s (%s)\n", device,
                         tgx_strerror(rc), tgx_last_cuda_error());
            std::exit(1);
        }
    }
    return engine;
}

GpuTrajectory::GpuTrajectory(const tgx_params& params, const char* shape, const char* logger_name)
    : ::trajectory_generator::Trajectory(params.dt), params_(params), records_(1, params), shape_(shape),
      logger_(rclcpp::get_logger(logger_name)) {}

GpuTrajectory::GpuTrajectory(const std::vector<tgx_params>& records, const char* shape, const char* logger_name)
    : ::trajectory_generator::Trajectory(records.at(0).dt), params_(records.at(0)), records_(records), shape_(shape),
      logger_(rclcpp::get_logger(logger_name)) {}

GpuTrajectory::~GpuTrajectory() {}

void GpuTrajectory::generateTraj(std::vector<Goal>& goals, std::unordered_map<int, std::string>& index_msgs,
                                 const rclcpp::Clock::SharedPtr& clock) {
    rclcpp::Time tstart = clock->now();
    tgx_engine* e = sharedEngine();

    // pass 1: the exact sample count (replays the reference's loops on the GPU)
    // (a Circle / Figure8 with more than 8 goal speeds is R > 1 records: the trajectory is row 0, the continuation rows
    //  carry no samples, only further index_msgs entries)
    const int64_t R = (int64_t)records_.size();
    std::vector<int32_t> counts((size_t)R, 0);
    std::vector<uint32_t> stats((size_t)R, 0u);
    int rc = tgx_count_host(e, records_.data(), R, nullptr, counts.data(), stats.data());
    if (rc != TGX_OK) die(logger_, "tgx_count_host", rc);
    int32_t n = counts[0];
    uint32_t status = stats[0];
    last_status_ = status;
    if (status & (TGX_ST_BAD_PARAM | TGX_ST_TOO_LONG)) {
        RCLCPP_ERROR(logger_, "Error: %s trajectory parameters rejected (status 0x%x)", shape_.c_str(), status);
        std::exit(1);
    }

    // pass 2: all samples into a page-locked SoA row, then repack to the reference's AoS messages
    const int64_t cap = std::max<int64_t>(4, ((int64_t)n + 3) / 4 * 4);
    double* row = staging().reserve(R * (int64_t)TGX_NCHAN * cap);
    if (!row) die(logger_, "tgx_alloc_host", TGX_ERR_NOMEM);
    std::vector<tgx_phases> phase_rows((size_t)R);
    std::vector<tgx_polyline_legs> leg_rows((size_t)R);
    // Bounce moves along z (Bounce.cpp:39-41) and ships all 14 planes; every other class uses the compact format
    const bool compact = params_.type != TGX_BOUNCE;
    if (compact)
        rc = tgx_generate_host_compact(e, records_.data(), R, nullptr, row, cap, counts.data(), stats.data(),
                                       phase_rows.data(), leg_rows.data());
    else
        rc = tgx_generate_host_legs(e, records_.data(), R, nullptr, row, cap, counts.data(), stats.data(),
                                    phase_rows.data(), leg_rows.data());
    if (rc != TGX_OK) die(logger_, "tgx_generate_host", rc);
    n = counts[0];
    status = stats[0];
    const tgx_polyline_legs& legs = leg_rows[0];
    last_status_ = status;

    const size_t base = goals.size();            // generateTraj APPENDS (Circle.cpp:41: push_back, keys size()-1)
    goals.reserve(base + (size_t)n);
    if (compact)
        for (int64_t k = 0; k < n; ++k) goals.push_back(goalFromCompact(row, cap, k, params_.alt));
    else
        for (int64_t k = 0; k < n; ++k) goals.push_back(goalFromPlanes(row, cap, k));
    if (TGX_IS_POLYLINE(params_.type)) {
        polylineMessages(shape_, params_.type, legs, (int)base, index_msgs);
    } else {
        for (const tgx_phases& phases : phase_rows)
            for (int i = 0; i < phases.n; ++i)
                index_msgs[(int)base + phases.key[i]] =
                    phaseText(shape_, params_.type, phases.kind[i], phases.value[i], phases.value2[i], false);
    }

    if (status & TGX_ST_VGOALS_NOT_INCREASING)   // Circle.cpp:57-59, Figure8.cpp:57-59
        RCLCPP_WARN(logger_, "Vels are not in increasing order, ignoring vels from the first to decrease...");
    if (status & TGX_ST_FINAL_V_NONZERO) {       // Circle.cpp:85-88
        RCLCPP_ERROR(logger_, "Error: final velocity is not zero");
        std::exit(1);
    }
    if (status & TGX_ST_LINE_END_NOT_B) {        // Line.cpp:76-79 (Boomerang.cpp:76-79, 126-129: "... is not A")
        RCLCPP_ERROR(logger_, "Error: final point is not B");
        std::exit(1);
    }
    RCLCPP_INFO(logger_, "Time to calculate the traj (s): %f", (clock->now() - tstart).seconds());
    RCLCPP_INFO(logger_, "Goal vector size = %lu", goals.size());
}

void GpuTrajectory::generateStopTraj(std::vector<Goal>& goals, std::unordered_map<int, std::string>& index_msgs,
                                     int& pub_index, const rclcpp::Clock::SharedPtr& clock) {
    rclcpp::Time tstart = clock->now();
    tgx_engine* e = sharedEngine();

    double from[TGX_NCHAN];
    goalToArray(goals[pub_index], from);         // the setpoint being braked from (Circle.cpp:140-143)

    // a braking ramp never has more samples than v / (a*dt) + 2; size the row from a count-only call
    const int64_t R = (int64_t)records_.size();        // continuation records brake nothing (rows 1.. stay empty)
    std::vector<int32_t> counts((size_t)R, 0);
    std::vector<uint32_t> stats((size_t)R, 0u);
    std::vector<double> from_rows((size_t)(R * TGX_NCHAN), 0.0);
    std::copy(from, from + TGX_NCHAN, from_rows.begin());
    double dummy[4 * TGX_NCHAN];
    int rc = tgx_stop_host(e, records_.data(), R, from_rows.data(), dummy, 0, counts.data(), stats.data(), nullptr);
    if (rc != TGX_OK) die(logger_, "tgx_stop_host", rc);
    int32_t n = counts[0];
    const int64_t cap = ((int64_t)n + 3) / 4 * 4;
    double* row = staging().reserve(R * (int64_t)TGX_NCHAN * (cap > 0 ? cap : 4));
    if (!row) die(logger_, "tgx_alloc_host", TGX_ERR_NOMEM);
    std::vector<tgx_phases> phase_rows((size_t)R);
    rc = tgx_stop_host(e, records_.data(), R, from_rows.data(), row, cap, counts.data(), stats.data(), phase_rows.data());
    if (rc != TGX_OK) die(logger_, "tgx_stop_host", rc);
    n = counts[0];
    const uint32_t status = stats[0];
    const tgx_phases& phases = phase_rows[0];
    last_status_ = status & ~(uint32_t)TGX_ST_TRUNCATED;

    std::vector<Goal> goals_tmp;
    std::unordered_map<int, std::string> index_msgs_tmp;
    goals_tmp.reserve((size_t)n);
    for (int64_t k = 0; k < n; ++k) goals_tmp.push_back(goalFromPlanes(row, cap, k));
    for (int i = 0; i < phases.n; ++i)
        index_msgs_tmp[phases.key[i]] =
            phaseText(shape_, params_.type, phases.kind[i], phases.value[i], phases.value2[i], true);

    goals = std::move(goals_tmp);                // Circle.cpp:162-164: replace, reset the publication index
    index_msgs = std::move(index_msgs_tmp);
    pub_index = 0;

    RCLCPP_INFO(logger_, "Time to calculate the braking traj (s): %f", (clock->now() - tstart).seconds());
    RCLCPP_INFO(logger_, "Goal vector size = %lu", goals.size());
}

bool GpuTrajectory::trajectoryInsideBounds(double xmin, double xmax, double ymin, double ymax, double zmin,
                                           double zmax) {
    tgx_limits lim;
    std::memset(&lim, 0, sizeof(lim));
    lim.box[0] = xmin; lim.box[1] = xmax; lim.box[2] = ymin; lim.box[3] = ymax; lim.box[4] = zmin; lim.box[5] = zmax;
    lim.check_box = 1;
    const int64_t R = (int64_t)records_.size();
    std::vector<int32_t> counts((size_t)R, 0);
    std::vector<uint32_t> stats((size_t)R, 0u);
    const int rc = tgx_count_host(sharedEngine(), records_.data(), R, &lim, counts.data(), stats.data());
    if (rc != TGX_OK) die(logger_, "tgx_count_host", rc);
    const uint32_t status = stats[0];
    if (status & TGX_ST_LINE_D2_NEGATIVE)        // Line.cpp:165-168
        RCLCPP_ERROR(logger_, "Line trajectory not feasible. Please increase accel, decrease v, or increase line length.");
    // the reference tests the geometry alone (Circle.cpp:171-179): parameters its samplers would choke on do not make
    // the box test fail, and the engine reports the box test for rejected records too
    return (status & TGX_ST_OUTSIDE_BOUNDS) == 0;
}

Goal GpuTrajectory::sampleGoal(double v, double accel, double s0, double s1) const {
    double out[TGX_NCHAN];
    const int rc = tgx_sample_host(sharedEngine(), &params_, v, accel, s0, s1, out);
    if (rc != TGX_OK) die(logger_, "tgx_sample_host", rc);
    return goalFromPlanes(out, 1, 0);
}

Circle::Circle(double alt, double r, double cx, double cy, std::vector<double> v_goals, double t_traj,
               double accel, double dt)
    : GpuTrajectory(orbitParams(TGX_CIRCLE, alt, r, cx, cy, v_goals, t_traj, accel, dt), "Circle", "circle_logger") {}

Goal Circle::createCircleGoal(double v, double accel, double theta) const { return sampleGoal(v, accel, theta, 0.0); }

Figure8::Figure8(double alt, double r, double cx, double cy, std::vector<double> v_goals, double t_traj,
                 double accel, double dt)
    : GpuTrajectory(orbitParams(TGX_FIGURE8, alt, r, cx, cy, v_goals, t_traj, accel, dt), "Figure8",
                    "figure8_logger") {}

Goal Figure8::createFigure8Goal(double v, double accel, double theta) const {
    return sampleGoal(v, accel, theta, 0.0);
}

namespace {
tgx_params lineParams(int type, double alt, const Eigen::Vector3d& A, const Eigen::Vector3d& B,
                      const std::vector<double>& v_goals, double a1, double a3, double dt) {
    if (v_goals.empty()) throw std::invalid_argument("tgx: Line needs v_goals[0]");
    tgx_params p;
    std::memset(&p, 0, sizeof(p));
    p.type = type;
    p.n_vgoals = 1;
    p.dt = dt;
    p.alt = alt;
    p.u.line.A[0] = A.x(); p.u.line.A[1] = A.y(); p.u.line.A[2] = A.z();
    p.u.line.B[0] = B.x(); p.u.line.B[1] = B.y(); p.u.line.B[2] = B.z();
    p.u.line.a1 = a1;
    p.u.line.a3 = a3;
    p.u.line.v_goal = v_goals[0];                // "for now just 1 element" (Line.hpp:57, Line.cpp:43)
    return p;
}
}  // namespace

Line::Line(double alt, Eigen::Vector3d A, Eigen::Vector3d B, std::vector<double> v_goals, double a1, double a3,
           double dt)
    : GpuTrajectory(lineParams(TGX_LINE, alt, A, B, v_goals, a1, a3, dt), "Line", "line_logger") {}

namespace {
Goal lineGoal(const tgx_params& params, const rclcpp::Logger& logger, double last_x, double last_y, double v,
              double accel, double theta) {
    // the line's own heading lives in the plan; an explicit theta is passed through the state slot of the call
    double out[TGX_NCHAN];
    tgx_params p = params;
    p.u.line.reserved[0] = theta;                // tgx_sample_host reads the explicit heading here for lines
    const int rc = tgx_sample_host(sharedEngine(), &p, v, accel, last_x, last_y, out);
    if (rc != TGX_OK) die(logger, "tgx_sample_host", rc);
    return goalFromPlanes(out, 1, 0);
}
}  // namespace

Goal Line::createLineGoal(double last_x, double last_y, double v, double accel, double theta) const {
    return lineGoal(params_, logger_, last_x, last_y, v, accel, theta);
}

// Boomerang.cpp announces itself with Line's texts and logger ("Line traj: ...", "line_logger", Boomerang.hpp:63).
Boomerang::Boomerang(double alt, Eigen::Vector3d A, Eigen::Vector3d B, std::vector<double> v_goals, double a1,
                     double a3, double dt)
    : GpuTrajectory(lineParams(TGX_BOOMERANG, alt, A, B, v_goals, a1, a3, dt), "Line", "line_logger") {}

Goal Boomerang::createLineGoal(double last_x, double last_y, double v, double accel, double theta) const {
    return lineGoal(params_, logger_, last_x, last_y, v, accel, theta);
}

// ---- constant-speed polyline family ---------------------------------------------------------------------------

namespace {
tgx_params polyParams(int type, double dt, double alt, double t_traj, const std::vector<double>& v_goals, double decel,
                      double orientation, std::initializer_list<double> geometry) {
    tgx_params p;
    std::memset(&p, 0, sizeof(p));
    p.type = type;
    p.dt = dt;
    p.alt = alt;
    p.u.poly.t_traj = t_traj;
    p.u.poly.v_goal = v_goals.empty() ? 1.0 : v_goals[0];   // Square.cpp:48, Bounce.cpp:26, M.cpp:39
    p.u.poly.decel = decel;
    p.u.poly.orientation = orientation;
    int i = 0;
    for (double g : geometry) p.u.poly.g[i++] = g;
    // cos / sin of the orientation from THIS host's libm, as the reference computes them (Square.cpp:37-38)
    tgx_polyline_finalize_host(&p, 1);
    return p;
}
}  // namespace

Goal GpuTrajectory::polylineGoal(double x, double y, double v, double accel, double heading, double z) const {
    double out[TGX_NCHAN];
    tgx_params p = params_;
    p.u.poly.g[5] = z;                           // tgx_plan_samples: z (Bounce) and the explicit heading
    p.u.poly.g[6] = heading;
    const int rc = tgx_sample_host(sharedEngine(), &p, v, accel, x, y, out);
    if (rc != TGX_OK) die(logger_, "tgx_sample_host", rc);
    return goalFromPlanes(out, 1, 0);
}

Square::Square(double alt, double side_length, double cx, double cy, double orientation, std::vector<double> v_goals,
               double t_traj, double accel, double dt)
    : GpuTrajectory(polyParams(TGX_SQUARE, dt, alt, t_traj, v_goals, accel, orientation, {side_length, cx, cy}),
                    "Square", "square_logger") {}
Goal Square::createSquareGoal(double x, double y, double v, double accel, double heading) const {
    return polylineGoal(x, y, v, accel, heading, 0.0);
}

Rectangle::Rectangle(double alt, double side_a, double side_b, double cx, double cy, double orientation,
                     std::vector<double> v_goals, double t_traj, double accel, double dt)
    : GpuTrajectory(polyParams(TGX_RECTANGLE, dt, alt, t_traj, v_goals, accel, orientation, {side_a, side_b, cx, cy}),
                    "Rectangle", "rectangle_logger") {}
Goal Rectangle::createRectangleGoal(double x, double y, double v, double accel, double heading) const {
    return polylineGoal(x, y, v, accel, heading, 0.0);
}

// a1 is stored by the reference class and never used (Reciprocating.cpp:13-19, :98).
Reciprocating::Reciprocating(double alt, Eigen::Vector3d A, Eigen::Vector3d B, std::vector<double> v_goals, double a1,
                             double a3, double t_traj, double dt)
    : GpuTrajectory(polyParams(TGX_RECIPROCATING, dt, alt, t_traj, v_goals, a3, 0.0,
                               {A.x(), A.y(), A.z(), B.x(), B.y(), B.z()}),
                    "Reciprocating", "reciprocating_logger") {
    (void)a1;
}
Goal Reciprocating::createReciprocatingGoal(double x, double y, double v, double accel, double heading) const {
    return polylineGoal(x, y, v, accel, heading, 0.0);
}

Bounce::Bounce(double cx, double cy, double Az, double Bz, std::vector<double> v_goals, double t_traj,
               double orientation, double dt)
    : GpuTrajectory(polyParams(TGX_BOUNCE, dt, 0.0, t_traj, v_goals, 0.0, orientation, {cx, cy, Az, Bz}), "Bounce",
                    "bounce_logger") {}
Goal Bounce::createBounceGoal(double x, double y, double z, double vz, double heading) const {
    return polylineGoal(x, y, vz, 0.0, heading, z);
}

M::M(double cx, double cy, double length, double width, double alt, std::vector<double> v_goals, double t_traj,
     double orientation, double dt)
    : GpuTrajectory(polyParams(TGX_M, dt, alt, t_traj, v_goals, 1.0, orientation, {cx, cy, length, width}), "M",
                    "m_logger") {}
Goal M::createMGoal(double x, double y, double v, double accel, double heading) const {
    return polylineGoal(x, y, v, accel, heading, 0.0);
}

I::I(double cx, double cy, double length, double width, double alt, std::vector<double> v_goals, double t_traj,
     double orientation, double dt)
    : GpuTrajectory(polyParams(TGX_I, dt, alt, t_traj, v_goals, 1.0, orientation, {cx, cy, length, width}), "I",
                    "i_logger") {}
Goal I::createIGoal(double x, double y, double v, double accel, double heading) const {
    return polylineGoal(x, y, v, accel, heading, 0.0);
}

T::T(double cx, double cy, double length, double width, double alt, std::vector<double> v_goals, double t_traj,
     double orientation, double dt)
    : GpuTrajectory(polyParams(TGX_T, dt, alt, t_traj, v_goals, 1.0, orientation, {cx, cy, length, width}), "T",
                    "t_logger") {}
Goal T::createTGoal(double x, double y, double v, double accel, double heading) const {
    return polylineGoal(x, y, v, accel, heading, 0.0);
}

}  // namespace TGX_DROPIN_NAMESPACE
