// plan.cu — planning kernels of libtgx: one thread per trajectory replays the reference's scalar recurrences.
//
// What is replayed, and why it must be a replay: the number of samples a reference trajectory has is decided
// by floating-point accumulation, not by a formula —
//     while (v < v_goal) v = std::min(v + accel_*dt_, v_goal);          Circle.cpp:47-49, Line.cpp:46-48, Figure8.cpp:47-49
//     while (current_t_traj_ < t_traj_) ... current_t_traj_ += dt_;     Circle.cpp:63-71, Line.cpp:57-62, Figure8.cpp:63-71
//     while (v > 0) v = std::max(v - accel_*dt_, 0.0);                  Circle.cpp:75-77, Line.cpp:65-66, Figure8.cpp:75-77
// (t_traj = 10, dt = 0.01 gives 1001 hold samples, not 1000.)  These kernels execute the same IEEE-754 double
// operations in the same order with the round-to-nearest intrinsics (__dadd_rn, __dmul_rn, __ddiv_rn,
// __dsqrt_rn), which nvcc never contracts into FMAs, so counts, phase boundaries and the (v, theta | x, y)
// state at every segment base are bit-identical to the reference's CPU result.  This file is additionally
// compiled with -fmad=false.
//
// Output of a plan: TrajRec / Seg / Tile tables (tgx_internal.cuh) that eval.cu consumes, plus the caller-visible
// counts, status bits and index_msgs (tgx_phases).
#include "tgx_internal.cuh"

namespace tgx {

namespace {

__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// libstdc++ std::min(a, b) / std::max(a, b)
__device__ __forceinline__ double std_min(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double std_max(double a, double b) { return (a < b) ? b : a; }

__device__ __forceinline__ bool finite_pos(double x) { return isfinite(x) && x > 0.0; }

// Same acceptance rule as the node-side validation (TrajectoryGenerator.cpp:184-195, 268-277) plus the
// conditions under which the reference's loops cannot terminate (dt <= 0, r <= 0, non-finite input).
__device__ bool params_ok(const tgx_params& p) {
    if (!finite_pos(p.dt) || !isfinite(p.alt)) return false;
    if (p.type == TGX_CIRCLE || p.type == TGX_FIGURE8) {
        const tgx_orbit_params& o = p.u.orbit;
        if (p.n_vgoals < 1 || p.n_vgoals > TGX_MAX_VGOALS) return false;
        if (!finite_pos(o.r) || !finite_pos(o.accel)) return false;
        if (!isfinite(o.cx) || !isfinite(o.cy) || !isfinite(o.t_traj)) return false;
        for (int i = 0; i < p.n_vgoals; ++i)
            if (!finite_pos(o.v_goals[i])) return false;
        return true;
    }
    if (p.type == TGX_LINE) {
        const tgx_line_params& l = p.u.line;
        for (int i = 0; i < 3; ++i)
            if (!isfinite(l.A[i]) || !isfinite(l.B[i])) return false;
        return finite_pos(l.v_goal) && finite_pos(l.a1) && finite_pos(l.a3);
    }
    return false;
}

// Collects what a replay produces.  In counting mode (segs == nullptr) it only counts.
struct Emitter {
    int tile_shift;
    int32_t traj;
    Seg* segs;            // this trajectory's slice of Seg[]   (nullptr: count only)
    Tile* tiles;          // this trajectory's slice of Tile[]
    int32_t seg_base;     // absolute index of segs[0]
    tgx_phases* ph;       // may be nullptr
    bool orbit;           // Circle / Figure8: Seg.s1 = theta increment per step, Seg.acc = exact theta of the last sample

    int nseg = 0;
    int ntile = 0;
    int nph = 0;
    int cur_tile = -1;
    int tile_seg_begin = 0;
    Seg cur;              // the open segment, kept in registers until closed

    __device__ void phase(int key, int kind, double value, double value2) {
        if (ph && nph < TGX_MAX_PHASES) {
            ph->key[nph] = key;
            ph->kind[nph] = kind;
            ph->value[nph] = value;
            ph->value2[nph] = value2;
        }
        ++nph;
    }
    __device__ void flush_tile() {
        if (cur_tile < 0) return;
        if (tiles) {
            Tile t;
            t.traj = traj;
            t.k_lo = cur_tile << tile_shift;
            t.seg_begin = seg_base + tile_seg_begin;
            t.nseg = nseg - tile_seg_begin;
            tiles[ntile] = t;
        }
        ++ntile;
    }
    // Open a segment whose base is sample kb (its first own sample is kb+1).
    __device__ void open(int kb, double vb, double dv, double vclamp, double s0, double s1, double acc) {
        const int t = (kb + 1) >> tile_shift;
        if (t != cur_tile) {
            flush_tile();
            cur_tile = t;
            tile_seg_begin = nseg;
        }
        cur.kb = kb;
        cur.n = 0;
        cur.flags = 0;
        cur.pad = 0;
        cur.vb = vb;
        cur.dv = dv;
        cur.vclamp = vclamp;
        cur.s0 = s0;
        cur.s1 = s1;
        cur.acc = acc;
    }
    // `last_state`: orbit only, the exactly replayed theta at sample k_last.
    __device__ void close(int k_last, bool clamp_last, double last_state) {
        if (segs) {
            cur.n = k_last - cur.kb;
            cur.flags = clamp_last ? kSegClampLast : 0;
            if (orbit) cur.acc = last_state;
            segs[nseg] = cur;
        }
        ++nseg;
    }
    // Optional breaks (exact-progression breaks inside a hold) are only taken while the tile has room left in the
    // evaluation kernel's shared-memory segment table; mandatory breaks (phases, tiles, ramp chunks) always fit.
    __device__ bool can_break() const { return nseg - tile_seg_begin < kMaxOptionalSegPerTile; }
    __device__ void finish() {
        flush_tile();
        cur_tile = -1;
        if (ph) ph->n = nph < TGX_MAX_PHASES ? nph : TGX_MAX_PHASES;
    }
};

// One velocity ramp of the reference (up: std::min clamp at v_goal; down: std::max clamp at 0), shared by all
// three classes.  STEP advances the class-specific state (theta, or x/y) for the new v.
// Returns false when the reference would never terminate / the sample guard is hit.
template <bool UP, class Step>
__device__ __forceinline__ bool ramp(double& v, double target, double adt, double dtr, int& k,
                                     int64_t max_samples, int tmask, Emitter& E, double acc, double& s0,
                                     double& s1, Step step) {
    bool open = false;
    while (UP ? (v < target) : (v > 0.0)) {
        const double vn = UP ? std_min(dadd(v, adt), target) : std_max(dsub(v, adt), 0.0);
        if (vn == v || (int64_t)k + 1 >= max_samples) return false;
        if (!open) {
            // orbit: Seg.s1 = theta increment per step at the base speed, (vb/r)*dt up to rounding
            E.open(k, v, UP ? adt : -adt, UP ? target : 0.0, s0, E.orbit ? dmul(v, dtr) : s1, acc);
            open = true;
        }
        v = vn;
        step(v);
        ++k;
        // k is the last sample of its tile (segments never straddle tiles), or the ramp chunk is full (bounds
        // the rounding drift of the closed form against the reference's running sums)
        if (((k + 1) & tmask) == 0 || k - E.cur.kb >= kRampChunk) {
            E.close(k, v == (UP ? target : 0.0), s0);
            open = false;
        }
    }
    if (open) E.close(k, v == (UP ? target : 0.0), s0);
    return true;
}

// The constant-speed phase: `while (current_t_traj_ < t) { ...; current_t_traj_ += dt_; }`.
//
// EXACT (orbits): with v constant the reference adds the same w = (v/r)*dt to theta every step.  While theta
// stays inside one binade every sum theta + w rounds by the same amount, so the reference's theta is an EXACT
// arithmetic progression with step d = fl(theta + w) - theta.  The segment is cut whenever d changes (a binade
// crossing; the crossing step itself becomes the segment's exactly stored last sample), so the evaluation
// kernel's theta_b + j*d reproduces the reference's running sum bit for bit.
template <bool EXACT, class Step>
__device__ __forceinline__ bool hold(double v, double t_hold, double dt, int& k, int64_t max_samples, int tmask,
                                     Emitter& E, double& s0, double& s1, Step step) {
    bool open = false;
    double cur = 0.0;
    double seg_d = 0.0;
    while (cur < t_hold) {
        if ((int64_t)k + 1 >= max_samples) return false;
        const double s0_old = s0, s1_old = s1;
        step(v);
        const double d = EXACT ? dsub(s0, s0_old) : 0.0;
        if (!open) {
            E.open(k, v, 0.0, v, s0_old, EXACT ? d : s1_old, 0.0);
            seg_d = d;
            open = true;
        }
        ++k;
        const double tn = dadd(cur, dt);
        if (tn == cur) return false;
        cur = tn;
        if (((k + 1) & tmask) == 0 || (EXACT && d != seg_d && E.can_break())) {
            E.close(k, false, s0);
            open = false;
        }
    }
    if (open) E.close(k, false, s0);
    return true;
}

// Circle::generateTraj (Circle.cpp:30-94) == Figure8::generateTraj (Figure8.cpp:30-94).
// STATE = false skips the theta recurrence (counts and status do not depend on it).
template <bool STATE>
__device__ int replay_orbit(const tgx_params& p, int64_t max_samples, Emitter& E, uint32_t& st) {
    const tgx_orbit_params& o = p.u.orbit;
    const double r = o.r, dt = p.dt;
    const double adt = dmul(o.accel, dt);
    const double dtr = ddiv(dt, r);
    const int tmask = (1 << E.tile_shift) - 1;
    double v = 0.0, th = 0.0, unused = 0.0;
    int k = 0;   // index of the last sample produced so far; sample 0 (v = 0, theta = 0) exists (:41)
    double v_cached = -1.0, w_cached = 0.0;
    auto step = [&](double vnew) {                           // omega = v/r_; theta += omega*dt_  (:50-51, :65-67)
        if (STATE) {
            if (vnew != v_cached) {                          // same v gives the same omega*dt: skip the division
                v_cached = vnew;
                w_cached = dmul(ddiv(vnew, r), dt);
            }
            th = dadd(th, w_cached);
        }
    };
    for (int g = 0; g < p.n_vgoals; ++g) {                   // :43
        const double vg = o.v_goals[g];
        E.phase(k, TGX_PH_ACCEL_TO, vg, 0.0);                // :45
        if (!ramp<true>(v, vg, adt, dtr, k, max_samples, tmask, E, 0.0, th, unused, step)) {   // :47-54
            st |= TGX_ST_TOO_LONG;
            return -1;
        }
        if (fabs(dsub(v, vg)) > 0.001) st |= TGX_ST_VGOALS_NOT_INCREASING;                // :57-59
        E.phase(k, TGX_PH_REACHED, vg, o.t_traj);            // :61-62
        if (!hold<STATE>(v, o.t_traj, dt, k, max_samples, tmask, E, th, unused, step)) {   // :63-71
            st |= TGX_ST_TOO_LONG;
            return -1;
        }
    }
    E.phase(k, TGX_PH_DECEL, 0.0, 0.0);                      // :74
    if (!ramp<false>(v, 0.0, adt, dtr, k, max_samples, tmask, E, 0.0, th, unused, step)) {   // :75-82
        st |= TGX_ST_TOO_LONG;
        return -1;
    }
    if (fabs(v) > 0.001) st |= TGX_ST_FINAL_V_NONZERO;       // :85-88
    E.phase(k, TGX_PH_STOPPED, 0.0, 0.0);                    // :89
    return k + 1;
}

// |B - A| with Eigen's reduction order x^2 + (y^2 + z^2) (Line.cpp:157,176) and Line::get_d2 (Line.cpp:175-181).
__device__ double line_d2(const tgx_line_params& l) {
    const double dx = dsub(l.B[0], l.A[0]);
    const double dy = dsub(l.B[1], l.A[1]);
    const double dz = dsub(l.B[2], l.A[2]);
    const double d = __dsqrt_rn(dadd(dmul(dx, dx), dadd(dmul(dy, dy), dmul(dz, dz))));
    const double vg = l.v_goal;
    const double d1 = ddiv(dmul(dmul(0.5, vg), vg), l.a1);
    const double d3 = ddiv(dmul(dmul(0.5, vg), vg), l.a3);
    return dsub(dsub(d, d1), d3);
}

// Line::Line (theta_, Line.cpp:24) + Line::generateTraj (Line.cpp:31-89).
__device__ int replay_line(const tgx_params& p, int64_t max_samples, Emitter& E, uint32_t& st, double& theta,
                           double& c, double& s) {
    const tgx_line_params& l = p.u.line;
    const double dt = p.dt;
    const int tmask = (1 << E.tile_shift) - 1;
    theta = atan2(dsub(l.B[1], l.A[1]), dsub(l.B[0], l.A[0]));   // :24
    sincos(theta, &s, &c);                                       // :93-94 (same value on every call)
    const double cc = c, ss = s;
    double v = 0.0;
    // sample 0: createLineGoal(A.x, A.y, 0, 0, theta)  (:40, :97-98)
    double x = dadd(l.A[0], dmul(dmul(v, cc), dt));
    double y = dadd(l.A[1], dmul(dmul(v, ss), dt));
    int k = 0;
    auto step = [&](double vnew) {                               // p = goals.back().p + v*c*dt  (:49, :97-98)
        x = dadd(x, dmul(dmul(vnew, cc), dt));
        y = dadd(y, dmul(dmul(vnew, ss), dt));
    };
    const double vg = l.v_goal;                                  // :43
    E.phase(k, TGX_PH_ACCEL_TO, vg, 0.0);                        // :44
    if (!ramp<true>(v, vg, dmul(l.a1, dt), 0.0, k, max_samples, tmask, E, l.a1, x, y, step)) {   // :46-50
        st |= TGX_ST_TOO_LONG;
        return -1;
    }
    const double d2 = line_d2(l);
    if (d2 < 0.0) st |= TGX_ST_LINE_D2_NEGATIVE;                 // the condition Line.cpp:165 reports
    const double t2 = ddiv(d2, vg);                              // :53
    E.phase(k, TGX_PH_REACHED, vg, t2);                          // :55-56
    if (!hold<false>(v, t2, dt, k, max_samples, tmask, E, x, y, step)) {                     // :57-62
        st |= TGX_ST_TOO_LONG;
        return -1;
    }
    E.phase(k, TGX_PH_DECEL, 0.0, 0.0);                          // :64
    if (!ramp<false>(v, 0.0, dmul(l.a3, dt), 0.0, k, max_samples, tmask, E, -l.a3, x, y, step)) {   // :65-68
        st |= TGX_ST_TOO_LONG;
        return -1;
    }
    if (fabs(dsub(l.B[0], x)) > 0.05 || fabs(dsub(l.B[1], y)) > 0.05) st |= TGX_ST_LINE_END_NOT_B;   // :71-79
    E.phase(k, TGX_PH_STOPPED, 0.0, 0.0);                        // :84
    return k + 1;
}

// Trajectory::isPointInsideBounds (Trajectory.hpp:50-57).
__device__ bool point_inside(const double* box, double x, double y, double z) {
    if (x < box[0] || x > box[1]) return false;
    if (y < box[2] || y > box[3]) return false;
    if (z < box[4] || z > box[5]) return false;
    return true;
}

// trajectoryInsideBounds: Circle.cpp:171-179, Figure8.cpp:169-177, Line.cpp:154-173.
__device__ bool inside_bounds(const tgx_params& p, const double* box) {
    if (p.type == TGX_LINE) {
        const tgx_line_params& l = p.u.line;
        if (line_d2(l) < 0.0) return false;
        return point_inside(box, l.A[0], l.A[1], l.A[2]) && point_inside(box, l.B[0], l.B[1], l.B[2]);
    }
    const tgx_orbit_params& o = p.u.orbit;
    return point_inside(box, dsub(o.cx, o.r), dsub(o.cy, o.r), p.alt) &&
           point_inside(box, dadd(o.cx, o.r), dadd(o.cy, o.r), p.alt);
}

struct PlanOut {
    int n;            // sample count, 0 if rejected
    uint32_t status;
    int nseg;
    int ntile;
};

// generateTraj plan of one trajectory.  FILL = false: count only; STATE = false: skip the theta replay (then the
// segment count is not meaningful: exact-progression breaks depend on theta).
template <bool FILL, bool STATE>
__device__ PlanOut plan_one(const tgx_params& p, int64_t max_samples, const tgx_limits* lim, Emitter& E,
                            TrajRec* rec) {
    PlanOut r{0, 0u, 0, 0};
    if (!params_ok(p)) {
        r.status = TGX_ST_BAD_PARAM;
    } else {
        int n;
        double theta = 0.0, c = 1.0, s = 0.0;
        if (p.type == TGX_LINE) n = replay_line(p, max_samples, E, r.status, theta, c, s);
        else n = replay_orbit<STATE>(p, max_samples, E, r.status);
        E.finish();
        if (lim && lim->check_box && !inside_bounds(p, lim->box)) r.status |= TGX_ST_OUTSIDE_BOUNDS;
        if (n > 0) {
            r.n = n;
            r.nseg = E.nseg;
            r.ntile = E.ntile;
        }
        if (FILL && rec) {
            TrajRec t;
            t.n = r.n;
            if (p.type == TGX_LINE) {
                t.type = TGX_LINE | kRecForceB;
                t.f[0] = c; t.f[1] = s; t.f[2] = theta; t.f[3] = p.alt; t.f[4] = p.dt;
                t.f[5] = p.u.line.B[0]; t.f[6] = p.u.line.B[1];
            } else {
                const tgx_orbit_params& o = p.u.orbit;
                t.type = p.type;
                t.f[0] = o.r; t.f[1] = o.cx; t.f[2] = o.cy; t.f[3] = p.alt;
                t.f[4] = ddiv(p.dt, o.r); t.f[5] = ddiv(1.0, o.r); t.f[6] = 0.0;
            }
            *rec = t;
        }
    }
    if (FILL && rec && r.n == 0) {
        TrajRec t;
        t.type = p.type & kRecTypeMask;
        t.n = 0;
        for (int i = 0; i < 7; ++i) t.f[i] = 0.0;
        *rec = t;
    }
    if (FILL && E.ph && r.n == 0) E.ph->n = 0;
    return r;
}

// generateStopTraj plan of one trajectory (Circle.cpp:132-169, Line.cpp:117-152, Figure8.cpp:130-167).
// Samples of a braking plan are numbered from 0 = first braking step, so the segment base is sample -1.
template <bool FILL>
__device__ PlanOut stop_one(const tgx_params& p, const double* from, int64_t max_samples, Emitter& E,
                            TrajRec* rec) {
    PlanOut r{0, 0u, 0, 0};
    TrajRec t;
    t.type = p.type & kRecTypeMask;
    t.n = 0;
    for (int i = 0; i < 7; ++i) t.f[i] = 0.0;
    if (!params_ok(p)) {
        r.status = TGX_ST_BAD_PARAM;
    } else {
        const int tmask = (1 << E.tile_shift) - 1;
        const double dt = p.dt;
        // 2D current (goal) vel: sqrt(pow(vx,2) + pow(vy,2))  (Circle.cpp:140-141, Line.cpp:124-125)
        double v = __dsqrt_rn(dadd(dmul(from[TGX_VX], from[TGX_VX]), dmul(from[TGX_VY], from[TGX_VY])));
        int k = -1;
        bool ok = true;
        E.phase(0, TGX_PH_PRESSED_END, 0.0, 0.0);                 // Circle.cpp:148, Line.cpp:132
        if (p.type == TGX_LINE) {
            const tgx_line_params& l = p.u.line;
            const double theta = atan2(from[TGX_VY], from[TGX_VX]);   // Line.cpp:126-127
            double s, c;
            sincos(theta, &s, &c);
            double x = from[TGX_PX], y = from[TGX_PY];
            const double adt = dmul(l.a3, dt);
            auto step = [&](double vnew) {
                x = dadd(x, dmul(dmul(vnew, c), dt));
                y = dadd(y, dmul(dmul(vnew, s), dt));
            };
            // the unconditional first braking step (Line.cpp:133-134), then `while (v > 0)` (:136-140)
            E.open(k, v, -adt, 0.0, x, y, -l.a3);
            v = std_max(dsub(v, adt), 0.0);
            step(v);
            ++k;
            bool open = true;
            if (((k + 1) & tmask) == 0) { E.close(k, v == 0.0, 0.0); open = false; }
            while (v > 0.0) {
                const double vn = std_max(dsub(v, adt), 0.0);
                if (vn == v || (int64_t)k + 1 >= max_samples) { ok = false; break; }
                if (!open) { E.open(k, v, -adt, 0.0, x, y, -l.a3); open = true; }
                v = vn;
                step(v);
                ++k;
                if (((k + 1) & tmask) == 0 || k - E.cur.kb >= kRampChunk) { E.close(k, v == 0.0, 0.0); open = false; }
            }
            if (ok && open) E.close(k, v == 0.0, 0.0);
            t.type = TGX_LINE;
            t.f[0] = c; t.f[1] = s; t.f[2] = theta; t.f[3] = p.alt; t.f[4] = dt;
        } else {
            const tgx_orbit_params& o = p.u.orbit;
            // current (goal) angle wrt the center (Circle.cpp:142-143; Figure8.cpp:140-141 uses the same formula)
            double th = atan2(dsub(from[TGX_PY], o.cy), dsub(from[TGX_PX], o.cx));
            double unused = 0.0;
            const double adt = dmul(o.accel, dt);
            auto step = [&](double vnew) { th = dadd(th, dmul(ddiv(vnew, o.r), dt)); };
            ok = ramp<false>(v, 0.0, adt, ddiv(dt, o.r), k, max_samples, tmask, E, 0.0, th, unused, step);   // :150-157
            t.type = p.type;
            t.f[0] = o.r; t.f[1] = o.cx; t.f[2] = o.cy; t.f[3] = p.alt;
            t.f[4] = ddiv(dt, o.r); t.f[5] = ddiv(1.0, o.r);
        }
        if (!ok) {
            r.status |= TGX_ST_TOO_LONG;
            E.nph = 0;
        } else {
            E.phase(k, TGX_PH_STOPPED, 0.0, 0.0);                 // key = size-1 (-1 when nothing was produced)
            r.n = k + 1;
        }
        E.finish();
        if (r.n > 0) {
            r.nseg = E.nseg;
            r.ntile = E.ntile;
        }
    }
    t.n = r.n;
    if (FILL && rec) *rec = t;
    return r;
}

__device__ __forceinline__ tgx_params load_params(const tgx_params* params, int64_t i) {
    // 128-byte record: eight 16-byte loads through the read-only path.
    tgx_params p;
    const double2* src = reinterpret_cast<const double2*>(params + i);
    double2* dst = reinterpret_cast<double2*>(&p);
#pragma unroll
    for (int q = 0; q < 8; ++q) dst[q] = __ldg(src + q);
    return p;
}

}  // namespace

// ---- kernels ------------------------------------------------------------------------------------------

// Counting pass: N_i, status_i and, with SEGS, the number of segments / tiles the fill pass will emit (which
// requires the full state replay, because exact-progression breaks depend on theta).
template <bool SEGS>
__global__ void __launch_bounds__(128)
plan_count_kernel(const tgx_params* __restrict__ params, const double* __restrict__ stop_from, int64_t n,
                  tgx_limits lim, int has_lim, int64_t max_samples, int tile_shift,
                  int32_t* __restrict__ counts, uint32_t* __restrict__ status, int32_t* __restrict__ nseg,
                  int32_t* __restrict__ ntile) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const tgx_params p = load_params(params, i);
    Emitter E{tile_shift, (int32_t)i, nullptr, nullptr, 0, nullptr, p.type != TGX_LINE};
    PlanOut r;
    if (stop_from) {
        double from[TGX_NCHAN];
#pragma unroll
        for (int c = 0; c < TGX_NCHAN; ++c) from[c] = stop_from[i * TGX_NCHAN + c];
        r = stop_one<false>(p, from, max_samples, E, nullptr);
    } else {
        r = plan_one<false, SEGS>(p, max_samples, has_lim ? &lim : nullptr, E, nullptr);
    }
    if (counts) counts[i] = r.n;
    if (status) status[i] = r.status;
    if (nseg) nseg[i] = r.nseg;
    if (ntile) ntile[i] = r.ntile;
}

// Fill pass: same replay, now writing the tables at the offsets the scans produced.
__global__ void __launch_bounds__(128)
plan_fill_kernel(const tgx_params* __restrict__ params, const double* __restrict__ stop_from, int64_t n,
                 tgx_limits lim, int has_lim, int64_t max_samples, int tile_shift,
                 const int32_t* __restrict__ plan_counts,
                 const int64_t* __restrict__ seg_off, const int64_t* __restrict__ tile_off,
                 TrajRec* __restrict__ recs, Seg* __restrict__ segs, Tile* __restrict__ tiles,
                 int32_t* __restrict__ counts, uint32_t* __restrict__ status, tgx_phases* __restrict__ phases) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const tgx_params p = load_params(params, i);
    // A trajectory the counting pass rejected owns no slice of Seg[] / Tile[]: replay it without writing tables.
    const bool keep = plan_counts[i] > 0;
    Emitter E{tile_shift, (int32_t)i, keep ? segs + seg_off[i] : nullptr, keep ? tiles + tile_off[i] : nullptr,
              (int32_t)seg_off[i], phases ? phases + i : nullptr, p.type != TGX_LINE};
    PlanOut r;
    if (stop_from) {
        double from[TGX_NCHAN];
#pragma unroll
        for (int c = 0; c < TGX_NCHAN; ++c) from[c] = stop_from[i * TGX_NCHAN + c];
        r = stop_one<true>(p, from, max_samples, E, recs + i);
        if (phases && r.status) phases[i].n = 0;
    } else {
        r = plan_one<true, true>(p, max_samples, has_lim ? &lim : nullptr, E, recs + i);
    }
    if (counts) counts[i] = r.n;
    if (status) status[i] = r.status;
}

// ---- host-side launchers (called from engine.cu) -------------------------------------------------------

cudaError_t launch_plan_count(const tgx_params* params, const double* stop_from, int64_t n, const tgx_limits* lim,
                              int64_t max_samples, int tile_shift, int32_t* counts, uint32_t* status,
                              int32_t* nseg, int32_t* ntile, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    tgx_limits l{};
    if (lim) l = *lim;
    const int threads = 128;
    const int64_t blocks = (n + threads - 1) / threads;
    if (nseg || ntile)
        plan_count_kernel<true><<<(unsigned)blocks, threads, 0, stream>>>(params, stop_from, n, l, lim ? 1 : 0,
                                                                         max_samples, tile_shift, counts, status,
                                                                         nseg, ntile);
    else
        plan_count_kernel<false><<<(unsigned)blocks, threads, 0, stream>>>(params, stop_from, n, l, lim ? 1 : 0,
                                                                          max_samples, tile_shift, counts, status,
                                                                          nseg, ntile);
    return cudaGetLastError();
}

cudaError_t launch_plan_fill(const tgx_params* params, const double* stop_from, int64_t n, const tgx_limits* lim,
                             int64_t max_samples, int tile_shift, const int32_t* plan_counts,
                             const int64_t* seg_off, const int64_t* tile_off, TrajRec* recs, Seg* segs, Tile* tiles,
                             int32_t* counts, uint32_t* status, tgx_phases* phases, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    tgx_limits l{};
    if (lim) l = *lim;
    const int threads = 128;
    const int64_t blocks = (n + threads - 1) / threads;
    plan_fill_kernel<<<(unsigned)blocks, threads, 0, stream>>>(params, stop_from, n, l, lim ? 1 : 0, max_samples,
                                                              tile_shift, plan_counts, seg_off, tile_off, recs, segs,
                                                              tiles, counts, status, phases);
    return cudaGetLastError();
}

}  // namespace tgx
