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
#include "replay_common.cuh"

namespace tgx {

namespace {

// Where a trajectory's parameter record sits in the batch: goal speeds beyond the eighth live in the continuation
// records that follow it (tgx.h: TGX_VGOALS_MORE).
struct Src {
    const tgx_params* params;
    int64_t i, n;
};

__device__ __forceinline__ double orbit_goal(const tgx_params& p, const Src& src, int g) {
    if (g < TGX_MAX_VGOALS) return p.u.orbit.v_goals[g];
    return __ldg(&src.params[src.i + (g >> 3)].u.orbit.v_goals[g & 7]);
}

// Same acceptance rule as the node-side validation (TrajectoryGenerator.cpp:184-195, 268-277) plus the
// conditions under which the reference's loops cannot terminate (dt <= 0, r == 0, non-finite input).
// check_goals = false: the per-time evaluation helpers (create*Goal) do not involve the goal speeds.
__device__ bool params_ok(const tgx_params& p, const Src& src, bool check_goals = true) {
    if (!finite_pos(p.dt) || !isfinite(p.alt)) return false;
    if (p.type == TGX_CIRCLE || p.type == TGX_FIGURE8) {
        const tgx_orbit_params& o = p.u.orbit;
        if (p.n_vgoals < 0 || p.n_vgoals > TGX_MAX_VGOALS_TOTAL) return false;
        // a negative radius is legal in the reference (omega = v / r_ < 0: the mirrored circle)
        if (!isfinite(o.r) || o.r == 0.0 || !finite_pos(o.accel)) return false;
        if (!isfinite(o.cx) || !isfinite(o.cy) || !isfinite(o.t_traj)) return false;
        if (!check_goals) return true;
        const int more = TGX_ORBIT_RECORDS(p.n_vgoals) - 1;
        if (src.i + more >= src.n) return false;
        for (int q = 1; q <= more; ++q)
            if (__ldg(&src.params[src.i + q].type) != TGX_VGOALS_MORE) return false;
        for (int g = 0; g < p.n_vgoals; ++g)
            if (!finite_pos(orbit_goal(p, src, g))) return false;
        return true;
    }
    if (p.type == TGX_LINE || p.type == TGX_BOOMERANG) {
        const tgx_line_params& l = p.u.line;
        for (int i = 0; i < 3; ++i)
            if (!isfinite(l.A[i]) || !isfinite(l.B[i])) return false;
        return finite_pos(l.v_goal) && finite_pos(l.a1) && finite_pos(l.a3);
    }
    return false;
}

// A continuation record is valid when it belongs to an orbit record with enough goal speeds to reach it.
__device__ bool continuation_has_owner(const Src& src) {
    for (int q = 1; q < TGX_MAX_VGOALS_TOTAL / TGX_MAX_VGOALS && src.i - q >= 0; ++q) {
        const int2 head = __ldg(reinterpret_cast<const int2*>(src.params + (src.i - q)));      // {type, n_vgoals}
        if (head.x == TGX_VGOALS_MORE) continue;
        return (head.x == TGX_CIRCLE || head.x == TGX_FIGURE8) && TGX_ORBIT_RECORDS(head.y) > q;
    }
    return false;
}

__device__ __forceinline__ bool is_line_like(int type) { return type == TGX_LINE || type == TGX_BOOMERANG; }
__device__ __forceinline__ bool is_orbit(int type) { return type == TGX_CIRCLE || type == TGX_FIGURE8; }
// Trajectories of one class take the same path through the replay: orbits with the same number of speed goals
// (1 .. TGX_MAX_VGOALS), lines, boomerangs, everything else.  < 32 classes: also a bit index in PlanStats.kinds.
__device__ __forceinline__ int replay_class(int type, int n_vgoals) {
    if (is_orbit(type)) return n_vgoals >= 1 && n_vgoals <= 16 ? n_vgoals : 0;
    return type == TGX_LINE ? 17 : (type == TGX_BOOMERANG ? 18 : 19);
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
    int seg_cap;          // slab mode: capacity of this trajectory's slice (writes beyond it are dropped and the
    int tile_cap;         //            plan is redone with exact offsets); otherwise INT_MAX
    // phase plans (PhaseRec): where each segment ends, the replayed state (orbit: the angle at the segment's end; line:
    // the position at its base), and what kind of segment it is
    int32_t* keys = nullptr;
    double* states = nullptr;
    uint64_t kinds = 0;        // orbit: 2 bits per segment, kPhaseKind*; line: 3 bits, kPhaseLine*
    int goal = 0;              // index of the speed goal the current ramp-up heads for
    int prev_kind = -1;
    bool ramp_split = false;   // some ramp was cut into several segments: not expressible as a PhaseRec

    int nseg = 0;
    int ntile = 0;
    int nph = 0;
    int max_tile_segs = 0;     // largest segment list of a tile (the evaluation kernel stages at most kMaxSegPerTile)
    int max_seg_len = 0;       // longest segment
    int cur_tile = -1;
    int tile_seg_begin = 0;
    Seg cur;              // the open segment, kept in registers until closed

    // Lanes of a warp that replay trajectories of the same class (same type, same number of speed goals) meet at a warp
    // barrier after every phase.  Without it a lane that leaves a ramp loop early runs ahead into the next phase on its own
    // — the compiler reconverges such loops (several exits) only at the end of the replay — and the warp executes every
    // phase once per straggler: ncu showed 9 of 32 lanes active per executed instruction.  `grp` is formed where the
    // lanes enter the replay (plan_one); every lane of it passes the same barriers, failed ones included.
    unsigned grp = 0;
    __device__ __forceinline__ void converge() const {
#ifndef TGX_NO_REPLAY_BARRIERS
        if (grp) __syncwarp(grp);
#endif
    }
    int ph_blocks = 1;         // tgx_phases rows this trajectory owns (its own and its continuation records')
    // entry nph goes to slot nph % 18 of row nph / 18 (tgx.h: TGX_VGOALS_MORE)
    __device__ void phase(int key, int kind, double value, double value2) {
        const int b = nph / TGX_MAX_PHASES, q = nph - b * TGX_MAX_PHASES;
        if (ph && b < ph_blocks) {
            ph[b].key[q] = key;
            ph[b].kind[q] = kind;
            ph[b].value[q] = value;
            ph[b].value2[q] = value2;
        }
        ++nph;
    }
    __device__ void set_phase_counts(int total) {
        if (!ph) return;
        for (int b = 0; b < ph_blocks; ++b) {
            const int left = total - b * TGX_MAX_PHASES;
            ph[b].n = left < 0 ? 0 : (left < TGX_MAX_PHASES ? left : TGX_MAX_PHASES);
        }
    }
    // The tile cur_tile is served by the segments tile_seg_begin .. seg_end - 1.
    __device__ void flush_tile(int seg_end) {
        if (cur_tile < 0) return;
        if (tiles && ntile < tile_cap) {
            Tile t;
            t.traj = traj;
            t.k_lo = cur_tile << tile_shift;
            t.seg_begin = seg_base + tile_seg_begin;
            t.nseg = seg_end - tile_seg_begin;
            tiles[ntile] = t;
        }
        if (seg_end - tile_seg_begin > max_tile_segs) max_tile_segs = seg_end - tile_seg_begin;
        ++ntile;
    }
    // Open a segment whose base is sample kb (its first own sample is kb+1).  Segments are cut where the REPLAY says
    // (phase boundaries, every kRebase / kRampChunk steps, exact-progression breaks), never at tile ends: a segment
    // may span several tiles, and a tile lists every segment that intersects it, so the samples do not depend on the
    // tile size.
    __device__ void open(int kb, double vb, double dv, double vclamp, double s0, double s1, double acc) {
        const int t = (kb + 1) >> tile_shift;
        if (t != cur_tile) {
            flush_tile(nseg);
            cur_tile = t;
            tile_seg_begin = nseg;
        }
        if (orbit) {
            const int kind = dv > 0.0 ? (goal & 1) : (dv == 0.0 ? kPhaseKindHold : kPhaseKindDown);
            if (kind != kPhaseKindHold && kind == prev_kind) ramp_split = true;
            prev_kind = kind;
            if (nseg < kPhaseMaxSegs) kinds |= (uint64_t)kind << (2 * nseg);
        } else {
            const int kind = dv > 0.0 ? kPhaseLineUp : (dv == 0.0 ? kPhaseLineHold : kPhaseLineDown);
            if (kind != kPhaseLineHold && kind == prev_kind) ramp_split = true;
            prev_kind = kind;
            if (nseg < kPhaseLineMaxSegs) {
                kinds |= (uint64_t)kind << (3 * nseg);
                if (states) {
                    states[2 * nseg] = s0;
                    states[2 * nseg + 1] = s1;
                }
            }
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
    __device__ void close(int k_last, bool clamp_last, double last_state, int extra_flags = 0) {
        if (segs && nseg < seg_cap) {
            cur.n = k_last - cur.kb;
            cur.flags = (clamp_last ? kSegClampLast : 0) | extra_flags;
            if (orbit) cur.acc = last_state;
            segs[nseg] = cur;
        }
        if (keys && nseg < kPhaseMaxSegs) {
            keys[nseg] = k_last;
            if (orbit) states[nseg] = last_state;
        }
        // (a forced end point was opened as a hold, dv = 0: kPhaseLineHold | 2 == kPhaseLineForced)
        if (!orbit && (extra_flags & kSegForcePos) && nseg < kPhaseLineMaxSegs) kinds |= (uint64_t)2 << (3 * nseg);
        if (k_last - cur.kb > max_seg_len) max_seg_len = k_last - cur.kb;
        // every further tile the segment reaches into starts its list with this segment
        const int t_last = k_last >> tile_shift;
        while (cur_tile < t_last) {
            flush_tile(nseg + 1);
            ++cur_tile;
            tile_seg_begin = nseg;
        }
        ++nseg;
    }
    // Optional breaks (exact-progression breaks inside a hold) are only taken while the tile has room left in the
    // evaluation kernel's shared-memory segment table; mandatory breaks (phases, tiles, ramp chunks) always fit.
    __device__ bool can_break() const { return nseg - tile_seg_begin < kMaxOptionalSegPerTile; }
    __device__ void finish() {
        flush_tile(nseg);
        cur_tile = -1;
        set_phase_counts(nph);
    }
};

// One velocity ramp of the reference (up: std::min clamp at v_goal; down: std::max clamp at 0), shared by all
// three classes.  Returns false when the reference would never terminate / the sample guard is hit.
//
// XR (exact ramps): every step is replayed; STEP advances the class state (theta, or x/y) with the reference's own
// operation sequence, so the state at every segment base is bit-identical to the reference's.
//
// !XR (fast ramps): v <- fl(v + a*dt) is itself an exact arithmetic progression inside a binade of v, so the ramp
// is advanced in exact jumps (the step COUNT and every v_k stay bit-identical to the reference), and the class
// state is advanced in closed form over each jump, s += c * sum(v): it then differs from the reference's running
// sum by that sum's own accumulated rounding, at most a few hundred half-ulps of theta (~1e-13 rad).
//
// v is the speed MAGNITUDE; sgn = -1 replays a Boomerang's return leg, whose speeds are the exact negatives
// (v = max(v - a*dt, -v_goal), Boomerang.cpp:97-101: round-to-nearest is symmetric).  The segment records, STEP and
// the coefficients c0 / c1 carry the sign.  force_xy (lines): the step on which the clamp fires ends a leg; the
// reference then overwrites that sample's position (Line.cpp:81-82), so it becomes a one-sample segment of its own
// holding the forced position.
template <bool UP, bool XR, bool TRACK0, bool TRACK1, class Step>
__device__ __forceinline__ bool ramp(double& v, double target, double adt, double dtr, int& k,
                                     int64_t max_samples, int tmask, Emitter& E, double acc, double& s0,
                                     double& s1, double c0, double c1, Step step, double sgn = 1.0,
                                     const double* force_xy = nullptr) {
    bool open = false;
    const double clampv = UP ? target : 0.0;
    // one division per ramp instead of two per run: the run-length estimates below multiply by 1 / (a*dt) and are then
    // settled by exact checks (see regular_run)
    const double inv_adt = XR ? 0.0 : ddiv(1.0, adt);
    while (UP ? (v < target) : (v > 0.0)) {
        const double vn = UP ? std_min(dadd(v, adt), target) : std_max(dsub(v, adt), 0.0);
        if (vn == v || (int64_t)k + 1 >= max_samples) return false;
        if (force_xy && vn == clampv) {
            // the leg's last step: its own segment, position forced
            if (open) {
                E.close(k, false, s0);
                open = false;
            }
            E.open(k, sgn * vn, 0.0, sgn * vn, force_xy[0], force_xy[1], acc);
            v = vn;
            if (XR) {
                step(sgn * v);
            } else {
                if (TRACK0) s0 = fma(c0, v, s0);
                if (TRACK1) s1 = fma(c1, v, s1);
            }
            ++k;
            E.close(k, false, s0, kSegForcePos);
            break;
        }
        if (!open) {
            // orbit: Seg.s1 = theta increment per step at the base speed, (vb/r)*dt up to rounding
            E.open(k, sgn * v, sgn * (UP ? adt : -adt), sgn * clampv, s0, E.orbit ? dmul(v, dtr) : s1, acc);
            open = true;
        }
        if (XR) {
            v = vn;
            step(sgn * v);
            ++k;
            // the ramp chunk is full (bounds the rounding drift of the closed form against the reference's running sums)
            if (k - E.cur.kb >= kRampChunk) {
                E.close(k, v == clampv, s0);
                open = false;
            }
        } else {
            const double inc = dsub(vn, v);                      // exact
            long long J = (vn == clampv) ? 0 : regular_run(v, vn, UP ? adt : -adt, inv_adt);
            J = min(J, (long long)(kRebase - (k + 1 - E.cur.kb)));  // a segment holds at most kRebase steps
            J = min(J, max_samples - 2 - (long long)k);            // the guard fires on the next real step
            if (J > 0) {
                // none of the jumped steps may reach the clamp: vn + J*inc strictly between 0 and target
                const double room = UP ? dsub(target, vn) : vn;
                const double q = ceil(dmul(room, inv_adt)) - 1.0;        // an estimate: the loop below is the exact test
                const long long Jc = q < 1.0 ? 0 : (q > 1e15 ? (1LL << 40) : (long long)q);
                J = min(J, Jc);
                while (J > 0) {
                    const double y = fma((double)J, inc, vn);
                    if (UP ? (y < target) : (y > 0.0)) break;
                    --J;
                }
            }
            if (J < 0) J = 0;
            const double fJ = (double)J;
            if (TRACK0 || TRACK1) {
                // sum of v over the real step and the J jumped ones: (J+1)*vn + inc*J(J+1)/2
                const double S = fma(inc, 0.5 * (fJ * (fJ + 1.0)), (fJ + 1.0) * vn);
                if (TRACK0) s0 = fma(c0, S, s0);
                if (TRACK1) s1 = fma(c1, S, s1);
            }
            v = fma(fJ, inc, vn);                                // exact
            k += 1 + (int)J;
            if (k - E.cur.kb >= kRebase) {
                E.close(k, v == clampv, s0);
                open = false;
            }
        }
    }
    if (open) E.close(k, v == clampv, s0);
    return true;
}

__global__ void build_cur_table_kernel(const tgx_params* __restrict__ params, int64_t max_samples,
                                       CurTable* __restrict__ tab) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const double dt = params[0].dt;
    tab->dt = dt;
    tab->stagnates = 0;
    int n = 0;
    long long m = 0;
    double cur = 0.0;
    if (isfinite(dt) && dt > 0.0) {
        while (n < kCurTableMax && m < max_samples) {
            double inc;
            long long cnt;
            if (!cur_run(cur, dt, inc, cnt)) {
                tab->stagnates = 1;
                break;
            }
            tab->m0[n] = m;
            tab->c0[n] = cur;
            tab->inc[n] = inc;
            tab->cnt[n] = cnt;
            ++n;
            cur = fma((double)cnt, inc, cur);     // exact
            m += cnt;
        }
    }
    tab->n = n;
    tab->m_end = m;
}

// The constant-speed phase: `while (current_t_traj_ < t) { ...; current_t_traj_ += dt_; }`
// (Circle.cpp:63-71, Line.cpp:57-62, Figure8.cpp:63-71).  a0 / a1 are the constants the reference adds to
// s0 / s1 on every step of the phase (orbit: s0 = theta, a0 = (v/r)*dt; line: s0 = x, s1 = y, a = (v*c)*dt, (v*s)*dt).
//
// The phase is NOT replayed step by step: its length comes from hold_steps(), and the running sums s0 / s1 are
// advanced in exact jumps over the steps for which regular_run() proves their increments constant, bounded by the
// tile end.  A phase of n steps costs O(number of binade crossings + number of tiles) iterations and still yields
// the reference's bit-exact state.
//
// EXACT (orbits): the segment is cut whenever theta's effective increment d changes (a binade crossing; the
// crossing step itself becomes the segment's exactly stored last sample), so the evaluation kernel's
// theta_b + j*d reproduces the reference's running sum bit for bit.
template <bool TRACK0, bool TRACK1, bool EXACT>
__device__ __forceinline__ bool hold(double v, double t_hold, double dt, int& k, int64_t max_samples, int tmask,
                                     Emitter& E, double& s0, double& s1, double a0, double a1,
                                     const CurTable* __restrict__ tab) {
    long long rem = hold_steps(t_hold, dt, (long long)max_samples - 1 - (long long)k, tab);
    if (rem < 0) return false;
    bool open = false;
    double seg_d = 0.0;
    // reciprocals of the per-step addends for regular_run's run-length estimates (0: not usable, divide)
    const double inv0 = (TRACK0 && rem > 1 && fabs(a0) > 0x1p-500 && fabs(a0) < 0x1p500) ? ddiv(1.0, fabs(a0)) : 0.0;
    const double inv1 = (TRACK1 && rem > 1 && fabs(a1) > 0x1p-500 && fabs(a1) < 0x1p500) ? ddiv(1.0, fabs(a1)) : 0.0;
    while (rem > 0) {
        // ---- one real step ----------------------------------------------------------------------------
        const double s0n = TRACK0 ? dadd(s0, a0) : s0;
        const double s1n = TRACK1 ? dadd(s1, a1) : s1;
        const double d0 = TRACK0 ? dsub(s0n, s0) : 0.0;      // exact effective increments
        const double d1 = TRACK1 ? dsub(s1n, s1) : 0.0;
        if (!open) {
            // Seg.s1: orbit = theta increment per step (the exact progression step when EXACT, else the nominal
            // (v/r)*dt); line = y at the base sample
            E.open(k, v, 0.0, v, s0, EXACT ? d0 : (E.orbit ? a0 : s1), 0.0);
            seg_d = d0;
            open = true;
        }
        long long J = rem - 1;
        if (TRACK0) J = min(J, regular_run(s0, s0n, a0, inv0));
        if (TRACK1) J = min(J, regular_run(s1, s1n, a1, inv1));
        s0 = s0n;
        s1 = s1n;
        ++k;
        --rem;
        if (k - E.cur.kb >= kRebase || (EXACT && d0 != seg_d && E.can_break())) {
            E.close(k, false, s0);
            open = false;
            continue;
        }
        // ---- exact jump over the regular run, inside the segment ---------------------------------------
        J = min(J, (long long)(kRebase - (k - E.cur.kb)));
        if (J > 0) {
            const double fJ = (double)J;
            if (TRACK0) s0 = fma(fJ, d0, s0);                // exact: every partial sum is representable
            if (TRACK1) s1 = fma(fJ, d1, s1);
            k += (int)J;
            rem -= J;
            if (k - E.cur.kb >= kRebase) {
                E.close(k, false, s0);
                open = false;
            }
        }
    }
    if (open) E.close(k, false, s0);
    return true;
}

// Division by a loop-invariant divisor, correctly rounded (== __ddiv_rn) with the reciprocal hoisted out of the
// ramp loops: y = RN(1/b); q0 = RN(a*y); two residual corrections q <- RN(q + RN(a - b*q)*y) (Markstein's
// sequence: the first correction makes q faithful, the second correctly rounded, provided b's significand is not
// all ones).  Outside a comfortable exponent range, or for that one significand, it falls back to __ddiv_rn.
struct InvDiv {
    double b, y;
    bool fast;
};

__device__ __forceinline__ InvDiv make_invdiv(double b) {
    InvDiv d;
    d.b = b;
    d.y = __drcp_rn(b);
    const long long bits = __double_as_longlong(b);
    const bool all_ones = (bits & 0x000fffffffffffffLL) == 0x000fffffffffffffLL;
    const double ab = fabs(b);
    d.fast = !all_ones && ab > 0x1p-200 && ab < 0x1p200;
    return d;
}

__device__ __forceinline__ double div_inv(double a, const InvDiv& d) {
    const double aa = fabs(a);
    if (d.fast && aa > 0x1p-200 && aa < 0x1p200) {
        double q = dmul(a, d.y);
        double r = fma(-q, d.b, a);
        q = fma(r, d.y, q);
        r = fma(-q, d.b, a);
        return fma(r, d.y, q);
    }
    return ddiv(a, d.b);
}

// Circle::generateTraj (Circle.cpp:30-94) == Figure8::generateTraj (Figure8.cpp:30-94).
// STATE = false skips the theta recurrence (counts and status do not depend on it); XR selects exact ramps.
template <bool STATE, bool XR>
__device__ int replay_orbit(const tgx_params& p, const Src& src, int64_t max_samples, Emitter& E, uint32_t& st,
                            const CurTable* __restrict__ tab) {
    const tgx_orbit_params& o = p.u.orbit;
    const double r = o.r, dt = p.dt;
    const double adt = dmul(o.accel, dt);
    const double dtr = ddiv(dt, r);
    const int tmask = (1 << E.tile_shift) - 1;
    double v = 0.0, th = 0.0, unused = 0.0;
    int k = 0;   // index of the last sample produced so far; sample 0 (v = 0, theta = 0) exists (:41)
    const InvDiv rdiv = make_invdiv(r);
    auto step = [&](double vnew) {                           // omega = v/r_; theta += omega*dt_  (:50-51, :79)
        if (STATE) th = dadd(th, dmul(div_inv(vnew, rdiv), dt));
    };
    // (a phase that fails — the sample guard, a speed that does not move — ends the replay, but the lane still walks
    //  through the remaining barriers so that its group stays in step)
    bool ok = true;
    for (int g = 0; g < p.n_vgoals; ++g) {                   // :43
        const double vg = orbit_goal(p, src, g);
        if (ok) {
            E.phase(k, TGX_PH_ACCEL_TO, vg, 0.0);            // :45
            E.goal = g;
            ok = ramp<true, XR, STATE, false>(v, vg, adt, dtr, k, max_samples, tmask, E, 0.0, th, unused, dtr, 0.0,
                                              step);                                            // :47-54
        }
        E.converge();
        if (ok) {
            if (fabs(dsub(v, vg)) > 0.001) st |= TGX_ST_VGOALS_NOT_INCREASING;            // :57-59
            E.phase(k, TGX_PH_REACHED, vg, o.t_traj);        // :61-62
            const double w = STATE ? dmul(div_inv(v, rdiv), dt) : 0.0;    // omega*dt_, the same on every step (:65-67)
            // holds are cut at every binade crossing of theta in BOTH planning modes: inside a segment the reference's
            // running sum is an exact arithmetic progression, so nothing drifts inside a hold however long it is (the
            // hold samples are the reference's running sum from the hold's first angle: bit for bit with exact ramps,
            // and off by the preceding ramps' closed-form rounding, ~1e-13 rad, with fast ones)
            ok = hold<STATE, false, STATE>(v, o.t_traj, dt, k, max_samples, tmask, E, th, unused, w, 0.0, tab);   // :63-71
        }
        E.converge();
    }
    if (ok) {
        E.phase(k, TGX_PH_DECEL, 0.0, 0.0);                  // :74
        ok = ramp<false, XR, STATE, false>(v, 0.0, adt, dtr, k, max_samples, tmask, E, 0.0, th, unused, dtr, 0.0,
                                           step);                                               // :75-82
    }
    E.converge();
    if (!ok) {
        st |= TGX_ST_TOO_LONG;
        return -1;
    }
    if (fabs(v) > 0.001) st |= TGX_ST_FINAL_V_NONZERO;       // :85-88
    E.phase(k, TGX_PH_STOPPED, 0.0, 0.0);                    // :89
    if (k == 0) {
        // no step at all (an empty v_goals vector): the trajectory is the start sample alone (:41), which needs a
        // segment of its own
        E.open(0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0);
        E.close(0, false, 0.0);
    }
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

// Line::Line (theta_, Line.cpp:24) + Line::generateTraj (Line.cpp:31-89); with `boomerang` the same profile is flown
// back from B to A with negative speeds (Boomerang::generateTraj, Boomerang.cpp:31-141).
template <bool XR>
__device__ int replay_line(const tgx_params& p, int64_t max_samples, Emitter& E, uint32_t& st, double& theta,
                           double& c, double& s, const CurTable* __restrict__ tab, bool boomerang) {
    const tgx_line_params& l = p.u.line;
    const double dt = p.dt;
    const int tmask = (1 << E.tile_shift) - 1;
    theta = atan2(dsub(l.B[1], l.A[1]), dsub(l.B[0], l.A[0]));   // :24
    sincos(theta, &s, &c);                                       // :93-94 (same value on every call)
    const double cc = c, ss = s;
    const double cdt = dmul(cc, dt), sdt = dmul(ss, dt);
    const double vg = l.v_goal;                                  // :43
    const double t2 = ddiv(line_d2(l), vg);                      // :53 (a negative d2 simply skips the cruise loop)
    double x = 0.0, y = 0.0;
    auto step = [&](double vnew) {                               // p = goals.back().p + v*c*dt  (:49, :97-98)
        x = dadd(x, dmul(dmul(vnew, cc), dt));
        y = dadd(y, dmul(dmul(vnew, ss), dt));
    };
    int k = 0;
    bool ok = true;                                              // (see replay_orbit: failed lanes keep to the barriers)
    for (int leg = 0; leg < (boomerang ? 2 : 1); ++leg) {
        const double sgn = leg == 0 ? 1.0 : -1.0;
        const double* from = leg == 0 ? l.A : l.B;
        const double* to = leg == 0 ? l.B : l.A;
        double v = 0.0;
        if (ok) {
            // the leg's first sample: createLineGoal(from.x, from.y, +-0, 0, theta)  (:40, Boomerang.cpp:90)
            x = dadd(from[0], dmul(dmul(sgn * v, cc), dt));
            y = dadd(from[1], dmul(dmul(sgn * v, ss), dt));
            if (leg == 1) {
                // sample 0 of the plan is served by the first ramp segment (j = 0); the return leg's first sample
                // needs a segment of its own, one step of speed -0 from B
                if ((int64_t)k + 1 >= max_samples) {
                    ok = false;
                } else {
                    E.open(k, sgn * v, 0.0, sgn * v, from[0], from[1], 0.0);
                    ++k;
                    E.close(k, false, 0.0);
                }
            }
        }
        if (ok) {
            E.phase(k, TGX_PH_ACCEL_TO, vg, 0.0);                // :44
            ok = ramp<true, XR, true, true>(v, vg, dmul(l.a1, dt), 0.0, k, max_samples, tmask, E, sgn * l.a1, x, y,
                                            sgn * cdt, sgn * sdt, step, sgn);                  // :46-50
        }
        E.converge();
        if (ok) {
            E.phase(k, TGX_PH_REACHED, vg, t2);                  // :55-56
            ok = hold<true, true, false>(sgn * v, t2, dt, k, max_samples, tmask, E, x, y, dmul(dmul(sgn * v, cc), dt),
                                         dmul(dmul(sgn * v, ss), dt), tab);                  // :57-62
        }
        E.converge();
        if (ok) {
            E.phase(k, TGX_PH_DECEL, 0.0, 0.0);                  // :64
            ok = ramp<false, XR, true, true>(v, 0.0, dmul(l.a3, dt), 0.0, k, max_samples, tmask, E, -sgn * l.a3, x, y,
                                             sgn * cdt, sgn * sdt, step, sgn, to);             // :65-68, forced :81-82
        }
        E.converge();
        // :71-79 / Boomerang.cpp:126-129 (exit(1) in the reference), checked on the replayed, not the forced, position
        if (ok && (fabs(dsub(to[0], x)) > 0.05 || fabs(dsub(to[1], y)) > 0.05)) st |= TGX_ST_LINE_END_NOT_B;
    }
    if (!ok) {
        st |= TGX_ST_TOO_LONG;
        return -1;
    }
    E.phase(k, TGX_PH_STOPPED, 0.0, 0.0);                        // :84, Boomerang.cpp:134
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
    if (is_line_like(p.type)) {
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
template <bool FILL, bool STATE, bool XR>
__device__ PlanOut plan_one(const tgx_params& p, const Src& src, int64_t max_samples, const tgx_limits* lim, Emitter& E,
                            TrajRec* rec, const CurTable* __restrict__ tab) {
    PlanOut r{0, 0u, 0, 0};
    if (is_orbit(p.type) && p.n_vgoals > TGX_MAX_VGOALS && p.n_vgoals <= TGX_MAX_VGOALS_TOTAL)
        E.ph_blocks = (int)min((int64_t)TGX_ORBIT_RECORDS(p.n_vgoals), src.n - src.i);
    if (TGX_IS_POLYLINE(p.type)) {
        r.status = TGX_ST_WRONG_PLANNER;   // planned by tgx_plan_polyline (polyline.cu)
    } else if (p.type == TGX_VGOALS_MORE) {
        // a continuation record: no trajectory of its own; its tgx_phases row belongs to the record it continues
        if (continuation_has_owner(src)) E.ph = nullptr;
        else r.status = TGX_ST_BAD_PARAM;
    } else if (!params_ok(p, src)) {
        r.status = TGX_ST_BAD_PARAM;
        // trajectoryInsideBounds tests the geometry alone (Circle.cpp:171-179): report it for rejected records too
        if (lim && lim->check_box && (is_orbit(p.type) || is_line_like(p.type)) && !inside_bounds(p, lim->box))
            r.status |= TGX_ST_OUTSIDE_BOUNDS;
    } else {
        int n;
        double theta = 0.0, c = 1.0, s = 0.0;
        // the lanes that arrive here together and replay the same class of trajectory (replay_orbit / replay_line)
        E.grp = __match_any_sync(__activemask(), (p.type << 8) | (is_orbit(p.type) ? p.n_vgoals : 0));
        if (is_line_like(p.type))
            n = replay_line<XR>(p, max_samples, E, r.status, theta, c, s, tab, p.type == TGX_BOOMERANG);
        else
            n = replay_orbit<STATE, XR>(p, src, max_samples, E, r.status, tab);
        E.finish();
        // the evaluation kernel stages at most kMaxSegPerTile segments per tile: a trajectory that would need more
        // (dozens of speed goals inside one tile) is rejected rather than evaluated from a truncated list
        if (n > 0 && E.max_tile_segs > kMaxSegPerTile) {
            r.status |= TGX_ST_TOO_LONG;
            n = -1;
        }
        if (lim && lim->check_box && !inside_bounds(p, lim->box)) {
            r.status |= TGX_ST_OUTSIDE_BOUNDS;
            // Line::trajectoryInsideBounds reports "not feasible" when d2 < 0 (Line.cpp:165-168)
            if (is_line_like(p.type) && line_d2(p.u.line) < 0.0) r.status |= TGX_ST_LINE_D2_NEGATIVE;
        }
        if (n > 0) {
            r.n = n;
            r.nseg = E.nseg;
            r.ntile = E.ntile;
        }
        if (FILL && rec) {
            TrajRec t;
            t.n = r.n;
            if (is_line_like(p.type)) {
                t.type = TGX_LINE;
                t.f[0] = c; t.f[1] = s; t.f[2] = theta; t.f[3] = p.alt; t.f[4] = p.dt;
                t.f[5] = 0.0; t.f[6] = 0.0;
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
    if (FILL && r.n == 0) E.set_phase_counts(0);
    return r;
}

// generateStopTraj plan of one trajectory (Circle.cpp:132-169, Line.cpp:117-152, Figure8.cpp:130-167).
// Samples of a braking plan are numbered from 0 = first braking step, so the segment base is sample -1.
template <bool FILL>
__device__ PlanOut stop_one(const tgx_params& p, const Src& src, const double* from, int64_t max_samples, Emitter& E,
                            TrajRec* rec) {
    PlanOut r{0, 0u, 0, 0};
    TrajRec t;
    t.type = p.type & kRecTypeMask;
    t.n = 0;
    for (int i = 0; i < 7; ++i) t.f[i] = 0.0;
    if (TGX_IS_POLYLINE(p.type)) {
        // Braking trajectories of the constant-speed polyline family: the position stays at the setpoint being braked
        // from, the speed ramps down along a fixed heading (kRecStatic records, see tgx_internal.cuh).
        if (!poly_params_ok(p)) {
            r.status = TGX_ST_BAD_PARAM;
        } else {
            const int tmask = (1 << E.tile_shift) - 1;
            int k = -1;
            bool ok = true;
            t.type = kRecStatic;
            E.phase(0, TGX_PH_PRESSED_END, 0.0, 0.0);                 // Square.cpp:123, Bounce.cpp:86
            if (p.type == TGX_BOUNCE) {
                // Bounce::generateStopTraj, Bounce.cpp:74-103: vz *= 0.8 until |vz| <= 0.01 (then exactly 0); one
                // one-sample segment per step, all inside the first tile (at most kMaxSegPerTile - 1 of them)
                double vz = from[TGX_VZ];
                t.f[0] = p.u.poly.g[0]; t.f[1] = p.u.poly.g[1]; t.f[2] = from[TGX_PZ];   // :82, :94
                t.f[3] = from[TGX_PSI];                                                   // :83
                t.f[4] = 0.0; t.f[5] = 0.0; t.f[6] = 1.0;
                if (!isfinite(vz)) ok = false;
                while (ok && fabs(vz) > 0.01) {                       // :91
                    if (k + 2 >= kMaxSegPerTile) { ok = false; break; }        // |vz| beyond ~1e4 m/s
                    vz = dmul(vz, 0.8);                               // :92
                    if (fabs(vz) < 0.01) vz = 0.0;                    // :93
                    E.open(k, vz, 0.0, vz, 0.0, 0.0, 0.0);
                    ++k;
                    E.close(k, false, 0.0);
                }
            } else {
                // Square.cpp:112-137 == Rectangle.cpp:113-137 == Reciprocating.cpp:81-107 (a3_) == M.cpp:88-113 ==
                // I.cpp:96-121 == T.cpp:94-119 (literal decel 1.0)
                double v = __dsqrt_rn(dadd(dmul(from[TGX_VX], from[TGX_VX]), dmul(from[TGX_VY], from[TGX_VY])));
                const double heading = atan2(from[TGX_VY], from[TGX_VX]);
                const double decel = (p.type == TGX_M || p.type == TGX_I || p.type == TGX_T) ? 1.0 : p.u.poly.decel;
                double sn, cs;
                sincos(heading, &sn, &cs);
                t.f[0] = from[TGX_PX]; t.f[1] = from[TGX_PY]; t.f[2] = p.alt;
                t.f[3] = heading;
                t.f[4] = cs; t.f[5] = sn; t.f[6] = 0.0;
                double u0 = 0.0, u1 = 0.0;
                auto step = [](double) {};
                ok = isfinite(v) && ramp<false, true, false, false>(v, 0.0, dmul(decel, p.dt), 0.0, k, max_samples,
                                                                    tmask, E, -decel, u0, u1, 0.0, 0.0, step);
            }
            if (!ok) {
                r.status |= TGX_ST_TOO_LONG;
                E.nph = 0;
            } else {
                E.phase(k, TGX_PH_STOPPED, 0.0, 0.0);                 // key = size-1 (-1 when nothing was produced)
                r.n = k + 1;
            }
            E.finish();
            if (r.n > 0 && E.max_tile_segs > kMaxSegPerTile) {
                r.status |= TGX_ST_TOO_LONG;
                r.n = 0;
            }
            if (r.n > 0) {
                r.nseg = E.nseg;
                r.ntile = E.ntile;
            }
        }
        t.n = r.n;
        if (FILL && rec) *rec = t;
        return r;
    }
    if (p.type == TGX_VGOALS_MORE && continuation_has_owner(src)) {
        E.ph = nullptr;                    // a continuation record: nothing to brake
    } else if (!params_ok(p, src)) {
        r.status = TGX_ST_BAD_PARAM;
    } else {
        const int tmask = (1 << E.tile_shift) - 1;
        const double dt = p.dt;
        // 2D current (goal) vel: sqrt(pow(vx,2) + pow(vy,2))  (Circle.cpp:140-141, Line.cpp:124-125)
        double v = __dsqrt_rn(dadd(dmul(from[TGX_VX], from[TGX_VX]), dmul(from[TGX_VY], from[TGX_VY])));
        int k = -1;
        bool ok = true;
        E.phase(0, TGX_PH_PRESSED_END, 0.0, 0.0);                 // Circle.cpp:148, Line.cpp:132
        if (is_line_like(p.type)) {
            const tgx_line_params& l = p.u.line;
            const double theta = atan2(from[TGX_VY], from[TGX_VX]);   // Line.cpp:126-127, Boomerang.cpp:178-179
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
            while (v > 0.0) {
                const double vn = std_max(dsub(v, adt), 0.0);
                if (vn == v || (int64_t)k + 1 >= max_samples) { ok = false; break; }
                if (!open) { E.open(k, v, -adt, 0.0, x, y, -l.a3); open = true; }
                v = vn;
                step(v);
                ++k;
                if (k - E.cur.kb >= kRampChunk) { E.close(k, v == 0.0, 0.0); open = false; }
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
            const InvDiv rdiv = make_invdiv(o.r);
            auto step = [&](double vnew) { th = dadd(th, dmul(div_inv(vnew, rdiv), dt)); };
            ok = ramp<false, true, true, false>(v, 0.0, adt, ddiv(dt, o.r), k, max_samples, tmask, E, 0.0, th, unused,
                                                0.0, 0.0, step);                               // :150-157
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
        if (r.n > 0 && E.max_tile_segs > kMaxSegPerTile) {
            r.status |= TGX_ST_TOO_LONG;
            r.n = 0;
        }
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

// Can this trajectory's plan be written as a PhaseRec?  (A rejected one can: n = 0.)
__device__ __forceinline__ bool phase_fits(const tgx_params& p, int n, const Emitter& E, int max_n) {
    if (!is_orbit(p.type) && p.type != TGX_LINE) return false;
    if (n <= 0) return true;
    if (E.ramp_split || n > max_n) return false;
    if (p.type == TGX_LINE) return E.nseg <= kPhaseLineMaxSegs;
    return p.n_vgoals <= kPhaseMaxGoals && E.nseg <= kPhaseMaxSegs;
}

}  // namespace

// ---- kernels ------------------------------------------------------------------------------------------

// Per-plan statistics the fill pass accumulates (one atomic per warp).
__device__ __forceinline__ void accumulate_stats(PlanStats* stats, int n, int nseg, int ntile, bool overflow,
                                                 bool line_like, int kind, int seg_len = 0, int tile_segs = 0,
                                                 bool phase_misfit = true) {
    if (!stats) return;
    const unsigned mask = __activemask();
    const unsigned misfit = __reduce_or_sync(mask, phase_misfit ? 1u : 0u);
    const int mlen = __reduce_max_sync(mask, seg_len);
    const int mts = __reduce_max_sync(mask, tile_segs);
    const unsigned kinds = __reduce_or_sync(mask, 1u << kind);
    const unsigned tot = __reduce_add_sync(mask, (unsigned)n);
    const int mseg = __reduce_max_sync(mask, nseg);
    const int mtile = __reduce_max_sync(mask, ntile);
    const int mn = __reduce_max_sync(mask, n);
    const unsigned tiles = __reduce_add_sync(mask, (unsigned)ntile);
    const unsigned ovf = __reduce_or_sync(mask, overflow ? 1u : 0u);
    const unsigned lin = __reduce_or_sync(mask, line_like ? 1u : 0u);
    if ((int)(threadIdx.x & 31) == __ffs(mask) - 1) {
        atomicAdd(&stats->total_samples, (unsigned long long)tot);
        atomicAdd(&stats->total_tiles, (unsigned long long)tiles);
        atomicMax(&stats->max_nseg, mseg);
        atomicMax(&stats->max_ntile, mtile);
        atomicMax(&stats->max_n, mn);
        atomicMax(&stats->max_seg_len, mlen);
        atomicMax(&stats->max_tile_segs, mts);
        if (ovf) atomicOr(&stats->overflow, 1);
        if (misfit && !*reinterpret_cast<volatile int*>(&stats->phase_misfit)) atomicOr(&stats->phase_misfit, 1);
        if (lin) atomicOr(&stats->has_line, 1);
        if ((*reinterpret_cast<volatile int*>(&stats->kinds) & (int)kinds) != (int)kinds) atomicOr(&stats->kinds, (int)kinds);
    }
}

// Replay order of a mixed batch: key[t] = replay class of trajectory t, idx[t] = t; a one-pass radix sort of the pairs
// (engine.cu) then hands neighbouring lanes trajectories of the same class.
__global__ void __launch_bounds__(256)
replay_keys_kernel(const tgx_params* __restrict__ params, int64_t n, uint8_t* __restrict__ key, int32_t* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int2 head = __ldg(reinterpret_cast<const int2*>(params + i));      // {type, n_vgoals}
    key[i] = (uint8_t)replay_class(head.x, head.y);
    idx[i] = (int32_t)i;
}

// Counting pass: N_i, status_i and, with SEGS, the number of segments / tiles the fill pass will emit (which
// requires the full state replay, because exact-progression breaks depend on theta).
template <bool SEGS, bool XR>
__global__ void __launch_bounds__(128, 8)
plan_count_kernel(const tgx_params* __restrict__ params, const double* __restrict__ stop_from, int64_t n,
                  tgx_limits lim, int has_lim, int64_t max_samples, int tile_shift,
                  const CurTable* __restrict__ tab,
                  int32_t* __restrict__ counts, uint32_t* __restrict__ status, int32_t* __restrict__ nseg,
                  int32_t* __restrict__ ntile) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const tgx_params p = load_params(params, i);
    const Src src{params, i, n};
    Emitter E{tile_shift, (int32_t)i, nullptr, nullptr, 0, nullptr, is_orbit(p.type), 0x7fffffff, 0x7fffffff};
    PlanOut r;
    if (stop_from) {
        double from[TGX_NCHAN];
#pragma unroll
        for (int c = 0; c < TGX_NCHAN; ++c) from[c] = stop_from[i * TGX_NCHAN + c];
        r = stop_one<false>(p, src, from, max_samples, E, nullptr);
    } else {
        r = plan_one<false, SEGS, XR>(p, src, max_samples, has_lim ? &lim : nullptr, E, nullptr, tab);
    }
    if (counts) counts[i] = r.n;
    if (status) status[i] = r.status;
    if (nseg) nseg[i] = r.nseg;
    if (ntile) ntile[i] = r.ntile;
}

// Fill pass: the replay that writes the tables.
//   exact-offset mode (seg_slab == 0): trajectory i owns Seg[seg_off[i] ..] and Tile[tile_off[i] ..], sized by the
//       counting pass and the scans (two replays per plan);
//   slab mode (seg_slab > 0): trajectory i owns the fixed slices Seg[i*seg_slab ..], Tile[i*tile_slab ..] — no
//       counting pass, no scans.  A trajectory that needs more than its slice sets stats->overflow (its extra
//       records are dropped) and the host redoes the plan in exact-offset mode.  Unused tile slots are written as
//       empty tiles (nseg = 0), which the evaluation kernel skips.
// (CTAs of 128 threads per SM the replay is compiled for.  Round 1 found 8 — 64 registers, 1.4 KB of spill code — better
//  than fewer spills; with the replay barriers and the reciprocal-based run lengths of round 2 the balance moved: config 4's
//  plan, 10^7 circles: 8 CTAs 5.70 ms, 6: 5.57, 5 (96 registers, 0.2 KB of spills): 5.50, 4 (126 registers, none): 5.66)
#ifndef TGX_FILL_CTAS
#define TGX_FILL_CTAS 5
#endif
template <bool XR>
__global__ void __launch_bounds__(128, TGX_FILL_CTAS)
plan_fill_kernel(const tgx_params* __restrict__ params, const double* __restrict__ stop_from, int64_t n,
                 tgx_limits lim, int has_lim, int64_t max_samples, int tile_shift,
                 const CurTable* __restrict__ tab, const int32_t* __restrict__ plan_counts,
                 const int64_t* __restrict__ seg_off, const int64_t* __restrict__ tile_off, int seg_slab,
                 int tile_slab, TrajRec* __restrict__ recs, Seg* __restrict__ segs, Tile* __restrict__ tiles,
                 int32_t* __restrict__ counts, uint32_t* __restrict__ status, int32_t* __restrict__ counts2,
                 uint32_t* __restrict__ status2, tgx_phases* __restrict__ phases, PlanStats* __restrict__ stats,
                 const int32_t* __restrict__ order) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    // order: thread t replays trajectory order[t] (a mixed batch sorted by replay class, so that the lanes of a warp
    // walk the same code: 4.5 of 32 lanes were active on average in config 3's unsorted replay); everything the
    // trajectory owns is indexed by i, so the tables do not depend on the order
    const int64_t i = order ? (int64_t)__ldg(order + t) : t;
    const tgx_params p = load_params(params, i);
    const Src src{params, i, n};
    const bool slab = seg_slab > 0;
    // exact-offset mode: a trajectory the counting pass rejected owns no slice: replay it without writing tables
    const bool keep = slab || plan_counts[i] > 0;
    const int64_t so = slab ? i * (int64_t)seg_slab : seg_off[i];
    const int64_t to = slab ? i * (int64_t)tile_slab : tile_off[i];
    TrajRec* rec_out = recs + i;
    Tile* tile_out = tiles + to;
    Emitter E{tile_shift, (int32_t)i, keep ? segs + so : nullptr, keep ? tile_out : nullptr, (int32_t)so,
              phases ? phases + i : nullptr, is_orbit(p.type), slab ? seg_slab : 0x7fffffff,
              slab ? tile_slab : 0x7fffffff};
    PlanOut r;
    if (stop_from) {
        double from[TGX_NCHAN];
#pragma unroll
        for (int c = 0; c < TGX_NCHAN; ++c) from[c] = stop_from[i * TGX_NCHAN + c];
        r = stop_one<true>(p, src, from, max_samples, E, rec_out);
        if (E.ph && r.status) E.ph->n = 0;
    } else {
        r = plan_one<true, true, XR>(p, src, max_samples, has_lim ? &lim : nullptr, E, rec_out, tab);
    }
    bool overflow = false;
    if (slab) {
        overflow = r.nseg > seg_slab || r.ntile > tile_slab;
        // E counted what the replay emitted even if the trajectory was rejected afterwards (r.ntile == 0 then)
        const int used = (r.n > 0 && !overflow) ? r.ntile : 0;
        for (int t = used; t < tile_slab; ++t) {
            Tile e;
            e.traj = (int32_t)i; e.k_lo = 0; e.seg_begin = (int32_t)so; e.nseg = 0;
            tile_out[t] = e;
        }
    }
    if (counts) counts[i] = r.n;
    if (status) status[i] = r.status;
    if (counts2) counts2[i] = r.n;
    if (status2) status2[i] = r.status;
    accumulate_stats(stats, r.n, r.nseg, r.ntile, overflow, is_line_like(p.type), replay_class(p.type, p.n_vgoals),
                     r.n > 0 ? E.max_seg_len : 0, r.n > 0 ? E.max_tile_segs : 0,
                     stop_from || !phase_fits(p, r.n, E, kPhaseMaxSamples));
}

// Phase plan: batches of short orbits (Circle / Figure8 with at most kPhaseMaxGoals speed goals, at most kPhaseMaxSegs
// segments) and plain lines (at most kPhaseLineMaxSegs segments), at most max_n samples each.  The same replay as
// plan_fill_kernel, but instead of TrajRec + Seg + Tile records it writes ONE self-contained PhaseRec per trajectory:
// where each segment ends, the replayed state there, the segment kinds and the constants of the parameter record.  The
// evaluation CTA rebuilds the table path's Seg records from it (build_phase_segment, eval.cu) — the samples are the same
// bits whichever way the batch was planned.  A trajectory that does not qualify sets stats->overflow and the host plans
// the batch with segment tables.  `order`: see plan_fill_kernel.
// (CTAs of 128 threads per SM, plan time per Mi trajectories of config 2 / config 3: 12 CTAs (40 registers, 2 KB of spill
//  code) 0.68 / 1.12 ms, 8: 0.62 / 0.99, 6: 0.58 / 0.96, 5 (96 registers): 0.56 / 0.96, 4 (112 registers, no spills): 0.60)
#ifndef TGX_PHASE_CTAS
#define TGX_PHASE_CTAS 5
#endif
__global__ void __launch_bounds__(128, TGX_PHASE_CTAS)
plan_phase_kernel(const tgx_params* __restrict__ params, int64_t n, tgx_limits lim, int has_lim,
                  int64_t max_samples, int tile_shift, int max_n, const CurTable* __restrict__ tab,
                  PhaseRec* __restrict__ phase, PhaseExt* __restrict__ phase_ext, int32_t* __restrict__ counts,
                  uint32_t* __restrict__ status, int32_t* __restrict__ counts2, uint32_t* __restrict__ status2,
                  tgx_phases* __restrict__ phases, PlanStats* __restrict__ stats, const int32_t* __restrict__ order) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = order ? (int64_t)__ldg(order + t) : t;
    const tgx_params p = load_params(params, i);
    PhaseRec rec;
    rec.n = 0;
    int32_t key[kPhaseMaxSegs];
    double th[kPhaseMaxSegs];
#pragma unroll
    for (int q = 0; q < kPhaseMaxSegs; ++q) {
        key[q] = 0;
        th[q] = 0.0;
    }
    const bool orbit = is_orbit(p.type);
    Emitter E{tile_shift, (int32_t)i, nullptr, nullptr, 0, phases ? phases + i : nullptr, orbit, 0x7fffffff,
              0x7fffffff, key, th};
    PlanOut r{0, 0u, 0, 0};
    TrajRec tr;
    bool overflow = !orbit && p.type != TGX_LINE;
    if (!overflow) {
        r = plan_one<true, true, false>(p, Src{params, i, n}, max_samples, has_lim ? &lim : nullptr, E, &tr, tab);
        overflow = !phase_fits(p, r.n, E, max_n);
        if (r.n > 0 && !overflow) rec.n = E.nseg;
    } else if (phases) {
        phases[i].n = 0;
    }
    rec.type = p.type;
    rec.kinds[0] = (uint32_t)E.kinds;
    rec.kinds[1] = (uint32_t)(E.kinds >> 32);
#pragma unroll
    for (int q = 0; q < kPhaseBaseSegs; ++q) {
        rec.key[q] = key[q];
        rec.th[q] = th[q];
    }
    if (rec.n > kPhaseBaseSegs) {
        PhaseExt x;
#pragma unroll
        for (int q = 0; q < kPhaseMaxSegs - kPhaseBaseSegs; ++q) {
            x.key[q] = key[kPhaseBaseSegs + q];
            x.th[q] = th[kPhaseBaseSegs + q];
        }
        const int4* src = reinterpret_cast<const int4*>(&x);
        int4* dst = reinterpret_cast<int4*>(phase_ext + i);
#pragma unroll
        for (int q = 0; q < (int)(sizeof(PhaseExt) / 16); ++q) dst[q] = src[q];
    }
    if (p.type == TGX_LINE) {
        const tgx_line_params& l = p.u.line;
        rec.c.lcos = tr.f[0]; rec.c.lsin = tr.f[1]; rec.c.ltheta = tr.f[2]; rec.c.lalt = tr.f[3]; rec.c.ldt = tr.f[4];
        rec.c.lvg = l.v_goal;
        rec.c.ladt1 = dmul(l.a1, p.dt);
        rec.c.ladt3 = dmul(l.a3, p.dt);
        rec.c.la1 = l.a1;
        rec.c.la3 = l.a3;
        rec.c.lspare[0] = rec.c.lspare[1] = 0.0;
    } else {
        const tgx_orbit_params& o = p.u.orbit;
        rec.c.dtr = orbit ? ddiv(p.dt, o.r) : 0.0;
        rec.c.rinv = orbit ? ddiv(1.0, o.r) : 0.0;
        rec.c.r = o.r; rec.c.cx = o.cx; rec.c.cy = o.cy; rec.c.alt = p.alt;
        rec.c.adt = dmul(o.accel, p.dt);
        rec.c.spare = 0.0;
#pragma unroll
        for (int q = 0; q < kPhaseMaxGoals; ++q) {
            rec.c.vg[q] = o.v_goals[q];
            // omega = v / r_; theta += omega * dt_ (Circle.cpp:65-67), as replay_orbit's hold passes it on
            rec.c.w[q] = orbit ? dmul(ddiv(o.v_goals[q], o.r), p.dt) : 0.0;
        }
    }
    // 256-byte record: sixteen 16-byte stores
    {
        const int4* src = reinterpret_cast<const int4*>(&rec);
        int4* dst = reinterpret_cast<int4*>(phase + i);
#pragma unroll
        for (int q = 0; q < (int)(sizeof(PhaseRec) / 16); ++q) dst[q] = src[q];
    }
    if (counts) counts[i] = r.n;
    if (status) status[i] = r.status;
    if (counts2) counts2[i] = r.n;
    if (status2) status2[i] = r.status;
    accumulate_stats(stats, r.n, 0, 0, overflow, is_line_like(p.type), replay_class(p.type, p.n_vgoals), 0, 0, overflow);
    }
}

// "Per-time evaluation": a one-sample plan per trajectory from an explicit state, i.e. the public helpers
// createCircleGoal(v, accel, theta) (Circle.cpp:96), createFigure8Goal (Figure8.cpp:96) and
// createLineGoal(last_x, last_y, v, accel, theta) (Line.cpp:91).  state[i] = {v, accel, s0, s1}: orbit s0 = theta;
// line s0 = last_x, s1 = last_y and the explicit heading theta is read from params[i].u.line.reserved[0].
__global__ void __launch_bounds__(128)
plan_samples_kernel(const tgx_params* __restrict__ params, const double* __restrict__ state, int64_t n,
                    TrajRec* __restrict__ recs, Seg* __restrict__ segs, Tile* __restrict__ tiles,
                    int32_t* __restrict__ counts, uint32_t* __restrict__ status) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const tgx_params p = load_params(params, i);
    const double v = state[4 * i + 0], accel = state[4 * i + 1], s0 = state[4 * i + 2], s1 = state[4 * i + 3];
    TrajRec t;
    Seg g;
    g.kb = -1; g.n = 1; g.flags = 0; g.pad = 0;
    g.vb = v; g.dv = 0.0; g.vclamp = v;
    t.n = 1;
    for (int q = 0; q < 7; ++q) t.f[q] = 0.0;
    const bool ok = TGX_IS_POLYLINE(p.type) ? poly_params_ok(p) : params_ok(p, Src{params, i, n}, false);
    if (TGX_IS_POLYLINE(p.type)) {
        // createSquareGoal(x, y, v, accel, heading) and its copies (Square.cpp:94-110); createBounceGoal(x, y, z, vz,
        // heading) (Bounce.cpp:54-72).  state = {v | vz, accel, x, y}; z = g[5] (Bounce), heading = g[6].
        const double heading = p.u.poly.g[6];
        double sn, cs;
        sincos(heading, &sn, &cs);
        t.type = kRecStatic;
        t.f[0] = s0; t.f[1] = s1; t.f[3] = heading;
        if (p.type == TGX_BOUNCE) {
            t.f[2] = p.u.poly.g[5];
            t.f[4] = 0.0; t.f[5] = 0.0; t.f[6] = 1.0;
            g.acc = 0.0;
        } else {
            t.f[2] = p.alt;
            t.f[4] = cs; t.f[5] = sn; t.f[6] = 0.0;
            g.acc = accel;
        }
        g.s0 = 0.0; g.s1 = 0.0;
    } else if (is_line_like(p.type)) {
        const double theta = p.u.line.reserved[0];
        double sn, cs;
        sincos(theta, &sn, &cs);
        t.type = TGX_LINE;
        t.f[0] = cs; t.f[1] = sn; t.f[2] = theta; t.f[3] = p.alt; t.f[4] = p.dt;
        g.s0 = s0; g.s1 = s1; g.acc = accel;
    } else {
        t.type = p.type & kRecTypeMask;
        t.f[0] = p.u.orbit.r; t.f[1] = p.u.orbit.cx; t.f[2] = p.u.orbit.cy; t.f[3] = p.alt;
        t.f[4] = ddiv(p.dt, p.u.orbit.r); t.f[5] = ddiv(1.0, p.u.orbit.r);
        g.s0 = s0; g.s1 = 0.0; g.acc = s0;    // the segment's (only) sample carries theta exactly
    }
    if (!ok) t.n = 0;
    recs[i] = t;
    segs[i] = g;
    Tile tl;
    tl.traj = (int32_t)i; tl.k_lo = 0; tl.seg_begin = (int32_t)i; tl.nseg = ok ? 1 : 0;
    tiles[i] = tl;
    if (counts) counts[i] = ok ? 1 : 0;
    if (status) status[i] = ok ? 0u : (uint32_t)TGX_ST_BAD_PARAM;
}

// Self-test of div_inv against __ddiv_rn on pseudo-random operands (splitmix64 streams): counts mismatches.
__global__ void __launch_bounds__(256)
selftest_division_kernel(int64_t n, uint64_t seed, int per_thread, unsigned long long* __restrict__ mismatches) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t z = seed + 0x9e3779b97f4a7c15ULL * (uint64_t)(i + 1);
    auto next = [&]() {
        z += 0x9e3779b97f4a7c15ULL;
        uint64_t x = z;
        x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
        x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
        return x ^ (x >> 31);
    };
    // divisor: random significand, exponent in [-8, 8]; every 16th thread gets a hard significand
    uint64_t mb = next() & 0x000fffffffffffffULL;
    if ((i & 15) == 0) mb = 0x000fffffffffffffULL - (next() & 3);       // all ones and its neighbours
    if ((i & 15) == 1) mb = next() & 7;                                  // just above a power of two
    const int eb = (int)(next() % 17) - 8;
    const double b = __longlong_as_double((long long)(((uint64_t)(1023 + eb) << 52) | mb));
    const InvDiv d = make_invdiv(b);
    unsigned long long bad = 0;
    for (int t = 0; t < per_thread; ++t) {
        const uint64_t ma = next() & 0x000fffffffffffffULL;
        const int ea = (int)(next() % 41) - 30;
        const double a = __longlong_as_double((long long)(((uint64_t)(1023 + ea) << 52) | ma));
        if (div_inv(a, d) != ddiv(a, b)) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// Ragged batches planned with fixed slices (one replay): most tile slots are empty, and launching one CTA per slot would
// waste the evaluation grid.  These two kernels turn the slots into the dense work list of an exact-offset plan
// (Tile.seg_begin already is the absolute position of the tile's segments inside the trajectory's slice).
__global__ void __launch_bounds__(256)
count_used_tiles_kernel(int64_t n, int tile_slab, const Tile* __restrict__ slots, int32_t* __restrict__ ntile) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = 0;
    for (int t = 0; t < tile_slab; ++t) c += slots[i * tile_slab + t].nseg > 0 ? 1 : 0;
    ntile[i] = c;
}

__global__ void __launch_bounds__(256)
compact_tiles_kernel(int64_t n, int tile_slab, const Tile* __restrict__ slots, const int64_t* __restrict__ tile_off,
                     Tile* __restrict__ dense) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t o = tile_off[i];
    for (int t = 0; t < tile_slab; ++t) {
        const Tile e = slots[i * tile_slab + t];
        if (e.nseg > 0) dense[o++] = e;            // the used slots of a trajectory are its first ones, in tile order
    }
}

// Chunk boundaries of tgx_generate: boundary b would fall at (b + 1) * chunk; it moves forward past continuation records
// (tgx.h: TGX_VGOALS_MORE), so that a Circle / Figure8 with more than 8 goal speeds is never cut from its goals.
__global__ void chunk_bounds_kernel(const tgx_params* __restrict__ params, int64_t n, int64_t chunk, int nb,
                                    int64_t* __restrict__ bounds) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    int64_t lo = (int64_t)(b + 1) * chunk;
    while (lo < n && __ldg(&params[lo].type) == TGX_VGOALS_MORE) ++lo;
    bounds[b] = lo < n ? lo : n;
}

cudaError_t launch_chunk_bounds(const tgx_params* params, int64_t n, int64_t chunk, int nb, int64_t* bounds,
                                cudaStream_t stream) {
    if (nb <= 0) return cudaSuccess;
    chunk_bounds_kernel<<<(nb + 63) / 64, 64, 0, stream>>>(params, n, chunk, nb, bounds);
    return cudaGetLastError();
}

// ---- host-side launchers (called from engine.cu) -------------------------------------------------------

cudaError_t launch_count_used_tiles(int64_t n, int tile_slab, const Tile* slots, int32_t* ntile, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    count_used_tiles_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, tile_slab, slots, ntile);
    return cudaGetLastError();
}

cudaError_t launch_compact_tiles(int64_t n, int tile_slab, const Tile* slots, const int64_t* tile_off, Tile* dense,
                                 cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    compact_tiles_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, tile_slab, slots, tile_off, dense);
    return cudaGetLastError();
}


cudaError_t launch_build_cur_table(const tgx_params* params, int64_t max_samples, void* table, cudaStream_t stream) {
    build_cur_table_kernel<<<1, 32, 0, stream>>>(params, max_samples, static_cast<CurTable*>(table));
    return cudaGetLastError();
}

size_t cur_table_bytes() { return sizeof(CurTable); }

cudaError_t launch_plan_count(const tgx_params* params, const double* stop_from, int64_t n, const tgx_limits* lim,
                              int64_t max_samples, int tile_shift, bool exact_ramps, const void* cur_table,
                              int32_t* counts, uint32_t* status, int32_t* nseg, int32_t* ntile,
                              cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    tgx_limits l{};
    if (lim) l = *lim;
    const int threads = 128;
    const int64_t blocks = (n + threads - 1) / threads;
    const CurTable* tab = static_cast<const CurTable*>(cur_table);
#define TGX_LAUNCH_COUNT(SEGS, XR)                                                                          \
    plan_count_kernel<SEGS, XR><<<(unsigned)blocks, threads, 0, stream>>>(params, stop_from, n, l, lim ? 1 : 0, \
                                                                         max_samples, tile_shift, tab, counts,  \
                                                                         status, nseg, ntile)
    if (nseg || ntile) {
        if (exact_ramps) TGX_LAUNCH_COUNT(true, true);
        else TGX_LAUNCH_COUNT(true, false);
    } else {
        TGX_LAUNCH_COUNT(false, false);   // counts and status only: the fast replay is exact for both
    }
#undef TGX_LAUNCH_COUNT
    return cudaGetLastError();
}

cudaError_t launch_plan_fill(const tgx_params* params, const double* stop_from, int64_t n, const tgx_limits* lim,
                             int64_t max_samples, int tile_shift, bool exact_ramps, const void* cur_table,
                             const int32_t* plan_counts, const int64_t* seg_off, const int64_t* tile_off,
                             int seg_slab, int tile_slab, TrajRec* recs, Seg* segs, Tile* tiles, int32_t* counts,
                             uint32_t* status, int32_t* counts2, uint32_t* status2, tgx_phases* phases,
                             PlanStats* stats, cudaStream_t stream, const int32_t* order) {
    if (n <= 0) return cudaSuccess;
    tgx_limits l{};
    if (lim) l = *lim;
    const int threads = 128;
    const int64_t blocks = (n + threads - 1) / threads;
    const CurTable* tab = static_cast<const CurTable*>(cur_table);
#define TGX_LAUNCH_FILL(XR)                                                                                       \
    plan_fill_kernel<XR><<<(unsigned)blocks, threads, 0, stream>>>(                                               \
        params, stop_from, n, l, lim ? 1 : 0, max_samples, tile_shift, tab, plan_counts, seg_off, tile_off,       \
        seg_slab, tile_slab, recs, segs, tiles, counts, status, counts2, status2, phases, stats, order)
    if (exact_ramps) TGX_LAUNCH_FILL(true);
    else TGX_LAUNCH_FILL(false);
#undef TGX_LAUNCH_FILL
    return cudaGetLastError();
}

cudaError_t launch_replay_keys(const tgx_params* params, int64_t n, uint8_t* key, int32_t* idx, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    replay_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(params, n, key, idx);
    return cudaGetLastError();
}

cudaError_t launch_plan_phase(const tgx_params* params, int64_t n, const tgx_limits* lim, int64_t max_samples,
                              int tile_shift, int max_n, const void* cur_table, PhaseRec* phase, PhaseExt* phase_ext,
                              int32_t* counts, uint32_t* status, int32_t* counts2, uint32_t* status2,
                              tgx_phases* phases, PlanStats* stats, cudaStream_t stream, const int32_t* order) {
    if (n <= 0) return cudaSuccess;
    tgx_limits l{};
    if (lim) l = *lim;
    const int cta = 128;
    const int64_t grid = (n + cta - 1) / cta;
    plan_phase_kernel<<<(unsigned)grid, cta, 0, stream>>>(
        params, n, l, lim ? 1 : 0, max_samples, tile_shift, max_n, static_cast<const CurTable*>(cur_table), phase,
        phase_ext, counts, status, counts2, status2, phases, stats, order);
    return cudaGetLastError();
}

cudaError_t launch_plan_samples(const tgx_params* params, const double* state, int64_t n, TrajRec* recs, Seg* segs,
                                Tile* tiles, int32_t* counts, uint32_t* status, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    plan_samples_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(params, state, n, recs, segs, tiles, counts,
                                                                         status);
    return cudaGetLastError();
}

cudaError_t launch_selftest_division(int64_t n, uint64_t seed, int per_thread, unsigned long long* mismatches,
                                     cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    selftest_division_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, seed, per_thread, mismatches);
    return cudaGetLastError();
}

}  // namespace tgx
