// replay_common.cuh — device helpers shared by the planning kernels (plan.cu, polyline.cu): non-contracted IEEE
// arithmetic, the exact skip-ahead for running sums, and the hold counter table.  Include inside namespace tgx; every
// translation unit that includes this must be compiled with -fmad=false.
#pragma once

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


// ---- skip-ahead for x <- fl(x + a) with a constant addend -------------------------------------------------
// While x stays inside one binade [2^E, 2^(E+1)] (same sign), every exact sum x + a is rounded to the same grid
// of spacing u = 2^(E-52), so each step adds exactly the same inc = fl(x + a) - x and the reference's running sum is
// an exact arithmetic progression — also when a is an odd multiple of u/2 (a tie), once x has an even significand.
// Given one real step x0 -> x1 this returns how many FURTHER steps are guaranteed to add exactly x1 - x0
// (0 when the step crossed a binade, started a tie run from an odd significand, or the values are zero / subnormal-ish).
// inv_a (optional): an approximation of 1 / |a| hoisted out of the caller's loop.  It only replaces the division in the
// ESTIMATE of the run length — the effective increment differs from a by less than half an ulp of x, so the estimate is
// off by at most one step — and the exact check below settles the result either way: an estimate that is too long is
// shortened, one that is a step short makes the caller take one more (equally exact) run.
__device__ __forceinline__ long long regular_run(double x0, double x1, double a, double inv_a = 0.0) {
    const long long i0 = __double_as_longlong(x0), i1 = __double_as_longlong(x1);
    if (((i0 ^ i1) >> 52) != 0) return 0;                 // sign or exponent changed: an irregular (crossing) step
    const int e = (int)((i1 >> 52) & 0x7ff);
    if (e < 64 || e == 0x7ff) return 0;                   // zero, subnormal, tiny or non-finite: step one by one
    const double inc = dsub(x1, x0);                      // exact
    if (inc == 0.0) return 1LL << 40;                     // |a| < u/2: x does not move while it stays in this binade
    const double lo = __longlong_as_double(i1 & 0x7ff0000000000000LL);   // 2^E
    // tie: |a| is an odd multiple of u/2 = 2^(E-53), the increments alternate.  In integers: |a| = ma * 2^(ea-1075) with
    // the 53-bit significand ma, so |a| / (u/2) = ma * 2^(ea-e+1); for ea >= e that is an even number (or >= 2^53: no tie
    // by definition), a subnormal |a| is below u/2, and otherwise it is ma >> s with s = e - ea - 1: an odd integer exactly
    // when ma has s trailing zeros.  (The same test used to be  r = |a| / (lo * 2^-53); r < 2^53 && r == rint(r) && odd(r):
    // a division, a rint and a 64-bit conversion, 17 % of the replay kernels' executed instructions.)
    {
        const unsigned long long ab = (unsigned long long)__double_as_longlong(a) & 0x7fffffffffffffffULL;
        const int ea = (int)(ab >> 52);
        if (ea > 0 && ea < e) {
            const unsigned long long ma = (ab & 0x000fffffffffffffULL) | 0x0010000000000000ULL;
            // Round-half-even sends a tie to the neighbour with the even significand and an even x stays even (its
            // increment is then an even number of ulps), so from an EVEN x0 on the run is as regular as any other; only
            // the one step that starts from an odd x0 is irregular.  (Returning 0 for every step of a tie binade made
            // one lane in a few replay its ramp step by step — 64 or 128 single steps — while the rest of its warp
            // waited: 9 of 32 lanes active per executed instruction.)
            if (__ffsll((long long)ma) - 1 == e - ea - 1 && (i0 & 1LL)) return 0;
        }
    }
    const double ax = fabs(x1), ai = fabs(inc);
    const bool growing = (inc > 0.0) == (x1 > 0.0);
    const double room = growing ? dsub(dmul(2.0, lo), ax) : dsub(ax, lo);   // exact distance to the binade edge
    const double q = floor(inv_a > 0.0 ? dmul(room, inv_a) : ddiv(room, ai));
    long long J = q > 1e15 ? (1LL << 40) : (long long)q;
    // exact check: after J further steps the value must still lie in [2^E, 2^(E+1)]
    while (J > 0) {
        const double y = fabs(fma((double)J, inc, x1));
        if (growing ? (y <= dmul(2.0, lo)) : (y >= lo)) break;
        --J;
    }
    return J;
}


// ---- the hold counter: current_t_traj_ += dt_ -----------------------------------------------------------------
// `double cur = 0; while (cur < t_hold) { ...; cur += dt; }` (Circle.cpp:62-71): the number of iterations is the
// smallest m with c_m >= t_hold where c_0 = 0, c_m = fl(c_{m-1} + dt).  That sequence depends on dt alone, and by
// the binade argument above it is piecewise an exact arithmetic progression: (m0, c0, inc, cnt) says "after m0
// steps cur == c0, and each of the next cnt steps adds exactly inc".  A batch almost always shares one dt
// (1/pub_freq, TrajectoryGenerator.cpp:171-172), so the runs are tabulated once per plan by a one-thread kernel and
// every trajectory looks its hold length up; a trajectory with a different dt walks the runs itself.
constexpr int kCurTableMax = 192;

struct CurTable {
    double dt;
    int32_t n;            // entries
    int32_t stagnates;    // after the last entry cur stops changing (dt < ulp/2): longer holds never terminate
    long long m_end;      // steps covered by the table
    long long m0[kCurTableMax];
    long long cnt[kCurTableMax];
    double c0[kCurTableMax];
    double inc[kCurTableMax];
};

// One run of the counter starting from (m, cur): returns false on stagnation.
__device__ __forceinline__ bool cur_run(double cur, double dt, double& inc, long long& cnt) {
    const double cn = dadd(cur, dt);
    if (cn == cur) return false;
    inc = dsub(cn, cur);
    cnt = 1 + regular_run(cur, cn, dt);
    return true;
}

// Smallest j in [1, cnt] with c0 + j*inc >= t, given that c0 < t and c0 + cnt*inc >= t (all partial sums exact).
__device__ __forceinline__ long long steps_to_reach(double c0, double inc, long long cnt, double t) {
    const double g = ceil(ddiv(dsub(t, c0), inc));
    long long j = g < 1.0 ? 1 : (g > (double)cnt ? cnt : (long long)g);
    while (j > 1 && fma((double)(j - 1), inc, c0) >= t) --j;
    while (j < cnt && fma((double)j, inc, c0) < t) ++j;
    return j;
}


// Number of hold iterations for t_hold, or -1 if the reference would not terminate within `limit` more samples.
__device__ long long hold_steps(double t_hold, double dt, long long limit, const CurTable* __restrict__ tab) {
    if (!(0.0 < t_hold)) return 0;
    if (tab && tab->dt == dt && tab->n > 0) {
        // largest entry whose start value is below t_hold (c0 is increasing, c0[0] = 0 < t_hold)
        int lo = 0, hi = tab->n - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (tab->c0[mid] < t_hold) lo = mid; else hi = mid - 1;
        }
        const double c0 = tab->c0[lo], inc = tab->inc[lo];
        const long long cnt = tab->cnt[lo];
        if (fma((double)cnt, inc, c0) >= t_hold) {
            const long long m = tab->m0[lo] + steps_to_reach(c0, inc, cnt, t_hold);
            return m <= limit ? m : -1;
        }
        // beyond the table: never terminates (stagnation), exceeds the guard, or the table was cut short
        if (tab->stagnates || tab->m_end > limit) return -1;
    }
    long long m = 0;
    double cur = 0.0;
    while (cur < t_hold) {
        double inc;
        long long cnt;
        if (m > limit || !cur_run(cur, dt, inc, cnt)) return -1;
        if (fma((double)cnt, inc, cur) >= t_hold) cnt = steps_to_reach(cur, inc, cnt, t_hold);
        cur = fma((double)cnt, inc, cur);
        m += cnt;
    }
    return m <= limit ? m : -1;
}


// Acceptance rule of the constant-speed polyline family (Square / Rectangle / Reciprocating / Bounce / M / I / T).
__device__ bool poly_params_ok(const tgx_params& p) {
    // "All velocities must be > 0" (TrajectoryGenerator.cpp:227-232, 306-311, 327-332, 349-354, 371-376), accel > 0
    // (:241-244, :251-254, :275-278); a leg of length 0 would make the reference divide 0 / 0.
    const tgx_polyline_params& q = p.u.poly;
    if (!finite_pos(p.dt) || !isfinite(p.alt)) return false;
    if (!finite_pos(q.v_goal) || !isfinite(q.t_traj) || !isfinite(q.orientation)) return false;
    for (int i = 0; i < 5; ++i)
        if (!isfinite(q.g[i])) return false;
    switch (p.type) {
        case TGX_SQUARE: return finite_pos(q.g[0]) && finite_pos(q.decel);
        case TGX_RECTANGLE: return finite_pos(q.g[0]) && finite_pos(q.g[1]) && finite_pos(q.decel);
        case TGX_RECIPROCATING:
            return isfinite(q.g[5]) && finite_pos(q.decel) && (q.g[0] != q.g[3] || q.g[1] != q.g[4]);
        case TGX_BOUNCE: return q.g[2] != q.g[3];
        default: return finite_pos(q.g[2]) && finite_pos(q.g[3]);   // M, I, T: length, width
    }
}

}  // namespace
}  // namespace tgx
