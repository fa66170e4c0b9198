// polyline.cu — planning and evaluation kernels for the constant-speed polyline family:
// Square (Square.cpp:21-92), Rectangle (Rectangle.cpp:20-93), Reciprocating (Reciprocating.cpp:13-60),
// Bounce (Bounce.cpp:19-52), M (M.cpp:13-67), I (I.cpp:19-75), T (T.cpp:19-73).
//
// What the reference does: it walks a short list of waypoints at constant speed; leg by leg it emits the samples
//     frac = (double)i / steps;  p = start + frac * (end - start)          i = i0 .. steps,  steps = ceil(d / (v*dt))
// while a running clock `t += dt` is below t_traj, laps repeating (and, for the letters and the reciprocating line,
// reversing) until the clock runs out.  So
//   - the sample COUNT is the number of `t += dt` iterations, the same floating-point counter as a hold phase of a
//     Circle (replay_common.cuh: hold_steps), plus a shape-specific 0 / 1;
//   - the samples are PERIODIC in k: one period is the list of legs, each contributing steps + 1 - i0 samples.
// The planner (one thread per trajectory) builds the waypoints and step counts with the reference's own operation
// order in non-contracted IEEE arithmetic (this file is compiled with -fmad=false) and writes kPolyRecs 64-byte
// records per trajectory; the evaluation kernel (one CTA per 1024-sample tile, like eval.cu) maps k to (leg, i) with one
// integer modulo and evaluates the reference's two-rounding interpolation, so positions are bit-identical.
#include <cuda_runtime.h>

#include "tgx_internal.cuh"
#include "replay_common.cuh"
#include "store.cuh"

namespace tgx {

namespace {

constexpr double kPi = 3.14159265358979323846;   // M_PI (Square.cpp:57)

// (int)std::ceil(x) for a step count: must lie in [1, 2^30].
__device__ __forceinline__ bool ceil_steps(double x, int& out) {
    const double c = ceil(x);
    if (!(c >= 1.0 && c <= 1073741824.0)) return false;
    out = (int)c;
    return true;
}

// (end - start).norm() of a 2-vector: sqrt(x*x + y*y)   (Square.cpp:69, Reciprocating.cpp:38, M.cpp:50)
__device__ __forceinline__ double norm2(double dx, double dy) {
    return __dsqrt_rn(dadd(dmul(dx, dx), dmul(dy, dy)));
}

struct PolyPlan {
    PolyHead head;
    int count[TGX_POLY_MAX_LEGS];
    uint32_t status;
};

// Writes record `slot` of this trajectory (four 16-byte stores); rec == nullptr: counting only.
__device__ __forceinline__ void put_leg(int4* rec, int slot, const PolyLeg& l) {
    if (!rec) return;
    const int4* src = reinterpret_cast<const int4*>(&l);
    int4* dst = rec + 4 * slot;
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = src[q];
}

__device__ __forceinline__ PolyLeg make_leg(double sx, double sy, double ex, double ey, double v, int steps, int i0) {
    PolyLeg l;
    l.sx = sx;
    l.sy = sy;
    l.dx = dsub(ex, sx);
    l.dy = dsub(ey, sy);
    l.heading = atan2(l.dy, l.dx);                      // atan2(end.y - start.y, end.x - start.x)
    double sn, cs;
    sincos(l.heading, &sn, &cs);
    l.vx = v * cs;                                      // goal.v.x = v * cos(heading)   (Square.cpp:100-101)
    l.vy = v * sn;
    l.steps = steps;
    l.i0 = i0;
    return l;
}

__device__ __forceinline__ PolyLeg make_point(double x, double y, double heading, double v) {
    // a one-sample record: i = 1 of 1 step over a zero-length leg, p = start + 1 * 0
    PolyLeg l;
    l.sx = x;
    l.sy = y;
    l.dx = 0.0;
    l.dy = 0.0;
    l.heading = heading;
    double sn, cs;
    sincos(heading, &sn, &cs);
    l.vx = v * cs;
    l.vy = v * sn;
    l.steps = 1;
    l.i0 = 1;
    return l;
}

// generateTraj of one polyline trajectory: leg records into `rec` (may be nullptr), structure into `pl`.
// Returns the sample count, 0 when the trajectory is empty or rejected (pl.status says which).
__device__ int poly_plan_one(const tgx_params& p, int64_t max_samples, const CurTable* __restrict__ tab, int4* rec,
                             PolyPlan& pl) {
    const tgx_polyline_params& q = p.u.poly;
    PolyHead& h = pl.head;
    h.type = p.type;
    h.n = 0;
    h.n_legs = 0;
    h.first_special = 0;
    h.last_special = 0;
    h.period = 0;
    h.pad[0] = h.pad[1] = 0;
    h.c0 = p.alt;
    h.c1 = 0.0;
    h.pad2[0] = h.pad2[1] = 0.0;
    pl.status = 0;
    for (int i = 0; i < TGX_POLY_MAX_LEGS; ++i) pl.count[i] = 0;
    if (!poly_params_ok(p)) {
        pl.status = TGX_ST_BAD_PARAM;
        return 0;
    }
    const double dt = p.dt, v = q.v_goal, T = q.t_traj;
    // the clock: number of `t += dt` iterations until t >= t_traj
    const long long H = hold_steps(T, dt, (long long)max_samples, tab);
    if (H < 0) {
        pl.status = TGX_ST_TOO_LONG;
        return 0;
    }
    double c = q.cos_o, s = q.sin_o;
    if (!(p.n_vgoals & TGX_POLY_TRIG_GIVEN)) sincos(q.orientation, &s, &c);
    long long N = 0;
    bool ok = true;

    if (p.type == TGX_SQUARE || p.type == TGX_RECTANGLE) {
        const bool rect = p.type == TGX_RECTANGLE;
        const double ha = ddiv(q.g[0], 2.0), hb = rect ? ddiv(q.g[1], 2.0) : ha;          // Square.cpp:27
        const double cx = rect ? q.g[2] : q.g[1], cy = rect ? q.g[3] : q.g[2];
        const double ux[4] = {-ha, ha, ha, -ha}, uy[4] = {hb, hb, -hb, -hb};               // :30-33
        double X[4], Y[4];
        for (int i = 0; i < 4; ++i) {                                                      // :39-44
            X[i] = dadd(dsub(dmul(c, ux[i]), dmul(s, uy[i])), cx);
            Y[i] = dadd(dadd(dmul(s, ux[i]), dmul(c, uy[i])), cy);
        }
        const double perimeter = rect ? dmul(2.0, dadd(q.g[0], q.g[1])) : dmul(4.0, q.g[0]);   // :49
        const double lapsd = ceil(ddiv(T, ddiv(perimeter, v)));                            // :50-51
        if (!(lapsd >= -2147483648.0 && lapsd <= 1073741824.0)) ok = false;
        long long per_lap = 0;
        for (int side = 0; side < 4 && ok; ++side) {
            const int nx = (side + 1) & 3;
            const double len = norm2(dsub(X[nx], X[side]), dsub(Y[nx], Y[side]));         // :69
            int steps = 0;
            ok = ceil_steps(ddiv(ddiv(len, v), dt), steps);                                // :70-71
            if (!ok) break;
            put_leg(rec, kPolySlotLeg0 + side, make_leg(X[side], Y[side], X[nx], Y[nx], v, steps, 1));
            pl.count[side] = steps;                                                        // step = 1 .. steps (:73)
            per_lap += steps;
        }
        if (ok) {
            h.n_legs = 4;
            h.first_special = 1;
            put_leg(rec, kPolySlotFirst, make_point(X[0], Y[0], dadd(q.orientation, kPi), v));   // :57-60
            const long long laps = lapsd > 0.0 ? (long long)lapsd : 0;
            const long long cap = laps * per_lap;              // < 2^62
            N = 1 + (cap < H ? cap : H);
        }
    } else if (p.type == TGX_RECIPROCATING) {
        const double Ax = q.g[0], Ay = q.g[1], Bx = q.g[3], By = q.g[4];
        const double th_fwd = atan2(dsub(By, Ay), dsub(Bx, Ax));                           // Reciprocating.cpp:17-18
        const double th_rev = atan2(dsub(Ay, By), dsub(Ax, Bx));
        const double dist = norm2(dsub(Bx, Ax), dsub(By, Ay));                             // :38
        int steps = 0;
        ok = ceil_steps(ddiv(dist, dmul(v, dt)), steps);                                   // :39
        if (ok) {
            PolyLeg f = make_leg(Ax, Ay, Bx, By, v, steps, 0);
            f.heading = th_fwd;
            PolyLeg r = make_leg(Bx, By, Ax, Ay, v, steps, 0);
            r.heading = th_rev;
            const PolyLeg atB = make_point(Bx, By, th_rev, 0.0);                           // :50-56 yaw flip, v = 0
            const PolyLeg atA = make_point(Ax, Ay, th_fwd, 0.0);
            put_leg(rec, kPolySlotLeg0 + 0, f);
            put_leg(rec, kPolySlotLeg0 + 1, atB);
            put_leg(rec, kPolySlotLeg0 + 2, r);
            put_leg(rec, kPolySlotLeg0 + 3, atA);
            pl.count[0] = steps + 1; pl.count[1] = 1; pl.count[2] = steps + 1; pl.count[3] = 1;
            h.n_legs = 4;
            // every leg pass is followed by an unconditional flip goal, also when t_traj cut the pass short
            const long long L = (long long)steps + 2;
            const long long rem = H % L;
            N = H + (rem ? 1 : 0);
            if (rem && H > 0) {
                h.last_special = 1;
                const bool fwd_pass = ((H - 1) % (2 * L)) < L;
                put_leg(rec, kPolySlotLast, fwd_pass ? atB : atA);
            }
        }
    } else if (p.type == TGX_BOUNCE) {
        const double Az = q.g[2], Bz = q.g[3];
        int steps = 0;
        ok = ceil_steps(ddiv(fabs(dsub(Bz, Az)), dmul(v, dt)), steps);                     // Bounce.cpp:35-36
        if (ok) {
            PolyLeg up, dn;
            up.sx = Az; up.sy = 0.0; up.dx = dsub(Bz, Az); up.dy = 0.0; up.heading = q.orientation;   // :32
            up.vx = (Bz > Az) ? v : -v; up.vy = 0.0; up.steps = steps; up.i0 = 0;                       // :40
            dn = up;
            dn.sx = Bz; dn.dx = dsub(Az, Bz); dn.vx = (Az > Bz) ? v : -v;
            put_leg(rec, kPolySlotLeg0 + 0, up);
            put_leg(rec, kPolySlotLeg0 + 1, dn);
            pl.count[0] = pl.count[1] = steps + 1;
            h.n_legs = 2;
            h.c0 = q.g[0];
            h.c1 = q.g[1];
            N = H;
        }
    } else {
        const double cx = q.g[0], cy = q.g[1], len = q.g[2], wid = q.g[3];
        const double hw = ddiv(wid, 2.0), hl = ddiv(len, 2.0);
        double bx[6], by[6];
        int np;
        if (p.type == TGX_M) {                                                              // M.cpp:20-26
            np = 5;
            bx[0] = -hw; by[0] = -hl; bx[1] = -hw; by[1] = hl; bx[2] = 0.0; by[2] = -hl;
            bx[3] = hw; by[3] = hl; bx[4] = hw; by[4] = -hl;
        } else if (p.type == TGX_I) {                                                       // I.cpp:28-35
            np = 6;
            bx[0] = -hw; by[0] = hl; bx[1] = hw; by[1] = hl; bx[2] = 0.0; by[2] = hl;
            bx[3] = 0.0; by[3] = -hl; bx[4] = -hw; by[4] = -hl; bx[5] = hw; by[5] = -hl;
        } else {                                                                            // T.cpp:28-33
            np = 4;
            bx[0] = -hw; by[0] = hl; bx[1] = hw; by[1] = hl; bx[2] = 0.0; by[2] = hl; bx[3] = 0.0; by[3] = -hl;
        }
        double X[6], Y[6];
        for (int i = 0; i < np; ++i) {                                                      // M.cpp:32-37
            X[i] = dadd(dsub(dmul(c, bx[i]), dmul(s, by[i])), cx);
            Y[i] = dadd(dadd(dmul(s, bx[i]), dmul(c, by[i])), cy);
        }
        const double vdt = dmul(v, dt);
        const int nseg = np - 1;
        for (int leg = 0; leg < 2 * nseg && ok; ++leg) {                                    // forward lap, then reversed
            const bool fwd = leg < nseg;
            const int sgm = fwd ? leg : leg - nseg;
            const int a = fwd ? sgm : np - 1 - sgm, b = fwd ? sgm + 1 : np - 2 - sgm;      // M.cpp:45-48
            int steps = 0;
            ok = ceil_steps(ddiv(norm2(dsub(X[b], X[a]), dsub(Y[b], Y[a])), vdt), steps);  // :50-51
            if (!ok) break;
            put_leg(rec, kPolySlotLeg0 + leg, make_leg(X[a], Y[a], X[b], Y[b], v, steps, 0));
            pl.count[leg] = steps + 1;                                                      // i = 0 .. steps (:52)
        }
        if (ok) {
            h.n_legs = 2 * nseg;
            N = H;
        }
    }
    if (!ok || N > max_samples) {
        pl.status = TGX_ST_TOO_LONG;
        h.n_legs = 0;
        h.first_special = h.last_special = 0;
        return 0;
    }
    long long period = 0;
    for (int i = 0; i < h.n_legs; ++i) period += pl.count[i];
    if (period > 0x7fffffffLL) {
        pl.status = TGX_ST_TOO_LONG;
        return 0;
    }
    h.period = (int)period;
    h.n = (int)N;
    return h.n;
}

// Trajectory::isPointInsideBounds (Trajectory.hpp:50-57).
__device__ bool poly_point_inside(const double* box, double x, double y, double z) {
    if (x < box[0] || x > box[1]) return false;
    if (y < box[2] || y > box[3]) return false;
    if (z < box[4] || z > box[5]) return false;
    return true;
}

// trajectoryInsideBounds of the family: Square.cpp:139-157, Rectangle.cpp:139-159, Reciprocating.cpp:109-122,
// Bounce.cpp:105-120, M.cpp:115-145, I.cpp:123-153, T.cpp:121-149.
__device__ bool poly_inside_bounds(const tgx_params& p, const double* box) {
    const tgx_polyline_params& q = p.u.poly;
    if (p.type == TGX_RECIPROCATING)
        return poly_point_inside(box, q.g[0], q.g[1], q.g[2]) && poly_point_inside(box, q.g[3], q.g[4], q.g[5]);
    if (p.type == TGX_BOUNCE)
        return poly_point_inside(box, q.g[0], q.g[1], q.g[2]) && poly_point_inside(box, q.g[0], q.g[1], q.g[3]);
    double c = q.cos_o, s = q.sin_o;
    if (!(p.n_vgoals & TGX_POLY_TRIG_GIVEN)) sincos(q.orientation, &s, &c);
    if (p.type == TGX_SQUARE || p.type == TGX_RECTANGLE) {
        const bool rect = p.type == TGX_RECTANGLE;
        const double ha = ddiv(q.g[0], 2.0), hb = rect ? ddiv(q.g[1], 2.0) : ha;
        const double cx = rect ? q.g[2] : q.g[1], cy = rect ? q.g[3] : q.g[2];
        const double ux[4] = {-ha, ha, ha, -ha}, uy[4] = {hb, hb, -hb, -hb};
        for (int i = 0; i < 4; ++i) {
            // cx_ + c * x - s * y,  cy_ + s * x + c * y   (Square.cpp:147-150)
            const double x = dsub(dadd(cx, dmul(c, ux[i])), dmul(s, uy[i]));
            const double y = dadd(dadd(cy, dmul(s, ux[i])), dmul(c, uy[i]));
            if (!poly_point_inside(box, x, y, p.alt)) return false;
        }
        return true;
    }
    const double cx = q.g[0], cy = q.g[1], len = q.g[2], wid = q.g[3];
    const double hw = ddiv(wid, 2.0), hl = ddiv(len, 2.0);
    double px[6], py[6];
    int np;
    if (p.type == TGX_M) {                                                                  // M.cpp:120-126
        np = 5;
        px[0] = dsub(cx, hw); py[0] = dsub(cy, hl); px[1] = dsub(cx, hw); py[1] = dadd(cy, hl);
        px[2] = cx; py[2] = dsub(cy, hl); px[3] = dadd(cx, hw); py[3] = dadd(cy, hl);
        px[4] = dadd(cx, hw); py[4] = dsub(cy, hl);
    } else if (p.type == TGX_I) {                                                           // I.cpp:128-135
        np = 6;
        px[0] = dsub(cx, hw); py[0] = dadd(cy, hl); px[1] = dadd(cx, hw); py[1] = dadd(cy, hl);
        px[2] = cx; py[2] = dadd(cy, hl); px[3] = cx; py[3] = dsub(cy, hl);
        px[4] = dsub(cx, hw); py[4] = dsub(cy, hl); px[5] = dadd(cx, hw); py[5] = dsub(cy, hl);
    } else {                                                                                // T.cpp:126-131
        np = 4;
        px[0] = dsub(cx, hw); py[0] = dadd(cy, hl); px[1] = dadd(cx, hw); py[1] = dadd(cy, hl);
        px[2] = cx; py[2] = dadd(cy, hl); px[3] = cx; py[3] = dsub(cy, hl);
    }
    for (int i = 0; i < np; ++i) {                                                          // M.cpp:131-143
        const double xs = dsub(px[i], cx), ys = dsub(py[i], cy);
        const double xr = dadd(dsub(dmul(c, xs), dmul(s, ys)), cx);
        const double yr = dadd(dadd(dmul(s, xs), dmul(c, ys)), cy);
        if (!poly_point_inside(box, xr, yr, p.alt)) return false;
    }
    return true;
}

__device__ __forceinline__ tgx_params poly_load_params(const tgx_params* params, int64_t i) {
    tgx_params p;
    const double2* src = reinterpret_cast<const double2*>(params + i);
    double2* dst = reinterpret_cast<double2*>(&p);
#pragma unroll
    for (int q = 0; q < 8; ++q) dst[q] = __ldg(src + q);
    return p;
}

}  // namespace

// ---- planning kernel ------------------------------------------------------------------------------------------
// One thread per trajectory.  recs == nullptr: counts / status / legs only (tgx_count).  A trajectory of another family
// gets TGX_ST_WRONG_PLANNER and no samples — or, with skip_foreign, is left alone entirely (its entries of counts /
// status / legs are not written: the host-buffer calls run both planners over a mixed batch).
__global__ void __launch_bounds__(128)
plan_poly_kernel(const tgx_params* __restrict__ params, int64_t n, tgx_limits lim, int has_lim, int64_t max_samples,
                 int tile_shift, const CurTable* __restrict__ tab, int4* __restrict__ recs,
                 int32_t* __restrict__ counts, uint32_t* __restrict__ status, int32_t* __restrict__ counts2,
                 uint32_t* __restrict__ status2, tgx_polyline_legs* __restrict__ legs, int32_t* __restrict__ ntile,
                 PlanStats* __restrict__ stats, int skip_foreign) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const tgx_params p = poly_load_params(params, i);
    int4* rec = recs ? recs + i * (int64_t)(4 * kPolyRecs) : nullptr;
    PolyPlan pl;
    int cnt = 0;
    const bool mine = TGX_IS_POLYLINE(p.type);
    if (mine) {
        cnt = poly_plan_one(p, max_samples, tab, rec, pl);
        if (has_lim && lim.check_box && !(pl.status & TGX_ST_BAD_PARAM) && !poly_inside_bounds(p, lim.box))
            pl.status |= TGX_ST_OUTSIDE_BOUNDS;
    } else {
        pl.head = PolyHead{};
        pl.head.type = p.type;
        pl.status = TGX_ST_WRONG_PLANNER;
        for (int q = 0; q < TGX_POLY_MAX_LEGS; ++q) pl.count[q] = 0;
    }
    if (rec) {
        const int4* src = reinterpret_cast<const int4*>(&pl.head);
#pragma unroll
        for (int q = 0; q < 4; ++q) rec[q] = src[q];
    }
    const int tiles = (cnt + (1 << tile_shift) - 1) >> tile_shift;
    if (mine || !skip_foreign) {
        if (counts) counts[i] = cnt;
        if (status) status[i] = pl.status;
        if (legs) {
            tgx_polyline_legs L;
            L.n = cnt;
            L.n_legs = pl.head.n_legs;
            L.first_special = pl.head.first_special;
            L.last_special = pl.head.last_special;
            L.period = pl.head.period;
            for (int q = 0; q < TGX_POLY_MAX_LEGS; ++q) L.count[q] = pl.count[q];
            L.reserved = 0;
            const int4* src = reinterpret_cast<const int4*>(&L);
            int4* dst = reinterpret_cast<int4*>(legs + i);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = src[q];
        }
    }
    if (counts2) counts2[i] = cnt;
    if (status2) status2[i] = mine ? pl.status : 0u;     // the engine's own copy feeds tgx_feasibility
    if (ntile) ntile[i] = tiles;
    if (stats) {
        const unsigned mask = __activemask();
        const unsigned tot = __reduce_add_sync(mask, (unsigned)cnt);
        const unsigned tt = __reduce_add_sync(mask, (unsigned)tiles);
        const int mt = __reduce_max_sync(mask, tiles);
        const int mn = __reduce_max_sync(mask, cnt);
        if ((int)(threadIdx.x & 31) == __ffs(mask) - 1) {
            atomicAdd(&stats->total_samples, (unsigned long long)tot);
            atomicAdd(&stats->total_tiles, (unsigned long long)tt);
            atomicMax(&stats->max_ntile, mt);
            atomicMax(&stats->max_n, mn);
        }
    }
}

// Ragged batches: the work list, one entry per (trajectory, tile).
__global__ void __launch_bounds__(256)
poly_tiles_kernel(int64_t n, const int32_t* __restrict__ ntile, const int64_t* __restrict__ tile_off, int tile_shift,
                  Tile* __restrict__ tiles) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int m = ntile[i];
    Tile* dst = tiles + tile_off[i];
    for (int t = 0; t < m; ++t) {
        Tile e;
        e.traj = (int32_t)i;
        e.k_lo = t << tile_shift;
        e.seg_begin = 0;
        e.nseg = 1;
        dst[t] = e;
    }
}

// ---- evaluation kernel -----------------------------------------------------------------------------------------
// One CTA per tile of THREADS*SPT consecutive samples of one trajectory, SPT adjacent samples per thread, channel by
// channel with one vector store per thread per channel (store.cuh), exactly like eval.cu.  Per sample: one integer
// modulo to find the position in the period, a scan over <= 10 leg lengths in shared memory, one correctly rounded
// division and two multiply-add pairs WITHOUT contraction (the reference's `start + frac * (end - start)` is two
// roundings per coordinate).  No transcendental on the per-sample path: heading, cos and sin are per leg.
// RECORDS: the tile is walked in PASSES passes of THREADS*SPT samples; in a pass a warp owns 32*SPT consecutive samples
// (lane l: l, l + 32, ...), stages whole records and sends them with TMA (RecTma, store.cuh), as eval.cu does.
template <int THREADS, int SPT, bool STORE, bool REDUCE, bool RECORDS = false, int PASSES = 1>
__global__ void __launch_bounds__(THREADS, 768 / THREADS)
eval_poly_kernel(PolyView pv, OutView out, double* __restrict__ max_v, double* __restrict__ max_a,
                 const __grid_constant__ RecOut ro = RecOut{}) {
    static_assert(RECORDS || PASSES == 1, "only the record mode walks a tile in passes");
    constexpr int TILE = THREADS * SPT * PASSES;
    constexpr int KS = RECORDS ? 32 : 1;                  // distance between a thread's samples
    extern __shared__ __align__(16) double2 s_dyn[];      // RECORDS: the warps' record staging areas (store.cuh)
    __shared__ __align__(16) int4 s_raw[4 * kPolyRecs];
    __shared__ double s_red[THREADS / 32];

    int traj, k_lo;
    if (pv.tiles) {
        const int4 tw = __ldg(reinterpret_cast<const int4*>(pv.tiles) + blockIdx.x);
        traj = tw.x;
        k_lo = tw.y;
    } else {
        traj = (int)(blockIdx.x / (unsigned)pv.tile_slab);
        k_lo = ((int)blockIdx.x - traj * pv.tile_slab) * TILE;
    }
    if (threadIdx.x < 4 * kPolyRecs) s_raw[threadIdx.x] = __ldg(pv.recs + (size_t)traj * (4 * kPolyRecs) + threadIdx.x);
    __syncthreads();
    const PolyHead& hd = *reinterpret_cast<const PolyHead*>(s_raw);
    const PolyLeg* legs = reinterpret_cast<const PolyLeg*>(s_raw) ;   // slot s is legs[s] (slot 0 is the head)
    const int n = hd.n;
    if (k_lo >= n) return;                                            // an empty slot (whole CTA)

    int limit = n;
    if (STORE && out.capacity < (int64_t)limit) limit = (int)out.capacity;
    // a trajectory's records start at its offset and end where its row capacity, or the record buffer, ends: offsets that
    // point outside the buffer write nothing (the TMA row coordinate is 32 bits: it must never wrap into the buffer)
    int64_t rec_off = 0;
    if (RECORDS) {
        if (ro.capacity < (int64_t)limit) limit = (int)ro.capacity;
        rec_off = ro.offset ? __ldg(ro.offset + traj) : (int64_t)traj * ro.stride;
        if (rec_off < 0 || rec_off >= ro.total) limit = 0;
        else if (rec_off + (int64_t)limit > ro.total) limit = (int)(ro.total - rec_off);
    }
    double best_v2 = 0.0;

    RecTma<SPT> stager;
    if (RECORDS) {
        if ((threadIdx.x & 31) == 0)
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&ro.tmap)) : "memory");
        // the warp's private staging area: 32*SPT records, 1024-byte aligned for the 128-byte swizzle
        const uint32_t dyn = ((uint32_t)__cvta_generic_to_shared(s_dyn) + 1023u) & ~1023u;
        stager.init(dyn + (uint32_t)(threadIdx.x >> 5) * (uint32_t)RecTma<SPT>::kBytesPerWarp, (int)threadIdx.x & 31);
    }

#pragma unroll 1
    for (int pass = 0; pass < PASSES; ++pass) {
    // first sample of this warp's block of 32*SPT (RECORDS) and of this thread
    const int wk0 = k_lo + pass * (THREADS * SPT) + ((int)threadIdx.x >> 5) * (32 * SPT);
    const int k0 = RECORDS ? wk0 + ((int)threadIdx.x & 31) : k_lo + SPT * (int)threadIdx.x;
    const int nvalid = (REDUCE ? n : limit) - k0;

    // RECORDS: all 32 lanes of a warp with any sample to write take part in staging the warp's records
    if (RECORDS ? (wk0 < limit) : (nvalid > 0 || (STORE && k0 < ((limit + (kFillAlign - 1)) & ~(kFillAlign - 1))))) {
        const int first = hd.first_special, n_legs = hd.n_legs;
        const bool bounce = hd.type == TGX_BOUNCE;
        // position of sample max(k, first) in the period: leg l, step i
        int l, i;
        auto locate = [&](int k) {
            int m = k - first;
            if (m < 0) m = 0;
            int cyc = (int)((unsigned)m % (unsigned)hd.period);
            l = 0;
            for (; l + 1 < n_legs; ++l) {
                const int c = legs[kPolySlotLeg0 + l].steps + 1 - legs[kPolySlotLeg0 + l].i0;
                if (cyc < c) break;
                cyc -= c;
            }
            i = legs[kPolySlotLeg0 + l].i0 + cyc;
        };
        locate(k0);

        double px[SPT], py[SPT], vx[SPT], vy[SPT], psi[SPT];
#pragma unroll
        for (int u = 0; u < SPT; ++u) {
            const int k = k0 + u * KS;
            if (RECORDS && u > 0) locate(k);                                      // the thread's samples are 32 apart
            const bool sp_first = first && k == 0;
            const bool sp_last = hd.last_special && k == n - 1;
            const int slot = sp_first ? kPolySlotFirst : (sp_last ? kPolySlotLast : kPolySlotLeg0 + l);
            const PolyLeg& L = legs[slot];
            const int ii = (sp_first || sp_last) ? L.i0 : i;
            const double frac = __ddiv_rn((double)ii, (double)L.steps);           // (double)i / steps
            px[u] = __dadd_rn(L.sx, __dmul_rn(frac, L.dx));                        // start + frac * (end - start)
            py[u] = __dadd_rn(L.sy, __dmul_rn(frac, L.dy));
            vx[u] = L.vx;
            vy[u] = L.vy;
            psi[u] = L.heading;
            if (!sp_first) {                                                       // advance along the pattern
                if (++i > legs[kPolySlotLeg0 + l].steps) {
                    l = (l + 1 == n_legs) ? 0 : l + 1;
                    i = legs[kPolySlotLeg0 + l].i0;
                }
            }
        }

        double* row = nullptr;
        int nst = 0, nfill = 0;
        if (STORE) {
            const int64_t toff = out.traj_offset ? __ldg(out.traj_offset + traj) : (int64_t)traj * out.traj_stride;
            row = out.base + toff + k0;
            nst = limit - k0;
            // the row's last 32-byte sector is completed with zeros when it lies inside the row's capacity (store.cuh)
            const int64_t lim4 = ((int64_t)limit + (kFillAlign - 1)) & ~(int64_t)(kFillAlign - 1);
            nfill = (int)((lim4 <= out.capacity ? lim4 : (int64_t)limit) - k0);
        }
        const uint32_t mask = out.channel_mask;
        const int64_t cs = out.chan_stride;
        if (RECORDS) stager.begin_pass();
#define TGX_STORE(CH, ARR)                                                                                \
    do {                                                                                                  \
        if (RECORDS) {                                                                                    \
            /* the previous pass's records must have left the staging area before the first write */     \
            if ((CH) == TGX_PX && pass > 0) stager.wait_read();                                           \
            stager.template put<(CH)>(ARR, ro);                                                           \
            if ((CH) == TGX_DPSI) {                                                                       \
                stager.put_tail(traj, k0, n);                                                             \
                stager.flush(&ro.tmap, ro.base + rec_off, rec_off, wk0, limit);                           \
            }                                                                                             \
        } else if (STORE && nfill > 0 && (mask & (1u << (CH)))) {                                         \
            store_channel<SPT>(row + (CH) * cs, ARR, nst, nfill);                                         \
        }                                                                                                 \
    } while (0)
        double o[SPT], z[SPT];
#pragma unroll
        for (int u = 0; u < SPT; ++u) z[u] = 0.0;
        if (bounce) {
            // Bounce::createBounceGoal, Bounce.cpp:54-72: p = (cx, cy, z), v = (0, 0, vz), psi = orientation
#pragma unroll
            for (int u = 0; u < SPT; ++u) o[u] = hd.c0;
            TGX_STORE(TGX_PX, o);
#pragma unroll
            for (int u = 0; u < SPT; ++u) o[u] = hd.c1;
            TGX_STORE(TGX_PY, o);
            TGX_STORE(TGX_PZ, px);
            TGX_STORE(TGX_VX, z);
            TGX_STORE(TGX_VY, z);
            TGX_STORE(TGX_VZ, vx);
        } else {
            // createSquareGoal and its copies (Square.cpp:94-110): p = (x, y, alt), v = v*(cos, sin) heading
            TGX_STORE(TGX_PX, px);
            TGX_STORE(TGX_PY, py);
#pragma unroll
            for (int u = 0; u < SPT; ++u) o[u] = hd.c0;
            TGX_STORE(TGX_PZ, o);
            TGX_STORE(TGX_VX, vx);
            TGX_STORE(TGX_VY, vy);
            TGX_STORE(TGX_VZ, z);
        }
        TGX_STORE(TGX_AX, z);            // accel = 0 on every generateTraj sample; jerk is never assigned
        TGX_STORE(TGX_AY, z);
        TGX_STORE(TGX_AZ, z);
        TGX_STORE(TGX_JX, z);
        TGX_STORE(TGX_JY, z);
        TGX_STORE(TGX_JZ, z);
        TGX_STORE(TGX_PSI, psi);
        TGX_STORE(TGX_DPSI, z);
#undef TGX_STORE
        if (REDUCE) {
#pragma unroll
            for (int u = 0; u < SPT; ++u)
                if (u < nvalid) best_v2 = max_nn(best_v2, fma(vx[u], vx[u], vy[u] * vy[u]));
        }
    }
    }   // pass
    if (RECORDS) stager.wait_read();               // the TMA unit must have read the staging area before the CTA exits

    if (REDUCE) {
        best_v2 = warp_max(best_v2);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) s_red[warp] = best_v2;
        __syncthreads();
        if (warp == 0) {
            double a = lane < THREADS / 32 ? s_red[lane] : 0.0;
            a = warp_max(a);
            if (lane == 0 && max_v) atomic_max_nonneg(max_v + traj, sqrt(a));
            // max |a| stays at the 0 the caller initialised it with: the family has no acceleration
        }
    }
}

// ---- host-side launchers (called from engine.cu) -----------------------------------------------------------------

cudaError_t launch_plan_poly(const tgx_params* params, int64_t n, const tgx_limits* lim, int64_t max_samples,
                             int tile_shift, const void* cur_table, void* recs, int32_t* counts, uint32_t* status,
                             int32_t* counts2, uint32_t* status2, tgx_polyline_legs* legs, int32_t* ntile,
                             PlanStats* stats, bool skip_foreign, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    tgx_limits l{};
    if (lim) l = *lim;
    plan_poly_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(
        params, n, l, lim ? 1 : 0, max_samples, tile_shift, static_cast<const CurTable*>(cur_table),
        static_cast<int4*>(recs), counts, status, counts2, status2, legs, ntile, stats, skip_foreign ? 1 : 0);
    return cudaGetLastError();
}

size_t poly_rec_bytes() { return (size_t)kPolyRecs * 64; }

cudaError_t launch_poly_tiles(int64_t n, const int32_t* ntile, const int64_t* tile_off, int tile_shift, Tile* tiles,
                              cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    poly_tiles_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, ntile, tile_off, tile_shift, tiles);
    return cudaGetLastError();
}

template <int THREADS, int SPT>
static cudaError_t launch_eval_poly_t(const PolyView& pv, int64_t ntiles, const OutView& out, bool store,
                                      double* max_v, double* max_a, cudaStream_t stream) {
    const bool reduce = max_v || max_a;
    const unsigned grid = (unsigned)ntiles;
    if (store && reduce)
        eval_poly_kernel<THREADS, SPT, true, true><<<grid, THREADS, 0, stream>>>(pv, out, max_v, max_a);
    else if (store)
        eval_poly_kernel<THREADS, SPT, true, false><<<grid, THREADS, 0, stream>>>(pv, out, max_v, max_a);
    else
        eval_poly_kernel<THREADS, SPT, false, true><<<grid, THREADS, 0, stream>>>(pv, out, max_v, max_a);
    return cudaGetLastError();
}

// Record mode: 2 samples per thread per pass, TILE / (2 * THREADS) passes.
template <int THREADS, int TILE>
static cudaError_t launch_eval_poly_records_t(const PolyView& pv, int64_t ntiles, const RecOut& ro,
                                              cudaStream_t stream) {
    constexpr int SPP = 2;
    auto kernel = eval_poly_kernel<THREADS, SPP, false, false, true, TILE / (SPP * THREADS)>;
    const int smem = (THREADS / 32) * RecTma<SPP>::kBytesPerWarp + 1024;   // whole records + alignment slack
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    kernel<<<(unsigned)ntiles, THREADS, smem, stream>>>(pv, OutView{}, nullptr, nullptr, ro);
    return cudaGetLastError();
}

cudaError_t launch_eval_poly_records(const PolyView& pv, int64_t ntiles, int tile_shift, int spt, const RecOut& ro,
                                     cudaStream_t stream) {
    if (ntiles <= 0) return cudaSuccess;
    if (ntiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    (void)spt;                 // record kernels: always 128 threads walking the tile in passes of 256 samples (eval.cu)
    const int tile = 1 << tile_shift;
    if (tile == 512) return launch_eval_poly_records_t<128, 512>(pv, ntiles, ro, stream);
    if (tile == 1024) return launch_eval_poly_records_t<128, 1024>(pv, ntiles, ro, stream);
    return cudaErrorInvalidConfiguration;
}

cudaError_t launch_eval_poly(const PolyView& pv, int64_t ntiles, int tile_shift, int spt, const OutView& out,
                             bool store, double* max_v, double* max_a, cudaStream_t stream) {
    if (ntiles <= 0) return cudaSuccess;
    if (ntiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    const int threads = (1 << tile_shift) / spt;
    if (threads == 128 && spt == 4) return launch_eval_poly_t<128, 4>(pv, ntiles, out, store, max_v, max_a, stream);
    if (threads == 256 && spt == 2) return launch_eval_poly_t<256, 2>(pv, ntiles, out, store, max_v, max_a, stream);
    if (threads == 256 && spt == 4) return launch_eval_poly_t<256, 4>(pv, ntiles, out, store, max_v, max_a, stream);
    return cudaErrorInvalidConfiguration;
}

}  // namespace tgx
