// transition.cu — the node's own setpoint recurrences around a trajectory, batched over vehicles.
//
//   TAKEOFF   TrajectoryGenerator.cpp:531-548    z <- saturate(z + vel_take*dt, 0, alt) until hovering
//   GOTO      :549-554, :574-586 -> simpleInterpolation :637-764 (both overloads are the same arithmetic)
//   LANDING   :588-599                            z <- z - vel_land*dt until z < 0
//   every tick: goal_.p <- saturate(goal_.p, room box) (:602-604), which feeds back into the next tick
//
// These are true recurrences (each tick reads the goal the previous tick published), a few hundred to a few thousand
// ticks long, so the parallelism is across vehicles: one thread replays one vehicle with non-contracted IEEE operations
// in the reference's order (this file is compiled with -fmad=false; sqrt and division are correctly rounded; there is
// no transcendental), which makes every record bit-identical to what the node publishes under perfect tracking.
// Each tick's 128-byte record leaves the thread as four 32-byte (full-sector) streaming stores into the vehicle's row.
// That store pattern — one line per vehicle per tick, as many row streams as vehicles — is what bounds the kernel: the
// same stores with no arithmetic at all take 0.87 of its time (tools/wbw/streams.cu).
#include <cuda_runtime.h>

#include "tgx_internal.cuh"

namespace tgx {

namespace {

constexpr double kPi = 3.14159265358979323846;   // M_PI

__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double std_min(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double std_max(double a, double b) { return (a < b) ? b : a; }

// TrajectoryGenerator::saturate, :773-780.
__device__ __forceinline__ double saturate(double val, double low, double high) {
    if (val > high) val = high;
    else if (val < low) val = low;
    return val;
}

// TrajectoryGenerator::wrap, :782-788.
__device__ __forceinline__ double wrap(double val) {
    if (val > kPi) val = dsub(val, dmul(2.0, kPi));
    if (val < -kPi) val = dadd(val, dmul(2.0, kPi));
    return val;
}

struct GoalState {
    double px, py, pz, vx, vy, psi, dpsi;
    bool power;
};

__device__ __forceinline__ void store_record(tgx_goal_record* dst, const GoalState& g, int traj, int k, int clamped,
                                             bool last) {
    // 16 doubles: p, v, a, j, psi, dpsi, {traj, k}, {power, mode_xy, mode_z, clamped, last}
    const unsigned long long w0 = (unsigned long long)(unsigned)traj | ((unsigned long long)(unsigned)k << 32);
    const unsigned long long w1 = (g.power ? 1ull : 0ull) | ((unsigned long long)clamped << 24) |
                                  ((unsigned long long)(last ? 1 : 0) << 32);
    double* p = reinterpret_cast<double*>(dst);
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(g.px), "d"(g.py), "d"(g.pz), "d"(g.vx)
                 : "memory");
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p + 4), "d"(g.vy), "d"(0.0), "d"(0.0), "d"(0.0)
                 : "memory");
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p + 8), "d"(0.0), "d"(0.0), "d"(0.0), "d"(0.0)
                 : "memory");
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p + 12), "d"(g.psi), "d"(g.dpsi),
                 "d"(__longlong_as_double((long long)w0)), "d"(__longlong_as_double((long long)w1))
                 : "memory");
}

__device__ __forceinline__ bool fin(double x) { return isfinite(x); }

__device__ bool transition_ok(const tgx_transition_params& t) {
    if (!(fin(t.dt) && t.dt > 0.0) || t.ticks < 0) return false;
    for (int i = 0; i < 3; ++i)
        if (!fin(t.start[i]) || !fin(t.dest[i])) return false;
    if (!fin(t.start_v[0]) || !fin(t.start_v[1]) || !fin(t.start_psi)) return false;
    if (!(fin(t.vel) && t.vel > 0.0)) return false;
    if (t.kind == TGX_TR_TAKEOFF) return true;
    if (t.kind == TGX_TR_LANDING) return fin(t.vel_yaw) && t.vel_yaw > 0.0;
    if (t.kind == TGX_TR_GOTO)
        return fin(t.dest_yaw) && fin(t.vel_yaw) && t.vel_yaw > 0.0 && fin(t.dist_thresh) && t.dist_thresh >= 0.0 &&
               fin(t.yaw_thresh) && t.yaw_thresh >= 0.0;
    return false;
}

}  // namespace

__global__ void __launch_bounds__(128)
transition_kernel(const tgx_transition_params* __restrict__ tparams, int64_t n, tgx_limits lim, int clamp,
                  int64_t max_samples, tgx_goal_record* __restrict__ records, int64_t rec_stride,
                  int64_t rec_capacity, int32_t* __restrict__ counts, uint32_t* __restrict__ status) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    tgx_transition_params t;
    {
        const double2* src = reinterpret_cast<const double2*>(tparams + i);
        double2* dst = reinterpret_cast<double2*>(&t);
#pragma unroll
        for (int q = 0; q < 8; ++q) dst[q] = __ldg(src + q);
    }
    uint32_t st = 0;
    // (32-bit tick counters: with 64-bit ones the compiler kept one of them in local memory and every tick reloaded it
    //  behind the store traffic — the LDL and the compare that waits for it were half of ncu's stall samples)
    int k = 0;
    const int cap = rec_capacity > 0x7fffffffLL ? 0x7fffffff : (int)rec_capacity;
    const int guard = max_samples > 0x7fffffffLL ? 0x7fffffff : (int)max_samples;
    if (!transition_ok(t)) {
        st = TGX_ST_BAD_PARAM;
    } else {
        GoalState g{t.start[0], t.start[1], t.start[2], t.start_v[0], t.start_v[1], t.start_psi, 0.0, true};
        tgx_goal_record* row = records ? records + i * rec_stride : nullptr;
        const int limit = t.ticks > 0 ? t.ticks : guard;
        double pose_z = g.pz;                       // perfect tracking: the pose is the previously published goal
        bool done = false;
        while (!done) {
            if (k >= limit) {
                if (t.ticks == 0) st |= TGX_ST_TOO_LONG;
                break;
            }
            bool ends = false;
            if (t.kind == TGX_TR_TAKEOFF) {
                const double alt = t.dest[2];
                // :540-548  (the hover test reads the pose; the goal is published unchanged on that tick)
                if (fabs(dsub(alt, pose_z)) < 0.10 && g.pz >= alt) ends = true;
                else g.pz = saturate(dadd(g.pz, dmul(t.vel, t.dt)), 0.0, alt);
            } else if (t.kind == TGX_TR_GOTO) {
                // simpleInterpolation, :637-699 / :702-764
                const double Dx = dsub(t.dest[0], g.px), Dy = dsub(t.dest[1], g.py);
                const double dist = __dsqrt_rn(dadd(dmul(Dx, Dx), dmul(Dy, Dy)));
                const double delta_yaw = wrap(dsub(t.dest_yaw, g.psi));
                const bool dist_far = dist > t.dist_thresh;
                const bool yaw_far = fabs(delta_yaw) > t.yaw_thresh;
                ends = !dist_far && !yaw_far;       // `finished`, computed from the CURRENT goal
                GoalState nx = g;
                nx.pz = t.dest[2];
                if (dist_far) {
                    const double c = ddiv(Dx, dist), s = ddiv(Dy, dist);
                    nx.px = dadd(g.px, dmul(dmul(c, t.vel), t.dt));
                    nx.py = dadd(g.py, dmul(dmul(s, t.vel), t.dt));
                    // `bool accel_for_vel = 0.1` is true, i.e. 1 (:653): the reference ramps the velocity by 1 * dt
                    nx.vx = std_min(dadd(g.vx, t.dt), dmul(c, t.vel));
                    nx.vy = std_min(dadd(g.vy, t.dt), dmul(s, t.vel));
                } else {
                    nx.px = t.dest[0];
                    nx.py = t.dest[1];
                    nx.vx = std_max(0.0, dsub(g.vx, t.dt));
                    nx.vy = std_max(0.0, dsub(g.vy, t.dt));
                }
                if (yaw_far) {
                    const double vy = delta_yaw >= 0.0 ? t.vel_yaw : -t.vel_yaw;
                    nx.psi = dadd(g.psi, dmul(vy, t.dt));
                    nx.dpsi = vy;
                } else {
                    nx.psi = t.dest_yaw;
                    nx.dpsi = 0.0;
                }
                g = nx;
            } else {
                // LANDING, :588-599
                const double vel_land = pose_z > dadd(t.dest[2], 0.4) ? t.vel : t.vel_yaw;
                g.pz = dsub(g.pz, dmul(vel_land, t.dt));
                if (g.pz < 0.0) {
                    g.power = false;
                    ends = true;
                }
            }
            int clamped = 0;
            if (clamp) {                            // :602-604, written back into goal_
                const double x = saturate(g.px, lim.box[0], lim.box[1]);
                const double y = saturate(g.py, lim.box[2], lim.box[3]);
                const double z = saturate(g.pz, lim.box[4], lim.box[5]);
                clamped = ((g.px > lim.box[1] || g.px < lim.box[0]) ? 1 : 0) |
                          ((g.py > lim.box[3] || g.py < lim.box[2]) ? 2 : 0) |
                          ((g.pz > lim.box[5] || g.pz < lim.box[4]) ? 4 : 0);
                g.px = x; g.py = y; g.pz = z;
            }
            done = ends && t.ticks == 0;
            const bool last = done || (t.ticks > 0 && k + 1 == limit);
            if (row && k < cap) store_record(row + k, g, (int)i, k, clamped, last);
            pose_z = g.pz;
            ++k;
        }
        if (k > cap && records) st |= TGX_ST_TRUNCATED;
    }
    if (counts) counts[i] = k;
    if (status) status[i] = st;
}

cudaError_t launch_transitions(const tgx_transition_params* tparams, int64_t n, const tgx_limits* lim,
                               int64_t max_samples, tgx_goal_record* records, int64_t rec_stride,
                               int64_t rec_capacity, int32_t* counts, uint32_t* status, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    tgx_limits l{};
    if (lim) l = *lim;
    transition_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(tparams, n, l, (lim && lim->check_box) ? 1 : 0,
                                                                     max_samples, records, rec_stride, rec_capacity,
                                                                     counts, status);
    return cudaGetLastError();
}

}  // namespace tgx
