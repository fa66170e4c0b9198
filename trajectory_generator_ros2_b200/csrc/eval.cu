// eval.cu — the evaluation kernel of libtgx: create{Circle,Line,Figure8}Goal for every (trajectory, k).
//
// One CTA per Tile (tile_size consecutive samples of one trajectory); phase plans (tgx_internal.cuh): one CTA per
// trajectory, which it walks in passes, its Seg records rebuilt from the trajectory's 256-byte PhaseRec.
// The CTA stages its trajectory's TrajRec (64 B) and the tile's Seg records (64 B each) in shared memory,
// every thread finds the segment its samples fall in, evaluates the closed form of the reference's
// recurrences inside that segment
//       v_k     = vb + j*dv                                   (j = k - kb; the clamped last step is exact)
//       S_k     = sum_{m=1..j} v_m = j*vb + dv*j(j+1)/2
//       theta_k = theta_b + S_k * (dt/r)          (Circle.cpp:50-51, Figure8.cpp:50-51: theta += (v/r)*dt)
//       p_k     = p_b + S_k * (cos|sin theta) * dt            (Line.cpp:97-98: p = last + v*c*dt)
// and then the reference's per-sample formulas (Circle.cpp:96-130, Line.cpp:91-115, Figure8.cpp:96-128) in
// fp64.  The kernel is store-bound: 112 B written per sample (128 B in record mode), ~0.3 B read.
//
// One kernel template, four ways out (template flags):
//   PTMA     struct-of-arrays planes through TMA (the default of tgx_eval for regular layouts): 128-thread CTAs walk
//            the tile in passes of 256 samples, a warp owns 64 consecutive samples per pass (lane l: l and l + 32),
//            stages [group][14 channels][32 samples] in its private shared memory and sends each group as one
//            UTMASTG.3D box of a [trajectory][channel][sample] tensor map over the caller's planes (PlaneTma, store.cuh)
//   STORE    the same planes with 128-bit (SPT=2) or 256-bit (SPT=4) streaming vector stores, one per thread per
//            channel, a thread owning SPT adjacent samples: irregular layouts (per-trajectory offsets, channel
//            subsets, short rows) and calls that also want the maxima
//   RECORDS  one clamped 128-byte tgx_goal_record per sample (tgx_eval_records), staged in the 128-byte TMA swizzle and
//            sent as 4 KiB UTMASTG.2D boxes (RecTma, store.cuh)
//   REDUCE   per-trajectory maxima of |v_k|^2 and |a_k|^2 (feasibility check, BASELINE.json configs 4-5): integer
//            REDUX.MAX warp reductions, one shared-memory hop per CTA and one atomicMax per tile; alone (no stores) it
//            runs as 64-thread CTAs walking the tile in passes, 16 samples per thread per reduction
#include <cuda_runtime.h>

#include "tgx_internal.cuh"
#include "store.cuh"

#ifndef TGX_REDUCE_LEGACY
#define TGX_REDUCE_LEGACY 0
#endif
#ifndef TGX_REDUCE_CTAS
#define TGX_REDUCE_CTAS 4      // the reduction-only instantiation is compiled for 4 * 256 threads per SM (64 registers; 5: 48 registers, spills in the slab / phase modes, 5.72 vs 5.77 ms)
#endif

namespace tgx {

namespace {

constexpr double kPiOver2 = 1.57079632679489661923;

// sin and cos of the orbit angle.  An orbit's angle is bounded by omega * t (the sample guard allows 2^24 samples of
// at most a few tenths of a radian each), so the argument reduction is three fused multiply-adds against a three-term
// split of pi/2 — each fma rounds once, the first one to a result of magnitude <= 1, so the reduced argument is off by
// ~1e-16 absolute for every |x| < 1e9 checked — followed by the fdlibm minimax polynomials on [-pi/4, pi/4] (< 1 ulp).
// No Payne-Hanek slow path, hence no stack frame.  Parity budget: |dp| <= 1e-9 m needs |d sin| <= 2e-10 at r = 5 m.
//
// The 17 constants live in __constant__ memory: a DFMA takes a constant-bank operand directly, whereas a 64-bit
// literal has to be materialised with two moves per use — in the reduction-only kernel (FP64- / issue-bound) those
// moves were 60 of the ~200 instructions per sample (ncu: IMAD 16 %, UMOV 14.5 % of all executed instructions).
__constant__ double kTrig[17] = {
    6.36619772367581382433e-01,                                       // 0: 2/pi
    6755399441055744.0,                                               // 1: 1.5 * 2^52
    -1.57079632679489655800e+00, -6.12323399573676603587e-17, -1.49738490485916983329e-33,   // 2-4: -pi/2, 3 terms
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,     // 5-10: sin, S6 .. S1
    -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,    // 11-16: cos, C6 .. C1
    2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02};

__device__ __forceinline__ void sincos_orbit(double x, double* sn, double* cs) {
    // k = rint(x * 2/pi) without a float->int conversion: adding 1.5 * 2^52 leaves k in the low mantissa bits
    const double shifted = fma(x, kTrig[0], kTrig[1]);
    const int q = __double2loint(shifted);                           // only the two low bits matter
    const double kd = shifted - kTrig[1];
    double r = fma(kd, kTrig[2], x);                                 // pi/2 = hi + mid + lo
    r = fma(kd, kTrig[3], r);
    r = fma(kd, kTrig[4], r);
    const double z = r * r;
    double ps = fma(z, kTrig[5], kTrig[6]);
    ps = fma(z, ps, kTrig[7]);
    ps = fma(z, ps, kTrig[8]);
    ps = fma(z, ps, kTrig[9]);
    ps = fma(z, ps, kTrig[10]);
    const double s = fma(r * z, ps, r);
    double pc = fma(z, kTrig[11], kTrig[12]);
    pc = fma(z, pc, kTrig[13]);
    pc = fma(z, pc, kTrig[14]);
    pc = fma(z, pc, kTrig[15]);
    pc = fma(z, pc, kTrig[16]);
    const double c = fma(z * z, pc, fma(z, -0.5, 1.0));
    const double a = (q & 1) ? c : s, b = (q & 1) ? s : c;           // quadrant: (s, c), (c, -s), (-s, -c), (-c, s)
    *sn = (q & 2) ? -a : a;
    *cs = ((q + 1) & 2) ? -b : b;
}

// Position of sample k inside segment sg: j = k - kb, the (double) step count fj the closed forms use, and v.
// On the step where the reference's std::min / std::max clamp fired (flag set, j == n) v is exactly the clamp
// value and the closed forms are evaluated at j-1 plus one clamped step.
struct SegPos {
    int j;
    bool last;     // j == n
    bool clamp;    // last && clamp flag
    double fj;     // clamp ? j-1 : j
    double tri;    // fj*(fj+1)/2, exact
    double v;
};

__device__ __forceinline__ SegPos seg_pos(const Seg& sg, int k) {
    SegPos q;
    q.j = k - sg.kb;
    q.last = (q.j == sg.n);
    q.clamp = q.last && (sg.flags & kSegClampLast);
    q.fj = (double)(q.clamp ? q.j - 1 : q.j);
    q.tri = 0.5 * (q.fj * (q.fj + 1.0));   // j(j+1) < 2^53: exact
    q.v = q.clamp ? sg.vclamp : fma(q.fj, sg.dv, sg.vb);
    return q;
}

}  // namespace

// The TrajRec of a phase plan's trajectory (what plan_one writes for the table path).
__device__ __forceinline__ TrajRec build_phase_rec(const PhasePlan& P, int n_total) {
    TrajRec r;
    r.type = P.type & kRecTypeMask;
    r.n = n_total;
    // both variants keep TrajRec.f[0..3] first; f[4] (dt / r | dt) and f[5] (1 / r | unused) follow
    r.f[0] = P.c.r; r.f[1] = P.c.cx; r.f[2] = P.c.cy; r.f[3] = P.c.alt;
    r.f[4] = P.c.dtr;
    r.f[5] = r.type == TGX_LINE ? 0.0 : P.c.rinv;
    r.f[6] = 0.0;
    return r;
}

// Phase plans: segment q of a trajectory from its self-contained record — exactly the Seg record plan.cu's replay
// writes into the segment table for the same trajectory (ramp(), hold()).
// Orbits (Circle / Figure8): the base speed is the level the preceding segments reached, the angle at both ends is the
// replayed one, and the per-step angle increment is rounded the way the planner rounds it for that kind of segment.  A
// ramp of n steps adds a*dt per step and clamps on its last step (Circle.cpp:47-54, 75-82); a hold keeps v (:63-71) and
// is one segment per binade of theta.
// Lines: ramp up from 0 to v_goal, clamped on its last step (Line.cpp:46-50); cruise (:57-62); ramp down, whose last step
// — the one the std::max clamp turns into 0 — is a one-sample segment of its own with the position forced to B
// (:65-68, :81-82).  The position at each segment's base is the replayed one.
// The lanes of a warp rebuild different segments of one trajectory: everything below is selects, no divergent branch
// (the trajectory type is the same for the whole CTA).
__device__ __forceinline__ void build_phase_segment(const PhasePlan& P, int q, Seg& sg, int& kend) {
    const int qp = q > 0 ? q - 1 : 0;
    kend = P.key[q];
    sg.kb = q ? P.key[qp] : 0;
    sg.n = kend - sg.kb;
    sg.pad = 0;
    const uint64_t kinds = (uint64_t)P.kinds[0] | ((uint64_t)P.kinds[1] << 32);
    if ((P.type & kRecTypeMask) == TGX_LINE) {
        const int kind = (int)(kinds >> (3 * q)) & 7;
        const double vg = P.c.lvg;
        sg.flags = kind == kPhaseLineUp ? kSegClampLast : (kind == kPhaseLineForced ? kSegForcePos : 0);
        sg.vb = (kind == kPhaseLineUp || kind == kPhaseLineForced) ? 0.0 : vg;
        sg.dv = kind == kPhaseLineUp ? P.c.ladt1 : (kind == kPhaseLineDown ? -P.c.ladt3 : 0.0);
        sg.vclamp = (kind == kPhaseLineUp || kind == kPhaseLineHold) ? vg : 0.0;
        sg.s0 = P.xy[q][0];
        sg.s1 = P.xy[q][1];
        sg.acc = kind == kPhaseLineUp ? P.c.la1 : (kind == kPhaseLineHold ? 0.0 : -P.c.la3);
        return;
    }
    // The speed at the start of segment q is the level the last ramp before it reached (holds, kind 0b10, keep it): the
    // highest two-bit field below q that is not a hold.  A ramp-down (0b11, v = 0) is put below segment 0 so that there
    // always is one.
    const uint64_t k2 = (kinds << 2) | 3ull;
    const uint64_t mask = (4ull << (2 * q)) - 1ull;              // the virtual field and segments 0 .. q-1
    const uint64_t below = k2 & mask;
    const uint64_t ramps = ((~below >> 1) | below) & 0x5555555555555555ull & mask;
    const int level = (int)(k2 >> (63 - __clzll((long long)ramps))) & 3;
    const bool moving = level != kPhaseKindDown;
    const double v = moving ? P.c.vg[level & 1] : 0.0;
    const double w = moving ? P.c.w[level & 1] : 0.0;
    const int kind = (int)(kinds >> (2 * q)) & 3;
    const bool hold = kind == kPhaseKindHold;
    const bool up = kind < kPhaseKindHold;
    const double thp = P.th[qp];
    const double thb = q ? thp : 0.0;
    sg.flags = hold ? 0 : kSegClampLast;
    sg.vb = v;
    sg.dv = hold ? 0.0 : (up ? P.c.adt : -P.c.adt);
    const double vup = P.c.vg[kind & 1];
    sg.vclamp = hold ? v : (up ? vup : 0.0);
    sg.s0 = thb;
    // hold: the exact progression step of theta += omega*dt with omega = v/r rounded first (Circle.cpp:65-67; plan.cu:
    // hold(), d0); ramp: v * (dt/r)  (plan.cu: ramp())
    const double s1h = __dsub_rn(__dadd_rn(thb, w), thb), s1r = __dmul_rn(v, P.c.dtr);
    sg.s1 = hold ? s1h : s1r;
    sg.acc = P.th[q];          // theta of the segment's last sample
}

// Stage a phase plan's trajectory: its PhaseRec (one round of loads), the PhaseExt row if it has one, then the Seg records
// and the TrajRec of the table path.  Returns the number of segments (0: rejected trajectory, whole CTA).
template <int NSEG>
__device__ __forceinline__ int stage_phase_plan(const TableView& tv, int traj, PhasePlan& P, Seg (&s_seg)[NSEG],
                                                int (&s_kend)[NSEG], TrajRec& s_rec) {
    // where the 16-byte chunks of the two records land in the shared-memory image (see PhasePlan)
    constexpr int kRecChunks = sizeof(PhaseRec) / 16, kExtChunks = sizeof(PhaseExt) / 16;
    constexpr int kThRec = offsetof(PhaseRec, th) / 16, kThPlan = offsetof(PhasePlan, th) / 16;
    constexpr int kKeyExtTo = offsetof(PhasePlan, key) / 16 + kPhaseBaseSegs * 4 / 16;
    constexpr int kThExt = offsetof(PhaseExt, th) / 16, kThExtTo = kThPlan + kPhaseBaseSegs * 8 / 16;
    static_assert(offsetof(PhaseRec, key) == offsetof(PhasePlan, key), "the head of the record is copied as it is");
    int4* dst = reinterpret_cast<int4*>(&P);
#ifdef TGX_EXPERIMENT_SAMEPACKET
    const int4* pphr = reinterpret_cast<const int4*>(tv.phase);      // bandwidth experiment: no DRAM reads
#else
    const int4* pphr = reinterpret_cast<const int4*>(tv.phase + traj);
#endif
    if (threadIdx.x < kRecChunks) {
        const int c = threadIdx.x;
        dst[c + (c >= kThRec ? kThPlan - kThRec : 0)] = __ldg(pphr + c);
    }
    __syncthreads();
    const int nseg = P.n;
    if (nseg <= 0) return 0;
    if (nseg > kPhaseBaseSegs) {                   // CTA-uniform, rare
        if (threadIdx.x < kExtChunks) {
            const int x = threadIdx.x;
            dst[x < kThExt ? kKeyExtTo + x : kThExtTo + (x - kThExt)] =
                __ldg(reinterpret_cast<const int4*>(tv.phase_ext + traj) + x);
        }
        __syncthreads();
    }
    if ((int)threadIdx.x < nseg) {
        Seg sg;
        int kend;
        build_phase_segment(P, threadIdx.x, sg, kend);
        s_seg[threadIdx.x] = sg;
        s_kend[threadIdx.x] = kend;
    } else if (threadIdx.x == 32) {                // (the second warp: the segments keep the first one busy)
        s_rec = build_phase_rec(P, P.key[nseg - 1] + 1);      // the trajectory's last sample ends its last segment
    }
    __syncthreads();
    return nseg;
}

// MODE 0: exact-offset plan, 1: slab plan (fixed per-trajectory slices), 2: phase plan (see TableView).
// RECORDS: instead of struct-of-arrays planes the kernel writes one clamped 128-byte tgx_goal_record per sample (the
// consumer side of SURVEY.md §8 f3 fused into the evaluation: 128 B written per sample instead of 112 + 112 + 128).
// The tile is walked in PASSES passes of THREADS*SPT samples; in a pass a warp owns 32*SPT consecutive samples, lane l
// the samples l, l + 32, ... of them, stages whole records in its private part of dynamic shared memory and sends them
// with TMA (RecTma, store.cuh).
// PTMA: the struct-of-arrays planes leave through TMA as well (PlaneTma, store.cuh): same passes and sample ownership as
// the record mode, ro.tmap then describes the caller's planes as a [trajectory][channel][sample] tensor.
template <int THREADS, int SPT, bool STORE, bool REDUCE, int MODE, bool RECORDS = false, int PASSES = 1,
          bool PTMA = false>
__global__ void __launch_bounds__(THREADS, (REDUCE && !STORE) ? TGX_REDUCE_CTAS * 256 / THREADS : 768 / THREADS)   // (PTMA: a 7th CTA per SM at 72 registers: 16.8 vs 16.6 ms)
eval_kernel(TableView tv, OutView out, double* __restrict__ max_v, double* __restrict__ max_a,
            const __grid_constant__ RecOut ro = RecOut{}) {
    constexpr bool STAGED = RECORDS || PTMA;       // samples staged in shared memory and sent by TMA
    static_assert(STAGED || (REDUCE && !STORE) || PASSES == 1 || MODE == 2,
                  "only the TMA and reduction-only modes walk a tile in passes (phase plans: the whole trajectory, always)");
    static_assert(!PTMA || (STORE && !REDUCE && !RECORDS), "PTMA is the plain store-only evaluation");
    constexpr bool SLAB = MODE == 1;
    constexpr int TILE = THREADS * SPT * PASSES;
    constexpr int KS = STAGED ? 32 : 1;           // distance between a thread's samples
    extern __shared__ __align__(16) double2 s_dyn[];
    __shared__ __align__(16) TrajRec s_rec;
    constexpr int NSEG = MODE == 2 ? kPhaseMaxSegs : kMaxSegPerTile;
    __shared__ __align__(16) Seg s_seg[NSEG];
    __shared__ int s_kend[NSEG];                    // last sample of each segment
    __shared__ int4 s_tile;
    __shared__ double s_red[2][THREADS / 32];

    // ---- stage the tile's constants in shared memory (16-byte chunks, one per thread) -------------------
    int traj, k_lo, nseg;
    int npass = PASSES;
    if (MODE == 2) {
        __shared__ __align__(16) PhasePlan s_plan;
        traj = (int)blockIdx.x;                    // one CTA per trajectory, which it walks in passes
        k_lo = 0;
        nseg = stage_phase_plan(tv, traj, s_plan, s_seg, s_kend, s_rec);
        if (nseg <= 0) return;                     // rejected trajectory (whole CTA)
        npass = (s_rec.n + THREADS * SPT - 1) / (THREADS * SPT);
    } else if (SLAB) {
        traj = (int)(blockIdx.x / (unsigned)tv.tile_slab);
        const int t = (int)blockIdx.x - traj * tv.tile_slab;
        const int4* pseg = reinterpret_cast<const int4*>(tv.segs + (size_t)traj * (size_t)tv.seg_slab);
        // round 1 (independent loads): record, this tile's directory entry and, for the first tile of a trajectory
        // (whose segments start the slice), the first kSlabSpecSegs segments
        {
            const int c = threadIdx.x;
            if (c < 4) {
                reinterpret_cast<int4*>(&s_rec)[c] = __ldg(reinterpret_cast<const int4*>(tv.recs + traj) + c);
            } else if (c == 4) {
                s_tile = __ldg(reinterpret_cast<const int4*>(tv.tiles) + blockIdx.x);
            } else if (t == 0 && c < 5 + 4 * kSlabSpecSegs) {
                const int q = c - 5;
                const int4 w = __ldg(pseg + q);
                reinterpret_cast<int4*>(s_seg)[q] = w;
                if ((q & 3) == 0) s_kend[q >> 2] = w.x + w.y;   // kb + n
            }
        }
        __syncthreads();
        const int4 tw = s_tile;
        if (tw.w <= 0) return;   // an empty slot (whole CTA)
        k_lo = tw.y;
        nseg = tw.w < kMaxSegPerTile ? tw.w : kMaxSegPerTile;
        const int have = (t == 0) ? kSlabSpecSegs : 0;
        if (nseg > have) {       // CTA-uniform: a later tile of a long trajectory, or more segments than speculated
            const int4* src = pseg + 4 * (tw.z - traj * tv.seg_slab);
            for (int q = 4 * have + (int)threadIdx.x; q < 4 * nseg; q += THREADS) {
                const int4 w = __ldg(src + q);
                reinterpret_cast<int4*>(s_seg)[q] = w;
                if ((q & 3) == 0) s_kend[q >> 2] = w.x + w.y;
            }
            __syncthreads();
        }
    } else {
        const int4 tw = __ldg(reinterpret_cast<const int4*>(tv.tiles) + blockIdx.x);   // {traj, k_lo, seg_begin, nseg}
        if (tw.w <= 0) return;
        traj = tw.x;
        k_lo = tw.y;
        nseg = tw.w < kMaxSegPerTile ? tw.w : kMaxSegPerTile;
        const int4* src = reinterpret_cast<const int4*>(tv.segs + tw.z);
        for (int t = threadIdx.x; t < 4 + 4 * nseg; t += THREADS) {
            if (t < 4) {
                reinterpret_cast<int4*>(&s_rec)[t] = __ldg(reinterpret_cast<const int4*>(tv.recs + traj) + t);
            } else {
                const int4 w = __ldg(src + (t - 4));
                reinterpret_cast<int4*>(s_seg)[t - 4] = w;
                if (((t - 4) & 3) == 0) s_kend[(t - 4) >> 2] = w.x + w.y;   // kb + n
            }
        }
        __syncthreads();
    }

    const int type = s_rec.type & kRecTypeMask;
    const int n = s_rec.n;
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

    double best_v2 = 0.0, best_a2 = 0.0;

    RecTma<SPT> stager;
    PlaneTma<SPT> pstager;
    if (STAGED && (threadIdx.x & 31) == 0)
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&ro.tmap)) : "memory");
    if (RECORDS) {
        // the warp's private staging area: 32*SPT records, 1024-byte aligned for the 128-byte swizzle
        const uint32_t dyn = ((uint32_t)__cvta_generic_to_shared(s_dyn) + 1023u) & ~1023u;
        stager.init(dyn + (uint32_t)(threadIdx.x >> 5) * (uint32_t)RecTma<SPT>::kBytesPerWarp, (int)threadIdx.x & 31);
    }
    // the row's last 32-byte sector is completed with zeros when it lies inside the row's capacity
    int fill_end = limit;
    if (PTMA) {
        const uint32_t dyn = ((uint32_t)__cvta_generic_to_shared(s_dyn) + 127u) & ~127u;
        pstager.init(dyn + (uint32_t)(threadIdx.x >> 5) * (uint32_t)PlaneTma<SPT>::kBytesPerWarp, (int)threadIdx.x & 31);
        const int64_t lim4 = ((int64_t)limit + (kFillAlign - 1)) & ~(int64_t)(kFillAlign - 1);
        if (lim4 <= out.capacity) fill_end = (int)lim4;
    }

#pragma unroll 1
    for (int pass = 0; pass < (MODE == 2 ? npass : PASSES); ++pass) {
    // first sample of this warp's block of 32*SPT (RECORDS) and of this thread
    const int wk0 = k_lo + pass * (THREADS * SPT) + ((int)threadIdx.x >> 5) * (32 * SPT);
    const int k0 = STAGED ? wk0 + ((int)threadIdx.x & 31) : k_lo + pass * (THREADS * SPT) + SPT * (int)threadIdx.x;
    const int nvalid = (REDUCE ? n : limit) - k0;   // samples this thread evaluates (may be <= 0)

    // RECORDS: every lane of a warp takes part in staging the warp's records, so lanes beyond the trajectory's end
    // walk through the block too (what they stage is never written)
    // (SPT = 2: the thread that owns the second half of the row's last sector zero-fills it, so it enters too)
    if (STAGED ? (wk0 < (PTMA ? fill_end : limit)) : (nvalid > 0 || (STORE && k0 < ((limit + (kFillAlign - 1)) & ~(kFillAlign - 1))))) {
        // ---- segment of each sample: count the segments that end before it (independent broadcast reads) ----
        int si[SPT];
        if (STAGED) {
#pragma unroll
            for (int u = 0; u < SPT; ++u) {
                int c = 0;
                for (int i = 0; i + 1 < nseg; ++i) c += (k0 + u * KS > s_kend[i]) ? 1 : 0;
                si[u] = c;
            }
        } else {
            int c = 0;
            for (int i = 0; i + 1 < nseg; ++i) c += (k0 > s_kend[i]) ? 1 : 0;
            si[0] = c;
#pragma unroll
            for (int u = 1; u < SPT; ++u) {
                while (c + 1 < nseg && k0 + u > s_kend[c]) ++c;     // a thread may straddle a boundary
                si[u] = c;
            }
        }

        double* row = nullptr;
        int nst = 0, nfill = 0;
        if (STORE) {
            const int64_t toff = out.traj_offset ? __ldg(out.traj_offset + traj) : (int64_t)traj * out.traj_stride;
            row = out.base + toff + (PTMA ? 0 : k0);
            nst = limit - k0;   // <= 0: nothing to store for this thread
            // the row's last 32-byte sector is completed with zeros when it lies inside the row's capacity
            const int64_t lim4 = ((int64_t)limit + (kFillAlign - 1)) & ~(int64_t)(kFillAlign - 1);
            nfill = (int)((lim4 <= out.capacity ? lim4 : (int64_t)limit) - k0);
        }
        const uint32_t mask = out.channel_mask;
        const int64_t cs = out.chan_stride;
        if (RECORDS) stager.begin_pass();
        // Channels are emitted in tgx_channel order (the record stager pairs 2c with 2c+1); the warp's records leave
        // after the trailing words.
#define TGX_STORE(CH, ARR)                                                                     \
    do {                                                                                       \
        if (RECORDS) {                                                                         \
            /* the previous pass's records must have left the staging area before the first write */ \
            if ((CH) == TGX_PX && pass > 0) stager.wait_read();                                \
            stager.template put<(CH)>(ARR, ro);                                                \
            if ((CH) == TGX_DPSI) {                                                            \
                stager.put_tail(traj, k0, n);                                                  \
                stager.flush(&ro.tmap, ro.base + rec_off, rec_off, wk0, limit);                \
            }                                                                                  \
        } else if (PTMA) {                                                                     \
            if ((CH) == TGX_PX && pass > 0) pstager.wait_read();                               \
            pstager.template put<(CH)>(ARR);                                                   \
            if ((CH) == TGX_DPSI) pstager.flush(&ro.tmap, traj, wk0, limit, fill_end, row, cs); \
        } else if (STORE && nfill > 0 && (mask & (1u << (CH)))) {                              \
            store_channel<SPT>(row + (CH) * cs, ARR, nst, nfill);                              \
        }                                                                                      \
    } while (0)

        double o[SPT], zero[SPT];
#pragma unroll
        for (int u = 0; u < SPT; ++u) zero[u] = 0.0;
        if (type == kRecStatic) {
            // ---- braking goals of the polyline family: createSquareGoal(last.p.x, last.p.y, v, -accel, heading) and
            //      its copies (Square.cpp:126-127, M.cpp:103-104), createBounceGoal(cx, cy, z, vz, heading)
            //      (Bounce.cpp:94): frozen position, velocity / acceleration along a fixed direction --------------
            const double dx = s_rec.f[4], dy = s_rec.f[5], dz = s_rec.f[6];
            double v[SPT], acc[SPT];
#pragma unroll
            for (int u = 0; u < SPT; ++u) {
                const Seg& sg = s_seg[si[u]];
                v[u] = seg_pos(sg, k0 + u * KS).v;
                acc[u] = sg.acc;
            }
#pragma unroll
            for (int u = 0; u < SPT; ++u) o[u] = s_rec.f[0];
            TGX_STORE(TGX_PX, o);
#pragma unroll
            for (int u = 0; u < SPT; ++u) o[u] = s_rec.f[1];
            TGX_STORE(TGX_PY, o);
#pragma unroll
            for (int u = 0; u < SPT; ++u) o[u] = s_rec.f[2];
            TGX_STORE(TGX_PZ, o);
#pragma unroll
            for (int u = 0; u < SPT; ++u) o[u] = v[u] * dx;
            TGX_STORE(TGX_VX, o);
#pragma unroll
            for (int u = 0; u < SPT; ++u) o[u] = v[u] * dy;
            TGX_STORE(TGX_VY, o);
#pragma unroll
            for (int u = 0; u < SPT; ++u) o[u] = v[u] * dz;
            TGX_STORE(TGX_VZ, o);
#pragma unroll
            for (int u = 0; u < SPT; ++u) o[u] = acc[u] * dx;
            TGX_STORE(TGX_AX, o);
#pragma unroll
            for (int u = 0; u < SPT; ++u) o[u] = acc[u] * dy;
            TGX_STORE(TGX_AY, o);
#pragma unroll
            for (int u = 0; u < SPT; ++u) o[u] = 0.0;
            TGX_STORE(TGX_AZ, o);
            TGX_STORE(TGX_JX, o);
            TGX_STORE(TGX_JY, o);
            TGX_STORE(TGX_JZ, o);
#pragma unroll
            for (int u = 0; u < SPT; ++u) o[u] = s_rec.f[3];
            TGX_STORE(TGX_PSI, o);
            TGX_STORE(TGX_DPSI, zero);
            if (REDUCE) {
#pragma unroll
                for (int u = 0; u < SPT; ++u)
                    if (u < nvalid) {
                        best_v2 = max_nn(best_v2, v[u] * v[u]);                       // the direction is a unit vector
                        best_a2 = max_nn(best_a2, acc[u] * acc[u] * (dx * dx + dy * dy));
                    }
            }
        } else if (type == TGX_LINE) {
            // ---- Line::createLineGoal, Line.cpp:91-115 ------------------------------------------------
            const double c = s_rec.f[0], s = s_rec.f[1], theta = s_rec.f[2], alt = s_rec.f[3], dt = s_rec.f[4];
            const double cdt = c * dt, sdt = s * dt;
            double v[SPT], acc[SPT], py[SPT];
#pragma unroll
            for (int u = 0; u < SPT; ++u) {
                const Seg& sg = s_seg[si[u]];
                const SegPos q = seg_pos(sg, k0 + u * KS);
                v[u] = q.v;
                acc[u] = (q.j == 0) ? 0.0 : sg.acc;   // sample 0 is createLineGoal(A.x, A.y, 0, accel = 0, theta) (:40)
                // S = sum of v over the segment's steps so far; p = p_base + S * (c|s) * dt   (:97-98)
                const double S = fma(sg.dv, q.tri, q.fj * sg.vb) + (q.clamp ? sg.vclamp : 0.0);
                const bool fb = (sg.flags & kSegForcePos) != 0;                // a leg's last goal, forced to B / A
                o[u] = fb ? sg.s0 : fma(S, cdt, sg.s0);
                py[u] = fb ? sg.s1 : fma(S, sdt, sg.s1);
            }
            TGX_STORE(TGX_PX, o);
            TGX_STORE(TGX_PY, py);
#pragma unroll
            for (int u = 0; u < SPT; ++u) o[u] = alt;
            TGX_STORE(TGX_PZ, o);
            double vx[SPT], vy[SPT], ax[SPT], ay[SPT];
#pragma unroll
            for (int u = 0; u < SPT; ++u) {
                vx[u] = v[u] * c;
                vy[u] = v[u] * s;
                ax[u] = acc[u] * c;
                ay[u] = acc[u] * s;
            }
            TGX_STORE(TGX_VX, vx);
            TGX_STORE(TGX_VY, vy);
            TGX_STORE(TGX_VZ, zero);
            TGX_STORE(TGX_AX, ax);
            TGX_STORE(TGX_AY, ay);
            TGX_STORE(TGX_AZ, zero);
            TGX_STORE(TGX_JX, zero);
            TGX_STORE(TGX_JY, zero);
            TGX_STORE(TGX_JZ, zero);
#pragma unroll
            for (int u = 0; u < SPT; ++u) o[u] = theta;
            TGX_STORE(TGX_PSI, o);
            TGX_STORE(TGX_DPSI, zero);
            if (REDUCE) {
#pragma unroll
                for (int u = 0; u < SPT; ++u)
                    if (u < nvalid) {
                        best_v2 = max_nn(best_v2, fma(vx[u], vx[u], vy[u] * vy[u]));
                        best_a2 = max_nn(best_a2, fma(ax[u], ax[u], ay[u] * ay[u]));
                    }
            }
        } else {
            const double r = s_rec.f[0], cx = s_rec.f[1], cy = s_rec.f[2], alt = s_rec.f[3];
            const double dtr = s_rec.f[4], rinv = s_rec.f[5];
            double v[SPT], th[SPT], sn[SPT], cn[SPT], om[SPT];
            constexpr bool SPLIT = REDUCE && !STORE;      // reduction only (the maxima do not see a speed of 1e-15)
            // theta_b + sum_{m<=j} (v_m / r) * dt  =  theta_b + j*w1 + j(j+1)/2 * (dv*dt/r); a hold (dv = 0) is the
            // reference's exact arithmetic progression; a segment's last sample carries the replayed theta.
            auto speed_and_angle = [&](const Seg& sg, int u) {
                const SegPos q = seg_pos(sg, k0 + u * KS);
                // Every sample of a ramp-down except its clamped last one has v > 0 in the reference (`while (v > 0)`,
                // Circle.cpp:75).  With round parameters the exact value of v one step before the end is 0 and the
                // reference's is its accumulated rounding (~1e-15); the single rounding of the closed form may land on
                // the other side of 0 and would turn a Figure8's atan2(vy, vx) yaw (Figure8.cpp:123) by pi.
                v[u] = (!SPLIT && sg.dv < 0.0 && !q.clamp && !(q.v > 0.0)) ? 1e-300 : q.v;
                th[u] = q.last ? sg.acc : fma(q.tri, sg.dv * dtr, fma(q.fj, sg.s1, sg.s0));
            };
            // The reduction-only kernel is issue-bound: a thread's samples almost always lie in one segment (si is
            // non-decreasing), so its record is read from shared memory once per thread, and speeds / angles are
            // computed before the trigonometry (5.77 -> 5.12 ms per Mi config-4 circles).  The store kernels are
            // latency-bound and 1.5 % slower that way (16.6 -> 16.85 ms): they keep one loop per sample.
            if (SPLIT) {
                if (si[0] == si[SPT - 1]) {
                    const Seg sg = s_seg[si[0]];
#pragma unroll
                    for (int u = 0; u < SPT; ++u) speed_and_angle(sg, u);
                } else {
#pragma unroll
                    for (int u = 0; u < SPT; ++u) speed_and_angle(s_seg[si[u]], u);
                }
            }
#pragma unroll
            for (int u = 0; u < SPT; ++u) {
                if (!SPLIT) speed_and_angle(s_seg[si[u]], u);
#ifdef TGX_EXPERIMENT_NOTRIG
                sn[u] = th[u] * 0.5; cn[u] = th[u] * 0.25;   // bandwidth experiment only: no trigonometry
#else
                // one routine for every instantiation (so tgx_eval's maxima equal tgx_feasibility's bit for bit): no slow
                // path, no stack frame, coefficients in constant memory.  The reduction-only kernel is issue-bound (ncu:
                // issue slots 71 % busy, FP64 pipe 32 %) and runs as fast with it at 64 registers / 4 CTAs per SM as with
                // the library's sincos (8.44 vs 8.45 ms per Mi config-4 circles); at 80 registers / 3 CTAs it takes 10.2 ms
                sincos_orbit(th[u], &sn[u], &cn[u]);
#endif
                om[u] = v[u] * rinv;                 // omega = v / r
            }
            if (type == TGX_CIRCLE) {
                // ---- Circle::createCircleGoal, Circle.cpp:96-130 ---------------------------------------
#pragma unroll
                for (int u = 0; u < SPT; ++u) o[u] = fma(r, cn[u], cx);
                TGX_STORE(TGX_PX, o);
#pragma unroll
                for (int u = 0; u < SPT; ++u) o[u] = fma(r, sn[u], cy);
                TGX_STORE(TGX_PY, o);
#pragma unroll
                for (int u = 0; u < SPT; ++u) o[u] = alt;
                TGX_STORE(TGX_PZ, o);
                double vx[SPT], vy[SPT], ax[SPT], ay[SPT];
#pragma unroll
                for (int u = 0; u < SPT; ++u) {
                    const double v2r = v[u] * om[u];          // v^2 / r
                    vx[u] = -v[u] * sn[u];
                    vy[u] = v[u] * cn[u];
                    ax[u] = -v2r * cn[u];                      // tangential term omitted as in :113-114
                    ay[u] = -v2r * sn[u];
                }
                TGX_STORE(TGX_VX, vx);
                TGX_STORE(TGX_VY, vy);
                TGX_STORE(TGX_VZ, zero);
                TGX_STORE(TGX_AX, ax);
                TGX_STORE(TGX_AY, ay);
                TGX_STORE(TGX_AZ, zero);
#pragma unroll
                for (int u = 0; u < SPT; ++u) o[u] = (v[u] * om[u]) * om[u] * sn[u];     // v^3/r^2 * s
                TGX_STORE(TGX_JX, o);
#pragma unroll
                for (int u = 0; u < SPT; ++u) o[u] = -((v[u] * om[u]) * om[u]) * cn[u];
                TGX_STORE(TGX_JY, o);
                TGX_STORE(TGX_JZ, zero);
#pragma unroll
                for (int u = 0; u < SPT; ++u) o[u] = th[u] + kPiOver2;                   // unwrapped (:125)
                TGX_STORE(TGX_PSI, o);
                TGX_STORE(TGX_DPSI, om);
                if (REDUCE) {
#pragma unroll
                    for (int u = 0; u < SPT; ++u)
                        if (u < nvalid) {
                            best_v2 = max_nn(best_v2, fma(vx[u], vx[u], vy[u] * vy[u]));
                            best_a2 = max_nn(best_a2, fma(ax[u], ax[u], ay[u] * ay[u]));
                        }
                }
            } else {
                // ---- Figure8::createFigure8Goal, Figure8.cpp:96-128 ------------------------------------
#pragma unroll
                for (int u = 0; u < SPT; ++u) o[u] = fma(r, sn[u], cx);
                TGX_STORE(TGX_PX, o);
#pragma unroll
                for (int u = 0; u < SPT; ++u) o[u] = fma(r * sn[u], cn[u], cy);
                TGX_STORE(TGX_PY, o);
#pragma unroll
                for (int u = 0; u < SPT; ++u) o[u] = alt;
                TGX_STORE(TGX_PZ, o);
                double vx[SPT], vy[SPT], ax[SPT], ay[SPT];
#pragma unroll
                for (int u = 0; u < SPT; ++u) {
                    const double rw = r * om[u];               // r_*omega (not v: :111)
                    vx[u] = rw * cn[u];
                    vy[u] = rw * (cn[u] * cn[u] - sn[u] * sn[u]);
                    ax[u] = -rw * om[u] * sn[u];
                    ay[u] = (-4.0 * r) * om[u] * om[u] * (sn[u] * cn[u]);
                }
                TGX_STORE(TGX_VX, vx);
                TGX_STORE(TGX_VY, vy);
                TGX_STORE(TGX_VZ, zero);
                TGX_STORE(TGX_AX, ax);
                TGX_STORE(TGX_AY, ay);
                TGX_STORE(TGX_AZ, zero);
                TGX_STORE(TGX_JX, zero);
                TGX_STORE(TGX_JY, zero);
                TGX_STORE(TGX_JZ, zero);
#pragma unroll
                for (int u = 0; u < SPT; ++u) o[u] = atan2(vy[u], vx[u]);               // face the velocity (:123)
                TGX_STORE(TGX_PSI, o);
                TGX_STORE(TGX_DPSI, om);
                if (REDUCE) {
#pragma unroll
                    for (int u = 0; u < SPT; ++u)
                        if (u < nvalid) {
                            best_v2 = max_nn(best_v2, fma(vx[u], vx[u], vy[u] * vy[u]));
                            best_a2 = max_nn(best_a2, fma(ax[u], ax[u], ay[u] * ay[u]));
                        }
                }
            }
        }
#undef TGX_STORE
    }
    }   // pass
    if (RECORDS) stager.wait_read();               // the TMA unit must have read the staging area before the CTA exits
    if (PTMA) pstager.wait_read();

    if (REDUCE) {
        // ---- per-trajectory max |v|, max |a|: warp shuffles -> shared -> one atomicMax per tile ------------
        best_v2 = warp_max(best_v2);
        best_a2 = warp_max(best_a2);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) {
            s_red[0][warp] = best_v2;
            s_red[1][warp] = best_a2;
        }
        __syncthreads();
        if (warp == 0) {
            double a = lane < THREADS / 32 ? s_red[0][lane] : 0.0;
            double b = lane < THREADS / 32 ? s_red[1][lane] : 0.0;
            a = warp_max(a);
            b = warp_max(b);
            if (lane == 0) {
                if (max_v) atomic_max_nonneg(max_v + traj, sqrt(a));
                if (max_a) atomic_max_nonneg(max_a + traj, sqrt(b));
            }
        }
    }
}

// ---- reduction-only evaluation (tgx_feasibility; BASELINE.json configs 4-5) ----------------------------------------------
// Per-trajectory maxima of |v_k| and |a_k| over the reference's own samples, nothing stored.  The arithmetic of every
// sample is the store kernels' (same closed forms, same sincos_orbit, same products: tgx_eval's maxima and
// tgx_feasibility's are the same bits); what differs is everything around it, because this kernel is bound by
// instruction issue and the FP64 pipe, not by HBM:
//   * a thread owns a run of CONSECUTIVE samples of the tile (ceil(samples in the tile / 64) of them), so it finds its
//     segment once, keeps the segment record in registers and only moves on where a segment ends — instead of a
//     search, a 64-byte shared-memory read and a "last sample of the segment?" select per sample;
//   * inside a segment the samples run through a branch-free body, four at a time for instruction-level parallelism
//     (one sample's sincos is a dependent chain of ~15 DFMA); a segment's last sample (clamped speed, replayed angle)
//     is the only special case and is handled apart;
//   * |v|^2 and |a|^2 are sums of squares, so the SIGNS the quadrant logic of sin / cos would fix do not matter: only the
//     swap of the two polynomials for odd quadrants is kept (4 selects instead of 12 per sample).
// 40 FP64 instructions per circle sample remain, plus ~15 others (was ~90).
struct RedSample {
    double v, th;
};

// sin / cos up to sign (see above): the magnitudes sincos_orbit returns, without its two sign selects.
__device__ __forceinline__ void sincos_orbit_unsigned(double x, double* sn, double* cs) {
    const double shifted = fma(x, kTrig[0], kTrig[1]);
    const int q = __double2loint(shifted);
    const double kd = shifted - kTrig[1];
    double r = fma(kd, kTrig[2], x);
    r = fma(kd, kTrig[3], r);
    r = fma(kd, kTrig[4], r);
    const double z = r * r;
    double ps = fma(z, kTrig[5], kTrig[6]);
    ps = fma(z, ps, kTrig[7]);
    ps = fma(z, ps, kTrig[8]);
    ps = fma(z, ps, kTrig[9]);
    ps = fma(z, ps, kTrig[10]);
    const double s = fma(r * z, ps, r);
    double pc = fma(z, kTrig[11], kTrig[12]);
    pc = fma(z, pc, kTrig[13]);
    pc = fma(z, pc, kTrig[14]);
    pc = fma(z, pc, kTrig[15]);
    pc = fma(z, pc, kTrig[16]);
    const double c = fma(z * z, pc, fma(z, -0.5, 1.0));
    *sn = (q & 1) ? c : s;
    *cs = (q & 1) ? s : c;
}

// |v|^2 and |a|^2 of one sample with the products of the store kernels' formulas (eval_kernel above).
template <int TYPE>
__device__ __forceinline__ void red_norms(const TrajRec& rec, double v, double th, double acc, double& v2, double& a2) {
    if (TYPE == TGX_LINE) {
        const double c = rec.f[0], s = rec.f[1];
        const double vx = v * c, vy = v * s, ax = acc * c, ay = acc * s;
        v2 = fma(vx, vx, vy * vy);
        a2 = fma(ax, ax, ay * ay);
    } else if (TYPE == kRecStatic) {
        const double dx = rec.f[4], dy = rec.f[5];
        v2 = v * v;
        a2 = acc * acc * (dx * dx + dy * dy);
    } else {
        double sn, cn;
        sincos_orbit_unsigned(th, &sn, &cn);
        const double om = v * rec.f[5];                  // omega = v / r
        if (TYPE == TGX_CIRCLE) {
            const double v2r = v * om;
            const double vx = v * sn, vy = v * cn, ax = v2r * cn, ay = v2r * sn;
            v2 = fma(vx, vx, vy * vy);
            a2 = fma(ax, ax, ay * ay);
        } else {
            const double r = rec.f[0];
            const double rw = r * om;
            const double vx = rw * cn, vy = rw * (cn * cn - sn * sn);
            const double ax = rw * om * sn, ay = (-4.0 * r) * om * om * (sn * cn);
            v2 = fma(vx, vx, vy * vy);
            a2 = fma(ax, ax, ay * ay);
        }
    }
}

// The samples [k, k_end) of one trajectory type, walking the tile's segment list forward from segment si.  Only the
// segment INDEX lives in a register and moves on (a two-instruction divergent loop) where a sample lies beyond the
// segment's end; the record's fields are read from shared memory where they are used and everything else is
// branch-free, so the lanes of a warp stay together through the long FP64 chains whatever segments they are in.
// (Two earlier shapes were slower than the generic kernel's 50 ms per 10^7 config-4 circles: peeling the segments' last
// samples into a path of their own, 57 ms — every warp serialised through the one-sample paths of the few lanes that met
// a segment end; and a register-resident copy of the record, 53 ms — 176 bytes of spills per thread, LSU pipe 40 % busy.)
#ifndef TGX_RED_U
#define TGX_RED_U 4            // samples a thread evaluates side by side
#endif
template <int TYPE>
__device__ __forceinline__ void red_run(const TrajRec& rec, const Seg* __restrict__ segs, const int* __restrict__ kends,
                                        int nseg, int si, int k, int k_end, double& best_v2, double& best_a2) {
    constexpr bool ORBIT = TYPE == TGX_CIRCLE || TYPE == TGX_FIGURE8;
    const double dtr = rec.f[4];
    // speed, angle and the `accel` argument of sample kk (seg_pos, and eval_kernel's speed_and_angle)
    auto state = [&](int kk, double& v, double& th, double& acc) {
        while (si + 1 < nseg && kk > kends[si]) ++si;
        const Seg& sg = segs[si];
        const SegPos q = seg_pos(sg, kk);
        v = q.v;
        th = 0.0;
        acc = sg.acc;
        if (ORBIT) th = q.last ? sg.acc : fma(q.tri, sg.dv * dtr, fma(q.fj, sg.s1, sg.s0));
        if (TYPE == TGX_LINE && q.j == 0) acc = 0.0;     // sample 0 is createLineGoal(A.x, A.y, 0, accel = 0, theta)
    };
    constexpr int U = TGX_RED_U;
    for (; k + U <= k_end; k += U) {
        double v[U], th[U], acc[U], v2[U], a2[U];
        while (si + 1 < nseg && k > kends[si]) ++si;
        const int k_last = kends[si];               // (also of the list's last segment: its last sample is special too)
        if (k + U - 1 <= k_last) {
            // the common case: one segment serves the whole group; its record is read once (four 16-byte LDS) and only
            // the group's last sample can be the segment's last one (clamped speed, replayed angle)
            const Seg sg = segs[si];
            const double ddth = sg.dv * dtr;
            const int j0 = k - sg.kb;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double fj = (double)(j0 + u);
                v[u] = fma(fj, sg.dv, sg.vb);
                th[u] = 0.0;
                // j(j+1)/2 is an exact integer below 2^53 however it is formed: fma(j, j, j)/2 == seg_pos's 0.5*(j*(j+1))
                if (ORBIT) th[u] = fma(0.5 * fma(fj, fj, fj), ddth, fma(fj, sg.s1, sg.s0));
                acc[u] = sg.acc;
            }
            if (k + U - 1 == k_last) {
                v[U - 1] = seg_pos(sg, k + U - 1).v;
                if (ORBIT) th[U - 1] = sg.acc;
            }
            if (TYPE == TGX_LINE && j0 == 0) acc[0] = 0.0;
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) state(k + u, v[u], th[u], acc[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) red_norms<TYPE>(rec, v[u], th[u], acc[u], v2[u], a2[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            best_v2 = max_nn(best_v2, v2[u]);
            best_a2 = max_nn(best_a2, a2[u]);
        }
    }
    for (; k < k_end; ++k) {
        double v, th, acc, v2, a2;
        state(k, v, th, acc);
        red_norms<TYPE>(rec, v, th, acc, v2, a2);
        best_v2 = max_nn(best_v2, v2);
        best_a2 = max_nn(best_a2, a2);
    }
}

template <int MODE, int TILE>
__global__ void __launch_bounds__(64, TGX_REDUCE_CTAS * 4)
reduce_kernel(TableView tv, double* __restrict__ max_v, double* __restrict__ max_a) {
    constexpr int THREADS = 64;
    __shared__ __align__(16) TrajRec s_rec;
    constexpr int NSEG = MODE == 2 ? kPhaseMaxSegs : kMaxSegPerTile;
    __shared__ __align__(16) Seg s_seg[NSEG];
    __shared__ int s_kend[NSEG];
    __shared__ int4 s_tile;
    __shared__ double s_red[2][THREADS / 32];

    // ---- stage the tile's constants (as eval_kernel does) -----------------------------------------------------------------
    int traj, k_lo, nseg;
    if (MODE == 2) {
        __shared__ __align__(16) PhasePlan s_plan;
        traj = (int)blockIdx.x;                    // one CTA per trajectory
        k_lo = 0;
        nseg = stage_phase_plan(tv, traj, s_plan, s_seg, s_kend, s_rec);
        if (nseg <= 0) return;
    } else {
        int4 tw;
        if (MODE == 1) {
            traj = (int)(blockIdx.x / (unsigned)tv.tile_slab);
            if (threadIdx.x == 0) s_tile = __ldg(reinterpret_cast<const int4*>(tv.tiles) + blockIdx.x);
            __syncthreads();
            tw = s_tile;
        } else {
            tw = __ldg(reinterpret_cast<const int4*>(tv.tiles) + blockIdx.x);   // {traj, k_lo, seg_begin, nseg}
            traj = tw.x;
        }
        if (tw.w <= 0) return;
        k_lo = tw.y;
        nseg = tw.w < kMaxSegPerTile ? tw.w : kMaxSegPerTile;
        const int4* src = reinterpret_cast<const int4*>(tv.segs + tw.z);
        for (int t = threadIdx.x; t < 4 + 4 * nseg; t += THREADS) {
            if (t < 4) {
                reinterpret_cast<int4*>(&s_rec)[t] = __ldg(reinterpret_cast<const int4*>(tv.recs + traj) + t);
            } else {
                const int4 w = __ldg(src + (t - 4));
                reinterpret_cast<int4*>(s_seg)[t - 4] = w;
                if (((t - 4) & 3) == 0) s_kend[(t - 4) >> 2] = w.x + w.y;
            }
        }
        __syncthreads();
    }

    const int type = s_rec.type & kRecTypeMask;
    const int n = s_rec.n;
    // the tile's samples in runs of `per` consecutive ones, one run per thread
    const int in_tile = MODE == 2 ? n : min(n - k_lo, TILE);      // (phase plan: the whole trajectory)
    int per = (in_tile + THREADS - 1) / THREADS;
    if (per > 2) per = (per + TGX_RED_U - 1) / TGX_RED_U * TGX_RED_U;      // whole groups (the unrolled body)
    const int k = k_lo + (int)threadIdx.x * per;
    const int k_end = min(k + per, k_lo + in_tile);
    double best_v2 = 0.0, best_a2 = 0.0;
    if (k < k_end) {
        int si = 0;
        for (int i = 0; i + 1 < nseg; ++i) si += (k > s_kend[i]) ? 1 : 0;
        if (type == TGX_CIRCLE) red_run<TGX_CIRCLE>(s_rec, s_seg, s_kend, nseg, si, k, k_end, best_v2, best_a2);
        else if (type == TGX_FIGURE8) red_run<TGX_FIGURE8>(s_rec, s_seg, s_kend, nseg, si, k, k_end, best_v2, best_a2);
        else if (type == TGX_LINE) red_run<TGX_LINE>(s_rec, s_seg, s_kend, nseg, si, k, k_end, best_v2, best_a2);
        else red_run<kRecStatic>(s_rec, s_seg, s_kend, nseg, si, k, k_end, best_v2, best_a2);
    }

    best_v2 = warp_max(best_v2);
    best_a2 = warp_max(best_a2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        s_red[0][warp] = best_v2;
        s_red[1][warp] = best_a2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double a = max_nn(s_red[0][0], s_red[0][1]), b = max_nn(s_red[1][0], s_red[1][1]);
        if (max_v) atomic_max_nonneg(max_v + traj, sqrt(a));
        if (max_a) atomic_max_nonneg(max_a + traj, sqrt(b));
    }
}

// Feasibility verdict per trajectory (BASELINE.json config 4): flag = max_v <= v_max && max_a <= a_max &&
// no status bit set (the plan's status already carries OUTSIDE_BOUNDS when a box was given).
__global__ void __launch_bounds__(256)
feasibility_finalize_kernel(int64_t n, const uint32_t* __restrict__ plan_status, const double* __restrict__ max_v,
                            const double* __restrict__ max_a, double v_max, double a_max,
                            uint8_t* __restrict__ flags, uint32_t* __restrict__ status_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t st = plan_status[i];
    if (max_v[i] > v_max) st |= TGX_ST_VMAX_EXCEEDED;
    if (max_a[i] > a_max) st |= TGX_ST_AMAX_EXCEEDED;
    if (flags) flags[i] = st == 0 ? 1 : 0;
    if (status_out) status_out[i] = st;
}

// ---- host-side launchers -----------------------------------------------------------------------------------

template <int THREADS, int SPT, int MODE>
static cudaError_t launch_eval_t(const TableView& tv, int64_t ntiles, const OutView& out, bool store, double* max_v,
                                 double* max_a, cudaStream_t stream) {
    const bool reduce = max_v || max_a;
    const unsigned grid = (unsigned)ntiles;
    if (store && reduce)
        eval_kernel<THREADS, SPT, true, true, MODE><<<grid, THREADS, 0, stream>>>(tv, out, max_v, max_a);
    else if (store)
        eval_kernel<THREADS, SPT, true, false, MODE><<<grid, THREADS, 0, stream>>>(tv, out, max_v, max_a);
    else if (TGX_REDUCE_LEGACY)
        // (the generic kernel's reduction-only instantiation, kept for A/B runs: 64-thread CTAs walking the tile in
        //  passes of 256 samples; per Mi config-4 circles 256 threads x 1 pass 7.76 ms, 128 x 2 6.21 ms, 64 x 4 5.72 ms)
        eval_kernel<64, 4, false, true, MODE, false, THREADS * SPT / 256><<<grid, 64, 0, stream>>>(tv, out, max_v, max_a);
    else
        reduce_kernel<MODE, THREADS * SPT><<<grid, 64, 0, stream>>>(tv, max_v, max_a);
    return cudaGetLastError();
}

// tile = 1 << tile_shift samples per CTA, spt samples per thread: threads per CTA = tile / spt in {128, 256}.
// Store-only evaluation with the planes leaving through TMA: 128-thread CTAs, 2 samples per thread per pass.
// (measured on 1 Mi circles, eval only: 128 threads x 2 samples 16.6 ms; 64 x 2 17.6; 256 x 2 17.1; 128 x 4 20.0;
//  64 x 4 17.2; the vector-store kernel 17.6)
template <int TILE, int MODE>
static cudaError_t launch_eval_ptma_t(const TableView& tv, int64_t ntiles, const OutView& out, const RecOut& ptma,
                                      cudaStream_t stream) {
    constexpr int THREADS = 128, SPP = 2;
    auto kernel = eval_kernel<THREADS, SPP, true, false, MODE, false, TILE / (SPP * THREADS), true>;
    const int smem = (THREADS / 32) * PlaneTma<SPP>::kBytesPerWarp + 128;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    kernel<<<(unsigned)ntiles, THREADS, smem, stream>>>(tv, out, nullptr, nullptr, ptma);
    return cudaGetLastError();
}

// ptma != nullptr: the caller's planes qualify for the TMA store path and ptma->tmap describes them (engine.cu).
cudaError_t launch_eval(const TableView& tv, int64_t ntiles, int tile_shift, int spt, const OutView& out, bool store,
                        double* max_v, double* max_a, cudaStream_t stream, const RecOut* ptma) {
    if (ntiles <= 0) return cudaSuccess;
    if (ntiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    const int threads = (1 << tile_shift) / spt;
    const int mode = tv.phase ? 2 : (tv.tile_slab > 0 ? 1 : 0);
    if (ptma && store && !max_v && !max_a) {
        const int tile = 1 << tile_shift;
#define TGX_PCASE(TL)                                                                           \
    if (tile == (TL))                                                                           \
        return mode == 2   ? launch_eval_ptma_t<TL, 2>(tv, ntiles, out, *ptma, stream)          \
               : mode == 1 ? launch_eval_ptma_t<TL, 1>(tv, ntiles, out, *ptma, stream)          \
                           : launch_eval_ptma_t<TL, 0>(tv, ntiles, out, *ptma, stream)
        TGX_PCASE(512);
        TGX_PCASE(1024);
#undef TGX_PCASE
    }
#define TGX_CASE(T, S)                                                                                     \
    if (threads == (T) && spt == (S))                                                                      \
        return mode == 2   ? launch_eval_t<T, S, 2>(tv, ntiles, out, store, max_v, max_a, stream)          \
               : mode == 1 ? launch_eval_t<T, S, 1>(tv, ntiles, out, store, max_v, max_a, stream)          \
                           : launch_eval_t<T, S, 0>(tv, ntiles, out, store, max_v, max_a, stream)
    TGX_CASE(128, 4);
    TGX_CASE(256, 2);
    TGX_CASE(256, 4);
#undef TGX_CASE
    return cudaErrorInvalidConfiguration;
}

// Record mode: 2 samples per thread per pass, TILE / (2 * THREADS) passes.
template <int THREADS, int TILE, int MODE>
static cudaError_t launch_eval_records_t(const TableView& tv, int64_t ntiles, const RecOut& ro, cudaStream_t stream) {
    constexpr int SPP = 2;
    auto kernel = eval_kernel<THREADS, SPP, false, false, MODE, true, TILE / (SPP * THREADS)>;
    const int smem = (THREADS / 32) * RecTma<SPP>::kBytesPerWarp + 1024;   // whole records + alignment slack
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    kernel<<<(unsigned)ntiles, THREADS, smem, stream>>>(tv, OutView{}, nullptr, nullptr, ro);
    return cudaGetLastError();
}

// Evaluation straight into clamped array-of-structs records (tgx_eval_records).  The record kernels always run as CTAs
// of 128 threads that walk their tile in passes of 256 samples, whatever the plan's tuning: 5 CTAs per SM whose
// prologues, arithmetic and TMA waits overlap (measured on 1 Mi circles: 256 threads x 2 passes 20.8 ms, 128 x 4
// 18.8 ms, 64 x 8 19.2 ms).
cudaError_t launch_eval_records(const TableView& tv, int64_t ntiles, int tile_shift, int spt, const RecOut& ro,
                                cudaStream_t stream) {
    (void)spt;
    if (ntiles <= 0) return cudaSuccess;
    if (ntiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    const int tile = 1 << tile_shift;
    const int mode = tv.phase ? 2 : (tv.tile_slab > 0 ? 1 : 0);
#define TGX_CASE(TL)                                                                            \
    if (tile == (TL))                                                                           \
        return mode == 2   ? launch_eval_records_t<128, TL, 2>(tv, ntiles, ro, stream)          \
               : mode == 1 ? launch_eval_records_t<128, TL, 1>(tv, ntiles, ro, stream)          \
                           : launch_eval_records_t<128, TL, 0>(tv, ntiles, ro, stream)
    TGX_CASE(512);
    TGX_CASE(1024);
#undef TGX_CASE
    return cudaErrorInvalidConfiguration;
}

cudaError_t launch_feasibility_finalize(int64_t n, const uint32_t* plan_status, const double* max_v,
                                        const double* max_a, double v_max, double a_max, uint8_t* flags,
                                        uint32_t* status_out, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int64_t blocks = (n + 255) / 256;
    feasibility_finalize_kernel<<<(unsigned)blocks, 256, 0, stream>>>(n, plan_status, max_v, max_a, v_max, a_max,
                                                                     flags, status_out);
    return cudaGetLastError();
}

}  // namespace tgx
