// store.cuh — vector stores of the struct-of-arrays output planes and the per-trajectory max reduction helpers,
// shared by the evaluation kernels (eval.cu, polyline.cu).
#pragma once

#include <cuda_runtime.h>

#include "tgx_internal.cuh"

namespace tgx {
namespace {

// POLICY: 0 = .cs (streaming, evict-first), 1 = default write-back, 2 = L1::no_allocate + L2::evict_first
template <int SPT, int POLICY>
struct VecStore;

template <int POLICY>
struct VecStore<2, POLICY> {
    static __device__ __forceinline__ void st(double* p, const double (&x)[2]) {
        if (POLICY == 0)
            asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(x[0]), "d"(x[1]) : "memory");
        else if (POLICY == 1)
            asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(x[0]), "d"(x[1]) : "memory");
        else
            asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(x[0]), "d"(x[1]) : "memory");
    }
};

template <int POLICY>
struct VecStore<4, POLICY> {
    static __device__ __forceinline__ void st(double* p, const double (&x)[4]) {
        if (POLICY == 0)
            asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(x[0]), "d"(x[1]), "d"(x[2]),
                         "d"(x[3]) : "memory");
        else if (POLICY == 1)
            asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(x[0]), "d"(x[1]), "d"(x[2]),
                         "d"(x[3]) : "memory");
        else
            asm volatile("st.global.L1::no_allocate.L2::evict_first.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p),
                         "d"(x[0]), "d"(x[1]), "d"(x[2]), "d"(x[3]) : "memory");
    }
};

#ifndef TGX_STORE_POLICY
#define TGX_STORE_POLICY 0
#endif

// A row's tail is zero-filled up to a multiple of kFillAlign samples (see store_channel).
#ifndef TGX_FILL_ALIGN
#define TGX_FILL_ALIGN 4
#endif
constexpr int kFillAlign = TGX_FILL_ALIGN;

// Store SPT adjacent samples of one channel.  nvalid = samples of this thread below the row's limit (may be <= 0),
// nfill = samples of this thread below the limit ROUNDED UP to a 32-byte sector (4 doubles) and inside the row's
// capacity.  A trajectory's last vector is written in full, its tail zero-filled: a partially written sector would make
// the memory system fetch the rest of it from DRAM before it can be written back, and those read-fills, one per channel
// per trajectory, cost a read/write bus turnaround each in the middle of the store stream (measured: 18.0 -> 17.3 ms on
// the 1 Mi-circle batch, 18.9 -> 16.7 ms where three quarters of the rows end inside a sector).
template <int SPT>
__device__ __forceinline__ void store_channel(double* p, const double (&x)[SPT], int nvalid, int nfill) {
    if (nvalid >= SPT) {
        VecStore<SPT, TGX_STORE_POLICY>::st(p, x);
    } else if (nfill >= SPT) {
        double y[SPT];
#pragma unroll
        for (int u = 0; u < SPT; ++u) y[u] = u < nvalid ? x[u] : 0.0;
        VecStore<SPT, TGX_STORE_POLICY>::st(p, y);
    } else {
#pragma unroll
        for (int u = 0; u < SPT; ++u)
            if (u < nvalid) __stcs(p + u, x[u]);
    }
}

// Running maximum of non-negative values: a compare and two selects instead of fmax's NaN-quieting sequence (DSETP.MAX,
// LOP3, FSEL, SEL, moves: ~8 instructions, twice per sample in the reduction kernels).  Like fmax it never lets a NaN
// sample replace the running value.
__device__ __forceinline__ double max_nn(double best, double x) { return x > best ? x : best; }

// Warp maximum of non-negative (non-NaN) doubles: their bit patterns are ordered like the values, so two integer warp
// reductions (REDUX.MAX on the high words, then on the low words of the lanes that hold the largest high word) replace
// five rounds of 64-bit shuffles + fmax (~30 instructions per sample of a 4-sample thread in the reduction kernels).
__device__ __forceinline__ double warp_max(double x) {
    const unsigned hi = (unsigned)__double2hiint(x), lo = (unsigned)__double2loint(x);
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return __hiloint2double((int)mh, (int)ml);
}

__device__ __forceinline__ void atomic_max_nonneg(double* addr, double x) {
    // For non-negative doubles the IEEE bit pattern is monotone in the value.
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(x));
}

// ---- record output (fused): one 128-byte tgx_goal_record per sample straight from the evaluation
// kernels (eval.cu, polyline.cu) -------------------------------------------------------------------------------
// History: the first version kept 4 adjacent samples per thread and transposed 64-byte record halves through an
// XOR-swizzled shared-memory buffer, each warp reading its records back with LDS and streaming them with 16-byte STG:
// 28.4 ms per 1.05e9 samples, L1TEX 87 % busy.  A variant without shared memory (every thread writing
// 32-byte pieces of its own records) took 32.3 ms.
// In record mode a thread does not own adjacent samples but samples 32 apart: in every pass a warp owns 32*SPT
// CONSECUTIVE samples (lane l: wk0 + l, wk0 + 32 + l, ...), so what the warp stages is already in record order.  Each
// lane writes its samples' 16 doubles as eight 16-byte chunks into the warp's private staging area, laid out exactly as
// TMA's 128-byte swizzle wants it (record row ri at ri*128, chunk c at ((c ^ (ri & 7)) << 4); the area is 1024-byte
// aligned): a quarter-warp of writers (8 consecutive lanes, one chunk index) hits 8 different 16-byte bank groups, so the
// st.shared.v2.f64 are conflict-free.  One elected lane then hands each group of 32 records (4 KiB, contiguous in
// global memory) to the TMA unit with cp.async.bulk.tensor: no LDS read-back and no STG — the staged bytes cross the
// L1/shared pipe once instead of three times, which was the limiter of the LDS + STG stager above (L1TEX 87 % busy at
// 4.7 TB/s).  Only the group that straddles the row's limit (at most one per trajectory) leaves through LDS + STG, since
// a TMA box cannot be cut at an arbitrary record.  Everything is warp-private: __syncwarp only, no CTA barrier.
template <int SPT>
struct RecTma {
    static constexpr int kBytesPerWarp = 32 * SPT * 128;
    uint32_t sbase;            // shared-window address of this warp's staging area (1024-byte aligned)
    int lane;
    double pend[SPT];          // the even channel of the open pair
    unsigned clamped[SPT];     // bit 0 / 1 / 2: p.x / p.y / p.z was saturated

    __device__ __forceinline__ void init(uint32_t warp_area, int lane_) {
        sbase = warp_area;
        lane = lane_;
    }
    __device__ __forceinline__ void begin_pass() {
#pragma unroll
        for (int u = 0; u < SPT; ++u) clamped[u] = 0;
    }
    // TrajectoryGenerator::saturate (:773-780): high is tested first, a NaN passes through.
    __device__ __forceinline__ double sat(double v, double lo, double hi, unsigned bit, unsigned& flags) {
        if (v > hi) { flags |= bit; return hi; }
        if (v < lo) { flags |= bit; return lo; }
        return v;
    }
    __device__ __forceinline__ void st_chunk(int u, int c, double a, double b) {
        // row u*32 + lane: (row & 7) == (lane & 7)
        const uint32_t addr = sbase + (uint32_t)((u * 32 + lane) * 128 + ((c ^ (lane & 7)) << 4));
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
    }
    // Channels must arrive in tgx_channel order.
    template <int CH>
    __device__ __forceinline__ void put(const double (&x)[SPT], const RecOut& ro) {
#pragma unroll
        for (int u = 0; u < SPT; ++u) {
            double y = x[u];
            if (CH <= TGX_PZ && ro.clamp) y = sat(x[u], ro.box[2 * CH], ro.box[2 * CH + 1], 1u << CH, clamped[u]);
            if ((CH & 1) == 0) pend[u] = y;
            else st_chunk(u, CH >> 1, pend[u], y);
        }
    }
    // The two trailing words of the record (chunk 7): {traj, k}, {power, modes, clamped, last}.  k of sample u = k0 + 32u.
    __device__ __forceinline__ void put_tail(int traj, int k0, int n) {
#pragma unroll
        for (int u = 0; u < SPT; ++u) {
            const int k = k0 + 32 * u;
            const unsigned long long w0 = (unsigned long long)(unsigned)traj | ((unsigned long long)(unsigned)k << 32);
            const unsigned long long w1 = 1ull | ((unsigned long long)clamped[u] << 24) |
                                          ((unsigned long long)(k == n - 1 ? 1 : 0) << 32);
            st_chunk(u, 7, __longlong_as_double((long long)w0), __longlong_as_double((long long)w1));
        }
    }
    // Send the warp's staged records: row = the trajectory's first record, grow = its index in the record buffer,
    // wk0 = the first sample the warp staged; samples >= limit are not written.  All 32 lanes must call it.
    __device__ __forceinline__ void flush(const CUtensorMap* tmap, tgx_goal_record* row, int64_t grow, int wk0,
                                          int limit) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> async-proxy reads
        __syncwarp();
#pragma unroll
        for (int u = 0; u < SPT; ++u) {
            const int kg = wk0 + 32 * u;                                 // warp-uniform
            if (kg + 32 <= limit) {
                if (lane == 0) {
                    const int c1 = (int)(grow + kg);
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                                     reinterpret_cast<uint64_t>(tmap)),
                                 "r"(0), "r"(c1), "r"(sbase + (uint32_t)(u * 4096))
                                 : "memory");
                }
            } else if (kg < limit) {
                // the group that straddles the limit: 32 rows x 8 chunks through LDS + STG, lane = (row mod 4, chunk)
                const int c = lane & 7;
                double2* out = reinterpret_cast<double2*>(row);
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int r = it * 4 + (lane >> 3);
                    if (kg + r < limit) {
                        const uint32_t addr = sbase + (uint32_t)((u * 32 + r) * 128 + ((c ^ (r & 7)) << 4));
                        double a, b;
                        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr) : "memory");
                        asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(out + (int64_t)(kg + r) * 8 + c), "d"(a),
                                     "d"(b) : "memory");
                    }
                }
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    // The staging area may be rewritten (or the CTA may exit) once the TMA unit has READ it.
    __device__ __forceinline__ void wait_read() {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
    }
};

// ---- plane output through TMA ----------------------------------------------------------------------------------
// The same sample ownership as the record mode (a warp owns 32*SPT consecutive samples per pass, lane l the samples
// l, l + 32, ...), staged as [group u][channel][32 samples]: consecutive lanes write consecutive doubles (conflict-free)
// and one group is exactly the box {32 samples, 14 channels, 1 trajectory} of the 3-D tensor map tgx_eval builds over the
// caller's planes, so a group leaves with ONE cp.async.bulk.tensor instead of 14 vector stores per thread.  The group
// that straddles the row's end keeps the contract of store_channel (valid samples, the last 32-byte sector completed
// with zeros, nothing beyond) by going through LDS + STG.
template <int SPT>
struct PlaneTma {
    static constexpr int kBoxBytes = TGX_NCHAN * 32 * 8;        // 3584 = 28 * 128
    static constexpr int kBytesPerWarp = SPT * kBoxBytes;
    uint32_t sbase;            // shared-window address of this warp's staging area (128-byte aligned)
    int lane;

    __device__ __forceinline__ void init(uint32_t warp_area, int lane_) {
        sbase = warp_area;
        lane = lane_;
    }
    template <int CH>
    __device__ __forceinline__ void put(const double (&x)[SPT]) {
#pragma unroll
        for (int u = 0; u < SPT; ++u)
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(sbase + (uint32_t)(u * kBoxBytes + CH * 256 + lane * 8)),
                         "d"(x[u]) : "memory");
    }
    // row = channel 0, sample 0 of the trajectory; cs = channel stride; samples >= limit are not written, except zeros up
    // to fill_end (limit rounded up to a sector when that fits the row).  All 32 lanes must call it.
    __device__ __forceinline__ void flush(const CUtensorMap* tmap, int traj, int wk0, int limit, int fill_end,
                                          double* row, int64_t cs) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> async-proxy reads
        __syncwarp();
#pragma unroll
        for (int u = 0; u < SPT; ++u) {
            const int kg = wk0 + 32 * u;                                 // warp-uniform
            if (kg + 32 <= limit) {
                if (lane == 0)
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(
                                     reinterpret_cast<uint64_t>(tmap)),
                                 "r"(kg), "r"(0), "r"(traj), "r"(sbase + (uint32_t)(u * kBoxBytes))
                                 : "memory");
            } else if (kg < fill_end) {
                // 14 channels x 8 quads of 4 samples, one (channel, quad) per lane per round
                for (int item = lane; item < TGX_NCHAN * 8; item += 32) {
                    const int ch = item >> 3, k = kg + 4 * (item & 7);
                    if (k < fill_end) {
                        const uint32_t addr = sbase + (uint32_t)(u * kBoxBytes + ch * 256 + (item & 7) * 32);
                        double x[4];
                        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x[0]), "=d"(x[1]) : "r"(addr) : "memory");
                        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x[2]), "=d"(x[3]) : "r"(addr + 16) : "memory");
                        store_channel<4>(row + ch * cs + k, x, limit - k, fill_end - k);
                    }
                }
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    __device__ __forceinline__ void wait_read() {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
    }
};

}  // namespace
}  // namespace tgx
