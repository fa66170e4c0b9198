// pack.cu — consumer side of the path: what the node publishes while it follows a trajectory.
//
//   goal_ = traj_goals_[pub_index_];                              TrajectoryGenerator.cpp:557
//   goal_.p.x = saturate(goal_.p.x, xmin_, xmax_);  (y, z alike)  :602-604, saturate() :773-780
//
// pack_goals_kernel turns the struct-of-arrays planes tgx_eval wrote into one 128-byte tgx_goal_record per
// (trajectory, sample) with the position clamped to the room box.  Pure data movement: 112 B read + 128 B written per
// sample, HBM-bound.  One CTA per 256 consecutive samples of one trajectory, one sample per thread:
//   - 14 coalesced 8-byte plane reads per thread (a warp reads 256 contiguous bytes per plane),
//   - the record is assembled in registers and written to shared memory as eight 16-byte chunks whose position is
//     XOR-swizzled with the record index, which makes both the column-wise writes and the row-wise reads conflict-free,
//   - each warp then streams its own 32 records (4 KB) to global memory with eight 512-byte coalesced stores.
#include <cuda_runtime.h>

#include "tgx_internal.cuh"

namespace tgx {

namespace {

constexpr int kPackThreads = 256;

// TrajectoryGenerator::saturate (TrajectoryGenerator.cpp:773-780): high is tested first; a NaN passes through.
__device__ __forceinline__ double saturate(double val, double low, double high, bool& hit) {
    if (val > high) {
        hit = true;
        return high;
    }
    if (val < low) {
        hit = true;
        return low;
    }
    return val;
}

__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }

__device__ __forceinline__ void st_stream(double2* p, double2 v) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

}  // namespace

__global__ void __launch_bounds__(kPackThreads)
pack_goals_kernel(OutView in, const int32_t* __restrict__ counts, int tiles_per_traj, tgx_limits lim, int clamp,
                  tgx_goal_record* __restrict__ records, int64_t rec_stride, const int64_t* __restrict__ rec_offset,
                  int64_t rec_capacity) {
    __shared__ double2 s_chunk[kPackThreads * 8];     // 32 KB: 256 records x 8 chunks of 16 bytes

    const int traj = (int)(blockIdx.x / (unsigned)tiles_per_traj);
    const int k_lo = ((int)blockIdx.x - traj * tiles_per_traj) * kPackThreads;
    int n = __ldg(counts + traj);
    if ((int64_t)n > rec_capacity) n = (int)rec_capacity;
    if ((int64_t)n > in.capacity) n = (int)in.capacity;
    if (k_lo >= n) return;                            // whole CTA

    const int t = threadIdx.x, k = k_lo + t;
    const int lane = t & 31, warp_base = t & ~31;
    if (k < n) {
        const int64_t toff = in.traj_offset ? __ldg(in.traj_offset + traj) : (int64_t)traj * in.traj_stride;
        const double* src = in.base + toff + k;
        double c[TGX_NCHAN];
#pragma unroll
        for (int q = 0; q < TGX_NCHAN; ++q) c[q] = ld_stream(src + q * in.chan_stride);
        bool hx = false, hy = false, hz = false;
        if (clamp) {
            c[TGX_PX] = saturate(c[TGX_PX], lim.box[0], lim.box[1], hx);
            c[TGX_PY] = saturate(c[TGX_PY], lim.box[2], lim.box[3], hy);
            c[TGX_PZ] = saturate(c[TGX_PZ], lim.box[4], lim.box[5], hz);
        }
        // the two trailing 8-byte words: {traj, k} and {power, mode_xy, mode_z, clamped, last, 0, 0, 0}
        const unsigned long long w0 = (unsigned long long)(unsigned)traj | ((unsigned long long)(unsigned)k << 32);
        const unsigned long long w1 = 1ull /* power = true (Circle.cpp:127) */
                                      | ((unsigned long long)((hx ? 1 : 0) | (hy ? 2 : 0) | (hz ? 4 : 0)) << 24)
                                      | ((unsigned long long)(k == __ldg(counts + traj) - 1 ? 1 : 0) << 32);
        double2* dst = s_chunk + t * 8;
        const int sw = t & 7;
#pragma unroll
        for (int q = 0; q < 7; ++q) dst[q ^ sw] = make_double2(c[2 * q], c[2 * q + 1]);
        dst[7 ^ sw] = make_double2(__longlong_as_double((long long)w0), __longlong_as_double((long long)w1));
    }
    __syncwarp();
    // the warp's 32 records are contiguous in the output: eight stores of 512 contiguous bytes (4 records) each
    const int64_t roff = rec_offset ? __ldg(rec_offset + traj) : (int64_t)traj * rec_stride;
    double2* out = reinterpret_cast<double2*>(records + roff + k_lo + warp_base);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int rl = 4 * j + (lane >> 3);            // record within the warp
        const int pos = lane & 7;                      // stored chunk position
        if (k_lo + warp_base + rl < n) {
            const double2 v = s_chunk[(warp_base + rl) * 8 + pos];
            st_stream(out + rl * 8 + (pos ^ (rl & 7)), v);
        }
    }
}

cudaError_t launch_pack_goals(const OutView& in, const int32_t* counts, int64_t n, const tgx_limits* lim,
                              tgx_goal_record* records, int64_t rec_stride, const int64_t* rec_offset,
                              int64_t rec_capacity, cudaStream_t stream) {
    const int64_t cap = rec_capacity < in.capacity ? rec_capacity : in.capacity;
    if (n <= 0 || cap <= 0) return cudaSuccess;
    const int64_t tiles_per_traj = (cap + kPackThreads - 1) / kPackThreads;
    const int64_t grid = n * tiles_per_traj;
    if (grid > 0x7fffffffLL || tiles_per_traj > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    tgx_limits l{};
    if (lim) l = *lim;
    pack_goals_kernel<<<(unsigned)grid, kPackThreads, 0, stream>>>(in, counts, (int)tiles_per_traj, l,
                                                                  (lim && lim->check_box) ? 1 : 0, records,
                                                                  rec_stride, rec_offset, rec_capacity);
    return cudaGetLastError();
}

}  // namespace tgx
