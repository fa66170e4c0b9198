// params_gen.cu — synthetic parameter records drawn on the device (BASELINE.json configs[3]-[4]).
//
// The 10^8-trajectory sweep of config 5 must not start with a 12.8 GB host->device copy (SURVEY.md §8d "Config 5"),
// so every shard draws its own records: one thread per trajectory, a counter-based generator (Philox4x32-10, Salmon
// et al., SC'11: key = seed, counter = (global trajectory index, draw number)), plain IEEE arithmetic compiled with
// -fmad=false.  Any record is therefore reproducible on the host from (seed, index) alone —
// trajectory_generator_ros2_b200/workloads.py: montecarlo_philox does the same arithmetic in numpy, and the parity
// tests draw their checked subset from it.  Nothing here samples a trajectory: this is workload preparation.
#include <cuda_runtime.h>

#include "tgx_internal.cuh"

namespace tgx {

namespace {

struct Philox4 {
    uint32_t x[4];
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    return Philox4{{c0, c1, c2, c3}};
}

// 53 random bits -> [0, 1): ((hi << 32 | lo) >> 11) * 2^-53, then lo + (hi - lo) * u with two roundings (numpy's
// Generator.uniform does the same two operations)
__device__ __forceinline__ double uniform(uint32_t lo32, uint32_t hi32, double a, double b) {
    const uint64_t bits = (((uint64_t)hi32 << 32) | lo32) >> 11;
    const double u = __dmul_rn((double)bits, 0x1p-53);
    return __dadd_rn(a, __dmul_rn(__dsub_rn(b, a), u));
}

// The config-4 distribution: circles, r ~ U[0.2, 5], centre ~ U[-2, 2]^2, alt ~ U[1, 2.5],
// v_goal ~ U[0.2, 8], accel ~ U[0.7, 2], dt = 0.01, t_traj = max(9.98 - 2 v / a, 0.5): ~1000 samples each.
__global__ void __launch_bounds__(256)
fill_montecarlo_kernel(uint64_t seed, int64_t first, int64_t n, tgx_params* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint64_t idx = (uint64_t)(first + t);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const uint32_t i0 = (uint32_t)idx, i1 = (uint32_t)(idx >> 32);
    const Philox4 a = philox4x32_10(i0, i1, 0u, 0u, k0, k1);
    const Philox4 b = philox4x32_10(i0, i1, 1u, 0u, k0, k1);
    const Philox4 c = philox4x32_10(i0, i1, 2u, 0u, k0, k1);
    tgx_params p;
    p.type = TGX_CIRCLE;
    p.n_vgoals = 1;
    p.dt = 0.01;
    p.u.orbit.r = uniform(a.x[0], a.x[1], 0.2, 5.0);
    p.u.orbit.cx = uniform(a.x[2], a.x[3], -2.0, 2.0);
    p.u.orbit.cy = uniform(b.x[0], b.x[1], -2.0, 2.0);
    p.alt = uniform(b.x[2], b.x[3], 1.0, 2.5);
    const double v = uniform(c.x[0], c.x[1], 0.2, 8.0);
    const double acc = uniform(c.x[2], c.x[3], 0.7, 2.0);
    p.u.orbit.accel = acc;
    const double hold = __dsub_rn(9.98, __ddiv_rn(__dmul_rn(2.0, v), acc));
    p.u.orbit.t_traj = hold < 0.5 ? 0.5 : hold;
    p.u.orbit.v_goals[0] = v;
#pragma unroll
    for (int g = 1; g < TGX_MAX_VGOALS; ++g) p.u.orbit.v_goals[g] = 0.0;
    const int4* src = reinterpret_cast<const int4*>(&p);
    int4* dst = reinterpret_cast<int4*>(out + t);
#pragma unroll
    for (int q = 0; q < 8; ++q) dst[q] = src[q];
}

}  // namespace

cudaError_t launch_fill_montecarlo(uint64_t seed, int64_t first, int64_t n, tgx_params* out, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    fill_montecarlo_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(seed, first, n, out);
    return cudaGetLastError();
}

}  // namespace tgx
