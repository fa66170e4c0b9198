// store.cuh — vector stores of the struct-of-arrays output planes and the per-trajectory max reduction helpers,
// shared by the evaluation kernels (eval.cu, polyline.cu).
#pragma once

#include <cuda_runtime.h>

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

// Store SPT adjacent samples of one channel; nvalid < SPT only on a trajectory's last, partial vector.
template <int SPT>
__device__ __forceinline__ void store_channel(double* p, const double (&x)[SPT], int nvalid) {
    if (nvalid >= SPT) {
        VecStore<SPT, TGX_STORE_POLICY>::st(p, x);
    } else {
#pragma unroll
        for (int u = 0; u < SPT; ++u)
            if (u < nvalid) __stcs(p + u, x[u]);
    }
}

__device__ __forceinline__ double warp_max(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

__device__ __forceinline__ void atomic_max_nonneg(double* addr, double x) {
    // For non-negative doubles the IEEE bit pattern is monotone in the value.
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(x));
}

}  // namespace
}  // namespace tgx
