// wbw.cu — write-bandwidth pattern probe (tools only, not product code).
// Emulates the store pattern of tgx::eval_kernel (a CTA writes a [14][ROW] slab) with no arithmetic, next to plain
// fills, to find what limits the sampling kernel's HBM write rate.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void st2(double* p, double a, double b) {
    asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st4(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// P0: plain fill, 128-bit stores, each thread 4 stores spaced by blockDim (like torch's vectorized kernel)
__global__ void fill128(double* out, size_t n) {
    size_t base = (size_t)blockIdx.x * blockDim.x * 8 + threadIdx.x * 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        size_t k = base + (size_t)i * blockDim.x * 2;
        if (k + 1 < n) st2(out + k, 1.5, 2.5);
    }
}
// P0b: plain fill, 256-bit stores
__global__ void fill256(double* out, size_t n) {
    size_t base = (size_t)blockIdx.x * blockDim.x * 16 + threadIdx.x * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        size_t k = base + (size_t)i * blockDim.x * 4;
        if (k + 3 < n) st4(out + k, 1.5, 2.5, 3.5, 4.5);
    }
}
// P1: slab pattern, CTA = 256 threads, thread writes 4 doubles (256-bit) to each of NCH rows of ROW doubles
template <int NCH, int ROW>
__global__ void slab256(double* out, int valid) {
    double* slab = out + (size_t)blockIdx.x * NCH * ROW + threadIdx.x * 4;
    if ((int)threadIdx.x * 4 + 3 >= valid) return;
#pragma unroll
    for (int c = 0; c < NCH; ++c) st4(slab + (size_t)c * ROW, 1.5 + c, 2.5, 3.5, 4.5);
}
// P2: slab pattern with 128-bit stores: CTA = 256 threads, two passes per row (each warp instr = 512 B contiguous)
template <int NCH, int ROW>
__global__ void slab128(double* out, int valid) {
    double* slab = out + (size_t)blockIdx.x * NCH * ROW + threadIdx.x * 2;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        if ((int)threadIdx.x * 2 + 1 < valid) st2(slab + (size_t)c * ROW, 1.5 + c, 2.5);
        if ((int)threadIdx.x * 2 + 513 < valid) st2(slab + (size_t)c * ROW + 512, 3.5, 4.5);
    }
}
// P3: slab pattern 256-bit, 128 threads per CTA, tile 512 (two CTAs per row)
template <int NCH, int ROW>
__global__ void slab256_t512(double* out, int valid) {
    const int half = blockIdx.x & 1;
    double* slab = out + (size_t)(blockIdx.x >> 1) * NCH * ROW + half * 512 + threadIdx.x * 4;
    if (half * 512 + (int)threadIdx.x * 4 + 3 >= valid) return;
#pragma unroll
    for (int c = 0; c < NCH; ++c) st4(slab + (size_t)c * ROW, 1.5 + c, 2.5, 3.5, 4.5);
}
// P4: persistent slab writer: grid = SMs * k CTAs, each loops over slabs (removes CTA launch overhead)
template <int NCH, int ROW>
__global__ void slab256_persist(double* out, int nslab, int valid) {
    for (int s = blockIdx.x; s < nslab; s += gridDim.x) {
        double* slab = out + (size_t)s * NCH * ROW + threadIdx.x * 4;
        if ((int)threadIdx.x * 4 + 3 < valid) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) st4(slab + (size_t)c * ROW, 1.5 + c, 2.5, 3.5, 4.5);
        }
    }
}
// P5: slab via shared memory + bulk async copy (TMA 1-D): each row of ROW doubles staged in smem, one
// cp.async.bulk.global.shared::cta per row issued by one thread.
template <int NCH, int ROW>
__global__ void slab_bulk(double* out, int valid_bytes_per_row) {
    extern __shared__ __align__(128) double sm[];   // NCH * ROW doubles
    double* slab = out + (size_t)blockIdx.x * NCH * ROW;
    for (int c = 0; c < NCH; ++c) {
        double4* dst = reinterpret_cast<double4*>(sm + (size_t)c * ROW) + threadIdx.x;
        *dst = make_double4(1.5 + c, 2.5, 3.5, 4.5);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < NCH) {
        const int c = threadIdx.x;
        unsigned saddr = (unsigned)__cvta_generic_to_shared(sm + (size_t)c * ROW);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(slab + (size_t)c * ROW), "r"(saddr),
                     "r"(valid_bytes_per_row)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// P6: P1 + a dependent-load prologue like eval_kernel's (tile descriptor -> record -> __syncthreads)
template <int NCH, int ROW>
__global__ void slab256_prologue(double* out, const int4* __restrict__ tiles, const double4* __restrict__ recs, int valid) {
    __shared__ double4 s_rec[36];
    const int4 tw = __ldg(tiles + blockIdx.x);
    if (threadIdx.x < 36) s_rec[threadIdx.x] = recs[(size_t)tw.x * 36 + threadIdx.x];
    __syncthreads();
    double* slab = out + (size_t)blockIdx.x * NCH * ROW + threadIdx.x * 4;
    if ((int)threadIdx.x * 4 + 3 >= valid) return;
    const double a = s_rec[threadIdx.x & 31].x;
#pragma unroll
    for (int c = 0; c < NCH; ++c) st4(slab + (size_t)c * ROW, a + c, 2.5, 3.5, 4.5);
}
// P7: P1 + FP64 work between the stores (WORK dependent DFMAs before each store)
template <int NCH, int ROW, int WORK>
__global__ void slab256_work(double* out, int valid, double seed) {
    double* slab = out + (size_t)blockIdx.x * NCH * ROW + threadIdx.x * 4;
    if ((int)threadIdx.x * 4 + 3 >= valid) return;
    double x = seed + threadIdx.x;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
#pragma unroll
        for (int w = 0; w < WORK; ++w) x = fma(x, 1.0000001, 1e-9);
        st4(slab + (size_t)c * ROW, x, 2.5, 3.5, 4.5);
    }
}
// P8: P1 with a long FP64 preamble (all the work first, then 14 back-to-back stores)
template <int NCH, int ROW, int WORK>
__global__ void slab256_work_first(double* out, int valid, double seed) {
    double* slab = out + (size_t)blockIdx.x * NCH * ROW + threadIdx.x * 4;
    if ((int)threadIdx.x * 4 + 3 >= valid) return;
    double x = seed + threadIdx.x;
#pragma unroll 8
    for (int w = 0; w < WORK * NCH; ++w) x = fma(x, 1.0000001, 1e-9);
#pragma unroll
    for (int c = 0; c < NCH; ++c) st4(slab + (size_t)c * ROW, x + c, 2.5, 3.5, 4.5);
}
// P9: P1 with the register footprint of eval_kernel (launch bounds 256 x 3 -> 3 CTAs per SM) emulated by dynamic smem
template <int NCH, int ROW>
__global__ void slab256_occ(double* out, int valid) {
    extern __shared__ double dummy[];
    double* slab = out + (size_t)blockIdx.x * NCH * ROW + threadIdx.x * 4;
    if ((int)threadIdx.x * 4 + 3 >= valid) return;
    if (valid < 0) dummy[threadIdx.x] = 1.0;
#pragma unroll
    for (int c = 0; c < NCH; ++c) st4(slab + (size_t)c * ROW, 1.5 + c, 2.5, 3.5, 4.5);
}

// V1: non-persistent, each CTA writes K consecutive slabs
template <int NCH, int ROW, int K>
__global__ void slab256_k(double* out, int valid) {
    if ((int)threadIdx.x * 4 + 3 >= valid) return;
#pragma unroll 1
    for (int j = 0; j < K; ++j) {
        double* slab = out + ((size_t)blockIdx.x * K + j) * NCH * ROW + threadIdx.x * 4;
#pragma unroll
        for (int c = 0; c < NCH; ++c) st4(slab + (size_t)c * ROW, 1.5 + c, 2.5, 3.5, 4.5);
    }
}
// V2: prologue loaded per warp (no CTA barrier)
template <int NCH, int ROW>
__global__ void slab256_prologue_warp(double* out, const int4* __restrict__ tiles, const double4* __restrict__ recs, int valid) {
    __shared__ double4 s_rec[8][36];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int4 tw = __ldg(tiles + blockIdx.x);
    s_rec[warp][lane] = recs[(size_t)tw.x * 36 + lane];
    if (lane < 4) s_rec[warp][32 + lane] = recs[(size_t)tw.x * 36 + 32 + lane];
    __syncwarp();
    double* slab = out + (size_t)tw.x * NCH * ROW + threadIdx.x * 4;
    if ((int)threadIdx.x * 4 + 3 >= valid) return;
    const double a = s_rec[warp][lane].x;
#pragma unroll
    for (int c = 0; c < NCH; ++c) st4(slab + (size_t)c * ROW, a + c, 2.5, 3.5, 4.5);
}
// V3: CTA handles K tiles; ALL the K prologue loads are issued before any store (loads ahead of the store queue)
template <int NCH, int ROW, int K>
__global__ void slab256_prefetch(double* out, const int4* __restrict__ tiles, const double4* __restrict__ recs, int valid) {
    __shared__ double4 s_rec[K][36];
    int4 tw[K];
#pragma unroll
    for (int j = 0; j < K; ++j) tw[j] = __ldg(tiles + (size_t)blockIdx.x * K + j);
#pragma unroll
    for (int j = 0; j < K; ++j)
        if (threadIdx.x < 36) s_rec[j][threadIdx.x] = recs[(size_t)tw[j].x * 36 + threadIdx.x];
    __syncthreads();
    if ((int)threadIdx.x * 4 + 3 >= valid) return;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        double* slab = out + (size_t)tw[j].x * NCH * ROW + threadIdx.x * 4;
        const double a = s_rec[j][threadIdx.x & 31].x;
#pragma unroll
        for (int c = 0; c < NCH; ++c) st4(slab + (size_t)c * ROW, a + c, 2.5, 3.5, 4.5);
    }
}
// V4: like P6 but only ONE 16-byte load in the prologue (tile descriptor), no second dependent load
template <int NCH, int ROW>
__global__ void slab256_oneload(double* out, const int4* __restrict__ tiles, int valid) {
    const int4 tw = __ldg(tiles + blockIdx.x);
    double* slab = out + (size_t)tw.x * NCH * ROW + threadIdx.x * 4;
    if ((int)threadIdx.x * 4 + 3 >= valid) return;
#pragma unroll
    for (int c = 0; c < NCH; ++c) st4(slab + (size_t)c * ROW, 1.5 + c + tw.y, 2.5, 3.5, 4.5);
}

// V5: ONE round of independent loads (address from blockIdx, as in a slab-mode plan): NB bytes per tile, per warp
template <int NCH, int ROW, int NB>
__global__ void slab256_oneround(double* out, const double4* __restrict__ recs, int valid) {
    __shared__ double4 s_rec[8][36];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane * 32 < NB) s_rec[warp][lane] = recs[(size_t)blockIdx.x * 36 + lane];
    if (NB > 1024 && lane < 4) s_rec[warp][32 + lane] = recs[(size_t)blockIdx.x * 36 + 32 + lane];
    __syncwarp();
    double* slab = out + (size_t)blockIdx.x * NCH * ROW + threadIdx.x * 4;
    if ((int)threadIdx.x * 4 + 3 >= valid) return;
    const double a = s_rec[warp][0].x;
#pragma unroll
    for (int c = 0; c < NCH; ++c) st4(slab + (size_t)c * ROW, a + c, 2.5, 3.5, 4.5);
}
// V6: two DEPENDENT rounds but tiny: tile (16 B) -> 64 B record
template <int NCH, int ROW>
__global__ void slab256_twosmall(double* out, const int4* __restrict__ tiles, const double4* __restrict__ recs, int valid) {
    const int4 tw = __ldg(tiles + blockIdx.x);
    const double4 r0 = recs[(size_t)tw.x * 36];
    const double4 r1 = recs[(size_t)tw.x * 36 + 1];
    double* slab = out + (size_t)tw.x * NCH * ROW + threadIdx.x * 4;
    if ((int)threadIdx.x * 4 + 3 >= valid) return;
#pragma unroll
    for (int c = 0; c < NCH; ++c) st4(slab + (size_t)c * ROW, r0.x + r1.y + c, 2.5, 3.5, 4.5);
}

// V7: the per-tile record fetched by the TMA unit (1-D bulk async copy + mbarrier) instead of LSU loads
template <int NCH, int ROW>
__global__ void slab256_tma(double* out, const double4* __restrict__ recs, int valid) {
    __shared__ __align__(128) double4 s_rec[36];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar);
    const unsigned dst_a = (unsigned)__cvta_generic_to_shared(s_rec);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(1152) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_a),
                     "l"(recs + (size_t)blockIdx.x * 36), "r"(1152), "r"(bar_a)
                     : "memory");
    }
    // all threads wait for phase 0
    {
        unsigned done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar_a), "r"(0) : "memory");
        }
    }
    double* slab = out + (size_t)blockIdx.x * NCH * ROW + threadIdx.x * 4;
    if ((int)threadIdx.x * 4 + 3 >= valid) return;
    const double a = s_rec[threadIdx.x & 31].x;
#pragma unroll
    for (int c = 0; c < NCH; ++c) st4(slab + (size_t)c * ROW, a + c, 2.5, 3.5, 4.5);
}

template <class F>
double timeit(F launch, int reps = 5) {
    for (int i = 0; i < 2; ++i) launch();
    CK(cudaDeviceSynchronize());
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(b);
    CK(cudaDeviceSynchronize());
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char** argv) {
    const int nslab = argc > 1 ? atoi(argv[1]) : (1 << 19);
    constexpr int NCH = 14, ROW = 1024;
    const size_t n = (size_t)nslab * NCH * ROW;
    double* out;
    CK(cudaMalloc(&out, n * sizeof(double)));
    const double gb = n * 8 / 1e9;
    const int valid = 1000;
    const double gbv = (double)nslab * NCH * valid * 8 / 1e9;
    printf("buffer %.1f GB, %d slabs of [14][1024] doubles\n", gb, nslab);
    double ms;
    ms = timeit([&] { fill128<<<(unsigned)((n + 1023) / 1024), 128>>>(out, n); });
    printf("P0  fill 128-bit, 128 thr                 : %7.3f ms  %7.1f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { fill256<<<(unsigned)((n + 2047) / 2048), 128>>>(out, n); });
    printf("P0b fill 256-bit, 128 thr                 : %7.3f ms  %7.1f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { slab256<NCH, ROW><<<nslab, 256>>>(out, 1024); });
    printf("P1  slab 256-bit, 256 thr, full rows      : %7.3f ms  %7.1f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { slab256<NCH, ROW><<<nslab, 256>>>(out, valid); });
    printf("P1v slab 256-bit, 256 thr, 1000 of 1024   : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
    ms = timeit([&] { slab128<NCH, ROW><<<nslab, 256>>>(out, 1024); });
    printf("P2  slab 128-bit x2, 256 thr, full rows   : %7.3f ms  %7.1f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { slab128<NCH, ROW><<<nslab, 256>>>(out, valid); });
    printf("P2v slab 128-bit x2, 256 thr, 1000        : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
    ms = timeit([&] { slab256_t512<NCH, ROW><<<nslab * 2, 128>>>(out, 1024); });
    printf("P3  slab 256-bit, 128 thr, tile 512       : %7.3f ms  %7.1f GB/s\n", ms, gb / ms * 1e3);
    for (int per_sm : {2, 4, 8}) {
        ms = timeit([&] { slab256_persist<NCH, ROW><<<148 * per_sm, 256>>>(out, nslab, 1024); });
        printf("P4  persistent slab 256-bit, %d CTA/SM      : %7.3f ms  %7.1f GB/s\n", per_sm, ms, gb / ms * 1e3);
    }
    {
        const int smem = NCH * ROW * 8;
        CK(cudaFuncSetAttribute(slab_bulk<NCH, ROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ms = timeit([&] { slab_bulk<NCH, ROW><<<nslab, 256, smem>>>(out, ROW * 8); });
        printf("P5  slab via smem + cp.async.bulk (TMA)   : %7.3f ms  %7.1f GB/s\n", ms, gb / ms * 1e3);
        ms = timeit([&] { slab_bulk<NCH, ROW><<<nslab, 256, smem>>>(out, 1000 * 8); });
        printf("P5v same, 8000 of 8192 bytes per row      : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
    }
    {
        int4* tiles; double4* recs;
        CK(cudaMalloc(&tiles, (size_t)nslab * 16));
        CK(cudaMalloc(&recs, (size_t)nslab * 36 * 32));
        CK(cudaMemset(recs, 0, (size_t)nslab * 36 * 32));
        int4* h = (int4*)malloc((size_t)nslab * 16);
        for (int i = 0; i < nslab; ++i) h[i] = make_int4(i, 0, i * 8, 8);
        CK(cudaMemcpy(tiles, h, (size_t)nslab * 16, cudaMemcpyHostToDevice));
        ms = timeit([&] { slab256_prologue<NCH, ROW><<<nslab, 256>>>(out, tiles, recs, valid); });
        printf("P6  P1v + dependent-load prologue + sync  : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
        {   // same prologue, but every tile reads the SAME record (always cache-resident: no DRAM reads at all)
            int4* tiles0; CK(cudaMalloc(&tiles0, (size_t)nslab * 16));
            CK(cudaMemset(tiles0, 0, (size_t)nslab * 16));
            ms = timeit([&] { slab256_prologue<NCH, ROW><<<nslab, 256>>>(out, tiles0, recs, valid); });
            printf("P6s P6 with one shared record (no DRAM rd): %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
            cudaFree(tiles0);
        }
        for (int mod : {256, 4096, 65536}) {   // records reused cyclically: L2-resident (not L1-resident) tables
            int4* tm; CK(cudaMalloc(&tm, (size_t)nslab * 16));
            int4* hm = (int4*)malloc((size_t)nslab * 16);
            for (int i = 0; i < nslab; ++i) hm[i] = make_int4((int)(((long long)i * 2654435761LL) % mod), 0, 0, 8);
            CK(cudaMemcpy(tm, hm, (size_t)nslab * 16, cudaMemcpyHostToDevice));
            ms = timeit([&] { slab256_prologue<NCH, ROW><<<nslab, 256>>>(out, tm, recs, valid); });
            printf("P6c P6 reading from a %5d-record table (%.1f MB, L2-resident): %7.3f ms  %7.1f GB/s\n", mod, mod * 1152 / 1e6, ms, gbv / ms * 1e3);
            cudaFree(tm); free(hm);
        }
        ms = timeit([&] { slab256_tma<NCH, ROW><<<nslab, 256>>>(out, recs, valid); });
        printf("V7  record fetched by TMA bulk copy + mbarrier (own record per tile): %7.3f ms  %7.1f GB/s\n", ms, gbv / ms * 1e3);
        ms = timeit([&] { slab256_prologue_warp<NCH, ROW><<<nslab, 256>>>(out, tiles, recs, valid); });
        printf("V2  prologue per warp, no CTA barrier     : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
        ms = timeit([&] { slab256_prefetch<NCH, ROW, 2><<<nslab / 2, 256>>>(out, tiles, recs, valid); });
        printf("V3  2 tiles per CTA, loads before stores  : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
        ms = timeit([&] { slab256_prefetch<NCH, ROW, 4><<<nslab / 4, 256>>>(out, tiles, recs, valid); });
        printf("V3  4 tiles per CTA, loads before stores  : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
        ms = timeit([&] { slab256_oneround<NCH, ROW, 64><<<nslab, 256>>>(out, recs, valid); });
        printf("V5  one round, 64 B per tile (per warp)   : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
        ms = timeit([&] { slab256_oneround<NCH, ROW, 256><<<nslab, 256>>>(out, recs, valid); });
        printf("V5  one round, 256 B per tile (per warp)  : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
        ms = timeit([&] { slab256_oneround<NCH, ROW, 512><<<nslab, 256>>>(out, recs, valid); });
        printf("V5  one round, 512 B per tile (per warp)  : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
        ms = timeit([&] { slab256_oneround<NCH, ROW, 1152><<<nslab, 256>>>(out, recs, valid); });
        printf("V5  one round, 1152 B per tile (per warp) : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
        ms = timeit([&] { slab256_twosmall<NCH, ROW><<<nslab, 256>>>(out, tiles, recs, valid); });
        printf("V6  two dependent rounds, 16 B -> 64 B    : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
        ms = timeit([&] { slab256_oneload<NCH, ROW><<<nslab, 256>>>(out, tiles, valid); });
        printf("V4  one 16-byte load, no barrier          : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
        cudaFree(tiles); cudaFree(recs); free(h);
    }
    ms = timeit([&] { slab256_k<NCH, ROW, 2><<<nslab / 2, 256>>>(out, valid); });
    printf("V1  2 slabs per CTA (no loads)            : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
    ms = timeit([&] { slab256_k<NCH, ROW, 8><<<nslab / 8, 256>>>(out, valid); });
    printf("V1  8 slabs per CTA (no loads)            : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
    ms = timeit([&] { slab256_k<NCH, ROW, 64><<<nslab / 64, 256>>>(out, valid); });
    printf("V1  64 slabs per CTA (no loads)           : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
    ms = timeit([&] { slab256_work<NCH, ROW, 8><<<nslab, 256>>>(out, valid, 1.0); });
    printf("P7a P1v + 8 dependent DFMA before each st : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
    ms = timeit([&] { slab256_work<NCH, ROW, 32><<<nslab, 256>>>(out, valid, 1.0); });
    printf("P7b P1v + 32 dependent DFMA before each st: %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
    ms = timeit([&] { slab256_work_first<NCH, ROW, 8><<<nslab, 256>>>(out, valid, 1.0); });
    printf("P8a P1v + 112 DFMA first, then 14 stores  : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
    ms = timeit([&] { slab256_work_first<NCH, ROW, 32><<<nslab, 256>>>(out, valid, 1.0); });
    printf("P8b P1v + 448 DFMA first, then 14 stores  : %7.3f ms  %7.1f GB/s (valid bytes)\n", ms, gbv / ms * 1e3);
    for (int kb : {24, 48, 72, 110}) {
        CK(cudaFuncSetAttribute(slab256_occ<NCH, ROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024));
        ms = timeit([&] { slab256_occ<NCH, ROW><<<nslab, 256, kb * 1024>>>(out, valid); });
        printf("P9  P1v at %d CTAs/SM (smem-limited)        : %7.3f ms  %7.1f GB/s (valid bytes)\n", 227 / kb, ms, gbv / ms * 1e3);
    }
    ms = timeit([&] { CK(cudaMemsetAsync(out, 0, n * 8)); });
    printf("M   cudaMemsetAsync                       : %7.3f ms  %7.1f GB/s\n", ms, gb / ms * 1e3);
    CK(cudaGetLastError());
    cudaFree(out);
    return 0;
}
