// streams.cu — probe: what does the transition kernel's WRITE PATTERN alone cost?  
// V threads each own a row of T 128-byte records and write one record per "tick" as four 32-byte streaming stores, with
// `work` dependent FP64 operations between ticks standing in for the recurrence; then the same bytes written so that a
// warp's 32 lanes cover one vehicle's consecutive records (what a warp-per-vehicle formulation would store).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/wbw/streams tools/wbw/streams.cu ; gpurun -- tools/wbw/streams
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void st4cs(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
// one thread per vehicle, one record per tick
template <int SECTORS>
__global__ void __launch_bounds__(128) per_thread(double* out, int V, int T, int work, double seed) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    double* row = out + (size_t)v * T * 16;
    double x = seed + v;
    for (int k = 0; k < T; ++k) {
        for (int w = 0; w < work; ++w) x = fma(x, 0.999999, 1e-9);
        st4cs(row + (size_t)k * 16, x, 1.0, 2.0, 3.0);
        if (SECTORS > 1) st4cs(row + (size_t)k * 16 + 12, 4.0, 5.0, 6.0, x);
        if (SECTORS > 2) st4cs(row + (size_t)k * 16 + 4, 7.0, 0.0, 0.0, 0.0);
        if (SECTORS > 3) st4cs(row + (size_t)k * 16 + 8, 0.0, 0.0, 0.0, 0.0);
    }
}
// one warp per vehicle: lane l writes sector (l & 3) of record 8*j + (l >> 2): 1 KiB contiguous per store instruction
__global__ void __launch_bounds__(128) per_warp(double* out, int V, int T, double seed) {
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (v >= V) return;
    double* row = out + (size_t)v * T * 16;
    for (int k = 0; k + 8 <= T; k += 8) st4cs(row + (size_t)(k + (lane >> 2)) * 16 + 4 * (lane & 3), seed + k, 1.0, 2.0, 3.0);
}
// one warp per 32 vehicles, all of whose rows advance together by R records per round (a thread-per-vehicle recurrence whose
// records are staged in shared memory for R ticks): a store instruction covers 8/R vehicles x R consecutive records
template <int R>
__global__ void __launch_bounds__(128) staged(double* out, int V, int T, double seed) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, j = threadIdx.x & 31;
    if (w * 32 >= V) return;
    constexpr int VPI = 8 / R;                       // vehicles per store instruction
    for (int k0 = 0; k0 + R <= T; k0 += R)
        for (int i = 0; i < 32 / VPI; ++i) {
            const int veh = w * 32 + i * VPI + j / (4 * R), tick = k0 + (j / 4) % R, sector = j & 3;
            st4cs(out + ((size_t)veh * T + tick) * 16 + 4 * sector, seed + tick, 1.0, 2.0, 3.0);
        }
}
int main(int argc, char** argv) {
    const int V = argc > 1 ? atoi(argv[1]) : 131072, T = argc > 2 ? atoi(argv[2]) : 1032;
    double* out;
    const size_t bytes = (size_t)V * T * 128;
    CK(cudaMalloc(&out, bytes));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    auto time = [&](auto launch) { launch(); CK(cudaDeviceSynchronize()); float best = 1e9f; for (int r = 0; r < 3; ++r) { CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); best = ms < best ? ms : best; } CK(cudaGetLastError()); return best; };
    printf("%d vehicles x %d records of 128 B = %.2f GB\n", V, T, bytes / 1e9);
    for (int work : {0, 50, 100, 200, 400}) {
        float ms = time([&] { per_thread<4><<<(V + 127) / 128, 128>>>(out, V, T, work, 1.0); });
        printf("one thread per vehicle, 4 sectors per tick, %3d dependent DFMA per tick : %7.3f ms  %7.1f GB/s\n", work, ms, bytes / ms / 1e6);
    }
    { float ms = time([&] { per_thread<2><<<(V + 127) / 128, 128>>>(out, V, T, 0, 1.0); });
      printf("one thread per vehicle, 2 sectors per tick (half the bytes), no work       : %7.3f ms  %7.1f GB/s (of the bytes written)\n", ms, bytes / 2 / ms / 1e6); }
    { float ms = time([&] { per_thread<1><<<(V + 127) / 128, 128>>>(out, V, T, 0, 1.0); });
      printf("one thread per vehicle, 1 sector per tick, no work                        : %7.3f ms  %7.1f GB/s (of the bytes written)\n", ms, bytes / 4 / ms / 1e6); }
    { float ms = time([&] { per_warp<<<(V * 32 + 127) / 128, 128>>>(out, V, T, 1.0); });
      printf("one warp per vehicle, 8 consecutive records per store instruction          : %7.3f ms  %7.1f GB/s\n", ms, (size_t)V * (T / 8 * 8) * 128 / ms / 1e6); }
    { float ms = time([&] { staged<8><<<(V + 127) / 128, 128>>>(out, V, T, 1.0); });
      printf("32 vehicles per warp advancing together, 8 records (1 KiB) per vehicle per round : %7.3f ms  %7.1f GB/s\n", ms, (size_t)V * (T / 8 * 8) * 128 / ms / 1e6); }
    { float ms = time([&] { staged<4><<<(V + 127) / 128, 128>>>(out, V, T, 1.0); });
      printf("32 vehicles per warp advancing together, 4 records (512 B) per vehicle per round : %7.3f ms  %7.1f GB/s\n", ms, (size_t)V * (T / 4 * 4) * 128 / ms / 1e6); }
    { float ms = time([&] { staged<2><<<(V + 127) / 128, 128>>>(out, V, T, 1.0); });
      printf("32 vehicles per warp advancing together, 2 records (256 B) per vehicle per round : %7.3f ms  %7.1f GB/s\n", ms, (size_t)V * (T / 2 * 2) * 128 / ms / 1e6); }
    { float ms = time([&] { staged<1><<<(V + 127) / 128, 128>>>(out, V, T, 1.0); });
      printf("32 vehicles per warp advancing together, 1 record (128 B) per vehicle per round  : %7.3f ms  %7.1f GB/s\n", ms, (size_t)V * T * 128 / ms / 1e6); }
    return 0;
}
