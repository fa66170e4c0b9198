// probe.cu — measurement probes that belong to the library because they must be hand-written kernels.
//
// tgx_probe_dfma: the FP64 roofline denominator of the reduction-only path (BASELINE.json configs[3]-[4]).  That path
// writes 17 bytes per trajectory, so it is bound by the FP64 pipe and by instruction issue, not by HBM; SURVEY.md §8d
// asks for it to be quoted "against a DFMA micro-benchmark measured on the same box" because MEASURED_PEAKS.json has no
// FP64 figure.  The kernel is the textbook peak probe: every thread owns kChains independent accumulator chains
// (no dependence between consecutive DFMAs of a thread until the chain wraps, 8 deep against a 4-8 cycle pipe),
// 1024 resident threads per SM, one full wave, enough iterations that launch overhead vanishes; nothing is read, one
// double per thread is written so the loop cannot be optimised away.
#include <cuda_runtime.h>

#include "tgx_internal.cuh"

namespace tgx {

namespace {

constexpr int kChains = 8;
constexpr int kUnroll = 16;      // DFMAs per chain per loop trip

__global__ void __launch_bounds__(256, 4)
dfma_probe_kernel(int trips, double a, double b, double* __restrict__ sink) {
    double x[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) x[c] = (double)(threadIdx.x + c) * 1e-3;
#pragma unroll 1
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
            for (int c = 0; c < kChains; ++c) x[c] = fma(x[c], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += x[c];
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

// One launch; returns the number of DFMA instructions (thread-level) it executes.
cudaError_t launch_dfma_probe(int ctas, int trips, double* sink, double* dfma_count, cudaStream_t stream) {
    dfma_probe_kernel<<<ctas, 256, 0, stream>>>(trips, 0.999999, 1e-9, sink);
    *dfma_count = (double)ctas * 256.0 * (double)trips * kUnroll * kChains;
    return cudaGetLastError();
}

}  // namespace tgx
