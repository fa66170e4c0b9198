// gather_flags.cpp — C++ driver of the multi-GPU split through the C-ABI alone (no Python, no torch):
// one process, every visible GPU, ncclCommInitAll behind tgx_comm_init_all, one stream per device.
//
// BASELINE.json configs[4] in miniature: a batch of n_total Monte-Carlo circles is sharded with tgx_shard_range,
// every GPU draws its own shard on the device (tgx_fill_montecarlo), plans it, reduces max |v| / max |a|
// (tgx_feasibility) and the 1-byte flags are all-gathered (tgx_gather_flags).  Checks: every device ends up with the
// same full vector, and that vector is the concatenation of the shards' local flags.  Run with an odd n_total as well
// so that the shards differ in length (the grouped-broadcast path).  Exit code 0 = pass.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tgx.h"

#define CK(call)                                                                                         \
    do {                                                                                                 \
        const int rc__ = (call);                                                                         \
        if (rc__ != TGX_OK) {                                                                            \
            std::fprintf(stderr, "%s failed: %s (%s | %s)\n", #call, tgx_strerror(rc__), tgx_last_cuda_error(), \
                         tgx_comm_last_error());                                                         \
            return 1;                                                                                    \
        }                                                                                                \
    } while (0)
#define CU(call)                                                                           \
    do {                                                                                   \
        const cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                          \
            std::fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(e__));       \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

static int run(int ndev, int64_t n_total) {
    std::vector<tgx_engine*> eng(ndev, nullptr);
    std::vector<tgx_comm*> comm(ndev, nullptr);
    std::vector<cudaStream_t> stream(ndev);
    std::vector<tgx_params*> d_params(ndev, nullptr);
    std::vector<uint8_t*> d_local(ndev, nullptr), d_all(ndev, nullptr);
    std::vector<int64_t> lo(ndev), hi(ndev);
    tgx_limits lim;
    std::memset(&lim, 0, sizeof(lim));
    const double box[6] = {-5, 5, -5, 5, -5, 5};
    std::memcpy(lim.box, box, sizeof(box));
    lim.check_box = 1;
    lim.v_max = 5.0;
    lim.a_max = 6.0;

    CK(tgx_comm_init_all(comm.data(), ndev, nullptr));
    for (int g = 0; g < ndev; ++g) {
        CU(cudaSetDevice(g));
        CK(tgx_create(&eng[g], g));
        CU(cudaStreamCreateWithFlags(&stream[g], cudaStreamNonBlocking));
        CK(tgx_shard_range(n_total, g, ndev, &lo[g], &hi[g]));
        const int64_t m = hi[g] - lo[g];
        CU(cudaMalloc(&d_params[g], (size_t)(m > 0 ? m : 1) * sizeof(tgx_params)));
        CU(cudaMalloc(&d_local[g], (size_t)(m > 0 ? m : 1)));
        CU(cudaMalloc(&d_all[g], (size_t)n_total));
        CU(cudaMemsetAsync(d_all[g], 0xff, (size_t)n_total, stream[g]));
    }
    // every GPU: draw, plan, reduce (asynchronous per device apart from the plan's own sizing sync)
    for (int g = 0; g < ndev; ++g) {
        CU(cudaSetDevice(g));
        const int64_t m = hi[g] - lo[g];
        CK(tgx_fill_montecarlo(eng[g], 1237, lo[g], m, d_params[g], stream[g]));
        CK(tgx_plan(eng[g], d_params[g], m, &lim, nullptr, nullptr, nullptr, nullptr, stream[g]));
        CK(tgx_feasibility(eng[g], &lim, d_local[g], nullptr, nullptr, nullptr, stream[g]));
    }
    // the exchange: one group, one call per device
    CK(tgx_comm_group_start());
    for (int g = 0; g < ndev; ++g) CK(tgx_gather_flags(comm[g], d_local[g], n_total, d_all[g], stream[g]));
    CK(tgx_comm_group_end());
    std::vector<uint8_t> want((size_t)n_total), got((size_t)n_total);
    for (int g = 0; g < ndev; ++g) {
        CU(cudaSetDevice(g));
        CU(cudaStreamSynchronize(stream[g]));
        CU(cudaMemcpy(want.data() + lo[g], d_local[g], (size_t)(hi[g] - lo[g]), cudaMemcpyDeviceToHost));
    }
    int64_t feasible = 0;
    for (int64_t i = 0; i < n_total; ++i) {
        if (want[i] > 1) {
            std::fprintf(stderr, "flag %lld is %d\n", (long long)i, want[i]);
            return 1;
        }
        feasible += want[i];
    }
    for (int g = 0; g < ndev; ++g) {
        CU(cudaSetDevice(g));
        CU(cudaMemcpy(got.data(), d_all[g], (size_t)n_total, cudaMemcpyDeviceToHost));
        if (std::memcmp(got.data(), want.data(), (size_t)n_total) != 0) {
            std::fprintf(stderr, "device %d: gathered flags differ from the shards' own\n", g);
            return 1;
        }
    }
    if (feasible == 0 || feasible == n_total) {
        std::fprintf(stderr, "degenerate sweep: %lld of %lld feasible\n", (long long)feasible, (long long)n_total);
        return 1;
    }
    std::printf("gather_flags ok: %d GPU(s), %lld trajectories, %lld feasible, shards %s\n", ndev, (long long)n_total,
                (long long)feasible, n_total % ndev == 0 ? "equal (ncclAllGather)" : "unequal (grouped ncclBroadcast)");
    for (int g = 0; g < ndev; ++g) {
        CU(cudaSetDevice(g));
        cudaFree(d_params[g]);
        cudaFree(d_local[g]);
        cudaFree(d_all[g]);
        cudaStreamDestroy(stream[g]);
        tgx_destroy(eng[g]);
        CK(tgx_comm_destroy(comm[g]));
    }
    return 0;
}

int main(int argc, char** argv) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        std::fprintf(stderr, "no CUDA device\n");
        return 2;
    }
    if (ndev > 8) ndev = 8;
    if (argc > 1 && std::atoi(argv[1]) > 0 && std::atoi(argv[1]) < ndev) ndev = std::atoi(argv[1]);
    int ver = 0;
    if (tgx_comm_nccl_version(&ver) != TGX_OK) {
        std::fprintf(stderr, "%s\n", tgx_comm_last_error());
        return 1;
    }
    std::printf("NCCL %d, %d device(s)\n", ver, ndev);
    if (run(ndev, (int64_t)1 << 18)) return 1;
    if (run(ndev, ((int64_t)1 << 18) + 3)) return 1;
    return 0;
}
