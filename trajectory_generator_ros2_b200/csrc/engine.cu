// engine.cu — host side of libtgx: the C-ABI declared in include/tgx.h.
//
// Owns the plan tables (TrajRec / Seg / Tile) in device memory, sequences the planning passes
// (count -> exclusive scans -> fill) and the evaluation kernel on the caller's stream, and provides the
// host-buffer convenience calls the drop-in C++ classes use (chunked, double-buffered H2D / eval / D2H).
// There is no CPU implementation of any sampler in this library: if CUDA is unavailable every entry point
// that would compute returns TGX_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iterator>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <pthread.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>

#include "tgx_internal.cuh"

namespace tgx {

cudaError_t launch_build_cur_table(const tgx_params* params, int64_t max_samples, void* table, cudaStream_t stream);
size_t cur_table_bytes();
cudaError_t launch_plan_count(const tgx_params* params, const double* stop_from, int64_t n, const tgx_limits* lim,
                              int64_t max_samples, int tile_shift, bool exact_ramps, const void* cur_table,
                              int32_t* counts, uint32_t* status, int32_t* nseg, int32_t* ntile,
                              cudaStream_t stream);
cudaError_t launch_plan_fill(const tgx_params* params, const double* stop_from, int64_t n, const tgx_limits* lim,
                             int64_t max_samples, int tile_shift, bool exact_ramps, const void* cur_table,
                             const int32_t* plan_counts, const int64_t* seg_off, const int64_t* tile_off,
                             int seg_slab, int tile_slab, TrajRec* recs, Seg* segs, Tile* tiles, int32_t* counts,
                             uint32_t* status, int32_t* counts2, uint32_t* status2, tgx_phases* phases,
                             PlanStats* stats, cudaStream_t stream, const int32_t* order = nullptr);
cudaError_t launch_chunk_bounds(const tgx_params* params, int64_t n, int64_t chunk, int nb, int64_t* bounds,
                                cudaStream_t stream);
cudaError_t launch_replay_keys(const tgx_params* params, int64_t n, uint8_t* key, int32_t* idx, cudaStream_t stream);
cudaError_t launch_count_used_tiles(int64_t n, int tile_slab, const Tile* slots, int32_t* ntile, cudaStream_t stream);
cudaError_t launch_compact_tiles(int64_t n, int tile_slab, const Tile* slots, const int64_t* tile_off, Tile* dense,
                                 cudaStream_t stream);
cudaError_t launch_plan_phase(const tgx_params* params, int64_t n, const tgx_limits* lim, int64_t max_samples,
                              int tile_shift, int max_n, const void* cur_table, PhaseRec* phase, PhaseExt* phase_ext,
                              int32_t* counts, uint32_t* status, int32_t* counts2, uint32_t* status2,
                              tgx_phases* phases, PlanStats* stats, cudaStream_t stream, const int32_t* order);
cudaError_t launch_eval(const TableView& tv, int64_t ntiles, int tile_shift, int spt, const OutView& out, bool store,
                        double* max_v, double* max_a, cudaStream_t stream, const RecOut* ptma = nullptr);
cudaError_t launch_feasibility_finalize(int64_t n, const uint32_t* plan_status, const double* max_v,
                                        const double* max_a, double v_max, double a_max, uint8_t* flags,
                                        uint32_t* status_out, cudaStream_t stream);

cudaError_t launch_dfma_probe(int ctas, int trips, double* sink, double* dfma_count, cudaStream_t stream);
cudaError_t launch_fill_montecarlo(uint64_t seed, int64_t first, int64_t n, tgx_params* out, cudaStream_t stream);
cudaError_t launch_selftest_division(int64_t n, uint64_t seed, int per_thread, unsigned long long* mismatches,
                                     cudaStream_t stream);
cudaError_t launch_plan_samples(const tgx_params* params, const double* state, int64_t n, TrajRec* recs, Seg* segs,
                                Tile* tiles, int32_t* counts, uint32_t* status, cudaStream_t stream);

cudaError_t launch_plan_poly(const tgx_params* params, int64_t n, const tgx_limits* lim, int64_t max_samples,
                             int tile_shift, const void* cur_table, void* recs, int32_t* counts, uint32_t* status,
                             int32_t* counts2, uint32_t* status2, tgx_polyline_legs* legs, int32_t* ntile,
                             PlanStats* stats, bool skip_foreign, cudaStream_t stream);
size_t poly_rec_bytes();
cudaError_t launch_poly_tiles(int64_t n, const int32_t* ntile, const int64_t* tile_off, int tile_shift, Tile* tiles,
                              cudaStream_t stream);
cudaError_t launch_pack_goals(const OutView& in, const int32_t* counts, int64_t n, const tgx_limits* lim,
                              tgx_goal_record* records, int64_t rec_stride, const int64_t* rec_offset,
                              int64_t rec_capacity, cudaStream_t stream);
cudaError_t launch_transitions(const tgx_transition_params* tparams, int64_t n, const tgx_limits* lim,
                               int64_t max_samples, tgx_goal_record* records, int64_t rec_stride,
                               int64_t rec_capacity, int32_t* counts, uint32_t* status, cudaStream_t stream);
cudaError_t launch_eval_records(const TableView& tv, int64_t ntiles, int tile_shift, int spt, const RecOut& ro,
                                cudaStream_t stream);
cudaError_t launch_eval_poly_records(const PolyView& pv, int64_t ntiles, int tile_shift, int spt, const RecOut& ro,
                                     cudaStream_t stream);
cudaError_t launch_eval_poly(const PolyView& pv, int64_t ntiles, int tile_shift, int spt, const OutView& out,
                             bool store, double* max_v, double* max_a, cudaStream_t stream);

}  // namespace tgx

namespace {

// Zero-fill on the SMs.  cudaMemsetAsync of a staging buffer this size is executed by a copy engine — the one the
// other slot's device-to-host copies are queued on — so the memset of chunk c+1, and with it that chunk's planning and
// evaluation, waited for ALL of chunk c's copies: no overlap at all, 0.46 ms of idle link per 14 ms chunk (events around
// the copy blocks showed it).  A kernel does not queue behind the copies.
__global__ void __launch_bounds__(256) zero_fill_kernel(int4* __restrict__ p, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) p[i] = make_int4(0, 0, 0, 0);
}

// A few dozen bytes from device memory to PINNED host memory, written by an SM instead of a copy engine.  The
// device-to-host copy engine works through its copies in submission order, so a cudaMemcpyAsync of a plan's 64 bytes of
// statistics queued behind the 0.8 GB of sample planes the other staging slot was shipping: tgx_plan's host
// synchronisation then lasted as long as those copies, and the chunks of a host-buffer call never overlapped (events
// around the copy blocks: every chunk's evaluation finished 0.46 ms AFTER the previous chunk's last copy).
__global__ void peek_kernel(const unsigned* __restrict__ src, volatile unsigned* __restrict__ host_dst, int words) {
    for (int i = threadIdx.x; i < words; i += blockDim.x) host_dst[i] = src[i];
    __threadfence_system();
}

cudaError_t peek_to_host(const void* d_src, void* h_dst_pinned, size_t bytes, cudaStream_t s) {
    void* mapped = nullptr;
    if ((bytes & 3u) || cudaHostGetDevicePointer(&mapped, h_dst_pinned, 0) != cudaSuccess) {
        cudaGetLastError();
        return cudaMemcpyAsync(h_dst_pinned, d_src, bytes, cudaMemcpyDeviceToHost, s);
    }
    peek_kernel<<<1, 32, 0, s>>>(static_cast<const unsigned*>(d_src), static_cast<volatile unsigned*>(mapped),
                                 (int)(bytes / 4));
    return cudaGetLastError();
}

__global__ void zero_words_kernel(unsigned* __restrict__ p, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) p[i] = 0u;
}

cudaError_t zero_fill(void* p, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(p) & 3u) || (bytes & 3u)) return cudaMemsetAsync(p, 0, bytes, s);
    if ((reinterpret_cast<uintptr_t>(p) & 15u) || (bytes & 15u)) {
        const size_t n4 = bytes / 4;
        zero_words_kernel<<<(unsigned)std::min<size_t>((n4 + 255) / 256, 148 * 8), 256, 0, s>>>(static_cast<unsigned*>(p), n4);
        return cudaGetLastError();
    }
    const size_t n16 = bytes / 16;
    const unsigned grid = (unsigned)std::min<size_t>((n16 + 255) / 256, 148 * 16);
    zero_fill_kernel<<<grid, 256, 0, s>>>(static_cast<int4*>(p), n16);
    return cudaGetLastError();
}

thread_local std::string g_last_cuda_error;

int cuda_fail(cudaError_t e, const char* what) {
    g_last_cuda_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return TGX_ERR_CUDA;
}

#define TGX_CUDA(call)                                   \
    do {                                                 \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

// A growable device buffer.
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int reserve(size_t want) {
        if (want <= bytes) return TGX_OK;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        const size_t grown = want + want / 8 + 256;
        cudaError_t e = cudaMalloc(&p, grown);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&p, want);
            if (e != cudaSuccess) {
                cudaGetLastError();
                g_last_cuda_error = "cudaMalloc(" + std::to_string(want) + " bytes) failed";
                return TGX_ERR_NOMEM;
            }
            bytes = want;
            return TGX_OK;
        }
        bytes = grown;
        return TGX_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

struct PinBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int reserve(size_t want) {
        if (want <= bytes) return TGX_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
        if (cudaHostAlloc(&p, want, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            g_last_cuda_error = "cudaHostAlloc(" + std::to_string(want) + " bytes) failed";
            return TGX_ERR_NOMEM;
        }
        bytes = want;
        return TGX_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
    }
};

// int32 -> int64 widening input iterator for the cub scans (sums of counts can exceed 2^31).
struct WideIter {
    using iterator_category = std::random_access_iterator_tag;
    using value_type = int64_t;
    using difference_type = int64_t;
    using pointer = const int64_t*;
    using reference = int64_t;
    const int32_t* p;
    __host__ __device__ int64_t operator*() const { return (int64_t)*p; }
    __host__ __device__ int64_t operator[](difference_type i) const { return (int64_t)p[i]; }
    __host__ __device__ WideIter operator+(difference_type i) const { return WideIter{p + i}; }
    __host__ __device__ WideIter operator-(difference_type i) const { return WideIter{p - i}; }
    __host__ __device__ difference_type operator-(const WideIter& o) const { return p - o.p; }
    __host__ __device__ WideIter& operator+=(difference_type i) { p += i; return *this; }
    __host__ __device__ WideIter& operator++() { ++p; return *this; }
    __host__ __device__ WideIter operator++(int) { WideIter t = *this; ++p; return t; }
    __host__ __device__ bool operator==(const WideIter& o) const { return p == o.p; }
    __host__ __device__ bool operator!=(const WideIter& o) const { return p != o.p; }
};

}  // namespace

static_assert(sizeof(tgx_params) == 128, "tgx_params must be 128 bytes");
static_assert(sizeof(tgx_polyline_params) == 13 * sizeof(double), "tgx_polyline_params must fill the union");
static_assert(sizeof(tgx_polyline_legs) == 64, "tgx_polyline_legs must be 64 bytes");
static_assert(sizeof(tgx_goal_record) == 128, "tgx_goal_record must be 128 bytes");
static_assert(sizeof(tgx_transition_params) == 128, "tgx_transition_params must be 128 bytes");

struct tgx_engine {
    int device = 0;
    int64_t max_samples = (int64_t)1 << 24;
    int tile_shift = 10;   // 1024 samples per tile
    bool host_fill_constants = true;   // host-buffer calls: ship 10 planes over PCIe, memset the 4 constant ones
    bool host_plane_major = false;     // host-buffer calls: h_out is [14][n][capacity] instead of [n][14][capacity]
    bool exact_ramps = false;   // plan mode: replay ramps step by step (bit-identical state) or in exact-v jumps
    int spt = 4;           // samples per thread: 2 -> 128-bit stores, 4 -> 256-bit stores (measured best on B200)
    int64_t launches = 0;

    // slab-mode planning (single replay): slice sizes learned from the previous exact-offset plan
    bool allow_slabs = true;
    bool slabs_ready = false;
    bool ragged_ready = false;                   // slices are known but the batch is ragged: slab fill + tile compaction
    bool plan_dense_tiles = false;               // current plan: exact-offset addressing through tiles_dense
    int64_t ragged_plans = 0;
    int seg_slab = 0, tile_slab = 0;             // slice sizes the NEXT slab plan will use
    int seg_slab_plan = 0, tile_slab_plan = 0;   // slice sizes of the CURRENT plan (if plan_packed)
    int64_t slab_plans = 0, exact_plans = 0;

    // the same for braking plans (tgx_plan_stop), whose segment / tile counts have nothing to do with generateTraj's:
    // swapped in for the duration of such a plan (LearnSwap in plan_common)
    struct Learned {
        bool slabs_ready = false, ragged_ready = false, mixed_batch = false, phase_ready = false;
        int seg_slab = 0, tile_slab = 0;
    } stop_learn;

    // per-trajectory scratch (capacity in trajectories)
    DevBuf cnt, nseg, ntile, status, seg_off, tile_off, recs, maxv, maxa, cub_tmp, totals, cur_table, stats;
    // tables
    DevBuf segs, tiles, packets, phase, phase_ext, tiles_dense;
    DevBuf order;                       // replay order of a mixed batch: keys in/out, indices in/out (10 bytes each)
    bool mixed_batch = false;           // the last plan saw more than one replay class: sort the next one by class
    bool plan_packed = false;                    // current plan is a slab plan
    bool plan_phase = false;                     // current plan is a phase plan
    bool allow_phase = true;
    bool plane_tma = true;              // tgx_eval may send the planes through TMA (tgx_set_store_path)
    bool phase_ready = false;
    // Which kernels consumed the current plan.  A plan that only fed the reduction kernel (feasibility sweeps) makes the
    // next plan prefer segment tables over phase records: that kernel is issue-bound, and rebuilding the segments from
    // a phase record costs it 2 % more instructions than copying them from a table (ncu, 10^6 config-4 circles: 3.32 vs
    // 3.22 G warp instructions, 4.17 vs 4.08 ms), more than the phase plan saves in planning time.  The store kernels
    // are HBM-bound and gain from the record's single round of loads.  Speed only: the samples do not depend on it.
    bool plan_stored = false, plan_reduced = false, reduce_only = false;
    int64_t phase_plans = 0;
    PinBuf h_totals;

    // polyline-family plans (tgx_plan_polyline): leg records, and the work list of a ragged batch
    DevBuf poly_recs, poly_tiles, h_legs[2];
    bool plan_poly = false;                      // current plan is a polyline plan
    bool poly_listed = false;                    // ... addressed through poly_tiles rather than blockIdx / tile_slab
    int poly_tile_slab = 0;
    int64_t poly_plans = 0;

    // current plan
    bool has_plan = false;
    int64_t plan_n = 0, plan_tiles = 0, plan_segs = 0, plan_samples = 0;

    // pipelined device-resident generation (tgx_generate): a second engine so that one chunk is planned while the
    // previous one is evaluated, each on its own stream
    tgx_engine* twin = nullptr;
    cudaStream_t gs[2] = {nullptr, nullptr};       // planning streams (highest priority), one per engine
    cudaStream_t ge = nullptr;                     // evaluation stream (lowest priority)
    cudaEvent_t gev_plan[2] = {nullptr, nullptr};  // engine i's plan is complete
    cudaEvent_t gev[2] = {nullptr, nullptr};       // engine i's evaluation is complete: its tables may be rewritten
    cudaEvent_t gev_in = nullptr;
    int64_t generate_calls = 0, generate_chunks = 0;
    // tgx_set_generate_profiling: an event pair around every evaluation launch of tgx_generate*
    bool prof_on = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev;
    size_t prof_used = 0;

    // host-buffer path
    cudaStream_t hs[2] = {nullptr, nullptr};
    cudaEvent_t hev[2] = {nullptr, nullptr};       // slot's D2H copies done -> its staging buffers are reusable
    cudaEvent_t hev_eval = nullptr;                // last evaluation done -> the shared plan tables are reusable
    // host topology of the host-buffer path (resolved once per engine, host_topology())
    bool topo_ready = false;
    int numa_node = -1;                            // NUMA node of the GPU's PCIe slot (-1: unknown / a single-node VM)
    int cpus_allowed = 0;                          // CPUs this process may run on (sched_getaffinity)
    int local_ranks = 1;                           // processes sharing the host (LOCAL_WORLD_SIZE)
    int filler_threads = 1;                        // host threads that write the constant planes
    cpu_set_t filler_cpus;                         // where they run: the allowed CPUs of the GPU's NUMA node
    bool filler_cpus_valid = false;
    DevBuf h_params[2], h_out[2], h_cnt[2], h_st[2], h_ph[2], h_from[2], h_rec[2];
    DevBuf h_params_all, h_from_all;               // a whole call's parameters, uploaded before the first D2H copy is queued
    PinBuf p_small;                                // pinned landing area of a call's counts / status / phases / legs
    cudaEvent_t hev_up = nullptr;
};

namespace {

tgx::TableView table_view(const tgx_engine* e) {
    tgx::TableView tv{};
    tv.recs = e->recs.as<tgx::TrajRec>();
    tv.segs = e->segs.as<tgx::Seg>();
    tv.tiles = e->plan_dense_tiles ? e->tiles_dense.as<tgx::Tile>() : e->tiles.as<tgx::Tile>();
    if (e->plan_phase) {
        tv.phase = e->phase.as<tgx::PhaseRec>();
        tv.phase_ext = e->phase_ext.as<tgx::PhaseExt>();
        tv.tile_slab = e->tile_slab_plan;
    } else if (e->plan_packed) {
        tv.seg_slab = e->seg_slab_plan;
        tv.tile_slab = e->tile_slab_plan;
    }
    return tv;
}

int ensure_traj_scratch(tgx_engine* e, int64_t n) {
    int rc;
    const size_t n1 = (size_t)n + 1;
    if ((rc = e->cnt.reserve(n1 * sizeof(int32_t)))) return rc;
    if ((rc = e->nseg.reserve(n1 * sizeof(int32_t)))) return rc;
    if ((rc = e->ntile.reserve(n1 * sizeof(int32_t)))) return rc;
    if ((rc = e->status.reserve(n1 * sizeof(uint32_t)))) return rc;
    if ((rc = e->seg_off.reserve(n1 * sizeof(int64_t)))) return rc;
    if ((rc = e->tile_off.reserve(n1 * sizeof(int64_t)))) return rc;
    if ((rc = e->recs.reserve(n1 * sizeof(tgx::TrajRec)))) return rc;
    if ((rc = e->totals.reserve(4 * sizeof(int64_t)))) return rc;
    if ((rc = e->h_totals.reserve(16 * sizeof(int64_t)))) return rc;
    if ((rc = e->cur_table.reserve(tgx::cur_table_bytes()))) return rc;
    if ((rc = e->stats.reserve(sizeof(tgx::PlanStats)))) return rc;
    return TGX_OK;
}

// counts -> plan tables.  `stop_from` selects the braking plan.
int plan_common(tgx_engine* e, const tgx_params* d_params, const double* d_stop_from, int64_t n,
                const tgx_limits* limits, int32_t* d_counts, uint32_t* d_status, tgx_phases* d_phases,
                int64_t* total_samples, cudaStream_t stream) {
    if (!e || n < 0 || (n > 0 && !d_params)) return TGX_ERR_INVALID;
    if (n > 0x7fffffffLL) return TGX_ERR_INVALID;
    TGX_CUDA(cudaSetDevice(e->device));
    if (e->plan_stored || e->plan_reduced) e->reduce_only = !e->plan_stored;
    e->plan_stored = e->plan_reduced = false;
    e->has_plan = false;
    e->plan_poly = false;
    e->plan_n = e->plan_tiles = e->plan_segs = e->plan_samples = 0;
    if (total_samples) *total_samples = 0;
    if (n == 0) {
        e->has_plan = true;
        return TGX_OK;
    }
    int rc = ensure_traj_scratch(e, n);
    if (rc) return rc;

    // braking plans learn and use their own slice sizes
    struct LearnSwap {
        tgx_engine* e;
        bool on;
        void swap() {
            auto& L = e->stop_learn;
            std::swap(e->slabs_ready, L.slabs_ready);
            std::swap(e->ragged_ready, L.ragged_ready);
            std::swap(e->mixed_batch, L.mixed_batch);
            std::swap(e->phase_ready, L.phase_ready);
            std::swap(e->seg_slab, L.seg_slab);
            std::swap(e->tile_slab, L.tile_slab);
        }
        LearnSwap(tgx_engine* e_, bool on_) : e(e_), on(on_) { if (on) swap(); }
        ~LearnSwap() { if (on) swap(); }
    } learn_swap(e, d_stop_from != nullptr);

    int32_t* cnt = e->cnt.as<int32_t>();
    int32_t* nseg = e->nseg.as<int32_t>();
    int32_t* ntile = e->ntile.as<int32_t>();
    uint32_t* st = e->status.as<uint32_t>();
    int64_t* seg_off = e->seg_off.as<int64_t>();
    int64_t* tile_off = e->tile_off.as<int64_t>();
    int64_t* totals = e->totals.as<int64_t>();
    tgx::PlanStats* d_stats = e->stats.as<tgx::PlanStats>();
    tgx::PlanStats* h_stats = reinterpret_cast<tgx::PlanStats*>(static_cast<int64_t*>(e->h_totals.p) + 4);

    // the hold-length table of this batch's dt (one thread, a few microseconds)
    const void* tab = nullptr;
    if (!d_stop_from) {
        TGX_CUDA(tgx::launch_build_cur_table(d_params, e->max_samples, e->cur_table.p, stream));
        e->launches += 1;
        tab = e->cur_table.p;
    }

    int64_t tot_samples = 0, tot_segs = 0, tot_tiles = 0;
    bool done = false;
    // Every fill pass measures the largest per-trajectory segment / tile counts: they size the slices of the NEXT plan
    // and decide which single-replay path it can take (dense slices, ragged slices + tile compaction, phase records).
    auto relearn = [&](int64_t total_tiles) {
        // slices grow to the largest batch seen and shrink only when a batch needs less than half of them, so that
        // alternating batches (the chunks of one large job) do not overflow each other's slices
        int seg_slab = (h_stats->max_nseg + 4 + 3) / 4 * 4;
        int tile_slab = std::max(h_stats->max_ntile, 1);
        if (e->seg_slab > seg_slab && e->seg_slab <= 2 * seg_slab) seg_slab = e->seg_slab;
        if (e->tile_slab > tile_slab && e->tile_slab <= 2 * tile_slab) tile_slab = e->tile_slab;
        const bool dense = n * (int64_t)tile_slab <= total_tiles + total_tiles / 4 + 1;
        // fixed slices cost n * seg_slab segment records: up to 8 GB always, beyond that (10^7 .. 10^8-trajectory
        // feasibility sweeps) only while they fit half of the memory that is free right now
        const int64_t slab_bytes = n * (int64_t)seg_slab * (int64_t)sizeof(tgx::Seg);
        int64_t budget = std::max<int64_t>((int64_t)8 << 30, (int64_t)e->segs.bytes);   // what is held already is affordable
        if (slab_bytes > budget) {
            // (asked only when the table would have to grow: cudaMemGetInfo costs milliseconds on a busy context)
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
                budget = std::max<int64_t>(budget, (int64_t)((free_b + e->segs.bytes) / 2));
        }
        const bool small = slab_bytes <= budget &&
                           n * (int64_t)seg_slab <= 0x7fffffffLL && n * (int64_t)tile_slab <= 0x7fffffffLL;
        e->slabs_ready = dense && small && total_tiles > 0;
        e->ragged_ready = !dense && small && total_tiles > 0;
        e->seg_slab = seg_slab;
        e->tile_slab = tile_slab;
        // phase records: every trajectory of the batch can be written as a PhaseRec (the fill pass checked)
        // (dense or ragged: a phase plan has no tile slots, one CTA walks a whole trajectory)
        e->phase_ready = e->allow_phase && !e->exact_ramps && !h_stats->phase_misfit && h_stats->max_n > 0;
        e->mixed_batch = (h_stats->kinds & (h_stats->kinds - 1)) != 0;      // more than one replay class
    };
    e->plan_phase = false;
    e->plan_dense_tiles = false;

    // A mixed batch is replayed in the order of its replay classes (orbits by number of speed goals, lines,
    // boomerangs): neighbouring lanes then walk the same code instead of diverging at every branch.  One key
    // kernel + a one-pass radix sort of (class, index) pairs; the tables are indexed by trajectory, so the
    // plan itself does not depend on the order.
    const int32_t* order = nullptr;
    auto replay_order = [&]() -> int {
        if (order || !e->mixed_batch || n < 256) return TGX_OK;
        int rc2;
        if ((rc2 = e->order.reserve((size_t)n * 10 + 64))) return rc2;
        uint8_t* key_in = e->order.as<uint8_t>();
        uint8_t* key_out = key_in + n;
        int32_t* idx_in = reinterpret_cast<int32_t*>(key_in + ((2 * n + 15) & ~(int64_t)15));
        int32_t* idx_out = idx_in + n;
        TGX_CUDA(tgx::launch_replay_keys(d_params, n, key_in, idx_in, stream));
        size_t need = 0;
        TGX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, key_in, key_out, idx_in, idx_out, (int)n, 0, 5, stream));
        if ((rc2 = e->cub_tmp.reserve(need))) return rc2;
        size_t tmp_bytes = e->cub_tmp.bytes;
        TGX_CUDA(cub::DeviceRadixSort::SortPairs(e->cub_tmp.p, tmp_bytes, key_in, key_out, idx_in, idx_out, (int)n, 0, 5,
                                                 stream));
        e->launches += 2;
        order = idx_out;
        return TGX_OK;
    };

    // ---- phase mode: batches of short orbits and plain lines.  One replay that writes a self-contained 256-byte record
    //      per trajectory instead of tables; the evaluation kernel rebuilds the table path's segments from it, bit for
    //      bit, one CTA per trajectory ----
    if (e->allow_phase && e->phase_ready && !e->exact_ramps && !d_stop_from &&
        !(e->reduce_only && (e->slabs_ready || e->ragged_ready))) {
        if ((rc = e->phase.reserve((size_t)n * sizeof(tgx::PhaseRec)))) return rc;
        if ((rc = e->phase_ext.reserve((size_t)n * sizeof(tgx::PhaseExt)))) return rc;
        if ((rc = replay_order())) return rc;
        TGX_CUDA(zero_fill(d_stats, sizeof(tgx::PlanStats), stream));
        TGX_CUDA(tgx::launch_plan_phase(d_params, n, limits, e->max_samples, e->tile_shift, tgx::kPhaseMaxSamples, tab,
                                        e->phase.as<tgx::PhaseRec>(), e->phase_ext.as<tgx::PhaseExt>(), d_counts,
                                        d_status, cnt, st, d_phases, d_stats, stream, order));
        e->launches += 1;
        TGX_CUDA(peek_to_host(d_stats, h_stats, sizeof(tgx::PlanStats), stream));
        TGX_CUDA(cudaStreamSynchronize(stream));
        if (!h_stats->overflow) {
            tot_samples = (int64_t)h_stats->total_samples;
            tot_segs = 0;
            tot_tiles = n;                // one CTA per trajectory
            done = true;
            e->plan_phase = true;
            e->plan_packed = false;
            e->tile_slab_plan = 0;
            e->phase_plans += 1;
            e->mixed_batch = (h_stats->kinds & (h_stats->kinds - 1)) != 0;
        } else {
            e->phase_ready = false;   // boomerangs, many speed goals, long trajectories: plan with segment tables
        }
    }

    // ---- slab mode: ONE replay, no scans; falls through to the exact-offset path if a slice overflows ----------
    if (!done && e->allow_slabs && (e->slabs_ready || e->ragged_ready)) {
        const int64_t need_segs = n * (int64_t)e->seg_slab, need_tiles = n * (int64_t)e->tile_slab;
        if (need_segs <= 0x7fffffffLL && need_tiles <= 0x7fffffffLL) {
            if ((rc = e->segs.reserve((size_t)need_segs * sizeof(tgx::Seg)))) return rc;
            if ((rc = e->tiles.reserve((size_t)need_tiles * sizeof(tgx::Tile)))) return rc;
            TGX_CUDA(zero_fill(d_stats, sizeof(tgx::PlanStats), stream));
            if ((rc = replay_order())) return rc;
            TGX_CUDA(tgx::launch_plan_fill(d_params, d_stop_from, n, limits, e->max_samples, e->tile_shift, e->exact_ramps,
                                           tab, nullptr, nullptr, nullptr, e->seg_slab, e->tile_slab,
                                           e->recs.as<tgx::TrajRec>(), e->segs.as<tgx::Seg>(),
                                           e->tiles.as<tgx::Tile>(), d_counts, d_status, cnt, st, d_phases, d_stats,
                                           stream, order));
            e->launches += 1;
            TGX_CUDA(peek_to_host(d_stats, h_stats, sizeof(tgx::PlanStats), stream));
            TGX_CUDA(cudaStreamSynchronize(stream));
            const bool sparse = (int64_t)h_stats->total_tiles + (int64_t)h_stats->total_tiles / 4 + 1 < need_tiles;
            if (!h_stats->overflow && sparse) {
                // Ragged batch: most tile slots are empty.  Keep the single replay, but hand the evaluation kernel a
                // dense work list instead of one CTA per slot: count the used slots, scan, compact (three tiny kernels).
                const int64_t used = (int64_t)h_stats->total_tiles;
                if ((rc = e->tiles_dense.reserve((size_t)std::max<int64_t>(used, 1) * sizeof(tgx::Tile)))) return rc;
                TGX_CUDA(tgx::launch_count_used_tiles(n, e->tile_slab, e->tiles.as<tgx::Tile>(), ntile, stream));
                TGX_CUDA(zero_fill(ntile + n, sizeof(int32_t), stream));
                WideIter tile_in{ntile};
                size_t need = 0;
                TGX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, tile_in, tile_off, (int)(n + 1), stream));
                if ((rc = e->cub_tmp.reserve(need))) return rc;
                size_t tmp_bytes = e->cub_tmp.bytes;
                TGX_CUDA(cub::DeviceScan::ExclusiveSum(e->cub_tmp.p, tmp_bytes, tile_in, tile_off, (int)(n + 1), stream));
                TGX_CUDA(tgx::launch_compact_tiles(n, e->tile_slab, e->tiles.as<tgx::Tile>(), tile_off,
                                                   e->tiles_dense.as<tgx::Tile>(), stream));
                e->launches += 2;
                tot_samples = (int64_t)h_stats->total_samples;
                tot_segs = need_segs;
                tot_tiles = used;
                done = true;
                e->plan_packed = false;
                e->plan_dense_tiles = true;
                e->ragged_plans += 1;
                e->slab_plans += 1;
                relearn(used);
            } else if (!h_stats->overflow) {
                tot_samples = (int64_t)h_stats->total_samples;
                tot_segs = need_segs;
                tot_tiles = need_tiles;
                done = true;
                e->plan_packed = true;
                e->seg_slab_plan = e->seg_slab;
                e->tile_slab_plan = e->tile_slab;
                e->slab_plans += 1;
                relearn((int64_t)h_stats->total_tiles);
            } else {
                e->slabs_ready = false;   // re-learn the slice sizes below
                e->ragged_ready = false;
            }
        }
    }

    if (!done) {
        e->plan_packed = false;
        // ---- exact-offset mode, pass 1: counts ------------------------------------------------------------------
        TGX_CUDA(tgx::launch_plan_count(d_params, d_stop_from, n, limits, e->max_samples, e->tile_shift,
                                        e->exact_ramps, tab, cnt, st, nseg, ntile, stream));
        e->launches += 1;
        // trailing zero so that an exclusive scan over n+1 items leaves the grand total in element n
        TGX_CUDA(zero_fill(nseg + n, sizeof(int32_t), stream));
        TGX_CUDA(zero_fill(ntile + n, sizeof(int32_t), stream));

        // pass 2: exclusive scans (segment and tile offsets) and the sample total
        WideIter seg_in{nseg}, tile_in{ntile}, cnt_in{cnt};
        size_t need = 0, t1 = 0, t2 = 0;
        TGX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t1, seg_in, seg_off, (int)(n + 1), stream));
        TGX_CUDA(cub::DeviceReduce::Sum(nullptr, t2, cnt_in, totals, (int)n, stream));
        need = std::max(t1, t2);
        if ((rc = e->cub_tmp.reserve(need))) return rc;
        size_t tmp_bytes = e->cub_tmp.bytes;
        TGX_CUDA(cub::DeviceScan::ExclusiveSum(e->cub_tmp.p, tmp_bytes, seg_in, seg_off, (int)(n + 1), stream));
        tmp_bytes = e->cub_tmp.bytes;
        TGX_CUDA(cub::DeviceScan::ExclusiveSum(e->cub_tmp.p, tmp_bytes, tile_in, tile_off, (int)(n + 1), stream));
        tmp_bytes = e->cub_tmp.bytes;
        TGX_CUDA(cub::DeviceReduce::Sum(e->cub_tmp.p, tmp_bytes, cnt_in, totals, (int)n, stream));
        // (the cub scan / reduce launches are library plumbing and are not counted in tgx_launch_count)

        int64_t* h = static_cast<int64_t*>(e->h_totals.p);
        TGX_CUDA(peek_to_host(totals, h + 0, sizeof(int64_t), stream));
        TGX_CUDA(peek_to_host(seg_off + n, h + 1, sizeof(int64_t), stream));
        TGX_CUDA(peek_to_host(tile_off + n, h + 2, sizeof(int64_t), stream));
        TGX_CUDA(cudaStreamSynchronize(stream));
        tot_samples = h[0];
        tot_segs = h[1];
        tot_tiles = h[2];
        if (tot_segs > 0x7fffffffLL || tot_tiles > 0x7fffffffLL) return TGX_ERR_CAPACITY;

        if ((rc = e->segs.reserve((size_t)std::max<int64_t>(tot_segs, 1) * sizeof(tgx::Seg)))) return rc;
        if ((rc = e->tiles.reserve((size_t)std::max<int64_t>(tot_tiles, 1) * sizeof(tgx::Tile)))) return rc;

        // pass 3: fill (also measures the per-trajectory maxima that size the slabs of the next plan)
        TGX_CUDA(zero_fill(d_stats, sizeof(tgx::PlanStats), stream));
        TGX_CUDA(tgx::launch_plan_fill(d_params, d_stop_from, n, limits, e->max_samples, e->tile_shift, e->exact_ramps,
                                       tab, cnt, seg_off, tile_off, 0, 0, e->recs.as<tgx::TrajRec>(),
                                       e->segs.as<tgx::Seg>(), e->tiles.as<tgx::Tile>(), d_counts, d_status, nullptr,
                                       nullptr, d_phases, d_stats, stream));
        e->launches += 1;
        e->exact_plans += 1;
        if (e->allow_slabs) {
            TGX_CUDA(peek_to_host(d_stats, h_stats, sizeof(tgx::PlanStats), stream));
            TGX_CUDA(cudaStreamSynchronize(stream));
            relearn(tot_tiles);
        }
    }

    e->has_plan = true;
    e->plan_n = n;
    e->plan_tiles = tot_tiles;
    e->plan_segs = tot_segs;
    e->plan_samples = tot_samples;
    if (total_samples) *total_samples = tot_samples;
    return TGX_OK;
}

// Polyline-family planning: one replay-free pass (leg records, counts, tile counts), then either slab addressing
// (dense batches: tile t of trajectory i is CTA i*slab + t) or a scanned work list (ragged batches).
int plan_polyline_common(tgx_engine* e, const tgx_params* d_params, int64_t n, const tgx_limits* limits,
                         int32_t* d_counts, uint32_t* d_status, tgx_polyline_legs* d_legs, int64_t* total_samples,
                         bool skip_foreign, cudaStream_t stream) {
    if (!e || n < 0 || (n > 0 && !d_params)) return TGX_ERR_INVALID;
    if (n > 0x7fffffffLL) return TGX_ERR_INVALID;
    TGX_CUDA(cudaSetDevice(e->device));
    e->has_plan = false;
    e->plan_n = e->plan_tiles = e->plan_segs = e->plan_samples = 0;
    if (total_samples) *total_samples = 0;
    e->plan_poly = true;
    e->plan_packed = e->plan_phase = false;
    if (n == 0) {
        e->has_plan = true;
        return TGX_OK;
    }
    int rc = ensure_traj_scratch(e, n);
    if (rc) return rc;
    if ((rc = e->poly_recs.reserve((size_t)n * tgx::poly_rec_bytes()))) return rc;
    int32_t* cnt = e->cnt.as<int32_t>();
    int32_t* ntile = e->ntile.as<int32_t>();
    uint32_t* st = e->status.as<uint32_t>();
    tgx::PlanStats* d_stats = e->stats.as<tgx::PlanStats>();
    tgx::PlanStats* h_stats = reinterpret_cast<tgx::PlanStats*>(static_cast<int64_t*>(e->h_totals.p) + 4);

    TGX_CUDA(tgx::launch_build_cur_table(d_params, e->max_samples, e->cur_table.p, stream));
    TGX_CUDA(zero_fill(d_stats, sizeof(tgx::PlanStats), stream));
    TGX_CUDA(tgx::launch_plan_poly(d_params, n, limits, e->max_samples, e->tile_shift, e->cur_table.p, e->poly_recs.p,
                                   d_counts, d_status, cnt, st, d_legs, ntile, d_stats, skip_foreign, stream));
    e->launches += 2;
    TGX_CUDA(peek_to_host(d_stats, h_stats, sizeof(tgx::PlanStats), stream));
    TGX_CUDA(cudaStreamSynchronize(stream));
    const int64_t tot_tiles = (int64_t)h_stats->total_tiles;
    const int64_t slab_tiles = n * (int64_t)h_stats->max_ntile;
    int64_t plan_tiles;
    if (slab_tiles <= tot_tiles + tot_tiles / 4 + 1 && slab_tiles <= 0x7fffffffLL) {
        e->poly_listed = false;
        e->poly_tile_slab = std::max(h_stats->max_ntile, 1);
        plan_tiles = tot_tiles > 0 ? slab_tiles : 0;
    } else {
        if (tot_tiles > 0x7fffffffLL) return TGX_ERR_CAPACITY;
        int64_t* tile_off = e->tile_off.as<int64_t>();
        TGX_CUDA(zero_fill(ntile + n, sizeof(int32_t), stream));
        WideIter tile_in{ntile};
        size_t need = 0;
        TGX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, tile_in, tile_off, (int)(n + 1), stream));
        if ((rc = e->cub_tmp.reserve(need))) return rc;
        size_t tmp_bytes = e->cub_tmp.bytes;
        TGX_CUDA(cub::DeviceScan::ExclusiveSum(e->cub_tmp.p, tmp_bytes, tile_in, tile_off, (int)(n + 1), stream));
        if ((rc = e->poly_tiles.reserve((size_t)std::max<int64_t>(tot_tiles, 1) * sizeof(tgx::Tile)))) return rc;
        TGX_CUDA(tgx::launch_poly_tiles(n, ntile, tile_off, e->tile_shift, e->poly_tiles.as<tgx::Tile>(), stream));
        e->launches += 1;
        e->poly_listed = true;
        e->poly_tile_slab = 0;
        plan_tiles = tot_tiles;
    }
    e->poly_plans += 1;
    e->has_plan = true;
    e->plan_n = n;
    e->plan_tiles = plan_tiles;
    e->plan_segs = 0;
    e->plan_samples = (int64_t)h_stats->total_samples;
    if (total_samples) *total_samples = e->plan_samples;
    return TGX_OK;
}

tgx::PolyView poly_view(const tgx_engine* e) {
    tgx::PolyView pv{};
    pv.recs = e->poly_recs.as<int4>();
    pv.tiles = e->poly_listed ? e->poly_tiles.as<tgx::Tile>() : nullptr;
    pv.tile_slab = e->poly_tile_slab;
    return pv;
}

// Evaluation of the current plan, whichever family planned it.
cudaError_t launch_current(tgx_engine* e, const tgx::OutView& out, bool store, double* max_v, double* max_a,
                           cudaStream_t s, const tgx::RecOut* ptma = nullptr) {
    (store ? e->plan_stored : e->plan_reduced) = true;         // what the plan is used for steers how the next one is made
    if (e->plan_poly)
        return tgx::launch_eval_poly(poly_view(e), e->plan_tiles, e->tile_shift, e->spt, out, store, max_v, max_a, s);
    return tgx::launch_eval(table_view(e), e->plan_tiles, e->tile_shift, e->spt, out, store, max_v, max_a, s, ptma);
}

// cuTensorMapEncodeTiled lives in libcuda; it is looked up through the runtime so libtgx.so needs no -lcuda.
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
tmap_encode_fn tmap_encoder() {
    static tmap_encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && fn &&
            q == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<tmap_encode_fn>(fn);
    }
    return encode;
}

// The caller's planes as TMA sees them: a 3-D fp64 tensor [trajectory][channel][sample] with the layout's strides,
// written in boxes of {32 samples, 14 channels, 1 trajectory} (PlaneTma in store.cuh).  Returns false when the layout
// does not qualify (per-trajectory offsets, a channel subset, rows shorter than a box, strides TMA cannot express): the
// vector-store kernel handles those.
bool make_plane_tmap(CUtensorMap* tmap, const tgx_layout* out, int64_t n) {
    tmap_encode_fn encode = tmap_encoder();
    if (!encode || out->d_traj_offset || n <= 0 || out->capacity < 32) return false;
    if (out->channel_mask && (out->channel_mask & 0x3fffu) != 0x3fffu) return false;
    if (out->traj_stride <= 0 || out->chan_stride <= 0) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)out->capacity, TGX_NCHAN, (cuuint64_t)n};
    const cuuint64_t strides[2] = {(cuuint64_t)out->chan_stride * 8, (cuuint64_t)out->traj_stride * 8};
    if (dims[0] > 0xffffffffull || dims[2] > 0x7fffffffull || strides[0] >= (1ull << 40) || strides[1] >= (1ull << 40))
        return false;
    const cuuint32_t box[3] = {32, TGX_NCHAN, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, out->d_base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int check_layout(const tgx_layout* out, int spt) {
    if (!out || !out->d_base || out->capacity < 0) return TGX_ERR_INVALID;
    const int64_t a = 4;   // doubles; keeps both store widths legal and rows sector-aligned
    (void)spt;
    if ((reinterpret_cast<uintptr_t>(out->d_base) & 31u) != 0) return TGX_ERR_ALIGNMENT;
    if (out->chan_stride % a != 0) return TGX_ERR_ALIGNMENT;
    if (!out->d_traj_offset && out->traj_stride % a != 0) return TGX_ERR_ALIGNMENT;
    return TGX_OK;
}

tgx::OutView make_view(const tgx_layout* out) {
    tgx::OutView v;
    v.base = out->d_base;
    v.traj_stride = out->traj_stride;
    v.chan_stride = out->chan_stride;
    v.traj_offset = out->d_traj_offset;
    v.capacity = out->capacity;
    v.channel_mask = out->channel_mask ? out->channel_mask : 0x3fffu;
    return v;
}

}  // namespace

// ---- host topology of the host-buffer path ------------------------------------------------------------------
// The end-to-end rate of the host-buffer calls is set by PCIe and by what the host's memory absorbs, so the host side
// is sized to the share of the machine this process owns: the CPUs it may run on (sched_getaffinity) divided by the
// processes that share the node (LOCAL_WORLD_SIZE, one per GPU under torchrun), and — where the platform exposes it —
// the NUMA node the GPU's PCIe slot hangs off, for the filler threads and for the pinned allocations.
static bool parse_cpulist(const char* text, cpu_set_t* set) {
    CPU_ZERO(set);
    bool any = false;
    const char* p = text;
    while (*p) {
        char* end = nullptr;
        long a = std::strtol(p, &end, 10);
        if (end == p) break;
        long b = a;
        p = end;
        if (*p == '-') {
            b = std::strtol(p + 1, &end, 10);
            if (end == p + 1) break;
            p = end;
        }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c)
            if (c >= 0) { CPU_SET((int)c, set); any = true; }
        while (*p == ',' || *p == ' ' || *p == '\n') ++p;
    }
    return any;
}

static bool read_small_file(const std::string& path, char* buf, size_t cap) {
    FILE* f = std::fopen(path.c_str(), "r");
    if (!f) return false;
    const size_t got = std::fread(buf, 1, cap - 1, f);
    std::fclose(f);
    buf[got] = 0;
    return got > 0;
}

static void host_topology(tgx_engine* e) {
    if (e->topo_ready) return;
    e->topo_ready = true;
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    int ncpu = 0;
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) ncpu = CPU_COUNT(&allowed);
    if (ncpu <= 0) {
        ncpu = (int)std::max(1u, std::thread::hardware_concurrency());
        for (int c = 0; c < ncpu && c < CPU_SETSIZE; ++c) CPU_SET(c, &allowed);
    }
    e->cpus_allowed = ncpu;
    int ranks = 1;
    for (const char* name : {"TGX_LOCAL_RANKS", "LOCAL_WORLD_SIZE"}) {
        const char* v = std::getenv(name);
        if (v && std::atoi(v) > 0) { ranks = std::atoi(v); break; }
    }
    e->local_ranks = ranks;
    // NUMA node of the GPU (sysfs; -1 on single-node hosts and in most VMs)
    char bus[32] = {0}, buf[4096];
    e->numa_node = -1;
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), e->device) == cudaSuccess) {
        for (char* c = bus; *c; ++c) *c = (char)std::tolower((unsigned char)*c);
        if (read_small_file(std::string("/sys/bus/pci/devices/") + bus + "/numa_node", buf, sizeof(buf)))
            e->numa_node = std::atoi(buf);
    } else {
        cudaGetLastError();
    }
    cpu_set_t local = allowed;
    if (e->numa_node >= 0 &&
        read_small_file("/sys/devices/system/node/node" + std::to_string(e->numa_node) + "/cpulist", buf, sizeof(buf))) {
        cpu_set_t node;
        if (parse_cpulist(buf, &node)) {
            cpu_set_t both;
            CPU_AND(&both, &node, &allowed);
            if (CPU_COUNT(&both) > 0) local = both;
        }
    }
    e->filler_cpus = local;
    e->filler_cpus_valid = true;
    // half of this process's share of the CPUs it may use, at most 8 (a memset stream saturates well before that)
    const int share = std::max(1, std::min(ncpu, CPU_COUNT(&local) * std::max(1, ranks)) / std::max(1, ranks));
    e->filler_threads = std::max(1, std::min(8, share / 2));
    if (const char* v = std::getenv("TGX_FILLER_THREADS"))
        if (std::atoi(v) > 0) e->filler_threads = std::min(64, std::atoi(v));
}

// MPOL_PREFERRED for the calling thread while `fn` allocates (no libnuma in the image: the raw system call)
template <class F>
static void with_preferred_node(int node, F fn) {
#ifdef SYS_set_mempolicy
    constexpr int kMpolDefault = 0, kMpolPreferred = 1;
    bool set = false;
    if (node >= 0 && node < 1024) {
        unsigned long mask[16] = {0};
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        set = syscall(SYS_set_mempolicy, kMpolPreferred, mask, sizeof(mask) * 8 + 1) == 0;
    }
    fn();
    if (set) syscall(SYS_set_mempolicy, kMpolDefault, nullptr, 0);
#else
    (void)node;
    fn();
#endif
}

extern "C" {

int tgx_version(void) { return TGX_VERSION; }

const char* tgx_strerror(int code) {
    switch (code) {
        case TGX_OK: return "ok";
        case TGX_ERR_INVALID: return "invalid argument";
        case TGX_ERR_CUDA: return "CUDA runtime error (see tgx_last_cuda_error)";
        case TGX_ERR_ALIGNMENT: return "output layout is not aligned for vector stores (32-byte base, strides multiple of 4 doubles)";
        case TGX_ERR_NO_PLAN: return "no current plan: call tgx_plan first";
        case TGX_ERR_NOMEM: return "out of memory";
        case TGX_ERR_CAPACITY: return "capacity exceeded";
        case TGX_ERR_COMM: return "NCCL error (see tgx_comm_last_error)";
        default: return "unknown error";
    }
}

const char* tgx_last_cuda_error(void) { return g_last_cuda_error.c_str(); }

int tgx_create(tgx_engine** out, int device) {
    if (!out) return TGX_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    cudaError_t err = cudaGetDeviceCount(&ndev);
    if (err != cudaSuccess) return cuda_fail(err, "cudaGetDeviceCount");
    if (ndev <= 0 || device < 0 || device >= ndev) {
        g_last_cuda_error = "no CUDA device " + std::to_string(device) + " (device count " + std::to_string(ndev) + ")";
        return TGX_ERR_CUDA;
    }
    TGX_CUDA(cudaSetDevice(device));
    tgx_engine* e = new (std::nothrow) tgx_engine();
    if (!e) return TGX_ERR_NOMEM;
    e->device = device;
    *out = e;
    return TGX_OK;
}

int tgx_destroy(tgx_engine* e) {
    if (!e) return TGX_OK;
    cudaSetDevice(e->device);
    if (e->twin) tgx_destroy(e->twin);
    e->twin = nullptr;
    for (int i = 0; i < 2; ++i) {
        if (e->gs[i]) cudaStreamDestroy(e->gs[i]);
        if (e->gev[i]) cudaEventDestroy(e->gev[i]);
        if (e->gev_plan[i]) cudaEventDestroy(e->gev_plan[i]);
    }
    if (e->ge) cudaStreamDestroy(e->ge);
    if (e->gev_in) cudaEventDestroy(e->gev_in);
    for (auto& pr : e->prof_ev) {
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    e->prof_ev.clear();
    DevBuf* bufs[] = {&e->cnt, &e->nseg, &e->ntile, &e->status, &e->seg_off, &e->tile_off, &e->recs, &e->maxv,
                      &e->maxa, &e->cub_tmp, &e->totals, &e->segs, &e->tiles, &e->cur_table, &e->stats, &e->packets,
                      &e->phase, &e->phase_ext, &e->poly_recs, &e->poly_tiles, &e->tiles_dense, &e->order};
    for (DevBuf* b : bufs) b->release();
    for (int i = 0; i < 2; ++i) {
        e->h_legs[i].release();
        e->h_rec[i].release();
        e->h_params[i].release();
        e->h_out[i].release();
        e->h_cnt[i].release();
        e->h_st[i].release();
        e->h_ph[i].release();
        e->h_from[i].release();
        if (e->hs[i]) cudaStreamDestroy(e->hs[i]);
        if (e->hev[i]) cudaEventDestroy(e->hev[i]);
    }
    if (e->hev_eval) cudaEventDestroy(e->hev_eval);
    if (e->hev_up) cudaEventDestroy(e->hev_up);
    e->h_params_all.release();
    e->h_from_all.release();
    e->p_small.release();
    e->h_totals.release();
    delete e;
    return TGX_OK;
}

int tgx_set_max_samples(tgx_engine* e, int64_t max_samples) {
    if (!e || max_samples < 1 || max_samples > 0x7ffffff0LL) return TGX_ERR_INVALID;
    e->max_samples = max_samples;
    return TGX_OK;
}

// Tuning knob used by bench.py sweeps: tile = 1 << tile_shift samples per CTA, spt samples per thread.
int tgx_set_tuning(tgx_engine* e, int tile_shift, int spt) {
    if (!e) return TGX_ERR_INVALID;
    if (spt != 2 && spt != 4) return TGX_ERR_INVALID;
    if (tile_shift < 9 || tile_shift > 10) return TGX_ERR_INVALID;
    const int threads = (1 << tile_shift) / spt;
    if (threads != 128 && threads != 256) return TGX_ERR_INVALID;
    e->tile_shift = tile_shift;
    e->spt = spt;
    e->has_plan = false;   // tile size is baked into a plan
    e->slabs_ready = false;
    e->ragged_ready = false;
    e->phase_ready = false;
    e->stop_learn = tgx_engine::Learned();
    return TGX_OK;
}

int tgx_set_plan_mode(tgx_engine* e, int exact_ramps) {
    if (!e) return TGX_ERR_INVALID;
    e->exact_ramps = exact_ramps != 0;
    e->has_plan = false;
    e->slabs_ready = false;   // segment counts differ between the modes (ramp chunks)
    e->ragged_ready = false;
    e->phase_ready = false;
    e->stop_learn = tgx_engine::Learned();
    return TGX_OK;
}

// Host-buffer calls: 1 (default) = evaluate and ship only the 10 varying planes and write the 4 constant planes
// (p.z = alt, v.z = a.z = j.z = 0) with host threads; 0 = evaluate and ship all 14 planes.
int tgx_set_host_fill(tgx_engine* e, int fill_constants_on_host) {
    if (!e) return TGX_ERR_INVALID;
    e->host_fill_constants = fill_constants_on_host != 0;
    return TGX_OK;
}

// Single-replay planning with per-trajectory slices sized from the previous plan (default on).  allow = 0 forces the
// two-replay exact-offset path for every plan.
// Host-buffer calls: 0 (default) = h_out[(i*14 + c)*capacity + k]; 1 = plane-major h_out[(c*n + i)*capacity + k].
int tgx_set_host_layout(tgx_engine* e, int plane_major) {
    if (!e) return TGX_ERR_INVALID;
    e->host_plane_major = plane_major != 0;
    return TGX_OK;
}

int tgx_set_slab_planning(tgx_engine* e, int allow) {
    if (!e) return TGX_ERR_INVALID;
    e->allow_slabs = allow != 0;
    e->slabs_ready = false;
    e->ragged_ready = false;
    e->has_plan = false;
    e->stop_learn = tgx_engine::Learned();
    return TGX_OK;
}

// Phase planning (default on): batches of Circle / Figure8 trajectories of at most 4096 samples are planned by a
// counting replay only and evaluated from the caller's parameter array (which must then stay valid and unchanged
// until the last tgx_eval / tgx_feasibility of that plan).  allow = 0 always plans with segment tables.
int tgx_set_phase_planning(tgx_engine* e, int allow) {
    if (!e) return TGX_ERR_INVALID;
    e->allow_phase = allow != 0;
    e->phase_ready = false;
    e->has_plan = false;
    return TGX_OK;
}

int64_t tgx_phase_plan_count(const tgx_engine* e) { return e ? e->phase_plans : 0; }

// Store path of tgx_eval: tma = 1 (default) sends the planes of a qualifying layout through TMA, 0 always uses the
// vector-store kernel.  Results are bit-identical.
int tgx_set_store_path(tgx_engine* e, int tma) {
    if (!e) return TGX_ERR_INVALID;
    e->plane_tma = tma != 0;
    return TGX_OK;
}

// How many plans so far took the single-replay / the two-replay path.
int tgx_plan_path_counts(const tgx_engine* e, int64_t* slab_plans, int64_t* exact_plans) {
    if (!e) return TGX_ERR_INVALID;
    if (slab_plans) *slab_plans = e->slab_plans;
    if (exact_plans) *exact_plans = e->exact_plans;
    return TGX_OK;
}

int64_t tgx_scratch_bytes(const tgx_engine* e) {
    if (!e) return 0;
    const DevBuf* bufs[] = {&e->cnt, &e->nseg, &e->ntile, &e->status, &e->seg_off, &e->tile_off, &e->recs, &e->maxv,
                            &e->maxa, &e->cub_tmp, &e->totals, &e->segs, &e->tiles, &e->cur_table, &e->stats,
                            &e->packets, &e->phase, &e->phase_ext, &e->poly_recs, &e->poly_tiles, &e->tiles_dense, &e->order};
    int64_t s = 0;
    for (const DevBuf* b : bufs) s += (int64_t)b->bytes;
    return s;
}

int64_t tgx_launch_count(const tgx_engine* e) { return e ? e->launches + (e->twin ? e->twin->launches : 0) : 0; }
int64_t tgx_plan_tiles(const tgx_engine* e) { return e && e->has_plan ? e->plan_tiles : 0; }
int64_t tgx_plan_segments(const tgx_engine* e) { return e && e->has_plan ? e->plan_segs : 0; }

int tgx_count(tgx_engine* e, const tgx_params* d_params, int64_t n, const tgx_limits* limits, int32_t* d_counts,
              uint32_t* d_status, void* stream) {
    if (!e || n < 0 || (n > 0 && !d_params)) return TGX_ERR_INVALID;
    if (n == 0) return TGX_OK;
    TGX_CUDA(cudaSetDevice(e->device));
    int rc = e->cur_table.reserve(tgx::cur_table_bytes());
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TGX_CUDA(tgx::launch_build_cur_table(d_params, e->max_samples, e->cur_table.p, s));
    TGX_CUDA(tgx::launch_plan_count(d_params, nullptr, n, limits, e->max_samples, e->tile_shift, false,
                                    e->cur_table.p, d_counts, d_status, nullptr, nullptr, s));
    // polyline-family trajectories (which the pass above marked WRONG_PLANNER) are counted by their own planner
    TGX_CUDA(tgx::launch_plan_poly(d_params, n, limits, e->max_samples, e->tile_shift, e->cur_table.p, nullptr,
                                   d_counts, d_status, nullptr, nullptr, nullptr, nullptr, nullptr, true, s));
    e->launches += 3;
    return TGX_OK;
}

int tgx_plan(tgx_engine* e, const tgx_params* d_params, int64_t n, const tgx_limits* limits, int32_t* d_counts,
             uint32_t* d_status, tgx_phases* d_phases, int64_t* total_samples, void* stream) {
    return plan_common(e, d_params, nullptr, n, limits, d_counts, d_status, d_phases, total_samples,
                       static_cast<cudaStream_t>(stream));
}

int tgx_plan_polyline(tgx_engine* e, const tgx_params* d_params, int64_t n, const tgx_limits* limits,
                      int32_t* d_counts, uint32_t* d_status, tgx_polyline_legs* d_legs, int64_t* total_samples,
                      void* stream) {
    return plan_polyline_common(e, d_params, n, limits, d_counts, d_status, d_legs, total_samples, false,
                                static_cast<cudaStream_t>(stream));
}

// cos / sin of `orientation` from the HOST libm, the one a reference build on this machine links (tgx.h:
// tgx_polyline_params).  Parameter preparation only: two libm calls per trajectory, no sampling.
int tgx_polyline_finalize_host(tgx_params* h_params, int64_t n) {
    if (n < 0 || (n > 0 && !h_params)) return TGX_ERR_INVALID;
    for (int64_t i = 0; i < n; ++i) {
        tgx_params& p = h_params[i];
        if (!TGX_IS_POLYLINE(p.type)) continue;
        p.u.poly.cos_o = std::cos(p.u.poly.orientation);
        p.u.poly.sin_o = std::sin(p.u.poly.orientation);
        p.n_vgoals |= TGX_POLY_TRIG_GIVEN;
    }
    return TGX_OK;
}

int tgx_plan_stop(tgx_engine* e, const tgx_params* d_params, int64_t n, const double* d_from, int32_t* d_counts,
                  uint32_t* d_status, tgx_phases* d_phases, int64_t* total_samples, void* stream) {
    if (n > 0 && !d_from) return TGX_ERR_INVALID;
    return plan_common(e, d_params, d_from, n, nullptr, d_counts, d_status, d_phases, total_samples,
                       static_cast<cudaStream_t>(stream));
}

int tgx_plan_samples(tgx_engine* e, const tgx_params* d_params, int64_t n, const double* d_state, int32_t* d_counts,
                     uint32_t* d_status, void* stream) {
    if (!e || n < 0 || (n > 0 && (!d_params || !d_state))) return TGX_ERR_INVALID;
    if (n > 0x7fffffffLL) return TGX_ERR_INVALID;
    TGX_CUDA(cudaSetDevice(e->device));
    e->has_plan = false;
    int rc = ensure_traj_scratch(e, n);
    if (rc) return rc;
    if ((rc = e->segs.reserve((size_t)std::max<int64_t>(n, 1) * sizeof(tgx::Seg)))) return rc;
    if ((rc = e->tiles.reserve((size_t)std::max<int64_t>(n, 1) * sizeof(tgx::Tile)))) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // the engine's own status copy feeds tgx_feasibility; keep it consistent
    TGX_CUDA(tgx::launch_plan_samples(d_params, d_state, n, e->recs.as<tgx::TrajRec>(), e->segs.as<tgx::Seg>(),
                                      e->tiles.as<tgx::Tile>(), e->cnt.as<int32_t>(), e->status.as<uint32_t>(), s));
    e->launches += 1;
    if (d_counts) TGX_CUDA(cudaMemcpyAsync(d_counts, e->cnt.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    if (d_status) TGX_CUDA(cudaMemcpyAsync(d_status, e->status.p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    e->has_plan = true;
    e->plan_packed = false;
    e->plan_phase = false;
    e->plan_poly = false;
    e->plan_n = n;
    e->plan_tiles = n;
    e->plan_segs = n;
    e->plan_samples = n;
    return TGX_OK;
}

int tgx_eval(tgx_engine* e, const tgx_layout* out, double* d_max_v, double* d_max_a, void* stream) {
    if (!e) return TGX_ERR_INVALID;
    if (!e->has_plan) return TGX_ERR_NO_PLAN;
    int rc = check_layout(out, e->spt);
    if (rc) return rc;
    TGX_CUDA(cudaSetDevice(e->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (e->plan_n == 0) return TGX_OK;
    if (d_max_v) TGX_CUDA(zero_fill(d_max_v, (size_t)e->plan_n * sizeof(double), s));
    if (d_max_a) TGX_CUDA(zero_fill(d_max_a, (size_t)e->plan_n * sizeof(double), s));
    if (e->plan_tiles == 0) return TGX_OK;
    tgx::RecOut ptma{};
    // (the polyline kernel has no per-sample transcendental and is faster with vector stores: 17.0 vs 17.7 ms per Mi
    //  T trajectories, so its planes never take the TMA path)
    const bool tma = e->plane_tma && !e->plan_poly && !d_max_v && !d_max_a && make_plane_tmap(&ptma.tmap, out, e->plan_n);
    TGX_CUDA(launch_current(e, make_view(out), true, d_max_v, d_max_a, s, tma ? &ptma : nullptr));
    e->launches += 1;
    return TGX_OK;
}

int tgx_feasibility(tgx_engine* e, const tgx_limits* limits, uint8_t* d_flags, double* d_max_v, double* d_max_a,
                    uint32_t* d_status, void* stream) {
    if (!e || !limits) return TGX_ERR_INVALID;
    if (!e->has_plan) return TGX_ERR_NO_PLAN;
    TGX_CUDA(cudaSetDevice(e->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t n = e->plan_n;
    if (n == 0) return TGX_OK;
    int rc;
    if (!d_max_v) {
        if ((rc = e->maxv.reserve((size_t)n * sizeof(double)))) return rc;
        d_max_v = e->maxv.as<double>();
    }
    if (!d_max_a) {
        if ((rc = e->maxa.reserve((size_t)n * sizeof(double)))) return rc;
        d_max_a = e->maxa.as<double>();
    }
    TGX_CUDA(zero_fill(d_max_v, (size_t)n * sizeof(double), s));
    TGX_CUDA(zero_fill(d_max_a, (size_t)n * sizeof(double), s));
    if (e->plan_tiles > 0) {
        tgx::OutView none{};
        TGX_CUDA(launch_current(e, none, false, d_max_v, d_max_a, s));
        e->launches += 1;
    }
    TGX_CUDA(tgx::launch_feasibility_finalize(n, e->status.as<uint32_t>(), d_max_v, d_max_a, limits->v_max,
                                              limits->a_max, d_flags, d_status, s));
    e->launches += 1;
    return TGX_OK;
}

// ---- pipelined generation of a device-resident batch -----------------------------------------------------------
// tgx_plan and tgx_eval of one batch are serial by construction (the evaluation needs the plan), and planning is a
// latency-bound replay that leaves the memory system idle while the evaluation is a store stream that leaves most issue
// slots idle.  tgx_generate cuts the batch into chunks and alternates them between the engine and a private twin:
// chunk c+1 is planned while chunk c is evaluated, so all planning but the first chunk's disappears behind the
// store-bound kernel.
//   * Evaluations run in order on ONE low-priority stream; each engine plans on its own HIGH-priority stream.  The
//     evaluation kernel fills every SM (6 CTAs x 80 registers), so without priorities the planner's CTAs of the other
//     stream are dispatched only when the evaluation grid is exhausted — no overlap at all (measured: 19.3 instead of
//     18.0 ms per Mi circles); with them every retiring evaluation CTA makes room for a waiting planner CTA first.
//   * Events order the rest: the evaluation of chunk c waits for its plan, the plan of chunk c+2 (same engine, same
//     tables) waits for the evaluation of chunk c.
// Measured on B200: the planning does disappear from the critical path (0.4 ms of 1.2 ms left per Mi
// circles), but the evaluation kernels that ran next to a planner take longer by what was hidden — 1 Mi circles 18.0 ms
// either way, the mixed batch 19.9 vs 18.2 ms, the 10^7-trajectory sweep 63.5 vs 63.1 ms — for every chunk size, CTA
// shape and residency of the planner tried.  The call is kept for callers that want one entry point; bench.py times
// tgx_plan + tgx_eval.
static int generate_setup(tgx_engine* e, cudaStream_t caller) {
    if (!e->twin) {
        int rc = tgx_create(&e->twin, e->device);
        if (rc) return rc;
    }
    tgx_engine* t = e->twin;
    if (t->tile_shift != e->tile_shift || t->spt != e->spt) {
        int rc = tgx_set_tuning(t, e->tile_shift, e->spt);
        if (rc) return rc;
    }
    if (t->exact_ramps != e->exact_ramps) tgx_set_plan_mode(t, e->exact_ramps);
    if (t->allow_slabs != e->allow_slabs) tgx_set_slab_planning(t, e->allow_slabs);
    if (t->allow_phase != e->allow_phase) tgx_set_phase_planning(t, e->allow_phase);
    t->plane_tma = e->plane_tma;
    t->max_samples = e->max_samples;
    if (!e->ge) {
        int least = 0, greatest = 0;       // numerically lower = higher priority
        TGX_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        for (int i = 0; i < 2; ++i) {
            TGX_CUDA(cudaStreamCreateWithPriority(&e->gs[i], cudaStreamNonBlocking, greatest));
            TGX_CUDA(cudaEventCreateWithFlags(&e->gev[i], cudaEventDisableTiming));
            TGX_CUDA(cudaEventCreateWithFlags(&e->gev_plan[i], cudaEventDisableTiming));
        }
        TGX_CUDA(cudaEventCreateWithFlags(&e->gev_in, cudaEventDisableTiming));
        TGX_CUDA(cudaStreamCreateWithPriority(&e->ge, cudaStreamNonBlocking, least));
    }
    // the chunks start after everything the caller has queued on its stream
    TGX_CUDA(cudaEventRecord(e->gev_in, caller));
    TGX_CUDA(cudaStreamWaitEvent(e->gs[0], e->gev_in, 0));
    TGX_CUDA(cudaStreamWaitEvent(e->gs[1], e->gev_in, 0));
    TGX_CUDA(cudaStreamWaitEvent(e->ge, e->gev_in, 0));
    return TGX_OK;
}

static int generate_finish(tgx_engine* e, cudaStream_t caller) {
    // every plan is followed by its evaluation on ge, so the end of ge is the end of the call
    TGX_CUDA(cudaEventRecord(e->gev_in, e->ge));
    TGX_CUDA(cudaStreamWaitEvent(caller, e->gev_in, 0));
    e->has_plan = false;
    if (e->twin) e->twin->has_plan = false;
    e->generate_calls += 1;
    return TGX_OK;
}

// Planning stream of chunk ci: its engine's tables are free once that engine's previous evaluation (chunk ci-2) is done.
static int generate_before_plan(tgx_engine* e, int64_t ci, cudaStream_t* plan_stream) {
    *plan_stream = e->gs[ci & 1];
    if (ci >= 2) TGX_CUDA(cudaStreamWaitEvent(*plan_stream, e->gev[ci & 1], 0));
    return TGX_OK;
}

static int generate_before_eval(tgx_engine* e, int64_t ci, cudaEvent_t* prof_end) {
    TGX_CUDA(cudaEventRecord(e->gev_plan[ci & 1], e->gs[ci & 1]));
    TGX_CUDA(cudaStreamWaitEvent(e->ge, e->gev_plan[ci & 1], 0));
    *prof_end = nullptr;
    if (e->prof_on) {
        if (e->prof_used == e->prof_ev.size()) {
            cudaEvent_t a = nullptr, b = nullptr;
            TGX_CUDA(cudaEventCreate(&a));
            TGX_CUDA(cudaEventCreate(&b));
            e->prof_ev.emplace_back(a, b);
        }
        TGX_CUDA(cudaEventRecord(e->prof_ev[e->prof_used].first, e->ge));
        *prof_end = e->prof_ev[e->prof_used].second;
        e->prof_used += 1;
    }
    return TGX_OK;
}

static int generate_after_eval(tgx_engine* e, int64_t ci, cudaEvent_t prof_end) {
    if (prof_end) TGX_CUDA(cudaEventRecord(prof_end, e->ge));
    TGX_CUDA(cudaEventRecord(e->gev[ci & 1], e->ge));
    return TGX_OK;
}

// Where the chunks of a pipelined call end: multiples of `chunk`, moved forward past continuation records so that no
// trajectory is separated from its goal speeds (one tiny kernel and one 8-byte-per-boundary read-back on the caller's
// stream, before anything else is queued).
static int generate_bounds(tgx_engine* e, const tgx_params* d_params, int64_t n, int64_t chunk, cudaStream_t caller,
                           std::vector<int64_t>& ends) {
    ends.clear();
    const int64_t nb = (n + chunk - 1) / chunk - 1;          // interior boundaries
    if (nb > 0) {
        if (nb > (1 << 20)) return TGX_ERR_INVALID;
        int rc = e->totals.reserve((size_t)std::max<int64_t>(nb, 4) * sizeof(int64_t));
        if (rc) return rc;
        std::vector<int64_t> h((size_t)nb);
        TGX_CUDA(tgx::launch_chunk_bounds(d_params, n, chunk, (int)nb, e->totals.as<int64_t>(), caller));
        TGX_CUDA(cudaMemcpyAsync(h.data(), e->totals.p, (size_t)nb * sizeof(int64_t), cudaMemcpyDeviceToHost, caller));
        TGX_CUDA(cudaStreamSynchronize(caller));
        for (int64_t v : h)
            if (v < n && (ends.empty() || v > ends.back())) ends.push_back(v);
    }
    ends.push_back(n);
    return TGX_OK;
}

static int64_t generate_chunk(int64_t n, int64_t chunk) {
    if (chunk <= 0) {
        // eight chunks hide 7/8 of the planning; at least 32 Ki trajectories (a few hundred CTAs per SM) per chunk so
        // that launch and synchronisation overheads stay small, at most 1 Mi so that the tables of a sweep stay small
        chunk = (n / 8 + 1023) / 1024 * 1024;
        chunk = std::max<int64_t>(chunk, 32768);
        chunk = std::min<int64_t>(chunk, (int64_t)1 << 20);
    }
    return std::min(chunk, std::max<int64_t>(n, 1));
}

int tgx_generate(tgx_engine* e, const tgx_params* d_params, int64_t n, const tgx_limits* limits, const tgx_layout* out,
                 int32_t* d_counts, uint32_t* d_status, tgx_phases* d_phases, int64_t chunk, int64_t* total_samples,
                 void* stream) {
    if (!e || n < 0 || (n > 0 && !d_params) || chunk < 0) return TGX_ERR_INVALID;
    int rc = check_layout(out, e->spt);
    if (rc) return rc;
    if (total_samples) *total_samples = 0;
    if (n == 0) return TGX_OK;
    TGX_CUDA(cudaSetDevice(e->device));
    cudaStream_t caller = static_cast<cudaStream_t>(stream);
    chunk = generate_chunk(n, chunk);
    std::vector<int64_t> ends;
    if ((rc = generate_bounds(e, d_params, n, chunk, caller, ends))) return rc;
    if ((rc = generate_setup(e, caller))) return rc;
    int64_t total = 0;
    int64_t ci = 0;
    for (int64_t lo = 0; ci < (int64_t)ends.size(); lo = ends[ci], ++ci) {
        const int64_t m = ends[ci] - lo;
        tgx_engine* eng = (ci & 1) ? e->twin : e;
        cudaStream_t ps;
        if ((rc = generate_before_plan(e, ci, &ps))) break;
        int64_t part = 0;
        rc = plan_common(eng, d_params + lo, nullptr, m, limits, d_counts ? d_counts + lo : nullptr,
                         d_status ? d_status + lo : nullptr, d_phases ? d_phases + lo : nullptr, &part, ps);
        if (rc) break;
        total += part;
        tgx_layout sub = *out;
        if (sub.d_traj_offset) sub.d_traj_offset += lo;
        else sub.d_base += lo * sub.traj_stride;
        cudaEvent_t prof_end;
        if ((rc = generate_before_eval(e, ci, &prof_end))) break;
        if ((rc = tgx_eval(eng, &sub, nullptr, nullptr, e->ge))) break;
        if ((rc = generate_after_eval(e, ci, prof_end))) break;
    }
    e->generate_chunks += ci;
    const int rc2 = generate_finish(e, caller);
    if (rc) return rc;
    if (total_samples) *total_samples = total;
    return rc2;
}

int tgx_generate_feasibility(tgx_engine* e, const tgx_params* d_params, int64_t n, const tgx_limits* limits,
                             uint8_t* d_flags, double* d_max_v, double* d_max_a, uint32_t* d_status, int64_t chunk,
                             int64_t* total_samples, void* stream) {
    if (!e || !limits || n < 0 || (n > 0 && !d_params) || chunk < 0) return TGX_ERR_INVALID;
    if (total_samples) *total_samples = 0;
    if (n == 0) return TGX_OK;
    TGX_CUDA(cudaSetDevice(e->device));
    cudaStream_t caller = static_cast<cudaStream_t>(stream);
    chunk = generate_chunk(n, chunk);
    std::vector<int64_t> ends;
    int rc = generate_bounds(e, d_params, n, chunk, caller, ends);
    if (rc) return rc;
    if ((rc = generate_setup(e, caller))) return rc;
    int64_t total = 0;
    int64_t ci = 0;
    for (int64_t lo = 0; ci < (int64_t)ends.size(); lo = ends[ci], ++ci) {
        const int64_t m = ends[ci] - lo;
        tgx_engine* eng = (ci & 1) ? e->twin : e;
        cudaStream_t ps;
        if ((rc = generate_before_plan(e, ci, &ps))) break;
        int64_t part = 0;
        rc = plan_common(eng, d_params + lo, nullptr, m, limits, nullptr, nullptr, nullptr, &part, ps);
        if (rc) break;
        total += part;
        cudaEvent_t prof_end;
        if ((rc = generate_before_eval(e, ci, &prof_end))) break;
        rc = tgx_feasibility(eng, limits, d_flags ? d_flags + lo : nullptr, d_max_v ? d_max_v + lo : nullptr,
                             d_max_a ? d_max_a + lo : nullptr, d_status ? d_status + lo : nullptr, e->ge);
        if (rc) break;
        if ((rc = generate_after_eval(e, ci, prof_end))) break;
    }
    e->generate_chunks += ci;
    const int rc2 = generate_finish(e, caller);
    if (rc) return rc;
    if (total_samples) *total_samples = total;
    return rc2;
}

// Profiling of tgx_generate / tgx_generate_feasibility: with on = 1 every evaluation launch is bracketed by an event
// pair on its stream; tgx_generate_profile waits for them, returns the sum of their durations and the number of
// launches since the last query, and starts over.
int tgx_set_generate_profiling(tgx_engine* e, int on) {
    if (!e) return TGX_ERR_INVALID;
    e->prof_on = on != 0;
    e->prof_used = 0;
    return TGX_OK;
}

int tgx_generate_profile(tgx_engine* e, double* eval_ms, int64_t* eval_launches) {
    if (!e) return TGX_ERR_INVALID;
    TGX_CUDA(cudaSetDevice(e->device));
    double sum = 0.0;
    for (size_t i = 0; i < e->prof_used; ++i) {
        TGX_CUDA(cudaEventSynchronize(e->prof_ev[i].second));
        float ms = 0.f;
        TGX_CUDA(cudaEventElapsedTime(&ms, e->prof_ev[i].first, e->prof_ev[i].second));
        sum += ms;
    }
    if (eval_ms) *eval_ms = sum;
    if (eval_launches) *eval_launches = (int64_t)e->prof_used;
    e->prof_used = 0;
    return TGX_OK;
}

// The record buffer as TMA sees it: a 2-D fp64 tensor [records][16], one 128-byte row per record, written in boxes of
// 32 records with the 128-byte shared-memory swizzle (RecTma in store.cuh stages in exactly that layout).  The outer
// extent is the number of records the buffer holds: TMA clips whatever a box would write beyond it, and the kernels never
// form a row coordinate outside [0, total).
static int make_record_tmap(CUtensorMap* tmap, tgx_goal_record* d_records, int64_t total) {
    tmap_encode_fn encode = tmap_encoder();
    if (!encode) {
        g_last_cuda_error = "cuTensorMapEncodeTiled is not available from this driver (tgx_eval_records needs TMA)";
        return TGX_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {16, (cuuint64_t)total};
    const cuuint64_t strides[1] = {sizeof(tgx_goal_record)};
    const cuuint32_t box[2] = {16, 32};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d_records, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        g_last_cuda_error = "cuTensorMapEncodeTiled failed for the record buffer (CUresult " + std::to_string((int)r) + ")";
        return TGX_ERR_CUDA;
    }
    return TGX_OK;
}

int tgx_eval_records(tgx_engine* e, const tgx_limits* limits, tgx_goal_record* d_records, int64_t rec_stride,
                     const int64_t* d_rec_offset, int64_t rec_capacity, void* stream) {
    if (!e || !d_records || rec_capacity < 0) return TGX_ERR_INVALID;
    if (!e->has_plan) return TGX_ERR_NO_PLAN;
    if ((reinterpret_cast<uintptr_t>(d_records) & 15u) != 0) return TGX_ERR_ALIGNMENT;
    TGX_CUDA(cudaSetDevice(e->device));
    if (e->plan_n == 0 || e->plan_tiles == 0 || rec_capacity == 0) return TGX_OK;
    // TMA addresses records by a 32-bit row coordinate (2^31 records = 275 GB, more than one GPU holds).  The buffer's
    // extent: n * rec_stride records, or — with per-trajectory offsets — rec_stride records in all
    if (rec_stride <= 0) return TGX_ERR_INVALID;
    const int64_t total = d_rec_offset ? rec_stride : e->plan_n * rec_stride;
    if (total > 0x7fffffffLL || (!d_rec_offset && rec_stride > 0x7fffffffLL)) return TGX_ERR_CAPACITY;
    tgx::RecOut ro{};
    ro.total = total;
    ro.base = d_records;
    ro.stride = rec_stride;
    ro.offset = d_rec_offset;
    ro.capacity = rec_capacity;
    ro.clamp = (limits && limits->check_box) ? 1 : 0;
    if (ro.clamp)
        for (int i = 0; i < 6; ++i) ro.box[i] = limits->box[i];
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int rc = make_record_tmap(&ro.tmap, d_records, total);
    if (rc) return rc;
    e->plan_stored = true;
    if (e->plan_poly)
        TGX_CUDA(tgx::launch_eval_poly_records(poly_view(e), e->plan_tiles, e->tile_shift, e->spt, ro, s));
    else
        TGX_CUDA(tgx::launch_eval_records(table_view(e), e->plan_tiles, e->tile_shift, e->spt, ro, s));
    e->launches += 1;
    return TGX_OK;
}

int tgx_pack_goals(tgx_engine* e, const tgx_layout* planes, const int32_t* d_counts, int64_t n,
                   const tgx_limits* limits, tgx_goal_record* d_records, int64_t rec_stride,
                   const int64_t* d_rec_offset, int64_t rec_capacity, void* stream) {
    if (!e || n < 0 || rec_capacity < 0 || (n > 0 && (!planes || !planes->d_base || !d_counts || !d_records)))
        return TGX_ERR_INVALID;
    if (n == 0 || rec_capacity == 0) return TGX_OK;
    if ((reinterpret_cast<uintptr_t>(d_records) & 15u) != 0) return TGX_ERR_ALIGNMENT;
    TGX_CUDA(cudaSetDevice(e->device));
    TGX_CUDA(tgx::launch_pack_goals(make_view(planes), d_counts, n, limits, d_records, rec_stride, d_rec_offset,
                                    rec_capacity, static_cast<cudaStream_t>(stream)));
    e->launches += 1;
    return TGX_OK;
}

int tgx_transitions(tgx_engine* e, const tgx_transition_params* d_tparams, int64_t n, const tgx_limits* limits,
                    tgx_goal_record* d_records, int64_t rec_stride, int64_t rec_capacity, int32_t* d_counts,
                    uint32_t* d_status, void* stream) {
    if (!e || n < 0 || rec_capacity < 0 || (n > 0 && !d_tparams)) return TGX_ERR_INVALID;
    if (n == 0) return TGX_OK;
    if (d_records && (reinterpret_cast<uintptr_t>(d_records) & 31u) != 0) return TGX_ERR_ALIGNMENT;
    TGX_CUDA(cudaSetDevice(e->device));
    TGX_CUDA(tgx::launch_transitions(d_tparams, n, limits, e->max_samples, d_records, rec_stride, rec_capacity,
                                     d_counts, d_status, static_cast<cudaStream_t>(stream)));
    e->launches += 1;
    return TGX_OK;
}

static int host_streams(tgx_engine* e);

int tgx_transitions_host(tgx_engine* e, const tgx_transition_params* h_tparams, int64_t n, const tgx_limits* limits,
                         tgx_goal_record* h_records, int64_t rec_capacity, int32_t* h_counts, uint32_t* h_status) {
    if (!e || n < 0 || rec_capacity < 0 || (n > 0 && !h_tparams)) return TGX_ERR_INVALID;
    if (n == 0) return TGX_OK;
    TGX_CUDA(cudaSetDevice(e->device));
    int rc = host_streams(e);
    if (rc) return rc;
    cudaStream_t s = e->hs[0];
    const bool want_rec = h_records && rec_capacity > 0;
    if ((rc = e->h_params[0].reserve((size_t)n * sizeof(tgx_transition_params)))) return rc;
    if ((rc = e->h_cnt[0].reserve((size_t)n * sizeof(int32_t)))) return rc;
    if ((rc = e->h_st[0].reserve((size_t)n * sizeof(uint32_t)))) return rc;
    if (want_rec && (rc = e->h_rec[0].reserve((size_t)(n * rec_capacity) * sizeof(tgx_goal_record)))) return rc;
    TGX_CUDA(cudaMemcpyAsync(e->h_params[0].p, h_tparams, (size_t)n * sizeof(tgx_transition_params),
                             cudaMemcpyHostToDevice, s));
    rc = tgx_transitions(e, e->h_params[0].as<tgx_transition_params>(), n, limits,
                         want_rec ? e->h_rec[0].as<tgx_goal_record>() : nullptr, rec_capacity,
                         want_rec ? rec_capacity : 0, e->h_cnt[0].as<int32_t>(), e->h_st[0].as<uint32_t>(), s);
    if (rc) return rc;
    if (want_rec)
        TGX_CUDA(cudaMemcpyAsync(h_records, e->h_rec[0].p, (size_t)(n * rec_capacity) * sizeof(tgx_goal_record),
                                 cudaMemcpyDeviceToHost, s));
    if (h_counts) TGX_CUDA(cudaMemcpyAsync(h_counts, e->h_cnt[0].p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (h_status) TGX_CUDA(cudaMemcpyAsync(h_status, e->h_st[0].p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    TGX_CUDA(cudaStreamSynchronize(s));
    return TGX_OK;
}

// FP64 peak of this GPU as a DFMA micro-benchmark (probe.cu): best of `reps` timed launches after one warm-up.
int tgx_probe_dfma(tgx_engine* e, int reps, double* dfma_per_s, double* ms_per_launch) {
    if (!e || !dfma_per_s || reps < 1) return TGX_ERR_INVALID;
    TGX_CUDA(cudaSetDevice(e->device));
    cudaDeviceProp prop;
    TGX_CUDA(cudaGetDeviceProperties(&prop, e->device));
    const int ctas = prop.multiProcessorCount * 4;          // one full wave of 4 x 256 threads per SM
    const int trips = 4096;                                 // x 128 DFMA per thread per trip: ~20 ms on a B200
    DevBuf sink;
    int rc = sink.reserve((size_t)ctas * 256 * sizeof(double));
    if (rc) return rc;
    cudaEvent_t a, b;
    TGX_CUDA(cudaEventCreate(&a));
    TGX_CUDA(cudaEventCreate(&b));
    double best_ms = 1e30, count = 0.0;
    cudaError_t err = cudaSuccess;
    for (int r = 0; r <= reps && err == cudaSuccess; ++r) {
        cudaEventRecord(a, nullptr);
        err = tgx::launch_dfma_probe(ctas, trips, sink.as<double>(), &count, nullptr);
        cudaEventRecord(b, nullptr);
        if (err == cudaSuccess) err = cudaEventSynchronize(b);
        float ms = 0.f;
        if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms > 0.f && (double)ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    sink.release();
    if (err != cudaSuccess) return cuda_fail(err, "tgx_probe_dfma");
    e->launches += reps + 1;
    *dfma_per_s = count / (best_ms * 1e-3);
    if (ms_per_launch) *ms_per_launch = best_ms;
    return TGX_OK;
}

// The end-to-end ceiling of the host-buffer calls on this machine, measured where they run: `reps` plain
// device->host copies of `bytes` into the caller's (pinned) buffer on the engine's copy stream, CUDA events.
int tgx_probe_d2h(tgx_engine* e, void* h_dst, int64_t bytes, int reps, double* seconds) {
    if (!e || !h_dst || bytes <= 0 || reps < 1 || !seconds) return TGX_ERR_INVALID;
    TGX_CUDA(cudaSetDevice(e->device));
    int rc = host_streams(e);
    if (rc) return rc;
    DevBuf src;
    if ((rc = src.reserve((size_t)bytes))) return rc;
    cudaStream_t s = e->hs[0];
    cudaEvent_t a = nullptr, b = nullptr;
    cudaError_t err = cudaMemsetAsync(src.p, 0x3c, (size_t)bytes, s);
    if (err == cudaSuccess) err = cudaEventCreate(&a);
    if (err == cudaSuccess) err = cudaEventCreate(&b);
    // warm-up: the first touch of a fresh pinned buffer is slower than the steady state
    if (err == cudaSuccess)
        err = cudaMemcpyAsync(h_dst, src.p, (size_t)std::min<int64_t>(bytes, (int64_t)256 << 20), cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess) err = cudaEventRecord(a, s);
    for (int r = 0; r < reps && err == cudaSuccess; ++r)
        err = cudaMemcpyAsync(h_dst, src.p, (size_t)bytes, cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess) err = cudaEventRecord(b, s);
    if (err == cudaSuccess) err = cudaEventSynchronize(b);
    float ms = 0.f;
    if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, a, b);
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
    src.release();
    if (err != cudaSuccess) return cuda_fail(err, "tgx_probe_d2h");
    *seconds = (double)ms * 1e-3;
    return TGX_OK;
}

int tgx_fill_montecarlo(tgx_engine* e, uint64_t seed, int64_t first_index, int64_t n, tgx_params* d_params,
                        void* stream) {
    if (!e || n < 0 || first_index < 0 || (n > 0 && !d_params)) return TGX_ERR_INVALID;
    if (n == 0) return TGX_OK;
    TGX_CUDA(cudaSetDevice(e->device));
    TGX_CUDA(tgx::launch_fill_montecarlo(seed, first_index, n, d_params, static_cast<cudaStream_t>(stream)));
    e->launches += 1;
    return TGX_OK;
}

int tgx_shard_range(int64_t n, int32_t rank, int32_t world, int64_t* lo, int64_t* hi) {
    if (n < 0 || world < 1 || rank < 0 || rank >= world || !lo || !hi) return TGX_ERR_INVALID;
    // floor(rank*n/world) without overflow for n < 2^62 / world
    *lo = (int64_t)(((__int128)n * rank) / world);
    *hi = (int64_t)(((__int128)n * (rank + 1)) / world);
    return TGX_OK;
}

// Debug aid: compares the planner's hoisted-reciprocal division with __ddiv_rn on n*per_thread pseudo-random
// operand pairs and returns the number of mismatches (must be 0).
int tgx_selftest_division(tgx_engine* e, int64_t n, uint64_t seed, int per_thread, uint64_t* mismatches) {
    if (!e || !mismatches || n < 0 || per_thread < 1) return TGX_ERR_INVALID;
    TGX_CUDA(cudaSetDevice(e->device));
    int rc = e->totals.reserve(4 * sizeof(int64_t));
    if (rc) return rc;
    unsigned long long* d = e->totals.as<unsigned long long>();
    TGX_CUDA(cudaMemset(d, 0, sizeof(unsigned long long)));
    TGX_CUDA(tgx::launch_selftest_division(n, seed, per_thread, d, nullptr));
    unsigned long long h = 0;
    TGX_CUDA(cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost));
    *mismatches = (uint64_t)h;
    return TGX_OK;
}

// ---- pinned host memory for callers that want asynchronous D2H (bench.py's e2e leg, the C++ drop-in) ----
void* tgx_alloc_host(int64_t bytes) {
    void* p = nullptr;
    if (bytes <= 0) return nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
// The same, with the pages preferably taken from the NUMA node the engine's GPU is attached to.
void* tgx_alloc_host_for(tgx_engine* e, int64_t bytes) {
    if (!e) return tgx_alloc_host(bytes);
    if (bytes <= 0 || cudaSetDevice(e->device) != cudaSuccess) return nullptr;
    host_topology(e);
    void* p = nullptr;
    with_preferred_node(e->numa_node, [&] { p = tgx_alloc_host(bytes); });
    return p;
}
void tgx_free_host(void* p) {
    if (p) cudaFreeHost(p);
}

int tgx_host_info(tgx_engine* e, tgx_host_info_t* out) {
    if (!e || !out) return TGX_ERR_INVALID;
    TGX_CUDA(cudaSetDevice(e->device));
    host_topology(e);
    out->numa_node = e->numa_node;
    out->cpus_allowed = e->cpus_allowed;
    out->local_ranks = e->local_ranks;
    out->filler_threads = e->filler_threads;
    out->filler_cpus = e->filler_cpus_valid ? CPU_COUNT(&e->filler_cpus) : 0;
    out->reserved[0] = out->reserved[1] = out->reserved[2] = 0;
    return TGX_OK;
}

// ---- host-buffer calls --------------------------------------------------------------------------------------

static int host_streams(tgx_engine* e) {
    for (int i = 0; i < 2; ++i) {
        if (!e->hs[i]) TGX_CUDA(cudaStreamCreateWithFlags(&e->hs[i], cudaStreamNonBlocking));
        if (!e->hev[i]) TGX_CUDA(cudaEventCreateWithFlags(&e->hev[i], cudaEventDisableTiming));
    }
    if (!e->hev_eval) TGX_CUDA(cudaEventCreateWithFlags(&e->hev_eval, cudaEventDisableTiming));
    if (!e->hev_up) TGX_CUDA(cudaEventCreateWithFlags(&e->hev_up, cudaEventDisableTiming));
    return TGX_OK;
}

// What a block of host-resident parameter records contains.
struct KindMix {
    bool classic = false, poly = false, bounce = false;
};

// Polyline-family records whose cos_o / sin_o the caller did not provide get the host libm's values (tgx.h:
// tgx_polyline_params) in a private copy; returns the array to upload (the caller's own when nothing had to change).
static const tgx_params* stage_params(const tgx_params* h_params, int64_t n, std::vector<tgx_params>& staged,
                                      KindMix* mix) {
    bool need_copy = false;
    KindMix m;
    for (int64_t i = 0; i < n; ++i) {
        const tgx_params& p = h_params[i];
        if (TGX_IS_POLYLINE(p.type)) {
            m.poly = true;
            if (p.type == TGX_BOUNCE) m.bounce = true;
            if (!(p.n_vgoals & TGX_POLY_TRIG_GIVEN)) need_copy = true;
        } else {
            m.classic = true;
        }
    }
    if (mix) *mix = m;
    if (!need_copy) return h_params;
    staged.assign(h_params, h_params + n);
    for (tgx_params& p : staged) {
        if (!TGX_IS_POLYLINE(p.type) || (p.n_vgoals & TGX_POLY_TRIG_GIVEN)) continue;
        p.u.poly.cos_o = std::cos(p.u.poly.orientation);
        p.u.poly.sin_o = std::sin(p.u.poly.orientation);
        p.n_vgoals |= TGX_POLY_TRIG_GIVEN;
    }
    return staged.data();
}

int tgx_count_host(tgx_engine* e, const tgx_params* h_params, int64_t n, const tgx_limits* limits,
                   int32_t* h_counts, uint32_t* h_status) {
    if (!e || n < 0 || (n > 0 && !h_params)) return TGX_ERR_INVALID;
    if (n == 0) return TGX_OK;
    TGX_CUDA(cudaSetDevice(e->device));
    int rc = host_streams(e);
    if (rc) return rc;
    cudaStream_t s = e->hs[0];
    if ((rc = e->h_params[0].reserve((size_t)n * sizeof(tgx_params)))) return rc;
    if ((rc = e->h_cnt[0].reserve((size_t)n * sizeof(int32_t)))) return rc;
    if ((rc = e->h_st[0].reserve((size_t)n * sizeof(uint32_t)))) return rc;
    std::vector<tgx_params> staged;
    const tgx_params* src = stage_params(h_params, n, staged, nullptr);
    TGX_CUDA(cudaMemcpyAsync(e->h_params[0].p, src, (size_t)n * sizeof(tgx_params), cudaMemcpyHostToDevice, s));
    rc = tgx_count(e, e->h_params[0].as<tgx_params>(), n, limits, e->h_cnt[0].as<int32_t>(),
                   e->h_st[0].as<uint32_t>(), s);
    if (rc) return rc;
    if (h_counts) TGX_CUDA(cudaMemcpyAsync(h_counts, e->h_cnt[0].p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (h_status) TGX_CUDA(cudaMemcpyAsync(h_status, e->h_st[0].p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    TGX_CUDA(cudaStreamSynchronize(s));
    return TGX_OK;
}

// Channels that vary along a trajectory (everything but p.z, v.z, a.z, j.z).
constexpr uint32_t kVaryingChannels = 0x3fffu & ~((1u << TGX_PZ) | (1u << TGX_VZ) | (1u << TGX_AZ) | (1u << TGX_JZ));

// One chunk of a host-buffer call: struct-of-arrays planes into the staging buffer, or (records != nullptr) clamped
// records straight from the evaluation kernel.
static int eval_chunk(tgx_engine* e, const tgx_layout* lay, const tgx_limits* limits, tgx_goal_record* records,
                      int64_t capacity, cudaStream_t s) {
    if (records) return tgx_eval_records(e, limits, records, capacity, nullptr, capacity, s);
    return tgx_eval(e, lay, nullptr, nullptr, s);
}

// Shared body of tgx_generate_host / tgx_stop_host.
// compact: the host buffer holds only the TGX_NCHAN_VARYING planes that vary along a trajectory (tgx.h:
// tgx_generate_host_compact); nothing is written for the constant ones.
static int host_run(tgx_engine* e, const tgx_params* h_params, const double* h_from, int64_t n,
                    const tgx_limits* limits, double* h_out, int64_t capacity, int32_t* h_counts,
                    uint32_t* h_status, tgx_phases* h_phases, tgx_polyline_legs* h_legs,
                    tgx_goal_record* h_records = nullptr, bool compact = false) {
    // h_records != nullptr: the samples stay on the device; what travels is one clamped 128-byte record per sample
    const auto t_entry = std::chrono::steady_clock::now();
    if (!e || n < 0 || (n > 0 && (!h_params || (!h_out && !h_records))) || capacity < 0) return TGX_ERR_INVALID;
    if (capacity % 4 != 0) return TGX_ERR_ALIGNMENT;
    if (n == 0) return TGX_OK;
    TGX_CUDA(cudaSetDevice(e->device));
    int rc = host_streams(e);
    if (rc) return rc;

    // chunk so that one device staging buffer stays <= ~1 GiB (two are in flight)
    const int64_t row_bytes = (int64_t)TGX_NCHAN * capacity * (int64_t)sizeof(double);
    int64_t chunk = row_bytes > 0 ? std::max<int64_t>(1, ((int64_t)1 << 30) / row_bytes) : n;
    chunk = std::min(chunk, n);
    // a chunk never ends between an orbit record and its continuation records (tgx.h: TGX_VGOALS_MORE): it may grow
    // by up to 7 records
    constexpr int64_t kMoreMax = TGX_MAX_VGOALS_TOTAL / TGX_MAX_VGOALS - 1;
    const int64_t chunk_cap = std::min(n, chunk + kMoreMax);

    for (int b = 0; b < 2 && (b == 0 || chunk < n); ++b) {
        if ((rc = e->h_params[b].reserve((size_t)chunk_cap * sizeof(tgx_params)))) return rc;
        if (!h_records && (rc = e->h_out[b].reserve((size_t)std::max<int64_t>(chunk_cap * row_bytes, 32)))) return rc;
        if ((rc = e->h_cnt[b].reserve((size_t)chunk_cap * sizeof(int32_t)))) return rc;
        if ((rc = e->h_st[b].reserve((size_t)chunk_cap * sizeof(uint32_t)))) return rc;
        if (h_phases && (rc = e->h_ph[b].reserve((size_t)chunk_cap * sizeof(tgx_phases)))) return rc;
        if (h_from && (rc = e->h_from[b].reserve((size_t)chunk_cap * TGX_NCHAN * sizeof(double)))) return rc;
        if (h_legs && (rc = e->h_legs[b].reserve((size_t)chunk_cap * sizeof(tgx_polyline_legs)))) return rc;
        if (h_records &&
            (rc = e->h_rec[b].reserve((size_t)std::max<int64_t>(chunk_cap * capacity, 1) * sizeof(tgx_goal_record))))
            return rc;
    }
    // Bounce moves along z (Bounce.cpp:39-41): its z-channels are not constants
    KindMix whole;
    for (int64_t i = 0; i < n && !whole.bounce; ++i)
        if (h_params[i].type == TGX_BOUNCE) whole.bounce = true;

    // The z-components are literal constants in the reference (p.z = alt_, v.z = a.z = j.z = 0: Circle.cpp:109-121,
    // Line.cpp:99-108, Figure8.cpp:110-119).  They are not worth 29 % of the PCIe traffic: the device evaluates and
    // ships the 10 varying planes, and host threads write the 4 constant rows of every trajectory meanwhile.
    if (compact && (whole.bounce || h_records || h_from)) return TGX_ERR_INVALID;
    // `fill`: the device evaluates and ships the varying planes only (always so in the compact format)
    const bool fill = (compact || e->host_fill_constants) && capacity > 0 && !whole.bounce && !h_records;
    const int64_t hplanes = compact ? TGX_NCHAN_VARYING : TGX_NCHAN;     // planes of the HOST buffer
    // Plane-major host buffers ([14][n][capacity]): every plane of a chunk is ONE contiguous run on both sides of the
    // bus, so the D2H copies are plain 1-D copies (52+ GB/s on a Gen5 x16 link) instead of 2-D copies of 16 KB runs
    // (46 GB/s).  The device staging buffer uses the same layout per chunk.
    const bool plane_major = e->host_plane_major && !h_records;
    std::vector<std::thread> fillers;
    if (fill && !compact) {
        // Sized to this process's share of the host (CPUs it may use / processes on the node) and bound to the CPUs of
        // the GPU's NUMA node: 8 ranks x 8 threads on a 32-core host only fought the DMA engines for memory bandwidth.
        host_topology(e);
        const int nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(e->filler_threads, n));
        const bool bind = e->filler_cpus_valid && e->numa_node >= 0;
        const cpu_set_t cpus = e->filler_cpus;
        for (int t = 0; t < nthreads; ++t) {
            const int64_t a = n * t / nthreads, z = n * (t + 1) / nthreads;
            fillers.emplace_back([=] {
                if (bind) pthread_setaffinity_np(pthread_self(), sizeof(cpus), &cpus);
                for (int64_t i = a; i < z; ++i) {
                    // element (i, c, k) of the host buffer
                    const int64_t ts = plane_major ? capacity : TGX_NCHAN * capacity;
                    const int64_t cs = plane_major ? n * capacity : capacity;
                    double* row = h_out + i * ts;
                    const double alt = h_params[i].alt;
                    double* pz = row + (int64_t)TGX_PZ * cs;
                    for (int64_t k = 0; k < capacity; ++k) pz[k] = alt;
                    std::memset(row + (int64_t)TGX_VZ * cs, 0, (size_t)capacity * sizeof(double));
                    std::memset(row + (int64_t)TGX_AZ * cs, 0, (size_t)capacity * sizeof(double));
                    std::memset(row + (int64_t)TGX_JZ * cs, 0, (size_t)capacity * sizeof(double));
                }
            });
        }
    }
    struct Joiner {
        std::vector<std::thread>& t;
        ~Joiner() { for (auto& x : t) if (x.joinable()) x.join(); }
    } joiner{fillers};

    // The parameters of the whole call go up FIRST, in one copy.  On these boxes a host-to-device copy queued while
    // device-to-host copies are in flight waits for them (events around the copy blocks: every chunk's H2D, and with it
    // its planning and evaluation, started when the previous chunk's last D2H copy had finished — 0.47 ms of idle link
    // per 14 ms chunk), so per-chunk uploads defeat the double buffering.  Up to 1 GiB of records (8 Mi trajectories);
    // larger calls keep the per-chunk uploads.
    const bool upfront = (size_t)n * sizeof(tgx_params) <= ((size_t)1 << 30);
    std::vector<tgx_params> staged_all;
    if (upfront) {
        if ((rc = e->h_params_all.reserve((size_t)n * sizeof(tgx_params)))) return rc;
        const tgx_params* src_all = stage_params(h_params, n, staged_all, nullptr);
        TGX_CUDA(cudaMemcpyAsync(e->h_params_all.p, src_all, (size_t)n * sizeof(tgx_params), cudaMemcpyHostToDevice, e->hs[0]));
        if (h_from) {
            if ((rc = e->h_from_all.reserve((size_t)n * TGX_NCHAN * sizeof(double)))) return rc;
            TGX_CUDA(cudaMemcpyAsync(e->h_from_all.p, h_from, (size_t)n * TGX_NCHAN * sizeof(double), cudaMemcpyHostToDevice,
                                     e->hs[0]));
        }
        TGX_CUDA(cudaEventRecord(e->hev_up, e->hs[0]));
        TGX_CUDA(cudaStreamWaitEvent(e->hs[1], e->hev_up, 0));
        if (!staged_all.empty()) TGX_CUDA(cudaEventSynchronize(e->hev_up));      // (the staged copy is a local)
    }
    // The per-trajectory results (counts, status, index_msgs, legs) land in PINNED memory of the engine and are handed to
    // the caller's arrays after the last chunk.  The caller's arrays are ordinary pageable memory, and a
    // cudaMemcpyAsync into pageable memory does not return before everything queued on its stream has finished: with
    // those four small copies queued behind a chunk's sample planes the host sat out every chunk's copies (14 ms)
    // before it could queue the next chunk, and the two staging slots never overlapped.
    const size_t small_bytes = (size_t)n * (sizeof(int32_t) + sizeof(uint32_t) + (h_phases ? sizeof(tgx_phases) : 0) +
                                            (h_legs ? sizeof(tgx_polyline_legs) : 0));
    if ((rc = e->p_small.reserve(small_bytes + 64))) return rc;
    // (8-byte alignment of every part: phases and legs first)
    char* sp = static_cast<char*>(e->p_small.p);
    tgx_phases* p_ph = reinterpret_cast<tgx_phases*>(sp);
    sp += h_phases ? (size_t)n * sizeof(tgx_phases) : 0;
    tgx_polyline_legs* p_legs = reinterpret_cast<tgx_polyline_legs*>(sp);
    sp += h_legs ? (size_t)n * sizeof(tgx_polyline_legs) : 0;
    int32_t* p_cnt = reinterpret_cast<int32_t*>(sp);
    sp += (size_t)n * sizeof(int32_t);
    uint32_t* p_st = reinterpret_cast<uint32_t*>(sp);
    static const bool trace = std::getenv("TGX_TRACE_D2H") != nullptr;
    std::vector<cudaEvent_t> tr_a, tr_b;
    std::vector<double> tr_host;
    const auto tr_t0 = std::chrono::steady_clock::now();
    auto tr_now = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tr_t0).count(); };
    cudaEvent_t tr_0 = nullptr;
    if (trace) { cudaEventCreate(&tr_0); cudaEventRecord(tr_0, e->hs[0]); }
    for (int64_t ci = 0, lo = 0, m = 0; lo < n; ++ci, lo += m) {
        const int b = (int)(ci & 1);
        cudaStream_t s = e->hs[b];
        m = std::min(chunk, n - lo);
        while (lo + m < n && m < chunk_cap && h_params[lo + m].type == TGX_VGOALS_MORE) ++m;
        // the staging buffers of slot b are free once the copies issued two chunks ago have completed
        if (ci >= 2) TGX_CUDA(cudaEventSynchronize(e->hev[b]));
        // the plan tables are shared by both slots: do not re-plan before the previous chunk's evaluation is done
        // (its D2H copy, the slow part, still overlaps with this chunk's planning and evaluation)
        if (ci >= 1) TGX_CUDA(cudaStreamWaitEvent(s, e->hev_eval, 0));
        KindMix mix;
        std::vector<tgx_params> staged;
        const tgx_params* d_par = e->h_params[b].as<tgx_params>();
        const double* d_from = h_from ? e->h_from[b].as<double>() : nullptr;
        if (upfront) {
            for (int64_t q = lo; q < lo + m; ++q) {
                const int ty = h_params[q].type;
                if (TGX_IS_POLYLINE(ty)) { mix.poly = true; if (ty == TGX_BOUNCE) mix.bounce = true; }
                else mix.classic = true;
            }
            d_par = e->h_params_all.as<tgx_params>() + lo;
            if (h_from) d_from = e->h_from_all.as<double>() + lo * TGX_NCHAN;
        } else {
            const tgx_params* src = stage_params(h_params + lo, m, staged, &mix);
            TGX_CUDA(cudaMemcpyAsync(e->h_params[b].p, src, (size_t)m * sizeof(tgx_params), cudaMemcpyHostToDevice, s));
            if (h_from)
                TGX_CUDA(cudaMemcpyAsync(e->h_from[b].p, h_from + lo * TGX_NCHAN, (size_t)m * TGX_NCHAN * sizeof(double),
                                         cudaMemcpyHostToDevice, s));
        }
        tgx_phases* d_ph = h_phases ? e->h_ph[b].as<tgx_phases>() : nullptr;
        tgx_polyline_legs* d_legs = h_legs ? e->h_legs[b].as<tgx_polyline_legs>() : nullptr;
        if (d_legs) TGX_CUDA(zero_fill(d_legs, (size_t)m * sizeof(tgx_polyline_legs), s));
        tgx_layout lay{};
        lay.d_base = e->h_out[b].as<double>();
        lay.traj_stride = plane_major ? capacity : TGX_NCHAN * capacity;
        lay.chan_stride = plane_major ? m * capacity : capacity;
        lay.capacity = capacity;
        lay.channel_mask = fill ? kVaryingChannels : 0;
        // the staging buffers are reused from call to call: clear the slot so that the padding (k >= N_i, and every
        // row of a rejected trajectory) reaches the caller as zeros, not as an earlier call's samples or records
        // (a zero-fill KERNEL at HBM rate, ~0.2 ms per GiB against ~20 ms of PCIe time for the same bytes — not
        //  cudaMemsetAsync, which a copy engine executes behind the other slot's device-to-host copies: zero_fill above)
        if (capacity > 0 && h_records)
            TGX_CUDA(zero_fill(e->h_rec[b].p, (size_t)(m * capacity) * sizeof(tgx_goal_record), s));
        else if (capacity > 0)
            TGX_CUDA(zero_fill(e->h_out[b].p, (size_t)(m * row_bytes), s));
        // planning synchronises stream s once; the other slot's D2H copies keep running meanwhile.  Braking plans take
        // every family in one pass; generateTraj plans route each family to its own planner (a mixed chunk is planned
        // and evaluated twice, each pass writing only its own trajectories' rows).
        const bool pass_classic = h_from || mix.classic;
        const bool pass_poly = !h_from && mix.poly;
        if (pass_classic) {
            if (h_from)
                rc = tgx_plan_stop(e, d_par, m, d_from, e->h_cnt[b].as<int32_t>(), e->h_st[b].as<uint32_t>(), d_ph,
                                   nullptr, s);
            else
                rc = tgx_plan(e, d_par, m, limits, e->h_cnt[b].as<int32_t>(), e->h_st[b].as<uint32_t>(), d_ph, nullptr, s);
            if (rc) return rc;
            if (capacity > 0 && (rc = eval_chunk(e, &lay, limits, h_records ? e->h_rec[b].as<tgx_goal_record>() : nullptr,
                                                 capacity, s)))
                return rc;
        }
        if (pass_poly) {
            if (d_ph && !pass_classic) TGX_CUDA(zero_fill(d_ph, (size_t)m * sizeof(tgx_phases), s));
            rc = plan_polyline_common(e, d_par, m, limits, e->h_cnt[b].as<int32_t>(), e->h_st[b].as<uint32_t>(), d_legs,
                                      nullptr, pass_classic, s);
            if (rc) return rc;
            if (capacity > 0 && (rc = eval_chunk(e, &lay, limits, h_records ? e->h_rec[b].as<tgx_goal_record>() : nullptr,
                                                 capacity, s)))
                return rc;
        }
        if (trace) { cudaEvent_t a; cudaEventCreate(&a); cudaEventRecord(a, s); tr_a.push_back(a); tr_host.push_back(tr_now()); }
        if (capacity > 0 && h_records) {
            TGX_CUDA(cudaEventRecord(e->hev_eval, s));
            TGX_CUDA(cudaMemcpyAsync(h_records + lo * capacity, e->h_rec[b].p,
                                     (size_t)(m * capacity) * sizeof(tgx_goal_record), cudaMemcpyDeviceToHost, s));
        } else if (capacity > 0 && plane_major) {
            TGX_CUDA(cudaEventRecord(e->hev_eval, s));
            for (int c = 0, q = 0; c < TGX_NCHAN; ++c) {
                if (fill && !(kVaryingChannels & (1u << c))) continue;
                const int64_t hc = compact ? q : c;      // plane index in the host buffer
                ++q;
                TGX_CUDA(cudaMemcpyAsync(h_out + (hc * n + lo) * capacity,
                                         e->h_out[b].as<double>() + (int64_t)c * m * capacity,
                                         (size_t)(m * capacity) * sizeof(double), cudaMemcpyDeviceToHost, s));
            }
        } else if (capacity > 0) {
            TGX_CUDA(cudaEventRecord(e->hev_eval, s));
            if (fill) {
                // the varying planes come in adjacent pairs (px,py | vx,vy | ax,ay | jx,jy | psi,dpsi): five 2-D copies,
                // each moving 2 rows of every trajectory of the chunk
                const size_t pitch = (size_t)row_bytes, width = (size_t)(2 * capacity) * sizeof(double);
                const size_t hpitch = (size_t)(hplanes * capacity) * sizeof(double);
                for (int q = 0; q < 5; ++q) {
                    const size_t off = (size_t)(3 * q) * (size_t)capacity;
                    const size_t hoff = (size_t)((compact ? 2 : 3) * q) * (size_t)capacity;
                    TGX_CUDA(cudaMemcpy2DAsync(h_out + lo * hplanes * capacity + hoff, hpitch,
                                               e->h_out[b].as<double>() + off, pitch, width, (size_t)m,
                                               cudaMemcpyDeviceToHost, s));
                }
            } else {
                TGX_CUDA(cudaMemcpyAsync(h_out + lo * TGX_NCHAN * capacity, e->h_out[b].p, (size_t)(m * row_bytes),
                                         cudaMemcpyDeviceToHost, s));
            }
        }
        TGX_CUDA(cudaMemcpyAsync(p_cnt + lo, e->h_cnt[b].p, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        TGX_CUDA(cudaMemcpyAsync(p_st + lo, e->h_st[b].p, (size_t)m * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        if (h_phases)
            TGX_CUDA(cudaMemcpyAsync(p_ph + lo, e->h_ph[b].p, (size_t)m * sizeof(tgx_phases), cudaMemcpyDeviceToHost, s));
        if (h_legs)
            TGX_CUDA(cudaMemcpyAsync(p_legs + lo, e->h_legs[b].p, (size_t)m * sizeof(tgx_polyline_legs),
                                     cudaMemcpyDeviceToHost, s));
        TGX_CUDA(cudaEventRecord(e->hev[b], s));
        if (trace) { cudaEvent_t z; cudaEventCreate(&z); cudaEventRecord(z, s); tr_b.push_back(z); tr_host.push_back(tr_now()); }
    }
    TGX_CUDA(cudaStreamSynchronize(e->hs[0]));
    TGX_CUDA(cudaStreamSynchronize(e->hs[1]));
    if (trace) {
        float t_prev_end = 0.f;
        for (size_t q = 0; q < tr_a.size(); ++q) {
            float ta = 0.f, tb = 0.f;
            cudaEventElapsedTime(&ta, tr_0, tr_a[q]);
            cudaEventElapsedTime(&tb, tr_0, tr_b[q]);
            std::fprintf(stderr, "[tgx d2h] chunk %zu: eval done %.3f ms, copies done %.3f ms (%.3f ms after the previous chunk's); host: eval queued %.3f, copies queued %.3f\n",
                         q, ta, tb, tb - t_prev_end, tr_host[2 * q], tr_host[2 * q + 1]);
            t_prev_end = tb;
            cudaEventDestroy(tr_a[q]);
            cudaEventDestroy(tr_b[q]);
        }
        cudaEventDestroy(tr_0);
    }
    for (auto& x : fillers) x.join();
    // hand the per-trajectory results over; TRUNCATED is a property of the caller's capacity, known only here
    if (h_status)
        for (int64_t i = 0; i < n; ++i) h_status[i] = p_st[i] | ((int64_t)p_cnt[i] > capacity ? (uint32_t)TGX_ST_TRUNCATED : 0u);
    if (h_counts) std::memcpy(h_counts, p_cnt, (size_t)n * sizeof(int32_t));
    if (h_phases) std::memcpy(h_phases, p_ph, (size_t)n * sizeof(tgx_phases));
    if (h_legs) std::memcpy(h_legs, p_legs, (size_t)n * sizeof(tgx_polyline_legs));
    if (trace) {
        static std::chrono::steady_clock::time_point last_return;
        static bool have_last = false;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[tgx d2h] call: %.3f ms inside (tracing starts %.3f ms after entry), %.3f ms since the previous call returned\n",
                     std::chrono::duration<double, std::milli>(now - t_entry).count(),
                     std::chrono::duration<double, std::milli>(tr_t0 - t_entry).count(),
                     have_last ? std::chrono::duration<double, std::milli>(t_entry - last_return).count() : 0.0);
        last_return = now;
        have_last = true;
    }
    return TGX_OK;
}

int tgx_sample_host(tgx_engine* e, const tgx_params* h_params, double v, double accel, double s0, double s1,
                    double* h_out14) {
    if (!e || !h_params || !h_out14) return TGX_ERR_INVALID;
    TGX_CUDA(cudaSetDevice(e->device));
    int rc = host_streams(e);
    if (rc) return rc;
    cudaStream_t s = e->hs[0];
    if ((rc = e->h_params[0].reserve(sizeof(tgx_params)))) return rc;
    if ((rc = e->h_from[0].reserve(4 * sizeof(double)))) return rc;
    if ((rc = e->h_out[0].reserve((size_t)TGX_NCHAN * 4 * sizeof(double)))) return rc;
    const double state[4] = {v, accel, s0, s1};
    TGX_CUDA(cudaMemcpyAsync(e->h_params[0].p, h_params, sizeof(tgx_params), cudaMemcpyHostToDevice, s));
    TGX_CUDA(cudaMemcpyAsync(e->h_from[0].p, state, sizeof(state), cudaMemcpyHostToDevice, s));
    rc = tgx_plan_samples(e, e->h_params[0].as<tgx_params>(), 1, e->h_from[0].as<double>(), nullptr, nullptr, s);
    if (rc) return rc;
    tgx_layout lay{};
    lay.d_base = e->h_out[0].as<double>();
    lay.traj_stride = TGX_NCHAN * 4;
    lay.chan_stride = 4;
    lay.capacity = 4;
    rc = tgx_eval(e, &lay, nullptr, nullptr, s);
    if (rc) return rc;
    double tmp[TGX_NCHAN * 4];
    TGX_CUDA(cudaMemcpyAsync(tmp, e->h_out[0].p, sizeof(tmp), cudaMemcpyDeviceToHost, s));
    TGX_CUDA(cudaStreamSynchronize(s));
    for (int c = 0; c < TGX_NCHAN; ++c) h_out14[c] = tmp[c * 4];
    return TGX_OK;
}

int tgx_generate_host(tgx_engine* e, const tgx_params* h_params, int64_t n, const tgx_limits* limits, double* h_out,
                      int64_t capacity, int32_t* h_counts, uint32_t* h_status, tgx_phases* h_phases) {
    return host_run(e, h_params, nullptr, n, limits, h_out, capacity, h_counts, h_status, h_phases, nullptr);
}

int tgx_generate_host_compact(tgx_engine* e, const tgx_params* h_params, int64_t n, const tgx_limits* limits,
                              double* h_out10, int64_t capacity, int32_t* h_counts, uint32_t* h_status,
                              tgx_phases* h_phases, tgx_polyline_legs* h_legs) {
    return host_run(e, h_params, nullptr, n, limits, h_out10, capacity, h_counts, h_status, h_phases, h_legs, nullptr,
                    true);
}

int tgx_generate_host_legs(tgx_engine* e, const tgx_params* h_params, int64_t n, const tgx_limits* limits,
                           double* h_out, int64_t capacity, int32_t* h_counts, uint32_t* h_status,
                           tgx_phases* h_phases, tgx_polyline_legs* h_legs) {
    return host_run(e, h_params, nullptr, n, limits, h_out, capacity, h_counts, h_status, h_phases, h_legs);
}

int tgx_generate_records_host(tgx_engine* e, const tgx_params* h_params, int64_t n, const tgx_limits* limits,
                              tgx_goal_record* h_records, int64_t rec_capacity, int32_t* h_counts,
                              uint32_t* h_status) {
    if (n > 0 && !h_records) return TGX_ERR_INVALID;
    return host_run(e, h_params, nullptr, n, limits, nullptr, rec_capacity, h_counts, h_status, nullptr, nullptr,
                    h_records);
}

int tgx_stop_host(tgx_engine* e, const tgx_params* h_params, int64_t n, const double* h_from, double* h_out,
                  int64_t capacity, int32_t* h_counts, uint32_t* h_status, tgx_phases* h_phases) {
    if (n > 0 && !h_from) return TGX_ERR_INVALID;
    return host_run(e, h_params, h_from, n, nullptr, h_out, capacity, h_counts, h_status, h_phases, nullptr);
}

}  // extern "C"
