// comm.cu — the one exchange step of the path, behind the C-ABI: an all-gather of the 1-byte feasibility flags.
//
// Batches are sharded over the GPUs of a box by contiguous index blocks (tgx_shard_range) and every shard is planned
// and reduced on its own GPU; the only inter-GPU traffic BASELINE.json's north star allows is "an optional NCCL gather
// of feasibility flags" (configs[4]).  It is 1 byte per trajectory — 10^8 bytes for the whole sweep, microseconds of
// NVLink time — so it is a plain ncclAllGather on the caller's stream, not a fused kernel.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): libtgx.so has no link-time dependency on it, a process that
// already carries an NCCL (torch's bundled one, same soname) shares that copy, and a host without NCCL still loads the
// library and only fails — loudly — in these entry points.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <new>
#include <string>
#include <type_traits>

#include "tgx_internal.cuh"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        const char* why = dlerror();
        api.error = std::string("NCCL is not available: ") + (why ? why : "dlopen(libnccl.so.2) failed");
        return api;
    }
    bool ok = true;
    auto bind = [&](auto& fn, const char* sym) {
        fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(api.handle, sym));
        if (!fn) {
            ok = false;
            api.error = std::string("NCCL symbol missing: ") + sym;
        }
    };
    bind(api.GetUniqueId, "ncclGetUniqueId");
    bind(api.CommInitRank, "ncclCommInitRank");
    bind(api.CommInitAll, "ncclCommInitAll");
    bind(api.CommDestroy, "ncclCommDestroy");
    bind(api.AllGather, "ncclAllGather");
    bind(api.Broadcast, "ncclBroadcast");
    bind(api.GroupStart, "ncclGroupStart");
    bind(api.GroupEnd, "ncclGroupEnd");
    bind(api.GetVersion, "ncclGetVersion");
    bind(api.GetErrorString, "ncclGetErrorString");
    if (!ok) {
        dlclose(api.handle);
        api.handle = nullptr;
    }
    return api;
}

thread_local std::string g_comm_error;

int fail(const std::string& what) {
    g_comm_error = what;
    return TGX_ERR_COMM;
}

int check(ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return TGX_OK;
    return fail(std::string(what) + ": " + nccl().GetErrorString(r));
}

}  // namespace

struct tgx_comm {
    ncclComm_t comm = nullptr;
    int world = 1;
    int rank = 0;
    int device = 0;
};

static_assert(sizeof(ncclUniqueId) == TGX_COMM_ID_BYTES, "tgx.h: TGX_COMM_ID_BYTES must be sizeof(ncclUniqueId)");

extern "C" {

const char* tgx_comm_last_error(void) { return g_comm_error.c_str(); }

int tgx_comm_nccl_version(int* version) {
    if (!version) return TGX_ERR_INVALID;
    NcclApi& api = nccl();
    if (!api.handle) return fail(api.error);
    return check(api.GetVersion(version), "ncclGetVersion");
}

int tgx_comm_unique_id(char id[TGX_COMM_ID_BYTES]) {
    if (!id) return TGX_ERR_INVALID;
    NcclApi& api = nccl();
    if (!api.handle) return fail(api.error);
    ncclUniqueId u;
    const int rc = check(api.GetUniqueId(&u), "ncclGetUniqueId");
    if (rc) return rc;
    std::memcpy(id, &u, sizeof(u));
    return TGX_OK;
}

int tgx_comm_init_rank(tgx_comm** out, int world, int rank, const char id[TGX_COMM_ID_BYTES], int device) {
    if (!out || !id || world < 1 || rank < 0 || rank >= world) return TGX_ERR_INVALID;
    *out = nullptr;
    NcclApi& api = nccl();
    if (!api.handle) return fail(api.error);
    if (cudaSetDevice(device) != cudaSuccess) {
        cudaGetLastError();
        return fail("cudaSetDevice(" + std::to_string(device) + ") failed");
    }
    tgx_comm* c = new (std::nothrow) tgx_comm();
    if (!c) return TGX_ERR_NOMEM;
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof(u));
    const int rc = check(api.CommInitRank(&c->comm, world, u, rank), "ncclCommInitRank");
    if (rc) {
        delete c;
        return rc;
    }
    c->world = world;
    c->rank = rank;
    c->device = device;
    *out = c;
    return TGX_OK;
}

int tgx_comm_init_all(tgx_comm** out, int ndev, const int* devices) {
    if (!out || ndev < 1 || ndev > 64) return TGX_ERR_INVALID;
    for (int i = 0; i < ndev; ++i) out[i] = nullptr;
    NcclApi& api = nccl();
    if (!api.handle) return fail(api.error);
    ncclComm_t comms[64];
    int devs[64];
    for (int i = 0; i < ndev; ++i) devs[i] = devices ? devices[i] : i;
    const int rc = check(api.CommInitAll(comms, ndev, devs), "ncclCommInitAll");
    if (rc) return rc;
    for (int i = 0; i < ndev; ++i) {
        tgx_comm* c = new (std::nothrow) tgx_comm();
        if (!c) return TGX_ERR_NOMEM;
        c->comm = comms[i];
        c->world = ndev;
        c->rank = i;
        c->device = devs[i];
        out[i] = c;
    }
    return TGX_OK;
}

int tgx_comm_destroy(tgx_comm* c) {
    if (!c) return TGX_OK;
    int rc = TGX_OK;
    if (c->comm && nccl().handle) {
        cudaSetDevice(c->device);
        rc = check(nccl().CommDestroy(c->comm), "ncclCommDestroy");
    }
    delete c;
    return rc;
}

int tgx_comm_group_start(void) {
    NcclApi& api = nccl();
    if (!api.handle) return fail(api.error);
    return check(api.GroupStart(), "ncclGroupStart");
}

int tgx_comm_group_end(void) {
    NcclApi& api = nccl();
    if (!api.handle) return fail(api.error);
    return check(api.GroupEnd(), "ncclGroupEnd");
}

int tgx_gather_flags(tgx_comm* c, const uint8_t* d_local, int64_t n_total, uint8_t* d_all, void* stream) {
    if (n_total < 0 || (n_total > 0 && !d_all)) return TGX_ERR_INVALID;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n_total == 0) return TGX_OK;
    if (!c || c->world == 1) {
        // one shard: the gather is a device-to-device copy
        if (!d_local) return TGX_ERR_INVALID;
        if (c && cudaSetDevice(c->device) != cudaSuccess) {
            cudaGetLastError();
            return fail("cudaSetDevice failed");
        }
        if (d_local != d_all &&
            cudaMemcpyAsync(d_all, d_local, (size_t)n_total, cudaMemcpyDeviceToDevice, s) != cudaSuccess) {
            cudaGetLastError();
            return fail("cudaMemcpyAsync (single-shard gather) failed");
        }
        return TGX_OK;
    }
    NcclApi& api = nccl();
    if (!api.handle) return fail(api.error);
    if (cudaSetDevice(c->device) != cudaSuccess) {
        cudaGetLastError();
        return fail("cudaSetDevice failed");
    }
    int64_t lo = 0, hi = 0;
    tgx_shard_range(n_total, c->rank, c->world, &lo, &hi);
    if (hi > lo && !d_local) return TGX_ERR_INVALID;
    if (n_total % c->world == 0) {
        // equal shards: ONE ncclAllGather, rank r's flags land at d_all[r * n_total / world ..]
        return check(api.AllGather(d_local, d_all, (size_t)(n_total / c->world), ncclUint8, c->comm, s),
                     "ncclAllGather");
    }
    // shards that differ by one trajectory: the all-gather-v idiom, one broadcast per shard inside a group
    int rc = check(api.GroupStart(), "ncclGroupStart");
    if (rc) return rc;
    for (int r = 0; r < c->world && rc == TGX_OK; ++r) {
        int64_t rlo = 0, rhi = 0;
        tgx_shard_range(n_total, r, c->world, &rlo, &rhi);
        if (rhi == rlo) continue;
        rc = check(api.Broadcast(r == c->rank ? (const void*)d_local : (const void*)(d_all + rlo), d_all + rlo,
                                 (size_t)(rhi - rlo), ncclUint8, r, c->comm, s),
                   "ncclBroadcast");
    }
    const int rc2 = check(api.GroupEnd(), "ncclGroupEnd");
    return rc ? rc : rc2;
}

}  // extern "C"
