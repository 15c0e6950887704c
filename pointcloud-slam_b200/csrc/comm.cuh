// b200reg — NCCL communicator for the sharded paths (global relocalization, batch map construction).
//
// One process per GPU; the unique id travels through whatever the host already has (torch.distributed,
// MPI, a file).  libnccl.so.2 is resolved at run time: if the process already carries an NCCL (e.g. the
// one PyTorch bundles) that copy is reused, so there are never two NCCL runtimes in one address space.
// The reference has no communication backend at all; this is new capability for
// BASELINE.json configs[3] / configs[4].
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

struct b200_comm {
    void* lib = nullptr;
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0, device = 0;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

namespace b200 {

inline void* nccl_lib() {
    static void* lib = nullptr;
    if (lib) return lib;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already in the process (PyTorch's)?
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    return lib;
}

#define NCCL_TRY(c, expr)                                                                                     \
    do {                                                                                                      \
        ncclResult_t _r = (expr);                                                                             \
        if (_r != ncclSuccess) return ::b200::fail(B200_ERR_NCCL, (c)->GetErrorString ? (c)->GetErrorString(_r) : "nccl error", __FILE__, __LINE__); \
    } while (0)

inline int32_t comm_allreduce_u64(b200_comm* c, const unsigned long long* d_in, unsigned long long* d_out, ncclRedOp_t op, cudaStream_t s) {
    NCCL_TRY(c, c->AllReduce(d_in, d_out, 1, ncclUint64, op, c->comm, s));
    return B200_OK;
}

}  // namespace b200

extern "C" {

inline int32_t b200_comm_unique_id_impl(uint8_t* id128) {
    if (!id128) B200_FAIL(B200_ERR_ARG, "null argument");
    void* lib = b200::nccl_lib();
    if (!lib) B200_FAIL(B200_ERR_NCCL, "libnccl.so.2 not found");
    auto get = (ncclResult_t(*)(ncclUniqueId*))dlsym(lib, "ncclGetUniqueId");
    if (!get) B200_FAIL(B200_ERR_NCCL, "ncclGetUniqueId not found");
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == B200_NCCL_ID_BYTES, "unique id size");
    if (get(&id) != ncclSuccess) B200_FAIL(B200_ERR_NCCL, "ncclGetUniqueId failed");
    memcpy(id128, &id, sizeof id);
    return B200_OK;
}

int32_t b200_comm_unique_id(uint8_t* id128) { return b200_comm_unique_id_impl(id128); }

int32_t b200_comm_init_rank(int32_t nranks, int32_t rank, const uint8_t* id128, int32_t device, b200_comm** out) {
    if (!id128 || !out || nranks < 1 || rank < 0 || rank >= nranks) B200_FAIL(B200_ERR_ARG, "bad argument");
    void* lib = b200::nccl_lib();
    if (!lib) B200_FAIL(B200_ERR_NCCL, "libnccl.so.2 not found");
    b200_comm* c = new b200_comm();
    c->lib = lib;
    c->nranks = nranks; c->rank = rank; c->device = device;
    auto init = (ncclResult_t(*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(lib, "ncclCommInitRank");
    c->AllReduce = (decltype(c->AllReduce))dlsym(lib, "ncclAllReduce");
    c->Broadcast = (decltype(c->Broadcast))dlsym(lib, "ncclBroadcast");
    c->AllGather = (decltype(c->AllGather))dlsym(lib, "ncclAllGather");
    c->CommDestroy = (decltype(c->CommDestroy))dlsym(lib, "ncclCommDestroy");
    c->GetErrorString = (decltype(c->GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!init || !c->AllReduce || !c->Broadcast || !c->AllGather || !c->CommDestroy) { delete c; B200_FAIL(B200_ERR_NCCL, "NCCL symbols missing"); }
    if (cudaSetDevice(device) != cudaSuccess) { delete c; B200_FAIL(B200_ERR_CUDA, "cudaSetDevice failed"); }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclResult_t r = init(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) {
        const char* msg = c->GetErrorString ? c->GetErrorString(r) : "ncclCommInitRank failed";
        delete c;
        B200_FAIL(B200_ERR_NCCL, msg);
    }
    *out = c;
    return B200_OK;
}

int32_t b200_comm_destroy(b200_comm* c) {
    if (!c) return B200_OK;
    if (c->comm && c->CommDestroy) c->CommDestroy(c->comm);
    delete c;
    return B200_OK;
}

}  // extern "C"
