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
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
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
