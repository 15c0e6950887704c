// b200reg — NCCL communicator entry points (see comm.cuh).
#include "comm.cuh"

extern "C" {

static int32_t b200_comm_unique_id_impl(uint8_t* id128) {
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
    c->Send = (decltype(c->Send))dlsym(lib, "ncclSend");
    c->Recv = (decltype(c->Recv))dlsym(lib, "ncclRecv");
    c->GroupStart = (decltype(c->GroupStart))dlsym(lib, "ncclGroupStart");
    c->GroupEnd = (decltype(c->GroupEnd))dlsym(lib, "ncclGroupEnd");
    c->CommDestroy = (decltype(c->CommDestroy))dlsym(lib, "ncclCommDestroy");
    c->GetErrorString = (decltype(c->GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!init || !c->AllReduce || !c->Broadcast || !c->AllGather || !c->Send || !c->Recv || !c->GroupStart || !c->GroupEnd || !c->CommDestroy) { delete c; B200_FAIL(B200_ERR_NCCL, "NCCL symbols missing"); }
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
