// NCCL plumbing for the coset-sharded prover (one process per GPU).  NCCL is resolved at run time with
// dlopen("libnccl.so.2") -- the copy torch already mapped when the host process uses torch.distributed -- so the
// library has no link-time dependency on a particular NCCL build.  Only two collectives sit on the data path:
// all-gather of 32-byte digests at every Merkle commit, and one all-reduce of the queried values.
#pragma once
#include <dlfcn.h>
#include <cstring>
#include <string>
#include <cuda_runtime.h>

namespace gs {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;

struct Nccl {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string error;
    bool load() {
        if (handle) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) { handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (handle) break; }
        if (!handle) { error = std::string("cannot load NCCL: ") + dlerror(); return false; }
#define GS_SYM(field, sym) field = reinterpret_cast<decltype(field)>(dlsym(handle, sym)); if (!field) { error = std::string("NCCL symbol missing: ") + sym; return false; }
        GS_SYM(GetUniqueId, "ncclGetUniqueId") GS_SYM(CommInitRank, "ncclCommInitRank") GS_SYM(CommDestroy, "ncclCommDestroy")
        GS_SYM(Send, "ncclSend") GS_SYM(Recv, "ncclRecv") GS_SYM(GroupStart, "ncclGroupStart") GS_SYM(GroupEnd, "ncclGroupEnd")
        GS_SYM(AllGather, "ncclAllGather") GS_SYM(AllReduce, "ncclAllReduce") GS_SYM(GetErrorString, "ncclGetErrorString")
#undef GS_SYM
        return true;
    }
};
static inline Nccl& nccl() { static Nccl n; return n; }
enum { GS_NCCL_UINT8 = 1, GS_NCCL_UINT32 = 3, GS_NCCL_SUM = 0 };   // ncclUint8, ncclUint32, ncclSum

// which rank owns evaluation position i = q*E + j, and where it lives there (cosets are dealt in contiguous ranges)
struct Shard {
    int rank = 0, world = 1, log_e = 0, log_el = 0;       // E = 2^log_e cosets, El = E / world per rank
    int j0() const { return rank << log_el; }
    long long to_global(long long i_loc) const { return ((i_loc >> log_el) << log_e) + j0() + (i_loc & ((1ll << log_el) - 1)); }
    int owner(long long i) const { return (int)((i & ((1ll << log_e) - 1)) >> log_el); }
    long long to_local(long long i) const { return ((i >> log_e) << log_el) + (i & ((1ll << log_el) - 1)); }
};

}  // namespace gs
