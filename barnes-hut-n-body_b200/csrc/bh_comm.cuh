// bh_comm.cuh — NCCL over NVLink 5 / NVSwitch, loaded at run time (dlopen "libnccl.so.2").
// No link-time dependency: a single-GPU user never needs NCCL; inside a torch process the
// already-loaded torch-bundled NCCL is reused.  Only the few entry points the step needs.
#ifndef BH_COMM_CUH
#define BH_COMM_CUH

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>
#include <string>

namespace bhcomm {

struct UniqueId { char internal[128]; };   // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* Comm;                         // ncclComm_t
constexpr int kSuccess = 0;                 // ncclSuccess
constexpr int kFloat64 = 8;                 // ncclFloat64
constexpr int kSum = 0;                     // ncclSum

struct Api {
    void* handle = nullptr;
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, Comm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
    bool ok = false;
};

inline Api& api() {
    static Api a;
    if (a.ok || !a.err.empty()) return a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        a.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) { a.err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return a; }
#define BH_SYM(field, name)                                                             \
    *(void**)(&a.field) = dlsym(a.handle, name);                                        \
    if (!a.field) { a.err = std::string("libnccl lacks ") + name; return a; }
    BH_SYM(GetUniqueId, "ncclGetUniqueId")
    BH_SYM(CommInitRank, "ncclCommInitRank")
    BH_SYM(CommDestroy, "ncclCommDestroy")
    BH_SYM(AllGather, "ncclAllGather")
    BH_SYM(AllReduce, "ncclAllReduce")
    BH_SYM(Broadcast, "ncclBroadcast")
    BH_SYM(Send, "ncclSend")
    BH_SYM(Recv, "ncclRecv")
    BH_SYM(GroupStart, "ncclGroupStart")
    BH_SYM(GroupEnd, "ncclGroupEnd")
    BH_SYM(GetErrorString, "ncclGetErrorString")
#undef BH_SYM
    a.ok = true;
    return a;
}

}  // namespace bhcomm
#endif  // BH_COMM_CUH
