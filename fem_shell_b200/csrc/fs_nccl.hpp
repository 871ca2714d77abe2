// fs_nccl.hpp -- NCCL entry points resolved at run time.
//
// The library is not linked against libnccl: a process that also loads PyTorch must end up with ONE
// libnccl.so.2 (torch bundles a newer NCCL than the system package and fails to import if an older
// one is already mapped).  dlopen by SONAME returns whichever copy the process already holds, or
// the system library in a plain C++ host program.  Only the multi-GPU entry points touch this.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace fs {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    bool ok = false;
};

inline NcclApi &nccl()
{
    static NcclApi api = [] {
        NcclApi a;
        a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.handle) a.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.handle) return a;
#define FS_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, name))
        FS_SYM(GetUniqueId, "ncclGetUniqueId");
        FS_SYM(CommInitRank, "ncclCommInitRank");
        FS_SYM(CommDestroy, "ncclCommDestroy");
        FS_SYM(GetErrorString, "ncclGetErrorString");
        FS_SYM(GroupStart, "ncclGroupStart");
        FS_SYM(GroupEnd, "ncclGroupEnd");
        FS_SYM(Send, "ncclSend");
        FS_SYM(Recv, "ncclRecv");
        FS_SYM(AllReduce, "ncclAllReduce");
        FS_SYM(AllGather, "ncclAllGather");
#undef FS_SYM
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.GetErrorString && a.GroupStart && a.GroupEnd &&
               a.Send && a.Recv && a.AllReduce && a.AllGather;
        return a;
    }();
    return api;
}

}  // namespace fs
