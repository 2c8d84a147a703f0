// NCCL through dlopen: libdbmm.so does not link against NCCL; the data-parallel entry points bind the handful of calls
// they need at first use, preferring the copy that is already loaded in the process (PyTorch's bundled libnccl.so.2).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace dbmm {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
    bool ok;
};

static NcclApi* nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        memset(&a, 0, sizeof(a));
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return a;
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
        a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
        a.AllGather = (decltype(a.AllGather))dlsym(h, "ncclAllGather");
        a.ReduceScatter = (decltype(a.ReduceScatter))dlsym(h, "ncclReduceScatter");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.AllGather && a.ReduceScatter && a.GetErrorString;
        return a;
    }();
    return &api;
}

#define DBMM_NCCL(call)                                                                           \
    do {                                                                                          \
        ncclResult_t r__ = (call);                                                                \
        if (r__ != ncclSuccess) {                                                                 \
            dbmm::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                  \
                            dbmm::nccl_api()->GetErrorString(r__));                               \
            return DBMM_ERR_CUDA;                                                                 \
        }                                                                                         \
    } while (0)

}  // namespace dbmm
