// nccl_dyn.hpp -- NCCL entry points resolved with dlopen at first use.
//
// libsmgpu.so must not carry a DT_NEEDED on libnccl: a host process may already hold a
// different NCCL build under the same SONAME (PyTorch bundles its own libnccl.so.2), and
// whichever copy is loaded first wins for the whole process.  Resolving lazily means a
// single-GPU user never loads NCCL at all, and a multi-GPU user shares the copy its host
// framework has already loaded.
#pragma once
#include <dlfcn.h>
#include <nccl.h>
#include <stdexcept>
#include <string>

namespace sm
{
struct NcclApi
{
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

inline const NcclApi &nccl()
{
    // function-local static: initialised once, thread-safe (several host threads may drive one GPU each)
    static const NcclApi loadedApi = [] {
        NcclApi api;
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h)
            h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h)
            throw std::runtime_error(std::string("cannot load libnccl.so.2: ") + dlerror());
        auto sym = [&](const char *name) {
            void *p = dlsym(h, name);
            if (!p)
                throw std::runtime_error(std::string("libnccl lacks symbol ") + name);
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.CommAbort = (decltype(api.CommAbort))sym("ncclCommAbort");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        return api;
    }();
    return loadedApi;
}
} // namespace sm
