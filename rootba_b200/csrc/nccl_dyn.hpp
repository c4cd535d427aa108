// NCCL bound at run time (dlopen) so the library has no link-time NCCL dependency: inside a Python
// process we reuse the libnccl.so.2 that torch already loaded, in a plain C++ host we load the system one.
#pragma once

#include <dlfcn.h>
#include <nccl.h>

namespace rba {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
};

inline NcclApi* nccl_api() {
  static NcclApi api;
  static int state = 0;  // 0 untried, 1 ok, -1 failed
  if (state == 0) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW);
    if (!h) { state = -1; return nullptr; }
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
    state = (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString) ? 1 : -1;
  }
  return state == 1 ? &api : nullptr;
}

}  // namespace rba
