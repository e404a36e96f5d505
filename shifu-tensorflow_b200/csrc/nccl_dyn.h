// NCCL is bound at run time (dlopen) so that (a) the C-ABI library loads on a host without NCCL /
// without a GPU, and (b) when the host process already carries an NCCL (e.g. PyTorch's bundled
// libnccl.so.2) the same copy is reused instead of a second one being mapped.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

namespace sb {

struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* NcclComm;

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*);
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int);
  int (*CommDestroy)(NcclComm);
  int (*AllReduce)(const void*, void*, size_t, int /*dtype*/, int /*op*/, NcclComm, cudaStream_t);
  int (*Broadcast)(const void*, void*, size_t, int /*dtype*/, int /*root*/, NcclComm, cudaStream_t);
  int (*CommGetAsyncError)(NcclComm, int*);
  const char* (*GetErrorString)(int);
  int (*GetVersion)(int*);
  bool ok;
};

enum { NCCL_INT64 = 4, NCCL_FLOAT32 = 7, NCCL_SUM = 0 };

inline NcclApi* nccl_api() {
  static NcclApi api = {};
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return nullptr;
  api.GetUniqueId = reinterpret_cast<int (*)(NcclUniqueId*)>(dlsym(h, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<int (*)(NcclComm*, int, NcclUniqueId, int)>(dlsym(h, "ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<int (*)(NcclComm)>(dlsym(h, "ncclCommDestroy"));
  api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t)>(dlsym(h, "ncclAllReduce"));
  api.Broadcast = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t)>(dlsym(h, "ncclBroadcast"));
  api.CommGetAsyncError = reinterpret_cast<int (*)(NcclComm, int*)>(dlsym(h, "ncclCommGetAsyncError"));
  api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
  api.GetVersion = reinterpret_cast<int (*)(int*)>(dlsym(h, "ncclGetVersion"));
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.Broadcast && api.GetErrorString;
  return api.ok ? &api : nullptr;
}

}  // namespace sb
