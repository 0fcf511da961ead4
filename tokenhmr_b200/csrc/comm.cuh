// The one exchange step of the sharded path: an in-place all-gather of the per-image outputs over
// NVLink / NVSwitch.  Every output field lives in ONE buffer of world * B rows; the engine of rank r writes its
// images straight into rows [r*B, (r+1)*B) (its thmr_outputs pointers point there), and thmr_allgather_outputs()
// issues one grouped ncclAllGather per field with sendbuff == recvbuff + rank * count (NCCL's in-place form): no
// pack / unpack copies, one NCCL kernel, stream-ordered and capturable in the same CUDA graph as the forward.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- inside a PyTorch process that is the copy torch already
// loaded), so libtokenhmr_b200.so has no link-time dependency on it and single-GPU users never need it.
#pragma once
#include <dlfcn.h>

#include "common.cuh"

namespace thmr {

// Minimal NCCL surface (nccl.h: ncclUniqueId is 128 opaque bytes, ncclFloat32 = 7, ncclSuccess = 0).
struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* NcclComm;
struct NcclApi {
  void* handle = nullptr;
  int (*GetVersion)(int*) = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
constexpr int kNcclFloat32 = 7;

inline NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char* names[] = {getenv("THMR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  }
  if (!h) return nullptr;
  bool ok = true;
  auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p) ok = false; return p; };
  api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  if (!ok) { dlclose(h); return nullptr; }
  api.handle = h;
  return &api;
}

#define THMR_NCCL(api, expr)                                                                            \
  do {                                                                                                  \
    int _r = (expr);                                                                                    \
    if (_r != 0) return ::thmr::fail(THMR_ERR_CUDA, "%s failed: %s", #expr, (api)->GetErrorString(_r)); \
  } while (0)

}  // namespace thmr

struct thmr_comm {
  thmr::NcclComm comm = nullptr;
  int nranks = 0, rank = 0, device = 0;
};
