// NCCL, loaded at run time (dlopen "libnccl.so.2": the copy a host process already carries, e.g. PyTorch's, or the
// system one) so that the library has no link-time dependency and single-GPU users never touch it.  Only what the
// sampler needs: one communicator per chain and a sum all-reduce of doubles on the chain's stream
// ("per iteration one ncclAllReduce(SUM, double, n + n^2 (+1))").
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <string>

namespace pm {
namespace host {

struct NcclApi {
  typedef struct { char internal[128]; } UniqueId;  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
  typedef void* Comm;                               // ncclComm_t
  enum { kSum = 0, kDouble = 8 };                   // ncclSum, ncclFloat64
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  void* lib = nullptr;
  std::string why;

  static NcclApi& get() {
    static NcclApi a = load();
    return a;
  }
  bool ok() const { return lib != nullptr; }
  std::string error(int rc) const { return GetErrorString ? GetErrorString(rc) : "nccl error " + std::to_string(rc); }

 private:
  static NcclApi load() {
    NcclApi a;
    const char* names[] = {getenv("PHYLOMAP_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm || !nm[0]) continue;
      a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib) break;
      a.why = dlerror();
    }
    if (!a.lib) return a;
    bool all = true;
    auto sym = [&](const char* s) { void* p = dlsym(a.lib, s); if (!p) { all = false; a.why = std::string("missing symbol ") + s; } return p; };
    a.GetUniqueId = reinterpret_cast<int (*)(UniqueId*)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<int (*)(Comm*, int, UniqueId, int)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<int (*)(Comm)>(sym("ncclCommDestroy"));
    a.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, Comm, cudaStream_t)>(sym("ncclAllReduce"));
    a.GetErrorString = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
    if (!all) a.lib = nullptr;
    return a;
  }
};

}  // namespace host
}  // namespace pm
