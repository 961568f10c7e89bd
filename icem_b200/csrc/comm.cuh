// NCCL plumbing for the per-iteration elite all-gather.  NCCL is resolved at run time
// with dlopen (the torch-bundled libnccl.so.2 when the host process already imported torch, else the
// system one), so the library itself links against nothing but the CUDA runtime.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdlib>
#include <stdexcept>
#include <string>

namespace icem {

struct CommError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct NcclApi {
  typedef struct { char internal[128]; } UniqueId;
  typedef void* CommT;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(CommT*, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(CommT) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, CommT, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  void* lib = nullptr;

  static NcclApi& get() {
    static NcclApi api;
    if (!api.lib) api.load();
    return api;
  }
  void load() {
    const char* override_path = std::getenv("ICEM_NCCL_LIB");
    const char* names[] = {override_path, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) throw CommError(std::string("cannot dlopen libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* s) {
      void* f = dlsym(lib, s);
      if (!f) throw CommError(std::string("missing NCCL symbol ") + s);
      return f;
    };
    GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
    AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
  }
  void check(int rc, const char* what) {
    if (rc != 0) throw CommError(std::string(what) + ": " + (GetErrorString ? GetErrorString(rc) : "nccl error"));
  }
};

struct Comm {
  NcclApi::CommT comm = nullptr;
  bool ready() const { return comm != nullptr; }
  static void unique_id(char out[128]) {
    NcclApi& n = NcclApi::get();
    NcclApi::UniqueId id;
    n.check(n.GetUniqueId(&id), "ncclGetUniqueId");
    for (int i = 0; i < 128; ++i) out[i] = id.internal[i];
  }
  void init(const char id_bytes[128], int world, int rank) {
    NcclApi& n = NcclApi::get();
    NcclApi::UniqueId id;
    for (int i = 0; i < 128; ++i) id.internal[i] = id_bytes[i];
    destroy();
    n.check(n.CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
  }
  // every rank contributes `bytes` bytes; recv holds world*bytes (rank-major)
  void all_gather(const void* send, void* recv, size_t bytes, cudaStream_t s) {
    if (!comm) throw CommError("icem_comm_init was not called");
    NcclApi& n = NcclApi::get();
    n.check(n.AllGather(send, recv, bytes, /*ncclInt8*/ 0, comm, s), "ncclAllGather");
  }
  void destroy() {
    if (comm) {
      NcclApi::get().CommDestroy(comm);
      comm = nullptr;
    }
  }
};

}  // namespace icem
