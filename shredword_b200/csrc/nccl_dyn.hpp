// nccl_dyn.hpp -- NCCL through dlopen, so that libtrainer.so has no link-time dependency on it and, inside a
// torch.distributed process, binds to the very libnccl.so.2 that torch already loaded (two NCCL copies in
// one process is asking for trouble). Only what the per-merge exchange needs: comm init and all-gather.
#pragma once

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <string>

#include "device_util.cuh"

namespace swb {

struct NcclUniqueId { char internal[128]; };  // == ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES 128)
typedef struct ncclComm *NcclComm;

class NcclApi {
 public:
  static NcclApi &get() { static NcclApi a; return a; }
  bool ok() const { return handle_ != nullptr; }
  const std::string &error() const { return err_; }

  int (*GetUniqueId)(NcclUniqueId *) = nullptr;
  int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int /*ncclDataType_t*/, NcclComm, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  static constexpr int kChar = 0;  // ncclInt8 / ncclChar

  void check(int rc, const char *what) const {
    if (rc != 0) throw Error(std::string("NCCL error in ") + what + ": " + (GetErrorString ? GetErrorString(rc) : "?"));
  }

 private:
  NcclApi() {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) { handle_ = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (handle_) break; }  // torch's copy, if loaded
    if (!handle_) for (const char *n : names) { handle_ = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (handle_) break; }
    if (!handle_) { err_ = std::string("libnccl.so.2 not found: ") + dlerror(); return; }
    auto sym = [&](const char *s) { void *p = dlsym(handle_, s); if (!p) err_ = std::string("NCCL symbol missing: ") + s; return p; };
    GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
    AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
    if (!err_.empty()) handle_ = nullptr;
  }
  void *handle_ = nullptr;
  std::string err_;
};

}  // namespace swb
