// sde_internal.h -- host-side helpers shared by the translation units of libsimplediffeq_cuda
// (hidden visibility: not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/simplediffeq_cuda.h"

namespace sde_host {

constexpr int kBlock = 128;   // threads per CTA of every kernel (one trajectory per thread)

extern thread_local std::string g_err;          // sde_last_error()
extern std::atomic<long long> g_launches;       // sde_launch_count()

// records the message for sde_last_error() and returns `code`
int fail(int code, const char* fmt, ...);

inline size_t esize(int dtype) { return dtype == SDE_F64 ? 8 : 4; }

// the per-device stream-ordered memory pool (sde_api.cu) and its retention cap
int device_pool(int dev, cudaMemPool_t* out);
size_t pool_keep_bytes();

// dynamic shared memory of the staged series writer (sde::StageCfg::kBytesPerWarp per warp); `user`: NVRTC system
// (honours the development knob SDE_TUNE_STAGE_ELEMS)
size_t staged_smem_bytes(int n_state, size_t es, int block, bool user);

// NVRTC: compile `program` (device headers are embedded in the library) to a cubin for sm_100a
int nvrtc_compile(const std::string& program, std::vector<char>* cubin, std::string* log, bool fmad = false);

}  // namespace sde_host

#define SDE_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return sde_host::fail(SDE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
