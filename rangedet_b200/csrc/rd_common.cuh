// Shared host/device helpers for librangedet_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <mutex>
#include <utility>

namespace rd {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Returns 0 if ok; records the message otherwise.
int check_launch(const char* what);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Opt a kernel in to `bytes` of dynamic shared memory.  cudaFuncAttributeMaxDynamicSharedMemorySize is a
// per-DEVICE attribute: the largest value already granted is remembered per (kernel, device) under a lock,
// so a process that drives several GPUs (or threads) neither skips the opt-in on the second device nor
// re-issues the driver call on every launch.  Returns cudaSuccess or the driver's error.
template <typename F>
inline cudaError_t smem_optin(F* fn, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> granted;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  size_t& g = granted[std::make_pair(reinterpret_cast<const void*>(fn), dev)];
  if (bytes <= g) return cudaSuccess;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) g = bytes;
  return e;
}

}  // namespace rd

#define RD_REQUIRE(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      rd::set_error(__VA_ARGS__);        \
      return 1;                          \
    }                                    \
  } while (0)

#define RD_CUDA(call)                                                                  \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      rd::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                    __LINE__);                                                         \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)
