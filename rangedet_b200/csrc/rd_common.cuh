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
bool pdl_enabled();   // programmatic dependent launch on (default) unless RD_PDL=0
int conv_t_tile();       // 0: tile width chosen per shape; 160 / 192 / 224 / 256: fixed by rd_set_conv_t (tests, A/B)
bool conv_t_enabled();   // transposed-orientation kernel for 3x3 / stride 1 / Cout 128 convolutions (default) unless RD_CONV_T=0

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

// ---- Programmatic dependent launch (PDL) ---------------------------------------------------------------------
// The training step is ~950 dependent kernels of 5-100 us.  Launched with the programmatic-stream-serialization
// attribute, kernel N+1 may start as soon as every CTA of kernel N has called pdl_trigger() (or exited): its CTAs take
// the SMs that kernel N's CTAs leave, run their prologue (barrier init, TMEM allocation, descriptor prefetch) and
// then block in pdl_wait() until kernel N has completed and flushed -- launch latency and prologue hide behind the
// tail of the predecessor.  Rules every kernel launched through rd::launch follows: pdl_trigger() first thing,
// pdl_wait() by ALL threads before the first access to global memory that another kernel writes or reads.
// (A kernel completes only after its own wait returned, so completion of N implies completion of N-1: transitive.)
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif

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
