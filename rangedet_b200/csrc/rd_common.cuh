// Shared host/device helpers for librangedet_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace rd {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Returns 0 if ok; records the message otherwise.
int check_launch(const char* what);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

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
