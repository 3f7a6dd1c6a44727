// Library-level entry points of librangedet_b200.so (see include/rangedet_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/rangedet_b200.h"
#include "rd_common.cuh"

namespace rd {

static thread_local char g_err[512] = "";
static thread_local uint64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches += (uint64_t)n; }

static int g_pdl = -1;   // -1: not decided yet (RD_PDL environment variable, default on)

bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("RD_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

static int g_conv_t = -1;

bool conv_t_enabled() {
  if (g_conv_t < 0) {
    const char* e = getenv("RD_CONV_T");
    g_conv_t = (e && e[0] == '0') ? 0 : 1;
  }
  return g_conv_t != 0;
}

int conv_t_tile() { return conv_t_enabled() && g_conv_t >= 128 ? g_conv_t : 0; }

}  // namespace rd

extern "C" {

int rd_version(void) { return 1; }

const char* rd_last_error(void) { return rd::g_err; }

uint64_t rd_launch_count(void) { return rd::g_launches; }

int rd_set_conv_t(int on) {
  const int prev = rd::conv_t_enabled() ? rd::g_conv_t : 0;
  rd::g_conv_t = (on == 160 || on == 192 || on == 224 || on == 256) ? on : (on ? 1 : 0);
  return prev;
}

int rd_set_pdl(int on) {
  const int prev = rd::pdl_enabled() ? 1 : 0;
  rd::g_pdl = on ? 1 : 0;
  return prev;
}

int rd_check_device(void) {
  // Result cached per device: cudaGetDeviceProperties costs ~1 ms and this guards every launch.
  static int cached[64];  // 0 = unknown, 1 = ok
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    rd::set_error("no CUDA device: %s (librangedet_b200 has no CPU fallback)", cudaGetErrorString(e));
    cudaGetLastError();
    return 1;
  }
  if (dev >= 0 && dev < 64 && cached[dev] == 1) return 0;
  int major = 0, minor = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (e != cudaSuccess) {
    rd::set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return 1;
  }
  if (major != 10) {
    rd::set_error("device %d is sm_%d%d; this library is built for sm_100a only", dev, major, minor);
    return 1;
  }
  if (dev >= 0 && dev < 64) cached[dev] = 1;
  return 0;
}

}  // extern "C"
