// Parameter plumbing of the training step (sm_100a): everything that touches the 9.1 M parameters once per
// step is ONE launch over a flat buffer instead of ~2 800 per-tensor launches.
//
//   rd_gather_f32_to_bf16   all bf16 operand copies of the conv weights (forward layout, transposed / flipped
//                           data-gradient layout, phase-grouped deconv layouts) from the flat fp32 masters
//   rd_gather_f32           all parameter gradients from the kernels' native output layouts ([tap][Cout'][Cin']
//                           weight-gradient tiles, per-channel BN sums ...) into the flat gradient buffer that is
//                           all-reduced (reference: hvd.DistributedOptimizer, tools/train.py:364-368)
//   rd_sgd_mom_update       MXNet `sgd_mom_update` with multi_precision masters (tools/train.py:306-319, :359-361):
//                             g   = clip(rescale_grad * grad, clip_gradient) + wd_i * w
//                             mom = momentum * mom - lr * g ;  w += mom
//                           wd_i = wd * wd_mult (0 for *_bias / *_beta, Optimizer.set_wd_mult); lr / momentum /
//                           rescale / clip are read from DEVICE memory so a captured CUDA graph follows the
//                           learning-rate schedule.
// All three are HBM-bound streaming kernels (index maps are int32, read once).
#include "../../include/rangedet_b200.h"
#include "act_type.cuh"
#include "rd_common.cuh"

namespace {

constexpr int OPT_THREADS = 256;

__global__ void __launch_bounds__(OPT_THREADS)
gather_act_kernel(const float* __restrict__ src, const int* __restrict__ idx, act_t* __restrict__ dst,
                   int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * OPT_THREADS * 2;
  for (int64_t i = ((int64_t)blockIdx.x * OPT_THREADS + threadIdx.x) * 2; i < n; i += stride) {
    const int i0 = __ldg(idx + i);
    const int i1 = i + 1 < n ? __ldg(idx + i + 1) : -1;
    const float v0 = i0 >= 0 ? __ldg(src + i0) : 0.f;
    const float v1 = i1 >= 0 ? __ldg(src + i1) : 0.f;
    if (i + 1 < n) {
      *reinterpret_cast<uint32_t*>(dst + i) = act::pack2(v0, v1);   // n even or tail handled below
    } else {
      dst[i] = act::from_float(v0);
    }
  }
}

__global__ void __launch_bounds__(OPT_THREADS)
gather_f32_kernel(const float* __restrict__ src, const int* __restrict__ idx, float* __restrict__ dst, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * OPT_THREADS;
  for (int64_t i = (int64_t)blockIdx.x * OPT_THREADS + threadIdx.x; i < n; i += stride) {
    const int j = __ldg(idx + i);
    dst[i] = j >= 0 ? __ldg(src + j) : 0.f;
  }
}

__global__ void __launch_bounds__(OPT_THREADS)
sgd_mom_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
               const float* __restrict__ wd, const float* __restrict__ hyper, int64_t n) {
  const float lr = __ldg(hyper), momentum = __ldg(hyper + 1), rescale = __ldg(hyper + 2), clip = __ldg(hyper + 3);
  const int64_t stride = (int64_t)gridDim.x * OPT_THREADS;
  for (int64_t i = (int64_t)blockIdx.x * OPT_THREADS + threadIdx.x; i < n; i += stride) {
    float gi = rescale * g[i];
    if (clip > 0.f) gi = fminf(fmaxf(gi, -clip), clip);   // element-wise clipping (MXNet clip_gradient)
    const float wi = w[i];
    gi += __ldg(wd + i) * wi;
    const float mi = momentum * m[i] - lr * gi;
    m[i] = mi;
    w[i] = wi + mi;
  }
}

inline unsigned grid_for(int64_t n, int per_thread) {
  const int64_t b = (n + (int64_t)OPT_THREADS * per_thread - 1) / ((int64_t)OPT_THREADS * per_thread);
  const int64_t cap = 148 * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

int RD_ACT_FN(rd_gather_f32_to_, )(const float* src, const int* idx, void* dst, int64_t n, rd_stream_t stream) {
  RD_REQUIRE(n >= 0, RD_ACT_FN_STR(rd_gather_f32_to_, ) ": negative n");
  if (n == 0) return 0;
  RD_REQUIRE(src && idx && dst, RD_ACT_FN_STR(rd_gather_f32_to_, ) ": null pointer");
  RD_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 3) == 0, RD_ACT_FN_STR(rd_gather_f32_to_, ) ": dst must be 4-byte aligned");
  if (rd_check_device()) return 1;
  gather_act_kernel<<<grid_for(n, 2), OPT_THREADS, 0, rd::as_stream(stream)>>>(src, idx, static_cast<act_t*>(dst), n);
  rd::count_launch();
  return rd::check_launch(RD_ACT_FN_STR(rd_gather_f32_to_, ) "");
}

#ifndef RD_ACT_F16   // storage-type independent entry points: compiled once, in the bf16 pass
int rd_gather_f32(const float* src, const int* idx, float* dst, int64_t n, rd_stream_t stream) {
  RD_REQUIRE(n >= 0, "rd_gather_f32: negative n");
  if (n == 0) return 0;
  RD_REQUIRE(src && idx && dst, "rd_gather_f32: null pointer");
  if (rd_check_device()) return 1;
  gather_f32_kernel<<<grid_for(n, 1), OPT_THREADS, 0, rd::as_stream(stream)>>>(src, idx, dst, n);
  rd::count_launch();
  return rd::check_launch("rd_gather_f32");
}

int rd_sgd_mom_update(float* weight, const float* grad, float* mom, const float* wd, const float* hyper, int64_t n,
                      rd_stream_t stream) {
  RD_REQUIRE(n >= 0, "rd_sgd_mom_update: negative n");
  if (n == 0) return 0;
  RD_REQUIRE(weight && grad && mom && wd && hyper, "rd_sgd_mom_update: null pointer");
  if (rd_check_device()) return 1;
  sgd_mom_kernel<<<grid_for(n, 1), OPT_THREADS, 0, rd::as_stream(stream)>>>(weight, grad, mom, wd, hyper, n);
  rd::count_launch();
  return rd::check_launch("rd_sgd_mom_update");
}
#endif  // RD_ACT_F16

}  // extern "C"
