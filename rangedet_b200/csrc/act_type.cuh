// Storage type of activations / operands in the NHWC convolution pipeline: bf16 (default) or fp16.
//
// The reference trains this graph with fp16 storage and loss scale 128 (config/rangedet/rangedet_veh_wo_aug_4_18e.py:
// 35-36; casts at rangedet/symbol/backbone/dla_backbone.py:136-137, meta_kernel.py:193-196); bf16 is the cfg-4 inference
// format.  Both are 2-byte types fed to the same tcgen05.mma kind::f16 with fp32 accumulation, so every kernel that
// touches stored activations is compiled TWICE from the same source: once as is (bf16) and once with -DRD_ACT_F16
// (rangedet_b200/build.py), into distinctly named namespaces and C-ABI entry points
//   RD_ACT_FN(rd_conv2d_nhwc_, )  ->  rd_conv2d_nhwc_bf16   |  rd_conv2d_nhwc_f16
//   RD_ACT_NS(conv)               ->  conv_bf16             |  conv_f16
// Entry points that do not depend on the storage type are compiled only in the bf16 pass (#ifndef RD_ACT_F16).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define RD_CAT3_(a, b, c) a##b##c
#define RD_CAT3(a, b, c) RD_CAT3_(a, b, c)
#ifdef RD_ACT_F16
typedef __half act_t;
#define RD_ACT_TAG f16
#define RD_ACT_TMA_TYPE CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#define RD_ACT_MMA_FMT 0u   // tcgen05 instruction descriptor A/B format: 0 = f16
#else
typedef __nv_bfloat16 act_t;
#define RD_ACT_TAG bf16
#define RD_ACT_TMA_TYPE CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#define RD_ACT_MMA_FMT 1u   // 1 = bf16
#endif
#define RD_ACT_FN(prefix, suffix) RD_CAT3(prefix, RD_ACT_TAG, suffix)
#define RD_ACT_NS(name) RD_CAT3(name, _, RD_ACT_TAG)
#define RD_STR_(x) #x
#define RD_STR(x) RD_STR_(x)
#define RD_ACT_FN_STR(prefix, suffix) RD_STR(RD_ACT_FN(prefix, suffix))

namespace act {

// two floats -> one packed 32-bit word of two stored elements (a in the low half = lower address), round to nearest
__device__ __forceinline__ uint32_t pack2(float a, float b) {
#ifdef RD_ACT_F16
  __half2 v = __floats2half2_rn(a, b);
#else
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
#endif
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void unpack2(uint32_t w, float& lo, float& hi) {
#ifdef RD_ACT_F16
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
  lo = f.x;
  hi = f.y;
#else
  lo = __uint_as_float(w << 16);
  hi = __uint_as_float(w & 0xffff0000u);
#endif
}

__device__ __forceinline__ float to_float(act_t v) {
#ifdef RD_ACT_F16
  return __half2float(v);
#else
  return __bfloat162float(v);
#endif
}

__device__ __forceinline__ act_t from_float(float v) {
#ifdef RD_ACT_F16
  return __float2half_rn(v);
#else
  return __float2bfloat16_rn(v);
#endif
}

}  // namespace act
