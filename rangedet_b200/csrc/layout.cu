// Layout conversions at the op boundaries of the Meta-Kernel inside the NHWC-bf16 pipeline (sm_100a).
//
// The reference keeps every tensor NCHW (MXNet), so the Meta-Kernel op boundary is (B,C,H,W) fp32 /
// (B,9C,H,W) with channel index c*9+k (/root/reference rangedet/symbol/backbone/meta_kernel.py:232-239);
// the convolution pipeline here is zero-haloed NHWC bf16.  These two kernels move a tensor across that
// boundary in one HBM pass each, through a shared-memory tile so that both sides are accessed with
// full 128-byte lines (a strided permute copy ran at 0.25 TB/s and was 12 % of the training step).
//   chmap 0: dst channel = src channel
//   chmap 1: tap-major <-> reference order: NHWC channel k*(C/9)+c  <->  NCHW channel c*9+k
#include "../../include/rangedet_b200.h"
#include "act_type.cuh"
#include "rd_common.cuh"

namespace RD_ACT_NS(lay) {

constexpr int TP = 32;  // pixels per tile
constexpr int TC = 64;  // channels per tile

__device__ __forceinline__ int map_channel(int c_nhwc, int C, int chmap) {
  if (chmap == 0) return c_nhwc;
  const int cc = C / 9;
  return (c_nhwc % cc) * 9 + c_nhwc / cc;
}

// src: haloed NHWC bf16 [N][H+2][W+2][Cs]; dst: NCHW fp32 [N][C][H][W]; channels [0, C) converted.
// Tile = 64 pixels x 64 channels: 16-byte loads (8 channels of one pixel; a warp covers 4 pixels x 128 B),
// transposed through shared memory, 16-byte stores (4 pixels of one channel; 16 lanes cover 256 contiguous bytes).
constexpr int TP2 = 64;
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const act_t* __restrict__ src, float* __restrict__ dst,
                                                          int N, int H, int W, int Cs, int C, int chmap) {
  __shared__ float tile[TC][TP2 + 1];
  const int wt = blockIdx.x, h = blockIdx.y % H, n = blockIdx.y / H;
  const int w0 = wt * TP2;
  const int64_t srow = (((int64_t)n * (H + 2) + h + 1) * (W + 2) + 1) * Cs;
  const bool vec_ld = (Cs % 8) == 0;                 // 16-byte aligned pixel rows
  const bool vec_st = (W % 4) == 0;                  // 16-byte aligned channel rows
  for (int c0 = 0; c0 < C; c0 += TC) {
    for (int e = threadIdx.x; e < TP2 * (TC / 8); e += 256) {
      const int p = e / (TC / 8), ch = (e % (TC / 8)) * 8;
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = 0.f;
      if (w0 + p < W && c0 + ch < C) {
        const act_t* q = src + srow + (int64_t)(w0 + p) * Cs + c0 + ch;
        if (vec_ld && c0 + ch + 8 <= Cs) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(q));
          const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            act::unpack2(w4[i], f[2 * i], f[2 * i + 1]);
          }
        } else {
          for (int i = 0; i < 8 && c0 + ch + i < Cs; ++i) f[i] = act::to_float(q[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) tile[ch + i][p] = f[i];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < TC * (TP2 / 4); e += 256) {
      const int c = e / (TP2 / 4), p = (e % (TP2 / 4)) * 4;
      if (c0 + c < C && w0 + p < W) {
        const int cd = map_channel(c0 + c, C, chmap);
        float* o = dst + (((int64_t)n * C + cd) * H + h) * W + w0 + p;
        if (vec_st && w0 + p + 4 <= W) {
          *reinterpret_cast<float4*>(o) = make_float4(tile[c][p], tile[c][p + 1], tile[c][p + 2], tile[c][p + 3]);
        } else {
          for (int i = 0; i < 4 && w0 + p + i < W; ++i) o[i] = tile[c][p + i];
        }
      }
    }
    __syncthreads();
  }
}

// src: NCHW fp32 [N][C][H][W]; dst: haloed NHWC bf16 [N][H+2][W+2][Cd] (interior, channels [0, C) written)
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ src, act_t* __restrict__ dst,
                                                          int N, int H, int W, int C, int Cd, int chmap) {
  __shared__ float tile[TC][TP + 1];
  const int wt = blockIdx.x, h = blockIdx.y % H, n = blockIdx.y / H;
  const int w0 = wt * TP;
  const int64_t drow = (((int64_t)n * (H + 2) + h + 1) * (W + 2) + 1) * Cd;
  for (int c0 = 0; c0 < C; c0 += TC) {
    for (int e = threadIdx.x; e < TC * TP; e += 256) {
      const int c = e / TP, p = e % TP;
      float v = 0.f;
      if (w0 + p < W && c0 + c < C) {
        const int cs = map_channel(c0 + c, C, chmap);
        v = src[(((int64_t)n * C + cs) * H + h) * W + w0 + p];
      }
      tile[c][p] = v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < TP * (TC / 2); e += 256) {
      const int p = e / (TC / 2), cp = e % (TC / 2);
      if (w0 + p < W && c0 + 2 * cp < C) {
        const float a = tile[2 * cp][p];
        if (c0 + 2 * cp + 1 < C) {
          *reinterpret_cast<uint32_t*>(dst + drow + (int64_t)(w0 + p) * Cd + c0 + 2 * cp) = act::pack2(a, tile[2 * cp + 1][p]);
        } else {
          dst[drow + (int64_t)(w0 + p) * Cd + c0 + 2 * cp] = act::from_float(a);
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace lay_<storage type>
namespace lay = RD_ACT_NS(lay);

extern "C" {

int RD_ACT_FN(rd_nhwc_, _to_nchw_f32)(const void* src_pad, float* dst, int N, int H, int W, int C_src, int C, int chmap,
                             rd_stream_t stream) {
  RD_REQUIRE(src_pad && dst, RD_ACT_FN_STR(rd_nhwc_, _to_nchw_f32) ": null pointer");
  RD_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C <= C_src && C_src % 2 == 0, RD_ACT_FN_STR(rd_nhwc_, _to_nchw_f32) ": bad shape");
  RD_REQUIRE(chmap == 0 || (chmap == 1 && C % 9 == 0), RD_ACT_FN_STR(rd_nhwc_, _to_nchw_f32) ": chmap 1 needs C %% 9 == 0");
  RD_REQUIRE((int64_t)N * H <= 65535, RD_ACT_FN_STR(rd_nhwc_, _to_nchw_f32) ": N*H too large");
  if (rd_check_device()) return 1;
  dim3 grid((W + lay::TP2 - 1) / lay::TP2, N * H);
  lay::nhwc_to_nchw_kernel<<<grid, 256, 0, rd::as_stream(stream)>>>(static_cast<const act_t*>(src_pad), dst, N, H, W,
                                                                    C_src, C, chmap);
  rd::count_launch();
  return rd::check_launch(RD_ACT_FN_STR(rd_nhwc_, _to_nchw_f32) "");
}

int RD_ACT_FN(rd_nchw_f32_to_nhwc_, )(const float* src, void* dst_pad, int N, int H, int W, int C, int C_dst, int chmap,
                             rd_stream_t stream) {
  RD_REQUIRE(src && dst_pad, RD_ACT_FN_STR(rd_nchw_f32_to_nhwc_, ) ": null pointer");
  RD_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C <= C_dst && C_dst % 2 == 0, RD_ACT_FN_STR(rd_nchw_f32_to_nhwc_, ) ": bad shape");
  RD_REQUIRE(chmap == 0 || (chmap == 1 && C % 9 == 0), RD_ACT_FN_STR(rd_nchw_f32_to_nhwc_, ) ": chmap 1 needs C %% 9 == 0");
  RD_REQUIRE((int64_t)N * H <= 65535, RD_ACT_FN_STR(rd_nchw_f32_to_nhwc_, ) ": N*H too large");
  if (rd_check_device()) return 1;
  dim3 grid((W + lay::TP - 1) / lay::TP, N * H);
  lay::nchw_to_nhwc_kernel<<<grid, 256, 0, rd::as_stream(stream)>>>(src, static_cast<act_t*>(dst_pad), N, H, W, C,
                                                                    C_dst, chmap);
  rd::count_launch();
  return rd::check_launch(RD_ACT_FN_STR(rd_nchw_f32_to_nhwc_, ) "");
}

#ifndef RD_ACT_F16
// dst[p][dst_off + c] = src[p][src_off + c], c < nchan, for every pixel p of two pixel-major tensors: 16-byte chunks,
// a warp covers consecutive chunks of consecutive pixels (channel concatenation without a strided framework copy)
namespace lay_any {
__global__ void __launch_bounds__(256) copy_channels_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                                            int64_t npix, int src_q, int src_off_q, int dst_q,
                                                            int dst_off_q, int nq) {
  const int64_t total = npix * nq;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t p = i / nq;
    const int q = (int)(i - p * nq);
    dst[p * dst_q + dst_off_q + q] = __ldg(src + p * src_q + src_off_q + q);
  }
}
}  // namespace lay_any

int rd_copy_channels_16b(const void* src, int src_ctotal, int src_off, void* dst, int dst_ctotal, int dst_off, int nchan,
                         int64_t npix, rd_stream_t stream) {
  RD_REQUIRE(src && dst, "rd_copy_channels_16b: null pointer");
  RD_REQUIRE(npix > 0 && nchan > 0 && nchan % 8 == 0 && src_ctotal % 8 == 0 && dst_ctotal % 8 == 0 && src_off % 8 == 0 &&
                 dst_off % 8 == 0 && src_off >= 0 && dst_off >= 0 && src_off + nchan <= src_ctotal &&
                 dst_off + nchan <= dst_ctotal,
             "rd_copy_channels_16b: channel counts and offsets must be multiples of 8 (2-byte elements) and in range");
  if (rd_check_device()) return 1;
  const int64_t total = npix * (nchan / 8);
  const int64_t want = (total + 255) / 256;
  const int grid = (int)(want < 148 * 16 ? want : 148 * 16);
  lay_any::copy_channels_kernel<<<grid, 256, 0, rd::as_stream(stream)>>>(
      static_cast<const uint4*>(src), static_cast<uint4*>(dst), npix, src_ctotal / 8, src_off / 8, dst_ctotal / 8,
      dst_off / 8, nchan / 8);
  rd::count_launch();
  return rd::check_launch("rd_copy_channels_16b");
}
#endif

}  // extern "C"
