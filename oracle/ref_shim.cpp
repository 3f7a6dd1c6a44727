// C-ABI shim around the REFERENCE's own functors / functions, compiled by
// oracle/build_ref.py into oracle/_ref/librd_ref.so.   TEST INFRASTRUCTURE ONLY.
//
// The three *_extract.h headers are derived at build time from the sources under
// /root/reference (see build_ref.py); they are never committed.  This file only
// supplies what the MXNet / pybind11 drivers around them would have supplied:
//   * mxnet_op::Kernel<OP,cpu>::Launch(s, N, args...) is `for i<N: OP::Map(i,args...)`
//     (reference call sites: decode_3d_bbox-inl.h:297-303, rotated_iou-inl.h:543-545)
//   * point4_wnms_4c (nms.h:781-794): copy, iota, sort by score desc, wnms_4c.
#include <math.h>
#include <cmath>
#include <cassert>
#include <cstring>
#include <algorithm>
#include <numeric>
#include <tuple>
#include <vector>

#define MSHADOW_XINLINE inline

namespace ref_decode {
#include "decode_extract.h"
}
#undef MACRO_MAX
#undef MACRO_MIN
namespace ref_riou {
#include "riou_extract.h"
}
#undef MACRO_MAX
#undef MACRO_MIN

// nms_3d.cu device helpers (operator_cxx/contrib/nms_3d.cu:29-378) and its two kernels (:380-468) compiled for the
// host.  Block emulation for nms_kernel_3d: blockIdx / threadIdx are globals, __shared__ is static storage and
// __syncthreads() is empty; each block is executed TWICE thread by thread -- the first sweep fills the shared tile,
// the second recomputes every mask word from the complete tile (the kernel's outputs are pure functions of its
// inputs and the tile, so the second sweep's writes are the kernel's result).
#define __device__
#define __global__
#define __shared__ static
#define __syncthreads()
#define DIVUP(m, n) ((m) / (n) + ((m) % (n) > 0))
namespace ref_nms3d {
struct Idx3 { int x, y, z; };
static Idx3 threadIdx, blockIdx;
const int THREADS_PER_BLOCK_NMS = sizeof(unsigned long long) * 8;   // nms_3d.cu:28
// CUDA's device-side min / max overloads for float are fminf / fmaxf (cuda math API); the host has no unqualified ones
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }
#include "nms3d_extract.h"
}
#undef __device__
#undef __global__
#undef __shared__
#undef __syncthreads

#include "nms_extract.h"

extern "C" {

// Decode3DBboxForward<cpu> restated: Fill(out,0) then Launch over B*N points.
void ref_decode_3d_bbox(const float* delta, const float* pc, float* out, long n_total,
                        int box_type, int is_bin) {
  std::memset(out, 0, sizeof(float) * 10 * n_total);
  if (is_bin) {
    for (long i = 0; i < n_total; ++i)
      ref_decode::Decode3DBboxBinKernelGPU::Map<float>((int)i, delta, pc, out, box_type);
  } else {
    for (long i = 0; i < n_total; ++i)
      ref_decode::Decode3DBboxKernelGPU::Map<float>((int)i, delta, pc, out, box_type);
  }
}

// RotatedIOUForward<cpu> restated: Fill(out,-1) then Launch over N1*N2 pairs.
void ref_rotated_iou(const float* b1, const float* b2, float* out, long n1, long n2,
                     int box_type, int use_omp) {
  const long total = n1 * n2;
  for (long i = 0; i < total; ++i) out[i] = -1.f;
  if (use_omp) {
#pragma omp parallel for schedule(static, 4096)
    for (long i = 0; i < total; ++i)
      ref_riou::RotateIoUKernelGPU::Map<float>((int)i, (int)n1, (int)n2, b1, b2, out, box_type);
  } else {
    for (long i = 0; i < total; ++i)
      ref_riou::RotateIoUKernelGPU::Map<float>((int)i, (int)n1, (int)n2, b1, b2, out, box_type);
  }
}

// iou_bev (volumetric, :342-368) / iou_normal (:370-378) of two 10-dim boxes, as nms_kernel_3d calls them (:420-426)
float ref_nms3d_iou(const float* box_a, const float* box_b, int normal_iou) {
  return normal_iou ? ref_nms3d::iou_normal(box_a, box_b) : ref_nms3d::iou_bev(box_a, box_b);
}

// NMS3DForward<gpu> (:470-534) restated around the reference's two kernels: per image the N x N/64 bitmask from
// nms_kernel_3d over a (N/64, N/64) grid of 64-thread blocks, then prepare_output_kernel_3d (one thread).
void ref_nms3d_kernels(const float* boxes, int B, int N, float thr, int max_keep, int normal_iou, int* keep_idx, float* out) {
  using namespace ref_nms3d;
  for (long i = 0; i < (long)B * max_keep; ++i) keep_idx[i] = -1;            // Fill(out_data[0], -1)  :497
  std::memset(out, 0, sizeof(float) * (size_t)B * max_keep * 10);            // Fill(out_data[1], 0)   :498
  const int col_blocks = DIVUP(N, THREADS_PER_BLOCK_NMS);
  std::vector<unsigned long long> mask((size_t)N * col_blocks), remv(col_blocks);
  for (int b = 0; b < B; ++b) {
    const float* bx = boxes + (long)b * N * 10;
    for (int by = 0; by < col_blocks; ++by)
      for (int bxi = 0; bxi < col_blocks; ++bxi) {
        ref_nms3d::blockIdx.x = bxi;
        ref_nms3d::blockIdx.y = by;
        for (int sweep = 0; sweep < 2; ++sweep)
          for (int t = 0; t < THREADS_PER_BLOCK_NMS; ++t) {
            ref_nms3d::threadIdx.x = t;
            nms_kernel_3d(N, thr, bx, mask.data(), normal_iou != 0);
          }
      }
    std::fill(remv.begin(), remv.end(), 0ULL);                                // cudaMemset(remv_dev, 0)  :519
    prepare_output_kernel_3d(N, max_keep, col_blocks, mask.data(), remv.data(), keep_idx + (long)b * max_keep, bx,
                             out + (long)b * max_keep * 10);
  }
}

float ref_single_overlap(const float* box1, const float* box2, int is3d) {
  trtplus::OverlapChecker oc;
  return oc.single_overlap(box1, box2, is3d != 0);
}

// point4_wnms_4c<float> restated without pybind11 arrays.  Returns K.
int ref_wnms_4c(const float* dets, int n, float thresh, float thresh_vote, int is3d,
                int hash_scale, float* out_dets, int* keep_inds) {
  std::vector<float> dets_v(dets, dets + (size_t)n * 12);
  const int dets_ndim = 12;
  std::vector<int> orders_v(dets_v.size() / dets_ndim);
  std::iota(orders_v.begin(), orders_v.end(), 0);
  std::sort(orders_v.begin(), orders_v.end(), [&](int i, int j) {
    return dets_v[i * dets_ndim + 11] > dets_v[j * dets_ndim + 11];
  });
  auto res = trtplus::wnms_4c<float>(dets_v, orders_v, thresh, thresh_vote, is3d != 0, hash_scale);
  const std::vector<float>& kd = std::get<0>(res);
  const std::vector<int>& ki = std::get<1>(res);
  std::copy(kd.begin(), kd.end(), out_dets);
  std::copy(ki.begin(), ki.end(), keep_inds);
  return (int)ki.size();
}

}  // extern "C"
