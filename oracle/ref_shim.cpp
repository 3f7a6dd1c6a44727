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

// nms_3d.cu device helpers compiled for the host (operator_cxx/contrib/nms_3d.cu:29-378)
#define __device__
namespace ref_nms3d {
// CUDA's device-side min / max overloads for float are fminf / fmaxf (cuda math API); the host has no unqualified ones
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }
#include "nms3d_extract.h"
}
#undef __device__

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
