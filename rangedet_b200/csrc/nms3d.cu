// NMS3D (hard NMS on 10-dim corner boxes with volumetric rotated IoU) for sm_100a.
//
// Replaces _contrib_NMS3D: NMS3DForward<gpu>, /root/reference operator_cxx/contrib/nms_3d.cu:470-534
// (nms_kernel_3d :380-434 builds an N x N/64 bitmask -- 313 MB cudaMalloc'ed per call at N = 50 000 --
// and prepare_output_kernel_3d :436-468 scans it with ONE GPU thread).  Same result, different
// structure: one CTA per image walks the kept boxes in score order; for each, all threads test the
// later unsuppressed boxes (expanded-AABB rejection, then the reference's polygon clipping) and set
// suppression bits in a shared-memory bitmap.  Only rows of kept boxes are evaluated, no N^2 mask.
// The AABB expansion covers check_in_box3d_anotherway's -1e-2 margin (:93-150): a corner may count as
// inside up to 0.01 / (|edge|^2 |next edge|) outside the box, so disjoint expanded boxes have IoU 0.
#include <math.h>

#include "../../include/rangedet_b200.h"
#include "rd_common.cuh"

namespace n3 {

constexpr int NT = 1024;
constexpr float EPS3 = 1e-8f;
struct Pt {
  float x, y;
};
__device__ __forceinline__ float cross2(Pt a, Pt b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ float cross3(Pt p1, Pt p2, Pt p0) {
  return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}
__device__ __forceinline__ bool rect_cross(Pt p1, Pt p2, Pt q1, Pt q2) {
  return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
         fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) && fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}
__device__ bool in_box(const float* box, Pt P) {  // check_in_box3d_anotherway :93-150
  const float MARGIN = -1e-2f;
  const Pt A = {box[0], box[1]}, B = {box[2], box[3]}, C = {box[4], box[5]}, D = {box[6], box[7]};
  const Pt AB = {B.x - A.x, B.y - A.y}, BC = {C.x - B.x, C.y - B.y}, CD = {D.x - C.x, D.y - C.y},
           DA = {A.x - D.x, A.y - D.y};
  const float cw = cross2(AB, BC);
  const Pt PA = {A.x - P.x, A.y - P.y};
  if (cross2(PA, AB) * cw < MARGIN) return false;
  const Pt PB = {B.x - P.x, B.y - P.y};
  if (cross2(PB, BC) * cw < MARGIN) return false;
  const Pt PC = {C.x - P.x, C.y - P.y};
  if (cross2(PC, CD) * cw < MARGIN) return false;
  const Pt PD = {D.x - P.x, D.y - P.y};
  if (cross2(PD, DA) * cw < MARGIN) return false;
  return true;
}
__device__ bool isect(Pt p1, Pt p0, Pt q1, Pt q0, Pt* ans) {  // intersection :152-181
  if (!rect_cross(p0, p1, q0, q1)) return false;
  const float s1 = cross3(q0, p1, p0), s2 = cross3(p1, q1, p0), s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
  if (!(s1 * s2 > 0 && s3 * s4 > 0)) return false;
  const float s5 = cross3(q1, p1, p0);
  if (fabsf(s5 - s1) > EPS3) {
    ans->x = __fdiv_rn(s5 * q0.x - s1 * q1.x, s5 - s1);
    ans->y = __fdiv_rn(s5 * q0.y - s1 * q1.y, s5 - s1);
  } else {
    const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
    const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
    const float Dd = a0 * b1 - a1 * b0;
    ans->x = __fdiv_rn(b0 * c1 - b1 * c0, Dd);
    ans->y = __fdiv_rn(a1 * c0 - a0 * c1, Dd);
  }
  return true;
}
__device__ __forceinline__ float area_of(const float* b) {  // get_area :193-198
  const float e1 = (b[0] - b[2]) * (b[0] - b[2]) + (b[1] - b[3]) * (b[1] - b[3]);
  const float e2 = (b[4] - b[2]) * (b[4] - b[2]) + (b[5] - b[3]) * (b[5] - b[3]);
  return sqrtf(e1 * e2);
}
__device__ float overlap(const float* a, const float* b) {  // box_overlap :218-336
  Pt ca[4], cb[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    ca[k] = {a[2 * k], a[2 * k + 1]};
    cb[k] = {b[2 * k], b[2 * k + 1]};
  }
  float px[24], py[24];
  float sx = 0.f, sy = 0.f;
  int cnt = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      Pt x;
      if (isect(ca[(i + 1) & 3], ca[i], cb[(j + 1) & 3], cb[j], &x)) {
        sx += x.x; sy += x.y;
        if (cnt < 24) { px[cnt] = x.x; py[cnt] = x.y; }
        cnt++;
      }
    }
  for (int k = 0; k < 4; ++k) {
    if (in_box(a, cb[k])) { sx += cb[k].x; sy += cb[k].y; if (cnt < 24) { px[cnt] = cb[k].x; py[cnt] = cb[k].y; } cnt++; }
    if (in_box(b, ca[k])) { sx += ca[k].x; sy += ca[k].y; if (cnt < 24) { px[cnt] = ca[k].x; py[cnt] = ca[k].y; } cnt++; }
  }
  const float cx = __fdiv_rn(sx, (float)cnt), cy = __fdiv_rn(sy, (float)cnt);
  const int n = cnt < 24 ? cnt : 24;
  float ang[24];
  for (int i = 0; i < n; ++i) ang[i] = atan2f(py[i] - cy, px[i] - cx);
  for (int j = 0; j < n - 1; ++j)
    for (int i = 0; i < n - j - 1; ++i)
      if (ang[i] > ang[i + 1]) {
        float t = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = t;
        t = px[i]; px[i] = px[i + 1]; px[i + 1] = t;
        t = py[i]; py[i] = py[i + 1]; py[i + 1] = t;
      }
  float area = 0.f;
  for (int k = 0; k < n - 1; ++k)
    area += (px[k] - px[0]) * (py[k + 1] - py[0]) - (py[k] - py[0]) * (px[k + 1] - px[0]);
  return fabsf(area) / 2.0f;
}
__device__ float iou_bev3d(const float* a, const float* b) {  // iou_bev :342-368
  const float ha = a[9] - a[8], hb = b[9] - b[8];
  float oh = fminf(a[9], b[9]) - fmaxf(a[8], b[8]);
  if (oh < 0.f) oh = 0.f;
  const float va = area_of(a) * ha, vb = area_of(b) * hb;
  const float vo = overlap(a, b) * oh;
  return __fdiv_rn(vo, fmaxf(va + vb - vo, EPS3));
}
__device__ float iou_normal(const float* a, const float* b) {  // :370-378
  const float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  const float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  const float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
  const float inter = w * h;
  const float Sa = (a[2] - a[0]) * (a[3] - a[1]), Sb = (b[2] - b[0]) * (b[3] - b[1]);
  return __fdiv_rn(inter, fmaxf(Sa + Sb - inter, EPS3));
}

// expanded AABB per box (see header); non-finite or degenerate boxes get an infinite box (never rejected)
__global__ void prep_kernel(const float* __restrict__ boxes, int64_t total, float4* __restrict__ aabb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float* b = boxes + i * 10;
  float x0 = b[0], x1 = b[0], y0 = b[1], y1 = b[1];
  bool fin = true;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    x0 = fminf(x0, b[2 * k]); x1 = fmaxf(x1, b[2 * k]);
    y0 = fminf(y0, b[2 * k + 1]); y1 = fmaxf(y1, b[2 * k + 1]);
    fin = fin && isfinite(b[2 * k]) && isfinite(b[2 * k + 1]);
  }
  const float e1 = (b[2] - b[0]) * (b[2] - b[0]) + (b[3] - b[1]) * (b[3] - b[1]);
  const float e2 = (b[4] - b[2]) * (b[4] - b[2]) + (b[5] - b[3]) * (b[5] - b[3]);
  const float l1 = sqrtf(e1), l2 = sqrtf(e2);
  // tolerance = 0.01 / (|edge|^2 |other edge|) for either edge family, doubled for rounding
  const float m = 2.2e-2f / fmaxf(fminf(e1 * l2, e2 * l1), 1e-30f);
  if (!fin || !(m < 1e3f)) aabb[i] = make_float4(-INFINITY, -INFINITY, INFINITY, INFINITY);
  else aabb[i] = make_float4(x0 - m, y0 - m, x1 + m, y1 + m);
}

__global__ void __launch_bounds__(NT, 1)
nms3d_kernel(const float* __restrict__ boxes, const float4* __restrict__ aabb, int N, float thr, int max_keep,
             int normal_iou, int* __restrict__ keep_idx, float* __restrict__ boxes_out) {
  extern __shared__ unsigned bitmap[];
  __shared__ float cur_box[10];
  __shared__ float4 cur_aabb;
  __shared__ int s_cur;
  const int t = threadIdx.x, b = blockIdx.x;
  const float* bx = boxes + (int64_t)b * N * 10;
  const float4* ab = aabb + (int64_t)b * N;
  int* kout = keep_idx + (int64_t)b * max_keep;
  float* bout = boxes_out + (int64_t)b * max_keep * 10;
  const int nwords = (N + 31) >> 5;
  for (int w = t; w < nwords; w += NT) bitmap[w] = 0u;
  for (int k = t; k < max_keep; k += NT) kout[k] = -1;                      // Fill(out[0], -1) :501
  for (int k = t; k < max_keep * 10; k += NT) bout[k] = 0.f;                // Fill(out[1], 0)  :502
  if (t == 0) s_cur = 0;
  __syncthreads();
  int kept = 0;
  while (true) {
    const int i = s_cur;
    if (i < 0 || i >= N || kept >= max_keep) break;
    if (t < 10) {
      const float v = bx[(int64_t)i * 10 + t];
      cur_box[t] = v;
      bout[kept * 10 + t] = v;
    }
    if (t == 32) { cur_aabb = ab[i]; kout[kept] = i; }
    __syncthreads();
    const float4 ai = cur_aabb;
    for (int j = i + 1 + t; j < N; j += NT) {
      if ((bitmap[j >> 5] >> (j & 31)) & 1u) continue;
      if (!normal_iou) {
        const float4 aj = ab[j];
        if (ai.z < aj.x || aj.z < ai.x || ai.w < aj.y || aj.w < ai.y) continue;
      }
      float bj[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) bj[k] = bx[(int64_t)j * 10 + k];
      const float v = normal_iou ? iou_normal(cur_box, bj) : iou_bev3d(cur_box, bj);
      if (v > thr) atomicOr(&bitmap[j >> 5], 1u << (j & 31));
    }
    __syncthreads();
    ++kept;
    if (t < 32) {  // next unsuppressed index > i
      int found = -1;
      const int w0 = (i + 1) >> 5;
      const unsigned first_mask = ~0u << ((i + 1) & 31);
      for (int wb = w0; wb < nwords && found < 0; wb += 32) {
        const int w = wb + t;
        unsigned free_bits = 0u;
        if (w < nwords) {
          free_bits = ~bitmap[w];
          if (w == w0) free_bits &= first_mask;
          if (w == nwords - 1 && (N & 31)) free_bits &= (1u << (N & 31)) - 1u;
        }
        const unsigned ball = __ballot_sync(0xffffffffu, free_bits != 0u);
        if (ball) {
          const int src = __ffs(ball) - 1;
          const unsigned fb = __shfl_sync(0xffffffffu, free_bits, src);
          found = ((wb + src) << 5) + (__ffs(fb) - 1);
        }
      }
      if (t == 0) s_cur = found;
    }
    __syncthreads();
  }
}

}  // namespace n3

extern "C" {

size_t rd_nms3d_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return (size_t)B * N * sizeof(float4);
}

int rd_nms3d(const float* boxes, int B, int N, float iou_thres, int max_keep, int normal_iou, int32_t* keep_idx,
             float* boxes_out, void* workspace, size_t workspace_bytes, rd_stream_t stream) {
  RD_REQUIRE(B >= 0 && N >= 0 && max_keep >= 0, "rd_nms3d: negative size");
  if (B == 0 || max_keep == 0) return 0;
  RD_REQUIRE(keep_idx && boxes_out, "rd_nms3d: null output pointer");
  if (rd_check_device()) return 1;
  cudaStream_t st = rd::as_stream(stream);
  if (N == 0) {
    RD_CUDA(cudaMemsetAsync(keep_idx, 0xff, sizeof(int32_t) * (size_t)B * max_keep, st));
    RD_CUDA(cudaMemsetAsync(boxes_out, 0, sizeof(float) * (size_t)B * max_keep * 10, st));
    return 0;
  }
  RD_REQUIRE(boxes != nullptr, "rd_nms3d: null boxes");
  RD_REQUIRE(workspace && workspace_bytes >= rd_nms3d_workspace_bytes(B, N), "rd_nms3d: workspace too small");
  const size_t bitmap_bytes = (size_t)((N + 31) / 32) * 4;
  RD_REQUIRE(bitmap_bytes <= 200 * 1024, "rd_nms3d: N=%d exceeds the shared-memory bitmap capacity (1.6M boxes)", N);
  float4* aabb = static_cast<float4*>(workspace);
  const int64_t total = (int64_t)B * N;
  n3::prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(boxes, total, aabb);
  RD_CUDA(rd::smem_optin(n3::nms3d_kernel, bitmap_bytes));   // per (kernel, device), see rd_common.cuh
  n3::nms3d_kernel<<<B, n3::NT, bitmap_bytes, st>>>(boxes, aabb, N, iou_thres, max_keep, normal_iou ? 1 : 0, keep_idx,
                                                    boxes_out);
  rd::count_launch(2);
  return rd::check_launch("rd_nms3d");
}

}  // extern "C"
