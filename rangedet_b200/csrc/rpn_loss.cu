// RPN loss head of the training graph, one fused pass over the pixels of a pyramid level (sm_100a).
//
// Reference behaviour reproduced (files under /root/reference):
//   rangedet/symbol/head/builder.py:155-197   get_iou_target: transpose(bbox_delta) -> Decode3DBbox ->
//                                             Custom 'batch_rotated_iou' -> stop_gradient
//   rangedet/symbol/head/loss.py:4-30         sigmoid_bce_loss_with_logits / vari_focal_loss
//   rangedet/symbol/head/builder.py:350-379   get_vfl_loss: * mask / (sum(mask)+1), MakeLoss(grad_scale)
//   rangedet/symbol/head/builder.py:381-422   get_normalize_reg_loss: smooth_l1(delta-target, scalar) *
//                                             weight * norm_weight / (sum(norm_weight)+1) * reg_loss_weight
// The reference runs this as ~45 MXNet ops per level (25 element-wise kernels, a (N,200) IoU matrix of
// 136 MB at level 0, a Python loop over the batch inside a CustomOp).  Here: one deterministic two-stage
// reduction for the two normalisers, then ONE kernel that reads every input once and writes the loss
// tensors (the graph outputs), the IoU target and the gradients w.r.t. cls_logit / reg_delta that
// MakeLoss would start back-propagation with.
//
// Per pixel: algorithmic bytes = read 4 (logit) + 32 (delta) + 12 (pc) + 4 (mask) + 96 (target, weight,
// norm weight) = 148 B, write 4 + 4 + 32 + 4 + 32 = 76 B; the 200 GT boxes live in shared memory.  The
// kernel is ALU/SFU-latency bound by the IoU target (200 AABB tests + a few clippings per pixel), not by HBM.
//
// Compiled with -fmad=false like decode_iou.cu: the decode + IoU arithmetic is the same operation sequence,
// so the IoU target is bit-identical to rd_decode_3d_bbox -> rd_batch_rotated_iou_max.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>

#include "../../include/rangedet_b200.h"
#include "iou_device.cuh"
#include "rd_common.cuh"

namespace {
using namespace rd_iou;

constexpr int LOSS_THREADS = 128;
constexpr int LOSS_GCHUNK = 256;
constexpr int LOSS_QCAP = 2048;   // queue of (pixel, GT) pairs per CTA and GT chunk
constexpr int SUM_BLOCKS = 296;   // 2 x 148 SMs
constexpr int SUM_THREADS = 256;

// ---- stage 1: partial sums of mask and reg_norm_weight (fixed partition -> deterministic) -------------
__global__ void __launch_bounds__(SUM_THREADS)
loss_norm_partial_kernel(const float* __restrict__ mask, int64_t n_mask, const float* __restrict__ nw,
                         int64_t n_nw, double* __restrict__ partial) {
  __shared__ double red[2][SUM_THREADS / 32];
  double a = 0.0, b = 0.0;
  const int64_t stride = (int64_t)gridDim.x * SUM_THREADS;
  for (int64_t i = (int64_t)blockIdx.x * SUM_THREADS + threadIdx.x; i < n_mask; i += stride) a += (double)__ldg(mask + i);
  for (int64_t i = (int64_t)blockIdx.x * SUM_THREADS + threadIdx.x; i < n_nw; i += stride) b += (double)__ldg(nw + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = a;
    red[1][threadIdx.x >> 5] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0.0, sb = 0.0;
    for (int w = 0; w < SUM_THREADS / 32; ++w) {
      sa += red[0][w];
      sb += red[1][w];
    }
    partial[blockIdx.x] = sa;
    partial[SUM_BLOCKS + blockIdx.x] = sb;
  }
}

__device__ __forceinline__ float sanitise_iou(float v) {  // batch_rotated_iou.py:43-46
  return (isnan(v) || isinf(v) || v > 1.0f || v < 0.0f) ? 0.f : v;
}

// Decode3DBboxKernelGPU::Map (decode_3d_bbox-inl.h:186-274), non-bin: same operation order as decode_kernel
// in decode_iou.cu.  v[0..7] = corners A,B,C,D (x,y), v[8] = z0, v[9] = z0 + h.
__device__ __forceinline__ void decode_box(const float* d, float px, float py, float* v) {
  const float az = atan2f(py, px);
  const float sa = sinf(az), ca = cosf(az);
  const float dx = d[0] * fabsf(d[0]);
  const float dy = d[1] * fabsf(d[1]);
  const float width = expf(d[2]);
  const float length = expf(d[3]);
  const float height = expf(d[7]);
  const float z0 = d[6];
  const float yaw = atan2f(d[5], d[4]) + az;
  const float cx = px + (dx * ca - dy * sa);
  const float cy = py + (dx * sa + dy * ca);
  const float s = sinf(yaw), c = cosf(yaw);
  const float hl = 0.5f * length, hw = 0.5f * width;
  // rot2(x, y): (x*c - y*s, x*s + y*c)
  v[0] = (hl * c - (-hw) * s) + cx;
  v[1] = (hl * s + (-hw) * c) + cy;
  v[2] = ((-hl) * c - (-hw) * s) + cx;
  v[3] = ((-hl) * s + (-hw) * c) + cy;
  v[4] = ((-hl) * c - hw * s) + cx;
  v[5] = ((-hl) * s + hw * c) + cy;
  v[6] = (hl * c - hw * s) + cx;
  v[7] = (hl * s + hw * c) + cy;
  v[8] = z0;
  v[9] = z0 + height;
}

struct LossParams {
  const float *cls_logit, *reg_delta, *pc, *gt, *mask, *reg_target, *reg_weight, *reg_norm_weight;
  const double* partial;
  float *iou_target, *cls_loss, *reg_loss, *d_cls, *d_reg;
  int64_t N;
  int G;
  float alpha, gamma, sl1_sigma2, cls_grad_scale, reg_loss_weight, reg_grad_scale;
  // IOF != 0 (rd_rpn_loss_nhwc_*): the head outputs and their gradients in the layout the head convolutions use -- zero-haloed
  // NHWC [B][H+2][W+2][cpad] of 2-byte floats, logit = channel 0 of cls_pad, deltas = channels 0..7 of reg_pad; the
  // gradients go to the same channels of dcls_pad / dreg_pad (other channels untouched: they stay zero)
  const void *cls_pad, *reg_pad;
  void *dcls_pad, *dreg_pad;
  int H, W, cpad;
};

// 2-byte storage <-> fp32 (IOF 1: bf16, 2: fp16)
template <int IOF>
__device__ __forceinline__ float ld2(const void* base, int64_t i) {
  if (IOF == 1) return __bfloat162float(static_cast<const __nv_bfloat16*>(base)[i]);
  return __half2float(static_cast<const __half*>(base)[i]);
}
template <int IOF>
__device__ __forceinline__ unsigned short st2(float v) {
  if (IOF == 1) { const __nv_bfloat16 h = __float2bfloat16_rn(v); return *reinterpret_cast<const unsigned short*>(&h); }
  const __half h = __float2half_rn(v);
  return *reinterpret_cast<const unsigned short*>(&h);
}

// ---- stage 2: one thread per pixel ----------------------------------------------------------------------
template <bool IOU_3D, int IOF>
__global__ void __launch_bounds__(LOSS_THREADS)
rpn_loss_kernel(const LossParams p) {
  __shared__ Box2 sg[IOU_3D ? 1 : LOSS_GCHUNK];
  __shared__ float sg7[IOU_3D ? LOSS_GCHUNK * 7 : 1];
  __shared__ float s_norm[2];
  __shared__ unsigned char s_dup[LOSS_GCHUNK];   // GT row bit-identical to the previous row (the fixed-length padding)
  __shared__ Box2 s_me[IOU_3D ? 1 : LOSS_THREADS];      // decoded boxes of this CTA's pixels (phase 2 reads any of them)
  __shared__ unsigned short s_q[IOU_3D ? 1 : LOSS_QCAP];   // queued (pixel << 8 | gt) pairs that need clipping
  __shared__ unsigned int s_best[LOSS_THREADS];
  __shared__ int s_qn;
  const int b = blockIdx.y;
  const int64_t n = (int64_t)blockIdx.x * LOSS_THREADS + threadIdx.x;
  const int64_t N = p.N;
  const bool active = n < N;
  if (threadIdx.x < 2) {  // every CTA folds the partials in the same order -> identical normalisers everywhere
    double s = 0.0;
    for (int i = 0; i < SUM_BLOCKS; ++i) s += p.partial[threadIdx.x * SUM_BLOCKS + i];
    s_norm[threadIdx.x] = (float)s + 1.0f;  // builder.py:365, :412
  }

  float d[8];
  float v[10];
  Box2 me;
  Rect mr;
  float mz = 0.f, mh = 0.f;
  if (active) {
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = 0.f;
    if (IOF == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) d[k] = __ldg(p.reg_delta + ((int64_t)b * 8 + k) * N + n);  // planar: the transpose is free
    } else {   // one 16-byte load: the eight deltas of this pixel
      const int hh = (int)(n / p.W), ww = (int)(n % p.W);
      const int64_t e = (((int64_t)b * (p.H + 2) + hh + 1) * (p.W + 2) + ww + 1) * p.cpad;
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(static_cast<const unsigned short*>(p.reg_pad) + e));
      const unsigned short* h8 = reinterpret_cast<const unsigned short*>(&q);
#pragma unroll
      for (int k = 0; k < 8; ++k) d[k] = ld2<IOF>(h8, k);
    }
    const float* q = p.pc + ((int64_t)b * N + n) * 3;
    decode_box(d, __ldg(q), __ldg(q + 1), v);
    if (!IOU_3D) {
      load_box8(v, me);
      s_me[threadIdx.x] = me;
    } else {  // to_box_type_7 (batch_rotated_iou.py:51-68) + yaw negation (:35-36)
      float b7[7];
      b7[0] = (((v[0] + v[2]) + v[4]) + v[6]) / 4.0f;
      b7[1] = (((v[1] + v[3]) + v[5]) + v[7]) / 4.0f;
      b7[2] = (v[8] + v[9]) / 2.0f;
      const float l0 = v[0] - v[2], l1 = v[1] - v[3], w0 = v[2] - v[4], w1 = v[3] - v[5];
      b7[3] = sqrtf(l0 * l0 + l1 * l1);
      b7[4] = sqrtf(w0 * w0 + w1 * w1);
      b7[5] = v[9] - v[8];
      b7[6] = -1.0f * atan2f(v[1] - v[3], v[0] - v[2]);
      load_rect7(b7, mr, &mz, &mh);
    }
  }
  float best = -INFINITY;
  s_best[threadIdx.x] = 0u;   // every sanitised IoU is >= +0 and at least GT row 0 is evaluated
  if (threadIdx.x == 0) s_qn = 0;
  const int G = p.G;
  for (int g0 = 0; g0 < G; g0 += LOSS_GCHUNK) {
    const int ng = min(LOSS_GCHUNK, G - g0);
    __syncthreads();
    // The reference pads GT to 200 rows with one constant box (rangedet/core/input.py:264-265).  max() over a
    // set does not change when bit-identical duplicates are dropped, so a row equal to its predecessor is skipped:
    // pixels without a return (pc = 0) decode to boxes around the origin that would otherwise be clipped against
    // every one of the ~170 padding rows.
    constexpr int GD = IOU_3D ? 7 : 8;
    for (int g = threadIdx.x; g < ng; g += LOSS_THREADS) {
      const float* r1 = p.gt + ((int64_t)b * G + g0 + g) * GD;
      bool dup = g0 + g > 0;
      if (dup) {
#pragma unroll
        for (int k = 0; k < GD; ++k) dup = dup && __float_as_uint(__ldg(r1 + k)) == __float_as_uint(__ldg(r1 + k - GD));
      }
      s_dup[g] = dup ? 1 : 0;
    }
    if (!IOU_3D) {
      for (int g = threadIdx.x; g < ng; g += LOSS_THREADS) {
        float w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = __ldg(p.gt + ((int64_t)b * G + g0 + g) * 8 + k);
        Box2 t;
        load_box8(w, t);
        sg[g] = t;
      }
    } else {
      for (int e = threadIdx.x; e < ng * 7; e += LOSS_THREADS) {
        float w = __ldg(p.gt + ((int64_t)b * G + g0) * 7 + e);
        if (e % 7 == 6) w = -1.0f * w;
        sg7[e] = w;
      }
    }
    __syncthreads();
    if (!IOU_3D) {
      // Two phases, so that the expensive part (polygon clipping, divergent: only a few percent of the (pixel, GT)
      // pairs survive the exact early-outs of iou_quads) is spread over the whole CTA instead of stalling the warp
      // of whichever pixel needs it:  (1) every pixel runs the early-outs inline and queues its surviving pairs,
      // (2) all threads clip queued pairs, one per thread, and fold the result into the pixel's maximum with an
      // atomicMax on the bit pattern (IoUs are non-negative floats: bit order == value order; max is
      // order-independent, so this stays deterministic).  A full queue falls back to clipping in place.
      if (active) {
        for (int g = 0; g < ng; ++g) {
          if (s_dup[g]) continue;   // uniform across the CTA
          const Box2& t = sg[g];
          if (me.area < R_EPS || t.area < R_EPS) continue;             // iou_quads would return +0
          if (me.convex && t.convex && aabb_disjoint(me, t)) continue;   // likewise
          const int pos = atomicAdd(&s_qn, 1);
          if (pos < LOSS_QCAP) {
            s_q[pos] = (unsigned short)((threadIdx.x << 8) | g);
          } else {
            float u = sanitise_iou(iou_quads(me, t));
            u = u > 0.f ? u : 0.f;
            atomicMax(&s_best[threadIdx.x], __float_as_uint(u));
          }
        }
      }
      __syncthreads();
      const int nq = min(s_qn, LOSS_QCAP);
      for (int i = threadIdx.x; i < nq; i += LOSS_THREADS) {
        const int e = s_q[i], px = e >> 8, g = e & 255;
        float u = sanitise_iou(iou_quads(s_me[px], sg[g]));
        u = u > 0.f ? u : 0.f;   // also folds -0.0 into +0.0 (its bit pattern would win the unsigned max)
        atomicMax(&s_best[px], __float_as_uint(u));
      }
      __syncthreads();
      if (threadIdx.x == 0) s_qn = 0;
    } else if (active) {
      for (int g = 0; g < ng; ++g) {
        if (s_dup[g]) continue;   // uniform across the CTA
        Rect rg;
        float gz, gh;
        load_rect7(sg7 + g * 7, rg, &gz, &gh);
        const float u = sanitise_iou(iou_rect7(mr, mz, mh, rg, gz, gh));
        best = u > best ? u : best;
      }
    }
  }
  if (!IOU_3D) best = __uint_as_float(s_best[threadIdx.x]);
  if (!active) return;
  const float t = best;  // IoU target in [0,1]
  const int64_t i1 = (int64_t)b * N + n;
  if (p.iou_target) p.iou_target[i1] = t;
  // element offset of this pixel's channel 0 in the haloed NHWC head tensors (IOF != 0)
  const int64_t epad = IOF ? (((int64_t)b * (p.H + 2) + (int)(n / p.W) + 1) * (p.W + 2) + (int)(n % p.W) + 1) * p.cpad : 0;

  // ---- varifocal loss (loss.py:4-30) and its derivative w.r.t. the logit ----
  {
    const float x = IOF ? ld2<IOF>(p.cls_pad, epad) : __ldg(p.cls_logit + i1);
    const float m = __ldg(p.mask + i1);
    const float pr = 1.0f / (1.0f + expf(-x));                    // mx.sym.sigmoid
    const float ge = x >= 0.f ? 1.f : 0.f;
    const float na = x - 2.0f * x * ge;                           // -|x|
    const float sp = log1pf(expf(na));                            // softrelu
    const float minus_log = -1.0f * x * ge - sp;                  // log(1 - p)
    const bool in_clip = pr >= 1e-6f && pr <= 1.0f - 1e-6f;
    const float pc_ = fminf(fmaxf(pr, 1e-6f), 1.0f - 1e-6f);
    const float log_p = logf(pc_);
    const float li = -1.0f * (0.5f * t * log_p + 0.5f * (1.0f - t) * minus_log) * 2.0f;      // loss_init
    const float dpr = pr * (1.0f - pr);
    const float dlogp = in_clip ? dpr / pc_ : 0.f;
    const float dminus = -ge - (1.0f / (1.0f + expf(-na))) * (1.0f - 2.0f * ge);              // = -p
    const float dli = -1.0f * (0.5f * t * dlogp + 0.5f * (1.0f - t) * dminus) * 2.0f;
    float loss, dloss;
    if (t > 0.f) {           // positive_mask
      loss = li * t;
      dloss = dli * t;
    } else if (t == 0.f) {   // negative_mask
      const float df = t - pr, a = fabsf(df);
      float pw, dpw;         // |.|^gamma and its derivative w.r.t. |.|
      if (p.gamma == 2.0f) {
        pw = a * a;
        dpw = 2.0f * a;
      } else {
        pw = powf(a, p.gamma);
        dpw = p.gamma * powf(a, p.gamma - 1.0f);
      }
      const float sgn = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
      loss = li * p.alpha * pw;
      dloss = dli * p.alpha * pw + li * p.alpha * dpw * sgn * (-dpr);
    } else {                 // NaN target cannot occur after sanitise; keep the reference's "neither mask" = 0
      loss = 0.f;
      dloss = 0.f;
    }
    const float w = m / s_norm[0];
    if (p.cls_loss) p.cls_loss[i1] = loss * m / s_norm[0];
    if (IOF) static_cast<unsigned short*>(p.dcls_pad)[epad] = st2<IOF>(p.cls_grad_scale * (dloss * w));
    else if (p.d_cls) p.d_cls[i1] = p.cls_grad_scale * (dloss * w);
  }

  // ---- normalised smooth-L1 (builder.py:381-422; mx smooth_l1: sigma^2 = scalar^2) ----
  const float s2 = p.sl1_sigma2, is2 = 1.0f / s2;
  unsigned short dq[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int64_t i8 = ((int64_t)b * 8 + k) * N + n;
    const float df = d[k] - __ldg(p.reg_target + i8);
    const float w = __ldg(p.reg_weight + i8) * __ldg(p.reg_norm_weight + i8);
    float l, g;
    if (df > is2) {
      l = df - 0.5f * is2;
      g = 1.f;
    } else if (df < -is2) {
      l = -df - 0.5f * is2;
      g = -1.f;
    } else {
      l = 0.5f * df * df * s2;
      g = s2 * df;
    }
    if (p.reg_loss) p.reg_loss[i8] = l * w / s_norm[1] * p.reg_loss_weight;
    if (IOF) dq[k] = st2<IOF>(p.reg_grad_scale * (g * w / s_norm[1] * p.reg_loss_weight));
    else if (p.d_reg) p.d_reg[i8] = p.reg_grad_scale * (g * w / s_norm[1] * p.reg_loss_weight);
  }
  if (IOF) {
    uint4 q;
    unsigned short* h8 = reinterpret_cast<unsigned short*>(&q);
#pragma unroll
    for (int k = 0; k < 8; ++k) h8[k] = dq[k];
    *reinterpret_cast<uint4*>(static_cast<unsigned short*>(p.dreg_pad) + epad) = q;
  }
}

}  // namespace

extern "C" {

size_t rd_rpn_loss_workspace_bytes(void) { return 2 * SUM_BLOCKS * sizeof(double); }

// iof 0: fp32 planar head tensors (the reference op boundary); 1 / 2: haloed NHWC bf16 / fp16 head tensors
static int loss_run(const char* who, int iof, const float* cls_logit, const float* reg_delta, const void* cls_pad,
                    const void* reg_pad, int H, int W, int cpad, const float* pc, const float* gt, const float* mask,
                    const float* reg_target, const float* reg_weight, const float* reg_norm_weight, int B, int64_t N, int G,
                    int iou_type, float alpha, float gamma, float smooth_l1_scalar, float cls_grad_scale,
                    float reg_loss_weight, float reg_grad_scale, float* iou_target, float* cls_loss, float* reg_loss,
                    float* d_cls, float* d_reg, void* dcls_pad, void* dreg_pad, void* workspace, size_t workspace_bytes,
                    rd_stream_t stream) {
  RD_REQUIRE(iou_type == 0 || iou_type == 1, "%s: iou_type must be 0 (bev) or 1 (3d)", who);
  RD_REQUIRE(B >= 0 && N >= 0 && G >= 1, "%s: bad sizes B=%d G=%d", who, B, G);
  RD_REQUIRE(smooth_l1_scalar > 0.f, "%s: smooth_l1_scalar must be positive", who);
  if (B == 0 || N == 0) return 0;
  RD_REQUIRE(pc && gt && mask && reg_target && reg_weight && reg_norm_weight, "%s: null input pointer", who);
  if (iof == 0) {
    RD_REQUIRE(cls_logit && reg_delta, "%s: null input pointer", who);
  } else {
    RD_REQUIRE(cls_pad && reg_pad && dcls_pad && dreg_pad, "%s: null head tensor", who);
    RD_REQUIRE(H > 0 && W > 0 && (int64_t)H * W == N && cpad >= 8 && cpad % 8 == 0, "%s: bad head tensor shape", who);
    RD_REQUIRE(((reinterpret_cast<uintptr_t>(reg_pad) | reinterpret_cast<uintptr_t>(dreg_pad)) & 15) == 0,
               "%s: head tensors must be 16-byte aligned", who);
  }
  RD_REQUIRE(B <= 65535, "%s: B too large", who);
  RD_REQUIRE(workspace && workspace_bytes >= rd_rpn_loss_workspace_bytes(), "%s: workspace too small", who);
  RD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "%s: workspace must be 8-byte aligned", who);
  if (rd_check_device()) return 1;
  cudaStream_t st = rd::as_stream(stream);
  double* partial = static_cast<double*>(workspace);
  loss_norm_partial_kernel<<<SUM_BLOCKS, SUM_THREADS, 0, st>>>(mask, (int64_t)B * N, reg_norm_weight,
                                                              (int64_t)B * 8 * N, partial);
  LossParams p;
  p.cls_logit = cls_logit; p.reg_delta = reg_delta; p.pc = pc; p.gt = gt; p.mask = mask;
  p.reg_target = reg_target; p.reg_weight = reg_weight; p.reg_norm_weight = reg_norm_weight;
  p.partial = partial;
  p.iou_target = iou_target; p.cls_loss = cls_loss; p.reg_loss = reg_loss; p.d_cls = d_cls; p.d_reg = d_reg;
  p.N = N; p.G = G;
  p.alpha = alpha; p.gamma = gamma; p.sl1_sigma2 = smooth_l1_scalar * smooth_l1_scalar;
  p.cls_grad_scale = cls_grad_scale; p.reg_loss_weight = reg_loss_weight; p.reg_grad_scale = reg_grad_scale;
  p.cls_pad = cls_pad; p.reg_pad = reg_pad; p.dcls_pad = dcls_pad; p.dreg_pad = dreg_pad;
  p.H = H; p.W = W; p.cpad = cpad;
  dim3 grid((unsigned)((N + LOSS_THREADS - 1) / LOSS_THREADS), (unsigned)B);
#define RD_LOSS_LAUNCH(I3, IO) rpn_loss_kernel<I3, IO><<<grid, LOSS_THREADS, 0, st>>>(p)
  if (iou_type == 0) {
    if (iof == 0) RD_LOSS_LAUNCH(false, 0); else if (iof == 1) RD_LOSS_LAUNCH(false, 1); else RD_LOSS_LAUNCH(false, 2);
  } else {
    if (iof == 0) RD_LOSS_LAUNCH(true, 0); else if (iof == 1) RD_LOSS_LAUNCH(true, 1); else RD_LOSS_LAUNCH(true, 2);
  }
#undef RD_LOSS_LAUNCH
  rd::count_launch(2);
  return rd::check_launch(who);
}

int rd_rpn_loss(const float* cls_logit, const float* reg_delta, const float* pc, const float* gt,
                const float* mask, const float* reg_target, const float* reg_weight,
                const float* reg_norm_weight, int B, int64_t N, int G, int iou_type, float alpha, float gamma,
                float smooth_l1_scalar, float cls_grad_scale, float reg_loss_weight, float reg_grad_scale,
                float* iou_target, float* cls_loss, float* reg_loss, float* d_cls, float* d_reg,
                void* workspace, size_t workspace_bytes, rd_stream_t stream) {
  return loss_run("rd_rpn_loss", 0, cls_logit, reg_delta, nullptr, nullptr, 0, 0, 0, pc, gt, mask, reg_target, reg_weight,
                  reg_norm_weight, B, N, G, iou_type, alpha, gamma, smooth_l1_scalar, cls_grad_scale, reg_loss_weight,
                  reg_grad_scale, iou_target, cls_loss, reg_loss, d_cls, d_reg, nullptr, nullptr, workspace, workspace_bytes,
                  stream);
}

int rd_rpn_loss_nhwc_bf16(const void* cls_pad, const void* reg_pad, int H, int W, int cpad, const float* pc, const float* gt,
                          const float* mask, const float* reg_target, const float* reg_weight, const float* reg_norm_weight,
                          int B, int G, int iou_type, float alpha, float gamma, float smooth_l1_scalar, float cls_grad_scale,
                          float reg_loss_weight, float reg_grad_scale, float* iou_target, float* cls_loss, float* reg_loss,
                          void* dcls_pad, void* dreg_pad, void* workspace, size_t workspace_bytes, rd_stream_t stream) {
  return loss_run("rd_rpn_loss_nhwc_bf16", 1, nullptr, nullptr, cls_pad, reg_pad, H, W, cpad, pc, gt, mask, reg_target,
                  reg_weight, reg_norm_weight, B, (int64_t)H * W, G, iou_type, alpha, gamma, smooth_l1_scalar, cls_grad_scale,
                  reg_loss_weight, reg_grad_scale, iou_target, cls_loss, reg_loss, nullptr, nullptr, dcls_pad, dreg_pad,
                  workspace, workspace_bytes, stream);
}

int rd_rpn_loss_nhwc_f16(const void* cls_pad, const void* reg_pad, int H, int W, int cpad, const float* pc, const float* gt,
                         const float* mask, const float* reg_target, const float* reg_weight, const float* reg_norm_weight,
                         int B, int G, int iou_type, float alpha, float gamma, float smooth_l1_scalar, float cls_grad_scale,
                         float reg_loss_weight, float reg_grad_scale, float* iou_target, float* cls_loss, float* reg_loss,
                         void* dcls_pad, void* dreg_pad, void* workspace, size_t workspace_bytes, rd_stream_t stream) {
  return loss_run("rd_rpn_loss_nhwc_f16", 2, nullptr, nullptr, cls_pad, reg_pad, H, W, cpad, pc, gt, mask, reg_target,
                  reg_weight, reg_norm_weight, B, (int64_t)H * W, G, iou_type, alpha, gamma, smooth_l1_scalar, cls_grad_scale,
                  reg_loss_weight, reg_grad_scale, iou_target, cls_loss, reg_loss, nullptr, nullptr, dcls_pad, dreg_pad,
                  workspace, workspace_bytes, stream);
}

}  // extern "C"
