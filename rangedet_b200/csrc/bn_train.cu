// Training-mode BatchNorm + ReLU + residual, forward and backward, over zero-haloed NHWC bf16 (sm_100a).
//
// Replaces what the reference gets from MXNet around every convolution of the DLA backbone / RPN head
// in training: mx.sym.BatchNorm with batch statistics per GPU (/root/reference mxnext/complicate.py:14,
// 32-43: eps 1e-5 + 1e-10, momentum 0.9, fix_gamma False, use_global_stats False), mx.sym.Activation
// relu (mxnext/simple.py:49) and the elementwise add of the residual branches
// (rangedet/symbol/backbone/dla_backbone.py:42-56 block: relu(bn2 + shortcut); :121-124 agg_stage:
// const + relu(bn(deconv))).  All of these are HBM-bound passes over the activation: each kernel here
// reads / writes every tensor exactly once with 16-byte accesses (8 channels per thread), per-channel
// reductions are two-stage and deterministic (per-block partial rows, then a fixed-order sum).
//
//   forward :  stats(z) -> finalize (mean, var, a = gamma*invstd, b = beta - mean*a, moving stats)
//              y = relu?(z*a + b + res_before) + res_after
//   backward:  g = dy * [pre-activation > 0]          (mask from y, or recomputed from z when res_after)
//              reduce  S1 = sum g, S2 = sum g*(z - mean)
//              dz = a * (g - S1/M - (z - mean) * invstd^2 * S2/M),  dgamma = invstd*S2,  dbeta = S1
#include <stdlib.h>

#include "../../include/rangedet_b200.h"
#include "act_type.cuh"
#include "rd_common.cuh"

namespace RD_ACT_NS(bn) {

constexpr int MAX_BLOCKS = 1184;  // 8 per SM: partial columns of the two-stage reductions
constexpr int NT = 256;

struct Geo {
  int N, H, W, C;
  int nseg, seg;      // each image row is cut into nseg segments of seg pixels (work units)
  int cgs, ppb;       // 8-channel groups per pixel, pixels per block iteration
};

__host__ __device__ inline int64_t pix_off(int n, int h, int w, int H, int W, int halo, int C) {
  return (((int64_t)n * (H + 2) + h + 1) * (W + 2 * halo) + w + halo) * C;
}

static Geo make_geo(int N, int H, int W, int C) {
  Geo g;
  g.N = N; g.H = H; g.W = W; g.C = C;
  g.cgs = C / 8;
  g.ppb = NT / g.cgs;
  if (g.ppb < 1) g.ppb = 1;
  int nseg = 1;
  while ((int64_t)N * H * nseg < 4 * MAX_BLOCKS && (W / (nseg * 2)) >= 4 * g.ppb) nseg *= 2;
  g.nseg = nseg;
  g.seg = (W + nseg - 1) / nseg;
  return g;
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    act::unpack2(w[i], f[2 * i], f[2 * i + 1]);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    w[i] = act::pack2(f[2 * i], f[2 * i + 1]);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// Block-level sum of 16 per-thread values over the threads that share a channel group; one partial COLUMN
// per block: partial[(which*C + channel) * MAX_BLOCKS + block] (so the finalize warps read contiguously).
__device__ __forceinline__ void block_reduce_store(const float (&s)[8], const float (&q)[8], const Geo& G,
                                                   float* __restrict__ partial) {
  __shared__ float red[NT * 16];
  const int t = threadIdx.x;
  const int cg = t % G.cgs, pl = t / G.cgs;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[(pl * G.cgs + cg) * 16 + i] = s[i];
    red[(pl * G.cgs + cg) * 16 + 8 + i] = q[i];
  }
  __syncthreads();
  const int nval = G.cgs * 16;  // = 2C
  for (int v = t; v < nval; v += blockDim.x) {
    float a = 0.f;
    for (int p = 0; p < G.ppb; ++p) a += red[p * nval + v];
    const int cgi = v / 16, i = v % 16;
    partial[(int64_t)((i / 8) * G.C + cgi * 8 + (i % 8)) * MAX_BLOCKS + blockIdx.x] = a;
  }
}

// ---- forward statistics: partial[block][2][C] = (sum z, sum z^2) ---------------------------------
__global__ void __launch_bounds__(NT) stats_kernel(const act_t* __restrict__ z, Geo G, int halo,
                                                   float* __restrict__ partial) {
  const int t = threadIdx.x, cg = t % G.cgs, pl = t / G.cgs;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  const int nunits = G.N * G.H * G.nseg;
  for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
    const int sg = u % G.nseg, row = u / G.nseg, h = row % G.H, n = row / G.H;
    const int w0 = sg * G.seg, w1 = min(G.W, w0 + G.seg);
    const act_t* base = z + pix_off(n, h, 0, G.H, G.W, halo, G.C) + cg * 8;
    int w = w0 + pl;
    for (; w + 3 * G.ppb < w1; w += 4 * G.ppb) {   // four independent 16-byte loads in flight per thread
      uint4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)(w + j * G.ppb) * G.C));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f[8];
        unpack8(v[j], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
      }
    }
    for (; w < w1; w += G.ppb) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(base + (int64_t)w * G.C)), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
    }
  }
  block_reduce_store(s, q, G, partial);
}

// One warp per channel: lanes stride over the per-block partials, fp64 combination, shuffle reduction.
__device__ __forceinline__ void channel_totals(const float* __restrict__ partial, int nblocks, int C, int c, double& s,
                                               double& q) {
  const int lane = threadIdx.x & 31;
  s = 0.0;
  q = 0.0;
  const float* ps = partial + (int64_t)c * MAX_BLOCKS;
  const float* pq = partial + (int64_t)(C + c) * MAX_BLOCKS;
  for (int b = lane; b < nblocks; b += 32) {
    s += (double)ps[b];
    q += (double)pq[b];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
}

// coef layout (fp32, 6 x C): a | b | mean | invstd | batch var (biased) | spare
__global__ void __launch_bounds__(256) fwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int C,
                                                           double count, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps, float momentum,
                                                           float* __restrict__ moving_mean,
                                                           float* __restrict__ moving_var, float* __restrict__ coef) {
  if (threadIdx.x == 0) rd::pdl_trigger();
  rd::pdl_wait();
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= C) return;
  double s, q;
  channel_totals(partial, nblocks, C, c, s, q);
  if ((threadIdx.x & 31) == 0) {
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    const float a = g * invstd;
    coef[c] = a;
    coef[C + c] = be - (float)mean * a;
    coef[2 * C + c] = (float)mean;
    coef[3 * C + c] = invstd;
    coef[4 * C + c] = (float)var;
    coef[5 * C + c] = (float)s;
    // MXNet BatchNorm aux update: moving = moving*momentum + batch*(1-momentum), biased batch variance
    if (moving_mean) moving_mean[c] = moving_mean[c] * momentum + (float)mean * (1.f - momentum);
    if (moving_var) moving_var[c] = moving_var[c] * momentum + (float)var * (1.f - momentum);
  }
}

// ---- forward apply: y = relu?(z*a + b + rb) + ra --------------------------------------------------
__global__ void __launch_bounds__(NT, 4) fwd_apply_kernel(const act_t* __restrict__ z,
                                                       const float* __restrict__ coef,
                                                       const act_t* __restrict__ rb,
                                                       const act_t* __restrict__ ra,
                                                       act_t* __restrict__ y, Geo G, int relu) {
  const int t = threadIdx.x, cg = t % G.cgs, pl = t / G.cgs;
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = coef[cg * 8 + i]; b[i] = coef[G.C + cg * 8 + i]; }
  const int nunits = G.N * G.H * G.nseg;
  for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
    const int sg = u % G.nseg, row = u / G.nseg, h = row % G.H, n = row / G.H;
    const int w0 = sg * G.seg, w1 = min(G.W, w0 + G.seg);
    const int64_t base = pix_off(n, h, 0, G.H, G.W, 1, G.C) + cg * 8;
    int w = w0 + pl;
    for (; w + G.ppb < w1; w += 2 * G.ppb) {   // two pixels per trip: all loads issued before the first use
      const int64_t o0 = base + (int64_t)w * G.C, o1 = o0 + (int64_t)G.ppb * G.C;
      const uint4 z0 = __ldg(reinterpret_cast<const uint4*>(z + o0)), z1 = __ldg(reinterpret_cast<const uint4*>(z + o1));
      uint4 rb0 = make_uint4(0, 0, 0, 0), rb1 = rb0, ra0 = rb0, ra1 = rb0;
      if (rb) {
        rb0 = __ldg(reinterpret_cast<const uint4*>(rb + o0));
        rb1 = __ldg(reinterpret_cast<const uint4*>(rb + o1));
      }
      if (ra) {
        ra0 = __ldg(reinterpret_cast<const uint4*>(ra + o0));
        ra1 = __ldg(reinterpret_cast<const uint4*>(ra + o1));
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float f[8], r[8];
        unpack8(j ? z1 : z0, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], a[i], b[i]);
        if (rb) {
          unpack8(j ? rb1 : rb0, r);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] += r[i];
        }
        if (relu) {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
        }
        if (ra) {
          unpack8(j ? ra1 : ra0, r);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] += r[i];
        }
        *reinterpret_cast<uint4*>(y + (j ? o1 : o0)) = pack8(f);
      }
    }
    for (; w < w1; w += G.ppb) {
      const int64_t o = base + (int64_t)w * G.C;
      float f[8], r[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(z + o)), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], a[i], b[i]);
      if (rb) {
        unpack8(__ldg(reinterpret_cast<const uint4*>(rb + o)), r);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] += r[i];
      }
      if (relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
      }
      if (ra) {
        unpack8(__ldg(reinterpret_cast<const uint4*>(ra + o)), r);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] += r[i];
      }
      *reinterpret_cast<uint4*>(y + o) = pack8(f);
    }
  }
}

// ---- backward ------------------------------------------------------------------------------------
// mask_mode: 0 no relu (g = dy), 1 mask = (y > 0) read from `ym`, 2 mask = (z*a + b > 0) recomputed
__device__ __forceinline__ void masked_grad(float (&g)[8], const float (&zf)[8], const act_t* ym,
                                            int64_t o, int mask_mode, const float (&a)[8], const float (&b)[8]) {
  if (mask_mode == 1) {
    float yf[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(ym + o)), yf);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = yf[i] > 0.f ? g[i] : 0.f;
  } else if (mask_mode == 2) {
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = fmaf(zf[i], a[i], b[i]) > 0.f ? g[i] : 0.f;
  }
}

__device__ __forceinline__ void masked_grad_v(float (&g)[8], const float (&zf)[8], const uint4& ymv, int mask_mode,
                                              const float (&a)[8], const float (&b)[8]) {
  if (mask_mode == 1) {
    float yf[8];
    unpack8(ymv, yf);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = yf[i] > 0.f ? g[i] : 0.f;
  } else if (mask_mode == 2) {
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = fmaf(zf[i], a[i], b[i]) > 0.f ? g[i] : 0.f;
  }
}

__global__ void __launch_bounds__(NT, 3) bwd_reduce_kernel(const act_t* __restrict__ dy,
                                                        const act_t* __restrict__ ym,
                                                        const act_t* __restrict__ z,
                                                        const float* __restrict__ coef, Geo G, int mask_mode,
                                                        float* __restrict__ partial) {
  const int t = threadIdx.x, cg = t % G.cgs, pl = t / G.cgs;
  float a[8], b[8], mean[8], s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = coef[cg * 8 + i];
    b[i] = coef[G.C + cg * 8 + i];
    mean[i] = coef[2 * G.C + cg * 8 + i];
    s[i] = q[i] = 0.f;
  }
  const int nunits = G.N * G.H * G.nseg;
  for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
    const int sg = u % G.nseg, row = u / G.nseg, h = row % G.H, n = row / G.H;
    const int w0 = sg * G.seg, w1 = min(G.W, w0 + G.seg);
    const int64_t base = pix_off(n, h, 0, G.H, G.W, 1, G.C) + cg * 8;
    int w = w0 + pl;
    for (; w + G.ppb < w1; w += 2 * G.ppb) {   // two pixels per trip: all loads issued before the first use
      const int64_t o0 = base + (int64_t)w * G.C, o1 = o0 + (int64_t)G.ppb * G.C;
      const uint4 vd0 = __ldg(reinterpret_cast<const uint4*>(dy + o0)), vz0 = __ldg(reinterpret_cast<const uint4*>(z + o0));
      const uint4 vd1 = __ldg(reinterpret_cast<const uint4*>(dy + o1)), vz1 = __ldg(reinterpret_cast<const uint4*>(z + o1));
      uint4 vm0 = make_uint4(0, 0, 0, 0), vm1 = vm0;
      if (mask_mode == 1) {
        vm0 = __ldg(reinterpret_cast<const uint4*>(ym + o0));
        vm1 = __ldg(reinterpret_cast<const uint4*>(ym + o1));
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float g[8], zf[8];
        unpack8(j ? vd1 : vd0, g);
        unpack8(j ? vz1 : vz0, zf);
        masked_grad_v(g, zf, j ? vm1 : vm0, mask_mode, a, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += g[i]; q[i] = fmaf(g[i], zf[i] - mean[i], q[i]); }
      }
    }
    for (; w < w1; w += G.ppb) {
      const int64_t o = base + (int64_t)w * G.C;
      float g[8], zf[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(dy + o)), g);
      unpack8(__ldg(reinterpret_cast<const uint4*>(z + o)), zf);
      masked_grad(g, zf, ym, o, mask_mode, a, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += g[i]; q[i] = fmaf(g[i], zf[i] - mean[i], q[i]); }
    }
  }
  block_reduce_store(s, q, G, partial);
}

// coef2 (fp32, 2 x C): c1 = S1/M | c2 = invstd^2 * S2/M ; also dgamma = invstd*S2, dbeta = S1
__global__ void __launch_bounds__(256) bwd_finalize_kernel(const float* __restrict__ partial, int nblocks, int C,
                                                           double count, const float* __restrict__ coef,
                                                           float* __restrict__ coef2, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta) {
  if (threadIdx.x == 0) rd::pdl_trigger();
  rd::pdl_wait();
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= C) return;
  double s, q;
  channel_totals(partial, nblocks, C, c, s, q);
  if ((threadIdx.x & 31) == 0) {
    const double invstd = (double)coef[3 * C + c];
    coef2[c] = (float)(s / count);
    coef2[C + c] = (float)(invstd * invstd * q / count);
    if (dgamma) dgamma[c] = (float)(invstd * q);
    if (dbeta) dbeta[c] = (float)s;
  }
}

__global__ void __launch_bounds__(NT, 3) bwd_apply_kernel(const act_t* __restrict__ dy,
                                                       const act_t* __restrict__ ym,
                                                       const act_t* __restrict__ z,
                                                       const float* __restrict__ coef,
                                                       const float* __restrict__ coef2, Geo G, int mask_mode,
                                                       act_t* __restrict__ dz, int dz_halo,
                                                       act_t* __restrict__ g_out) {
  const int t = threadIdx.x, cg = t % G.cgs, pl = t / G.cgs;
  float a[8], b[8], mean[8], c1[8], c2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = coef[cg * 8 + i];
    b[i] = coef[G.C + cg * 8 + i];
    mean[i] = coef[2 * G.C + cg * 8 + i];
    c1[i] = coef2[cg * 8 + i];
    c2[i] = coef2[G.C + cg * 8 + i];
  }
  const int nunits = G.N * G.H * G.nseg;
  for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
    const int sg = u % G.nseg, row = u / G.nseg, h = row % G.H, n = row / G.H;
    const int w0 = sg * G.seg, w1 = min(G.W, w0 + G.seg);
    const int64_t base = pix_off(n, h, 0, G.H, G.W, 1, G.C) + cg * 8;
    const int64_t obase = pix_off(n, h, 0, G.H, G.W, dz_halo, G.C) + cg * 8;
    int w = w0 + pl;
    for (; w + G.ppb < w1; w += 2 * G.ppb) {   // two pixels per trip: all loads issued before the first use
      const int64_t o0 = base + (int64_t)w * G.C, o1 = o0 + (int64_t)G.ppb * G.C;
      const uint4 vd0 = __ldg(reinterpret_cast<const uint4*>(dy + o0)), vz0 = __ldg(reinterpret_cast<const uint4*>(z + o0));
      const uint4 vd1 = __ldg(reinterpret_cast<const uint4*>(dy + o1)), vz1 = __ldg(reinterpret_cast<const uint4*>(z + o1));
      uint4 vm0 = make_uint4(0, 0, 0, 0), vm1 = vm0;
      if (mask_mode == 1) {
        vm0 = __ldg(reinterpret_cast<const uint4*>(ym + o0));
        vm1 = __ldg(reinterpret_cast<const uint4*>(ym + o1));
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float g[8], zf[8], d[8];
        unpack8(j ? vd1 : vd0, g);
        unpack8(j ? vz1 : vz0, zf);
        masked_grad_v(g, zf, j ? vm1 : vm0, mask_mode, a, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = a[i] * (g[i] - c1[i] - (zf[i] - mean[i]) * c2[i]);
        const int wj = w + j * G.ppb;
        *reinterpret_cast<uint4*>(dz + obase + (int64_t)wj * G.C) = pack8(d);
        if (g_out) *reinterpret_cast<uint4*>(g_out + (j ? o1 : o0)) = pack8(g);
      }
    }
    for (; w < w1; w += G.ppb) {
      const int64_t o = base + (int64_t)w * G.C;
      float g[8], zf[8], d[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(dy + o)), g);
      unpack8(__ldg(reinterpret_cast<const uint4*>(z + o)), zf);
      masked_grad(g, zf, ym, o, mask_mode, a, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = a[i] * (g[i] - c1[i] - (zf[i] - mean[i]) * c2[i]);
      *reinterpret_cast<uint4*>(dz + obase + (int64_t)w * G.C) = pack8(d);
      if (g_out) *reinterpret_cast<uint4*>(g_out + o) = pack8(g);
    }
  }
}

// ---- y = x0 + x1 (gradient accumulation where a tensor has two consumers) -------------------------
__global__ void __launch_bounds__(NT) add_kernel(const act_t* __restrict__ x0,
                                                 const act_t* __restrict__ x1,
                                                 act_t* __restrict__ y, Geo G) {
  const int t = threadIdx.x, cg = t % G.cgs, pl = t / G.cgs;
  const int nunits = G.N * G.H * G.nseg;
  for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
    const int sg = u % G.nseg, row = u / G.nseg, h = row % G.H, n = row / G.H;
    const int w0 = sg * G.seg, w1 = min(G.W, w0 + G.seg);
    const int64_t base = pix_off(n, h, 0, G.H, G.W, 1, G.C) + cg * 8;
    for (int w = w0 + pl; w < w1; w += G.ppb) {
      const int64_t o = base + (int64_t)w * G.C;
      float f[8], r[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(x0 + o)), f);
      unpack8(__ldg(reinterpret_cast<const uint4*>(x1 + o)), r);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += r[i];
      *reinterpret_cast<uint4*>(y + o) = pack8(f);
    }
  }
}

#include "bn_stream.cuh"

static int check_shape(const char* what, int N, int H, int W, int C) {
  RD_REQUIRE(N > 0 && H > 0 && W > 0, "%s: bad shape", what);
  RD_REQUIRE(C >= 8 && C % 8 == 0 && C <= 1024, "%s: C must be a multiple of 8, <= 1024 (got %d)", what, C);
  return 0;
}

static int grid_for(const Geo& g) {
  const int64_t units = (int64_t)g.N * g.H * g.nseg;
  return (int)(units < MAX_BLOCKS ? units : MAX_BLOCKS);
}

}  // namespace bn_<storage type>
namespace bn = RD_ACT_NS(bn);
// the per-channel finalize kernels ((C + 7) / 8 blocks of 8 warps), launched with the PDL attribute
#define RD_CUDA_LAUNCH_FINALIZE(kernel, stream_, ...) \
  RD_CUDA(rd::launch(kernel, dim3((C + 7) / 8), dim3(256), 0, stream_, __VA_ARGS__))

extern "C" {

#ifndef RD_ACT_F16   // storage-type independent: defined once, by the bf16 pass
size_t rd_bn_workspace_bytes(int C) { return (size_t)bn::MAX_BLOCKS * 2 * (size_t)(C > 0 ? C : 0) * sizeof(float); }
#endif

int RD_ACT_FN(rd_bn_train_stats_nhwc_, )(const void* z_pad, int N, int H, int W, int C, const float* gamma, const float* beta,
                                float eps, float momentum, float* moving_mean, float* moving_var, float* coef,
                                void* workspace, size_t workspace_bytes, rd_stream_t stream) {
  if (bn::check_shape("rd_bn_train_stats", N, H, W, C)) return 1;
  RD_REQUIRE(z_pad && coef && workspace, "rd_bn_train_stats: null pointer");
  RD_REQUIRE(workspace_bytes >= rd_bn_workspace_bytes(C), "rd_bn_train_stats: workspace too small");
  if (rd_check_device()) return 1;
  const bn::Geo g = bn::make_geo(N, H, W, C);
  int grid = bn::grid_for(g);
  cudaStream_t s = rd::as_stream(stream);
  if (bn::stream_enabled()) {
    const bn::SGeo sg = bn::make_sgeo(N, H, W, C, 1);
    const size_t smem = bn::sgeo_smem(sg, 1);
    if (bn::stream_prepare(bn::s_stats_kernel, smem)) return 1;
    grid = bn::stream_grid(sg);
    RD_CUDA(rd::launch(bn::s_stats_kernel, dim3(grid), dim3(bn::SNT), smem, s, static_cast<const act_t*>(z_pad), sg,
                       static_cast<float*>(workspace), bn::stream_rev(0)));
  } else {
    bn::stats_kernel<<<grid, g.ppb * g.cgs, 0, s>>>(static_cast<const act_t*>(z_pad), g, 1,
                                                     static_cast<float*>(workspace));
  }
  RD_CUDA_LAUNCH_FINALIZE(bn::fwd_finalize_kernel, s, static_cast<const float*>(workspace), grid, C,
                                                          (double)N * H * W, gamma, beta, eps, momentum, moving_mean,
                                                          moving_var, coef);
  rd::count_launch(2);
  return rd::check_launch("rd_bn_train_stats");
}

#ifndef RD_ACT_F16   // storage-type independent: defined once, by the bf16 pass
// Second half of rd_bn_train_stats on partial sums produced elsewhere (the fused epilogue of rd_conv2d_nhwc_*_stats):
// partial[(which*C + c) * 1184 + slot], slot < nslots.
int rd_bn_train_finalize(const float* partial, int nslots, int N, int H, int W, int C, const float* gamma, const float* beta,
                         float eps, float momentum, float* moving_mean, float* moving_var, float* coef, rd_stream_t stream) {
  if (bn::check_shape("rd_bn_train_finalize", N, H, W, C)) return 1;
  RD_REQUIRE(partial && coef, "rd_bn_train_finalize: null pointer");
  RD_REQUIRE(nslots > 0 && nslots <= bn::MAX_BLOCKS, "rd_bn_train_finalize: nslots %d out of range", nslots);
  if (rd_check_device()) return 1;
  RD_CUDA_LAUNCH_FINALIZE(bn::fwd_finalize_kernel, rd::as_stream(stream), partial, nslots, C, (double)N * H * W, gamma, beta, eps,
                                                                         momentum, moving_mean, moving_var, coef);
  rd::count_launch();
  return rd::check_launch("rd_bn_train_finalize");
}
#endif

int RD_ACT_FN(rd_bn_act_fwd_nhwc_, )(const void* z_pad, const float* coef, const void* res_before, const void* res_after,
                            void* y_pad, int N, int H, int W, int C, int relu, rd_stream_t stream) {
  if (bn::check_shape("rd_bn_act_fwd", N, H, W, C)) return 1;
  RD_REQUIRE(z_pad && coef && y_pad, "rd_bn_act_fwd: null pointer");
  if (rd_check_device()) return 1;
  const bn::Geo g = bn::make_geo(N, H, W, C);
  const int64_t units = (int64_t)g.N * g.H * g.nseg;
  const int grid = (int)(units < 8 * 148 ? units : 8 * 148);
  if (bn::stream_enabled()) {
    const int nt = 1 + (res_before ? 1 : 0) + (res_after ? 1 : 0);
    const bn::SGeo sg = bn::make_sgeo(N, H, W, C, nt);
    const size_t smem = bn::sgeo_smem(sg, nt);
    if (bn::stream_prepare(bn::s_fwd_apply_kernel, smem)) return 1;
    RD_CUDA(rd::launch(bn::s_fwd_apply_kernel, dim3(bn::stream_grid(sg)), dim3(bn::SNT), smem, rd::as_stream(stream),
                       static_cast<const act_t*>(z_pad), coef, static_cast<const act_t*>(res_before),
                       static_cast<const act_t*>(res_after), static_cast<act_t*>(y_pad), sg, relu ? 1 : 0,
                       bn::stream_rev(1)));
  } else {
    bn::fwd_apply_kernel<<<grid, g.ppb * g.cgs, 0, rd::as_stream(stream)>>>(
        static_cast<const act_t*>(z_pad), coef, static_cast<const act_t*>(res_before),
        static_cast<const act_t*>(res_after), static_cast<act_t*>(y_pad), g, relu ? 1 : 0);
  }
  rd::count_launch();
  return rd::check_launch("rd_bn_act_fwd");
}

int RD_ACT_FN(rd_bn_act_bwd_nhwc_, )(const void* dy_pad, const void* y_mask_pad, const void* z_pad, const float* coef,
                            int mask_mode, void* dz_pad, int dz_halo_w, void* g_out_pad, float* dgamma,
                            float* dbeta, int N, int H, int W, int C, void* workspace, size_t workspace_bytes,
                            rd_stream_t stream) {
  if (bn::check_shape("rd_bn_act_bwd", N, H, W, C)) return 1;
  RD_REQUIRE(dy_pad && z_pad && coef && dz_pad && workspace, "rd_bn_act_bwd: null pointer");
  RD_REQUIRE(mask_mode >= 0 && mask_mode <= 2, "rd_bn_act_bwd: mask_mode must be 0, 1 or 2");
  RD_REQUIRE(mask_mode != 1 || y_mask_pad, "rd_bn_act_bwd: mask_mode 1 needs the forward output");
  RD_REQUIRE(dz_halo_w >= 1 && dz_halo_w <= 8, "rd_bn_act_bwd: dz_halo_w must be in [1,8]");
  RD_REQUIRE(workspace_bytes >= rd_bn_workspace_bytes(C) + 2 * (size_t)C * sizeof(float),
             "rd_bn_act_bwd: workspace too small");
  if (rd_check_device()) return 1;
  const bn::Geo g = bn::make_geo(N, H, W, C);
  const int grid = bn::grid_for(g);
  cudaStream_t s = rd::as_stream(stream);
  float* partial = static_cast<float*>(workspace);
  float* coef2 = partial + (size_t)bn::MAX_BLOCKS * 2 * C;
  const act_t* dy = static_cast<const act_t*>(dy_pad);
  const act_t* ym = static_cast<const act_t*>(y_mask_pad);
  const act_t* z = static_cast<const act_t*>(z_pad);
  if (bn::stream_enabled()) {
    const int nt = mask_mode == 1 ? 3 : 2;
    const bn::SGeo sg = bn::make_sgeo(N, H, W, C, nt);
    const size_t smem = bn::sgeo_smem(sg, nt);
    if (bn::stream_prepare(bn::s_bwd_reduce_kernel, smem) || bn::stream_prepare(bn::s_bwd_apply_kernel, smem)) return 1;
    const int sgrid = bn::stream_grid(sg);
    RD_CUDA(rd::launch(bn::s_bwd_reduce_kernel, dim3(sgrid), dim3(bn::SNT), smem, s, dy, ym, z, coef, sg, mask_mode, partial,
                       bn::stream_rev(2)));
    RD_CUDA_LAUNCH_FINALIZE(bn::bwd_finalize_kernel, s, partial, sgrid, C, (double)N * H * W, coef, coef2, dgamma,
                                                            dbeta);
    RD_CUDA(rd::launch(bn::s_bwd_apply_kernel, dim3(sgrid), dim3(bn::SNT), smem, s, dy, ym, z, coef, (const float*)coef2, sg,
                       mask_mode, static_cast<act_t*>(dz_pad), dz_halo_w, static_cast<act_t*>(g_out_pad), bn::stream_rev(3)));
  } else {
    bn::bwd_reduce_kernel<<<grid, g.ppb * g.cgs, 0, s>>>(dy, ym, z, coef, g, mask_mode, partial);
    RD_CUDA_LAUNCH_FINALIZE(bn::bwd_finalize_kernel, s, partial, grid, C, (double)N * H * W, coef, coef2, dgamma,
                                                            dbeta);
    const int64_t units = (int64_t)g.N * g.H * g.nseg;
    const int grid2 = (int)(units < 8 * 148 ? units : 8 * 148);
    bn::bwd_apply_kernel<<<grid2, g.ppb * g.cgs, 0, s>>>(dy, ym, z, coef, coef2, g, mask_mode,
                                                         static_cast<act_t*>(dz_pad), dz_halo_w,
                                                         static_cast<act_t*>(g_out_pad));
  }
  rd::count_launch(3);
  return rd::check_launch("rd_bn_act_bwd");
}

// rd_bn_act_bwd with the two sums already accumulated elsewhere (the epilogue of the convolution that produced dy:
// rd_conv2d_nhwc_*_bwdstats): finalize + apply only, one full read of dy and z less.
int RD_ACT_FN(rd_bn_act_bwd_apply_nhwc_, )(const void* dy_pad, const void* y_mask_pad, const void* z_pad, const float* coef,
                                  int mask_mode, const float* sums_partial, int sums_slots, void* dz_pad, int dz_halo_w,
                                  void* g_out_pad, float* dgamma, float* dbeta, int N, int H, int W, int C, void* workspace,
                                  size_t workspace_bytes, rd_stream_t stream) {
  if (bn::check_shape("rd_bn_act_bwd_apply", N, H, W, C)) return 1;
  RD_REQUIRE(dy_pad && z_pad && coef && dz_pad && workspace && sums_partial, "rd_bn_act_bwd_apply: null pointer");
  RD_REQUIRE(sums_slots > 0 && sums_slots <= bn::MAX_BLOCKS, "rd_bn_act_bwd_apply: bad number of partial slots (%d)", sums_slots);
  RD_REQUIRE(mask_mode >= 0 && mask_mode <= 2, "rd_bn_act_bwd_apply: mask_mode must be 0, 1 or 2");
  RD_REQUIRE(mask_mode != 1 || y_mask_pad, "rd_bn_act_bwd_apply: mask_mode 1 needs the forward output");
  RD_REQUIRE(dz_halo_w >= 1 && dz_halo_w <= 8, "rd_bn_act_bwd_apply: dz_halo_w must be in [1,8]");
  RD_REQUIRE(workspace_bytes >= 2 * (size_t)C * sizeof(float), "rd_bn_act_bwd_apply: workspace too small");
  if (rd_check_device()) return 1;
  cudaStream_t s = rd::as_stream(stream);
  float* coef2 = static_cast<float*>(workspace);
  const act_t* dy = static_cast<const act_t*>(dy_pad);
  const act_t* ym = static_cast<const act_t*>(y_mask_pad);
  const act_t* z = static_cast<const act_t*>(z_pad);
  RD_CUDA_LAUNCH_FINALIZE(bn::bwd_finalize_kernel, s, sums_partial, sums_slots, C, (double)N * H * W, coef, coef2, dgamma, dbeta);
  if (bn::stream_enabled()) {
    const int nt = mask_mode == 1 ? 3 : 2;
    const bn::SGeo sg = bn::make_sgeo(N, H, W, C, nt);
    const size_t smem = bn::sgeo_smem(sg, nt);
    if (bn::stream_prepare(bn::s_bwd_apply_kernel, smem)) return 1;
    RD_CUDA(rd::launch(bn::s_bwd_apply_kernel, dim3(bn::stream_grid(sg)), dim3(bn::SNT), smem, s, dy, ym, z, coef, (const float*)coef2,
                       sg, mask_mode, static_cast<act_t*>(dz_pad), dz_halo_w, static_cast<act_t*>(g_out_pad), bn::stream_rev(3)));
  } else {
    const bn::Geo g = bn::make_geo(N, H, W, C);
    const int64_t units = (int64_t)g.N * g.H * g.nseg;
    const int grid2 = (int)(units < 8 * 148 ? units : 8 * 148);
    bn::bwd_apply_kernel<<<grid2, g.ppb * g.cgs, 0, s>>>(dy, ym, z, coef, coef2, g, mask_mode, static_cast<act_t*>(dz_pad),
                                                         dz_halo_w, static_cast<act_t*>(g_out_pad));
  }
  rd::count_launch(2);
  return rd::check_launch("rd_bn_act_bwd_apply");
}

int RD_ACT_FN(rd_channel_sums_nhwc_, )(const void* x_pad, int N, int H, int W, int C, float* sums, void* workspace,
                              size_t workspace_bytes, rd_stream_t stream) {
  if (bn::check_shape("rd_channel_sums", N, H, W, C)) return 1;
  RD_REQUIRE(x_pad && sums && workspace, "rd_channel_sums: null pointer");
  RD_REQUIRE(workspace_bytes >= rd_bn_workspace_bytes(C) + 6 * (size_t)C * sizeof(float),
             "rd_channel_sums: workspace too small");
  if (rd_check_device()) return 1;
  const bn::Geo g = bn::make_geo(N, H, W, C);
  const int grid = bn::grid_for(g);
  cudaStream_t s = rd::as_stream(stream);
  float* partial = static_cast<float*>(workspace);
  float* coef = partial + (size_t)bn::MAX_BLOCKS * 2 * C;
  bn::stats_kernel<<<grid, g.ppb * g.cgs, 0, s>>>(static_cast<const act_t*>(x_pad), g, 1, partial);
  RD_CUDA_LAUNCH_FINALIZE(bn::fwd_finalize_kernel, s, partial, grid, C, (double)N * H * W, nullptr, nullptr, 1e-5f,
                                                          0.f, nullptr, nullptr, coef);
  RD_CUDA(cudaMemcpyAsync(sums, coef + 5 * (size_t)C, (size_t)C * sizeof(float), cudaMemcpyDeviceToDevice, s));
  rd::count_launch(2);
  return rd::check_launch("rd_channel_sums");
}

int RD_ACT_FN(rd_add_nhwc_, )(const void* x0_pad, const void* x1_pad, void* y_pad, int N, int H, int W, int C,
                     rd_stream_t stream) {
  if (bn::check_shape(RD_ACT_FN_STR(rd_add_nhwc_, ) "", N, H, W, C)) return 1;
  RD_REQUIRE(x0_pad && x1_pad && y_pad, RD_ACT_FN_STR(rd_add_nhwc_, ) ": null pointer");
  if (rd_check_device()) return 1;
  const bn::Geo g = bn::make_geo(N, H, W, C);
  const int64_t units = (int64_t)g.N * g.H * g.nseg;
  const int grid = (int)(units < 8 * 148 ? units : 8 * 148);
  bn::add_kernel<<<grid, g.ppb * g.cgs, 0, rd::as_stream(stream)>>>(static_cast<const act_t*>(x0_pad),
                                                                    static_cast<const act_t*>(x1_pad),
                                                                    static_cast<act_t*>(y_pad), g);
  rd::count_launch();
  return rd::check_launch(RD_ACT_FN_STR(rd_add_nhwc_, ) "");
}

}  // extern "C"
