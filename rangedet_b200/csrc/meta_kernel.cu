// Meta-Kernel (dynamic 3x3 convolution) for sm_100a -- impl 1: fused CUDA-core fp32 kernels.
//
// Replaces MetaKernel.meta_baseline_bias, /root/reference rangedet/symbol/backbone/
// meta_kernel.py:166-240, which the reference runs as ~10 MXNet library ops that materialise
// six (B, 27..576, H*W) intermediates in HBM.  Here one kernel per direction:
//
//   forward   : coords tile -> relative xyz -> hidden (3->32, ReLU) in shared memory -> per-tap
//               32->C register-tiled mini-GEMM -> x data neighbour -> 9C output planes
//   grad_data : same MLP evaluated on the mirrored tap (gather form, no atomics)
//   grad_mlp  : recompute hidden, gw = grad_out*data, three small reductions (gW1, ghid, gW0)
//               accumulated in registers by persistent CTAs, deterministic 2-stage reduce
//
// HBM traffic at the op boundary (fp32): fwd 2572 B/pixel, bwd 2828 B/pixel algorithmic
// (SURVEY.md 8d); this implementation reads grad_out twice in bwd (5132 B/pixel).
// The TMA + tcgen05 warp-specialised variant (impl 3, the default for C == 64) lives in meta_kernel_ws*.cu.
#include "../../include/rangedet_b200.h"
#include "rd_common.cuh"

namespace mk {

constexpr int HID = 32;  // hidden width of the coordinate MLP (channel_list[0], config :95-103)
constexpr int CCH = 3;   // coordinate channels
constexpr int NT = 256;  // threads per CTA
constexpr int MAXC = 64;

// ------------------------------------------------------------------------------------------
// shared helpers
// ------------------------------------------------------------------------------------------
// coordinate tile with a one-pixel zero halo: rows h-1..h+1, columns w0-1..w0+TW
template <int TW>
__device__ __forceinline__ void load_coord_tile(float* cs, const float* __restrict__ coord, int b,
                                                int h, int w0, int H, int W) {
  constexpr int ROW = TW + 2;
  for (int e = threadIdx.x; e < 3 * CCH * ROW; e += NT) {
    const int col = e % ROW;
    const int d = (e / ROW) % CCH;
    const int r = e / (ROW * CCH);
    const int hh = h + r - 1, ww = w0 + col - 1;
    float v = 0.f;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(coord + (((int64_t)b * CCH + d) * H + hh) * W + ww);
    cs[(r * CCH + d) * ROW + col] = v;
  }
}

// params -> shared: W1 transposed to [j][c], {W0[j][0..2], b0[j]} packed as float4
__device__ __forceinline__ void load_params(float* w1t, float* b1s, float4* w0b,
                                            const float* __restrict__ w0, const float* __restrict__ b0,
                                            const float* __restrict__ w1, const float* __restrict__ b1,
                                            int C) {
  for (int e = threadIdx.x; e < C * HID; e += NT) {
    const int c = e / HID, j = e % HID;
    w1t[j * C + c] = __ldg(w1 + e);
  }
  for (int e = threadIdx.x; e < C; e += NT) b1s[e] = __ldg(b1 + e);
  for (int j = threadIdx.x; j < HID; j += NT)
    w0b[j] = make_float4(__ldg(w0 + j * 3 + 0), __ldg(w0 + j * 3 + 1), __ldg(w0 + j * 3 + 2), __ldg(b0 + j));
}

// hidden[j][px] = relu(W0[j] . rel + b0[j]) for one pixel, all 32 units
template <int TW>
__device__ __forceinline__ void hidden_column(float* hid, const float4* w0b, int px, float r0, float r1,
                                              float r2) {
#pragma unroll
  for (int j = 0; j < HID; ++j) {
    const float4 w = w0b[j];
    float z = w.w;
    z = fmaf(w.x, r0, z);
    z = fmaf(w.y, r1, z);
    z = fmaf(w.z, r2, z);
    hid[j * TW + px] = fmaxf(z, 0.f);
  }
}

// acc[i][cc] = sum_j hid[j][lane+32i] * W1[c0+cc][j]   (PXT pixels x 8 channels per thread)
template <int TW, int PXT>
__device__ __forceinline__ void weight_tile(float (&acc)[PXT][8], const float* hid, const float* w1t,
                                            int C, int c0, int lane) {
#pragma unroll
  for (int i = 0; i < PXT; ++i)
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) acc[i][cc] = 0.f;
#pragma unroll 4
  for (int j = 0; j < HID; ++j) {
    const float4 wa = *reinterpret_cast<const float4*>(w1t + j * C + c0);
    const float4 wb = *reinterpret_cast<const float4*>(w1t + j * C + c0 + 4);
    const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
    for (int i = 0; i < PXT; ++i) {
      const float hv = hid[j * TW + lane + 32 * i];
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) acc[i][cc] = fmaf(hv, wv[cc], acc[i][cc]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
constexpr int F_PXT = 8;
constexpr int F_TW = 32 * F_PXT;  // 256 pixels of one image row per CTA

struct FwdSmem {
  alignas(16) float w1t[HID * MAXC];
  alignas(16) float b1s[MAXC];
  alignas(16) float4 w0b[HID];
  alignas(16) float cs[3 * CCH * (F_TW + 2)];
  alignas(16) float hid[HID * F_TW];
};

__global__ void __launch_bounds__(NT, 2)
meta_fwd_kernel(const float* __restrict__ data, const float* __restrict__ coord,
                const float* __restrict__ w0, const float* __restrict__ b0,
                const float* __restrict__ w1, const float* __restrict__ b1, float* __restrict__ out,
                int B, int C, int H, int W, int tiles_w) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& S = *reinterpret_cast<FwdSmem*>(smem_raw);
  constexpr int ROW = F_TW + 2;
  const int tile = blockIdx.x;
  const int wt = tile % tiles_w;
  const int h = (tile / tiles_w) % H;
  const int b = tile / (tiles_w * H);
  const int w0px = wt * F_TW;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

  load_params(S.w1t, S.b1s, S.w0b, w0, b0, w1, b1, C);
  load_coord_tile<F_TW>(S.cs, coord, b, h, w0px, H, W);
  __syncthreads();

  const float c0 = S.cs[(1 * CCH + 0) * ROW + t + 1];
  const float c1 = S.cs[(1 * CCH + 1) * ROW + t + 1];
  const float c2 = S.cs[(1 * CCH + 2) * ROW + t + 1];
  const int ncg = C >> 3;
  const int64_t plane = (int64_t)H * W;

  for (int k = 0; k < 9; ++k) {
    const int dy = k / 3 - 1, dx = k % 3 - 1;
    {  // phase A: hidden activations of tap k for pixel t
      const int col = t + 1 + dx, r = dy + 1;
      const float r0 = S.cs[(r * CCH + 0) * ROW + col] - c0;
      const float r1 = S.cs[(r * CCH + 1) * ROW + col] - c1;
      const float r2 = S.cs[(r * CCH + 2) * ROW + col] - c2;
      hidden_column<F_TW>(S.hid, S.w0b, t, r0, r1, r2);
    }
    __syncthreads();
    const int nh = h + dy;
    const bool row_ok = nh >= 0 && nh < H;
    for (int cg = warp; cg < ncg; cg += NT / 32) {
      float acc[F_PXT][8];
      weight_tile<F_TW, F_PXT>(acc, S.hid, S.w1t, C, cg * 8, lane);
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        const int c = cg * 8 + cc;
        const float bias = S.b1s[c];
        const float* dplane = data + ((int64_t)b * C + c) * plane + (int64_t)nh * W;
        float* oplane = out + (((int64_t)b * C + c) * 9 + k) * plane + (int64_t)h * W;
#pragma unroll
        for (int i = 0; i < F_PXT; ++i) {
          const int w = w0px + lane + 32 * i;
          const int nw = w + dx;
          if (w < W) {
            float dv = 0.f;
            if (row_ok && nw >= 0 && nw < W) dv = __ldg(dplane + nw);
            __stcs(oplane + w, dv * (acc[i][cc] + bias));
          }
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// backward w.r.t. data (gather form)
//   grad_data[q,c] = sum_k' go[p, c*9 + (8-k')] * MLP(coord[q] - coord[p])[c],  p = q + delta(k')
// ------------------------------------------------------------------------------------------
constexpr int G_PXT = 4;
constexpr int G_TW = 32 * G_PXT;  // 128

struct GdSmem {
  alignas(16) float w1t[HID * MAXC];
  alignas(16) float b1s[MAXC];
  alignas(16) float4 w0b[HID];
  alignas(16) float cs[3 * CCH * (G_TW + 2)];
  alignas(16) float hid[HID * G_TW];
};

__global__ void __launch_bounds__(NT, 2)
meta_bwd_data_kernel(const float* __restrict__ go, const float* __restrict__ coord,
                     const float* __restrict__ w0, const float* __restrict__ b0,
                     const float* __restrict__ w1, const float* __restrict__ b1,
                     float* __restrict__ gdata, int B, int C, int H, int W, int tiles_w) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GdSmem& S = *reinterpret_cast<GdSmem*>(smem_raw);
  constexpr int ROW = G_TW + 2;
  const int tile = blockIdx.x;
  const int wt = tile % tiles_w;
  const int h = (tile / tiles_w) % H;
  const int b = tile / (tiles_w * H);
  const int w0px = wt * G_TW;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

  load_params(S.w1t, S.b1s, S.w0b, w0, b0, w1, b1, C);
  load_coord_tile<G_TW>(S.cs, coord, b, h, w0px, H, W);
  __syncthreads();

  const int ncg = C >> 3;  // <= 8 == number of warps: one channel group per warp
  const int cg = warp;
  const int64_t plane = (int64_t)H * W;
  float gd[G_PXT][8];
#pragma unroll
  for (int i = 0; i < G_PXT; ++i)
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) gd[i][cc] = 0.f;

  // phase A work split: thread -> (pixel t%128, hidden half t/128)
  const int apx = t & (G_TW - 1), ahalf = t >> 7;
  const float c0 = S.cs[(1 * CCH + 0) * ROW + apx + 1];
  const float c1 = S.cs[(1 * CCH + 1) * ROW + apx + 1];
  const float c2 = S.cs[(1 * CCH + 2) * ROW + apx + 1];

  for (int k = 0; k < 9; ++k) {
    const int dy = k / 3 - 1, dx = k % 3 - 1;
    {
      const int col = apx + 1 + dx, r = dy + 1;
      // rel seen by pixel p for its tap (8-k), whose neighbour is q: coord[q] - coord[p]
      const float r0 = c0 - S.cs[(r * CCH + 0) * ROW + col];
      const float r1 = c1 - S.cs[(r * CCH + 1) * ROW + col];
      const float r2 = c2 - S.cs[(r * CCH + 2) * ROW + col];
#pragma unroll
      for (int jj = 0; jj < HID / 2; ++jj) {
        const int j = ahalf * (HID / 2) + jj;
        const float4 w = S.w0b[j];
        float z = w.w;
        z = fmaf(w.x, r0, z);
        z = fmaf(w.y, r1, z);
        z = fmaf(w.z, r2, z);
        S.hid[j * G_TW + apx] = fmaxf(z, 0.f);
      }
    }
    __syncthreads();
    if (cg < ncg) {
      const int ph = h + dy;
      const bool row_ok = ph >= 0 && ph < H;
      float acc[G_PXT][8];
      weight_tile<G_TW, G_PXT>(acc, S.hid, S.w1t, C, cg * 8, lane);
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        const int c = cg * 8 + cc;
        const float bias = S.b1s[c];
        const float* gplane = go + (((int64_t)b * C + c) * 9 + (8 - k)) * plane + (int64_t)ph * W;
#pragma unroll
        for (int i = 0; i < G_PXT; ++i) {
          const int w = w0px + lane + 32 * i;
          const int pw = w + dx;
          float g = 0.f;
          if (row_ok && w < W && pw >= 0 && pw < W) g = __ldg(gplane + pw);
          gd[i][cc] = fmaf(g, acc[i][cc] + bias, gd[i][cc]);
        }
      }
    }
    __syncthreads();
  }
  if (cg < ncg) {
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      float* gp = gdata + ((int64_t)b * C + cg * 8 + cc) * plane + (int64_t)h * W;
#pragma unroll
      for (int i = 0; i < G_PXT; ++i) {
        const int w = w0px + lane + 32 * i;
        if (w < W) gp[w] = gd[i][cc];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward w.r.t. the MLP parameters
// ------------------------------------------------------------------------------------------
constexpr int P_TW = 128;
constexpr int P_GRID = 296;       // persistent CTAs (2 per SM on a 148-SM B200)
constexpr int GWS = MAXC + 4;     // row stride of gw[m][c]
constexpr int HS = HID + 4;       // row stride of hid[m][j] / ghid[m][j]
constexpr int P_NOUT_MAX = MAXC * HID + MAXC + HID * 4;

struct GpSmem {
  alignas(16) float w1n[MAXC * HID];  // W1[c][j] (natural layout)
  alignas(16) float4 w0b[HID];
  alignas(16) float cs[3 * CCH * (P_TW + 2)];
  alignas(16) float gw[P_TW * GWS];
  alignas(16) float hid[P_TW * HS];
  alignas(16) float ghid[P_TW * HS];
  alignas(16) float rel[P_TW * 4];
};

__global__ void __launch_bounds__(NT, 2)
meta_bwd_param_kernel(const float* __restrict__ go, const float* __restrict__ data,
                      const float* __restrict__ coord, const float* __restrict__ w0,
                      const float* __restrict__ b0, const float* __restrict__ w1,
                      float* __restrict__ partial, int B, int C, int H, int W, int tiles_w,
                      int ntiles) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GpSmem& S = *reinterpret_cast<GpSmem*>(smem_raw);
  constexpr int ROW = P_TW + 2;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int64_t plane = (int64_t)H * W;

  for (int e = t; e < C * HID; e += NT) S.w1n[e] = __ldg(w1 + e);
  for (int j = t; j < HID; j += NT)
    S.w0b[j] = make_float4(__ldg(w0 + j * 3 + 0), __ldg(w0 + j * 3 + 1), __ldg(w0 + j * 3 + 2), __ldg(b0 + j));

  // persistent register accumulators
  float aW1[4][4];  // gW1[c = 4*cgB + cc][j = 4*jgB + jj], over this thread's half of the pixels
  float aB1[4];
  float aW0 = 0.f;  // gW0[j][d] (d==3 -> gb0[j]) over this thread's half of the pixels
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    aB1[a] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) aW1[a][c] = 0.f;
  }
  // phase A mapping
  const int apx = t & (P_TW - 1), ahalf = t >> 7;
  // phase B1 mapping (ghid GEMM): 4 pixels x 4 hidden units
  const int jg1 = lane & 7, mg1 = warp * 4 + (lane >> 3);
  // phase B2 mapping (gW1): 4 channels x 4 hidden units over 64 pixels
  const int half2 = t >> 7, cgB = (t & 127) & 15, jgB = (t & 127) >> 4;
  const bool b2_on = cgB * 4 < C;
  // phase C mapping (gW0/gb0)
  const int jC = t & 31, dC = (t >> 5) & 3, halfC = t >> 7;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int wt = tile % tiles_w;
    const int h = (tile / tiles_w) % H;
    const int b = tile / (tiles_w * H);
    const int w0px = wt * P_TW;
    __syncthreads();
    load_coord_tile<P_TW>(S.cs, coord, b, h, w0px, H, W);
    __syncthreads();
    const float c0 = S.cs[(1 * CCH + 0) * ROW + apx + 1];
    const float c1 = S.cs[(1 * CCH + 1) * ROW + apx + 1];
    const float c2 = S.cs[(1 * CCH + 2) * ROW + apx + 1];
    const int w = w0px + apx;

    for (int k = 0; k < 9; ++k) {
      const int dy = k / 3 - 1, dx = k % 3 - 1;
      {  // ---- phase A
        const int col = apx + 1 + dx, r = dy + 1;
        const float r0 = S.cs[(r * CCH + 0) * ROW + col] - c0;
        const float r1 = S.cs[(r * CCH + 1) * ROW + col] - c1;
        const float r2 = S.cs[(r * CCH + 2) * ROW + col] - c2;
        if (ahalf == 0) *reinterpret_cast<float4*>(S.rel + apx * 4) = make_float4(r0, r1, r2, 1.0f);
#pragma unroll
        for (int jj = 0; jj < HID / 2; ++jj) {
          const int j = ahalf * (HID / 2) + jj;
          const float4 wv = S.w0b[j];
          float z = wv.w;
          z = fmaf(wv.x, r0, z);
          z = fmaf(wv.y, r1, z);
          z = fmaf(wv.z, r2, z);
          S.hid[apx * HS + j] = fmaxf(z, 0.f);
        }
        const int nh = h + dy, nw = w + dx;
        const bool ok = w < W && nh >= 0 && nh < H && nw >= 0 && nw < W;
        const int chalf = C >> 1;
        for (int cc = 0; cc < chalf; ++cc) {
          const int c = ahalf * chalf + cc;
          float g = 0.f;
          if (ok) {
            const float gv = __ldg(go + (((int64_t)b * C + c) * 9 + k) * plane + (int64_t)h * W + w);
            const float dv = __ldg(data + ((int64_t)b * C + c) * plane + (int64_t)nh * W + nw);
            g = gv * dv;
          }
          S.gw[apx * GWS + c] = g;
        }
      }
      __syncthreads();
      {  // ---- phase B1: ghid[m][j] = (hid>0) * sum_c gw[m][c] W1[c][j]
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.f;
        for (int c4 = 0; c4 < C; c4 += 4) {
          float4 wr[4];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
            wr[cc] = *reinterpret_cast<const float4*>(S.w1n + (c4 + cc) * HID + jg1 * 4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 g = *reinterpret_cast<const float4*>(S.gw + (mg1 * 4 + i) * GWS + c4);
            const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              acc[i][0] = fmaf(gv[cc], wr[cc].x, acc[i][0]);
              acc[i][1] = fmaf(gv[cc], wr[cc].y, acc[i][1]);
              acc[i][2] = fmaf(gv[cc], wr[cc].z, acc[i][2]);
              acc[i][3] = fmaf(gv[cc], wr[cc].w, acc[i][3]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int m = mg1 * 4 + i;
          const float4 hv = *reinterpret_cast<const float4*>(S.hid + m * HS + jg1 * 4);
          float4 o;
          o.x = hv.x > 0.f ? acc[i][0] : 0.f;
          o.y = hv.y > 0.f ? acc[i][1] : 0.f;
          o.z = hv.z > 0.f ? acc[i][2] : 0.f;
          o.w = hv.w > 0.f ? acc[i][3] : 0.f;
          *reinterpret_cast<float4*>(S.ghid + m * HS + jg1 * 4) = o;
        }
      }
      if (b2_on) {  // ---- phase B2: gW1 += gw^T hid ; gb1 += sum gw
        const int m0 = half2 * (P_TW / 2);
#pragma unroll 4
        for (int mm = 0; mm < P_TW / 2; ++mm) {
          const int m = m0 + mm;
          const float4 g = *reinterpret_cast<const float4*>(S.gw + m * GWS + cgB * 4);
          const float4 hv = *reinterpret_cast<const float4*>(S.hid + m * HS + jgB * 4);
          const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            aW1[cc][0] = fmaf(gv[cc], hv.x, aW1[cc][0]);
            aW1[cc][1] = fmaf(gv[cc], hv.y, aW1[cc][1]);
            aW1[cc][2] = fmaf(gv[cc], hv.z, aW1[cc][2]);
            aW1[cc][3] = fmaf(gv[cc], hv.w, aW1[cc][3]);
            aB1[cc] += gv[cc];
          }
        }
      }
      __syncthreads();
      {  // ---- phase C: gW0[j][d] += sum_m ghid[m][j] * rel[m][d]   (rel[m][3] = 1 -> gb0)
        const int m0 = halfC * (P_TW / 2);
#pragma unroll 8
        for (int mm = 0; mm < P_TW / 2; ++mm) {
          const int m = m0 + mm;
          aW0 = fmaf(S.ghid[m * HS + jC], S.rel[m * 4 + dC], aW0);
        }
      }
      __syncthreads();
    }
  }
  // write partials: two "virtual CTAs" (pixel halves) per CTA
  const int nout = C * HID + C + HID * 4;
  {
    float* p = partial + (int64_t)(blockIdx.x * 2 + half2) * nout;
    if (b2_on) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c = cgB * 4 + cc;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) p[c * HID + jgB * 4 + jj] = aW1[cc][jj];
        if (jgB == 0) p[C * HID + c] = aB1[cc];
      }
    }
  }
  {
    float* p = partial + (int64_t)(blockIdx.x * 2 + halfC) * nout;
    p[C * HID + C + jC * 4 + dC] = aW0;
  }
}

__global__ void meta_bwd_param_reduce_kernel(const float* __restrict__ partial, int nparts, int C,
                                             float* __restrict__ gw0, float* __restrict__ gb0,
                                             float* __restrict__ gw1, float* __restrict__ gb1) {
  const int nout = C * HID + C + HID * 4;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= nout) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += partial[(int64_t)p * nout + o];  // fixed order: deterministic
  if (o < C * HID) gw1[o] = s;
  else if (o < C * HID + C) gb1[o - C * HID] = s;
  else {
    const int r = o - C * HID - C, j = r >> 2, d = r & 3;
    if (d < 3) gw0[j * 3 + d] = s;
    else gb0[j] = s;
  }
}

}  // namespace mk

// implemented in meta_kernel_ws.cu (impl 3)
int rd_meta_kernel_fwd_ws(const float* data, const float* coord, const float* w0, const float* b0, const float* w1,
                          const float* b1, float* out, int B, int C, int H, int W, cudaStream_t stream);
int rd_meta_kernel_bwd_data_ws(const float* grad_out, const float* coord, const float* w0, const float* b0,
                               const float* w1, const float* b1, float* grad_data, int B, int C, int H, int W,
                               cudaStream_t stream);
int rd_meta_kernel_bwd_params_ws(const float* grad_out, const float* data, const float* coord, const float* w0,
                                 const float* b0, const float* w1, float* partial, int* nparts, int B, int C, int H,
                                 int W, cudaStream_t stream);

int rd_meta_kernel_bwd_data_ws_nhwc(const void* grad_out_pad, int gof, const float* coord, const float* w0, const float* b0,
                                    const float* w1, const float* b1, float* grad_data, int B, int C, int H, int W,
                                    cudaStream_t stream);
int rd_meta_kernel_bwd_params_ws_nhwc(const void* grad_out_pad, int gof, const float* data, const float* coord,
                                      const float* w0, const float* b0, const float* w1, float* partial, int* nparts, int B,
                                      int C, int H, int W, cudaStream_t stream);

// Backward from the haloed NHWC tap-major gradient (the layout rd_meta_kernel_fwd_nhwc_* writes and the BatchNorm backward
// of the training graph hands back): both kernels read the 2-byte tensor directly.
static int meta_bwd_nhwc(const char* who, const void* grad_out_pad, int gof, const float* data, const float* coord,
                         const float* w0, const float* b0, const float* w1, const float* b1, float* grad_data, float* grad_w0,
                         float* grad_b0, float* grad_w1, float* grad_b1, void* workspace, size_t workspace_bytes, int B,
                         int C, int H, int W, rd_stream_t stream) {
  RD_REQUIRE(B > 0 && H > 0 && W > 0, "%s: bad shape B=%d H=%d W=%d", who, B, H, W);
  RD_REQUIRE(C == 64 && W % 4 == 0, "%s: needs C == 64 and W %% 4 == 0 (got C=%d W=%d)", who, C, W);
  RD_REQUIRE(grad_out_pad && data && coord && w0 && b0 && w1 && b1, "%s: null input pointer", who);
  const bool want_params = grad_w0 || grad_b0 || grad_w1 || grad_b1;
  RD_REQUIRE(!want_params || (grad_w0 && grad_b0 && grad_w1 && grad_b1), "%s: the four parameter gradients go together", who);
  RD_REQUIRE(grad_data || want_params, "%s: nothing to compute", who);
  RD_REQUIRE((reinterpret_cast<uintptr_t>(grad_out_pad) & 15) == 0 && (reinterpret_cast<uintptr_t>(data) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(grad_data) & 15) == 0,
             "%s: tensors must be 16-byte aligned", who);
  if (rd_check_device()) return 1;
  cudaStream_t st = rd::as_stream(stream);
  if (grad_data && rd_meta_kernel_bwd_data_ws_nhwc(grad_out_pad, gof, coord, w0, b0, w1, b1, grad_data, B, C, H, W, st)) return 1;
  if (!want_params) return 0;
  RD_REQUIRE(workspace && workspace_bytes >= rd_meta_kernel_bwd_workspace_bytes(B, C, H, W), "%s: workspace too small (%zu < %zu)",
             who, workspace_bytes, rd_meta_kernel_bwd_workspace_bytes(B, C, H, W));
  float* partial = static_cast<float*>(workspace);
  int nparts = 0;
  if (rd_meta_kernel_bwd_params_ws_nhwc(grad_out_pad, gof, data, coord, w0, b0, w1, partial, &nparts, B, C, H, W, st)) return 1;
  const int nout = C * mk::HID + C + mk::HID * 4;
  mk::meta_bwd_param_reduce_kernel<<<(nout + 255) / 256, 256, 0, st>>>(partial, nparts, C, grad_w0, grad_b0, grad_w1, grad_b1);
  rd::count_launch();
  return rd::check_launch(who);
}

extern "C" {

int rd_meta_kernel_bwd_nhwc_bf16(const void* grad_out_pad, const float* data, const float* coord, const float* w0,
                                 const float* b0, const float* w1, const float* b1, float* grad_data, float* grad_w0,
                                 float* grad_b0, float* grad_w1, float* grad_b1, void* workspace, size_t workspace_bytes,
                                 int B, int C, int H, int W, rd_stream_t stream) {
  return meta_bwd_nhwc("rd_meta_kernel_bwd_nhwc_bf16", grad_out_pad, 1, data, coord, w0, b0, w1, b1, grad_data, grad_w0, grad_b0,
                       grad_w1, grad_b1, workspace, workspace_bytes, B, C, H, W, stream);
}
int rd_meta_kernel_bwd_nhwc_f16(const void* grad_out_pad, const float* data, const float* coord, const float* w0,
                                const float* b0, const float* w1, const float* b1, float* grad_data, float* grad_w0,
                                float* grad_b0, float* grad_w1, float* grad_b1, void* workspace, size_t workspace_bytes,
                                int B, int C, int H, int W, rd_stream_t stream) {
  return meta_bwd_nhwc("rd_meta_kernel_bwd_nhwc_f16", grad_out_pad, 2, data, coord, w0, b0, w1, b1, grad_data, grad_w0, grad_b0,
                       grad_w1, grad_b1, workspace, workspace_bytes, B, C, H, W, stream);
}

int rd_meta_kernel_fwd(const float* data, const float* coord, const float* w0, const float* b0,
                       const float* w1, const float* b1, float* out, int B, int C, int H, int W,
                       int impl, rd_stream_t stream) {
  RD_REQUIRE(B >= 0 && H > 0 && W > 0, "rd_meta_kernel_fwd: bad shape B=%d H=%d W=%d", B, H, W);
  RD_REQUIRE(C > 0 && C % 8 == 0 && C <= mk::MAXC,
             "rd_meta_kernel_fwd: C must be a multiple of 8 and <= %d (got %d)", mk::MAXC, C);
  RD_REQUIRE(impl >= 0 && impl <= 3, "rd_meta_kernel_fwd: impl must be 0..3");
  if (B == 0) return 0;
  RD_REQUIRE(data && coord && w0 && b0 && w1 && b1 && out, "rd_meta_kernel_fwd: null pointer");
  if (rd_check_device()) return 1;
  if (impl == 0)  // default: TMA/tcgen05 kernel when its preconditions hold, generic fp32 kernel otherwise
    impl = (C == 64 && W % 4 == 0 && ((reinterpret_cast<uintptr_t>(data) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) ? 3 : 1;
  if (impl == 3) return rd_meta_kernel_fwd_ws(data, coord, w0, b0, w1, b1, out, B, C, H, W, rd::as_stream(stream));
  RD_REQUIRE(impl != 2, "rd_meta_kernel_fwd: impl 2 (first tcgen05 version) was removed; use 0 (default), 1 (fp32) or 3 (TMA + tcgen05)");
  const int tiles_w = (W + mk::F_TW - 1) / mk::F_TW;
  const int64_t ntiles = (int64_t)B * H * tiles_w;
  RD_REQUIRE(ntiles <= 0x7fffffffLL, "rd_meta_kernel_fwd: too many tiles");
  const size_t smem = sizeof(mk::FwdSmem);
  RD_CUDA(cudaFuncSetAttribute(mk::meta_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mk::meta_fwd_kernel<<<(unsigned)ntiles, mk::NT, smem, rd::as_stream(stream)>>>(
      data, coord, w0, b0, w1, b1, out, B, C, H, W, tiles_w);
  rd::count_launch();
  return rd::check_launch("rd_meta_kernel_fwd");
}

size_t rd_meta_kernel_bwd_workspace_bytes(int B, int C, int H, int W) {
  (void)B; (void)H; (void)W;
  if (C <= 0 || C > mk::MAXC) return 0;
  return (size_t)mk::P_GRID * 2 * (size_t)(C * mk::HID + C + mk::HID * 4) * sizeof(float);
}

int rd_meta_kernel_bwd_data(const float* grad_out, const float* coord, const float* w0, const float* b0,
                            const float* w1, const float* b1, float* grad_data, int B, int C, int H, int W,
                            int impl, rd_stream_t stream) {
  RD_REQUIRE(B >= 0 && H > 0 && W > 0, "rd_meta_kernel_bwd_data: bad shape B=%d H=%d W=%d", B, H, W);
  RD_REQUIRE(C > 0 && C % 8 == 0 && C <= mk::MAXC,
             "rd_meta_kernel_bwd_data: C must be a multiple of 8 and <= %d (got %d)", mk::MAXC, C);
  RD_REQUIRE(impl >= 0 && impl <= 3, "rd_meta_kernel_bwd_data: impl must be 0..3");
  if (B == 0) return 0;
  RD_REQUIRE(grad_out && coord && w0 && b0 && w1 && b1 && grad_data, "rd_meta_kernel_bwd_data: null pointer");
  if (rd_check_device()) return 1;
  if (impl == 0)
    impl = (C == 64 && W % 4 == 0 && ((reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(grad_data)) & 15) == 0) ? 3 : 1;
  if (impl == 3)
    return rd_meta_kernel_bwd_data_ws(grad_out, coord, w0, b0, w1, b1, grad_data, B, C, H, W, rd::as_stream(stream));
  const int tiles_w = (W + mk::G_TW - 1) / mk::G_TW;
  const int64_t ntiles = (int64_t)B * H * tiles_w;
  RD_REQUIRE(ntiles <= 0x7fffffffLL, "rd_meta_kernel_bwd_data: too many tiles");
  const size_t smem = sizeof(mk::GdSmem);
  RD_CUDA(cudaFuncSetAttribute(mk::meta_bwd_data_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mk::meta_bwd_data_kernel<<<(unsigned)ntiles, mk::NT, smem, rd::as_stream(stream)>>>(
      grad_out, coord, w0, b0, w1, b1, grad_data, B, C, H, W, tiles_w);
  rd::count_launch();
  return rd::check_launch("rd_meta_kernel_bwd_data");
}

int rd_meta_kernel_bwd_params(const float* grad_out, const float* data, const float* coord, const float* w0,
                              const float* b0, const float* w1, const float* b1, float* grad_w0,
                              float* grad_b0, float* grad_w1, float* grad_b1, void* workspace,
                              size_t workspace_bytes, int B, int C, int H, int W, int impl,
                              rd_stream_t stream) {
  (void)b1;
  RD_REQUIRE(B >= 0 && H > 0 && W > 0, "rd_meta_kernel_bwd_params: bad shape B=%d H=%d W=%d", B, H, W);
  RD_REQUIRE(C > 0 && C % 8 == 0 && C <= mk::MAXC,
             "rd_meta_kernel_bwd_params: C must be a multiple of 8 and <= %d (got %d)", mk::MAXC, C);
  RD_REQUIRE(impl >= 0 && impl <= 3, "rd_meta_kernel_bwd_params: impl must be 0..3");
  RD_REQUIRE(grad_w0 && grad_b0 && grad_w1 && grad_b1, "rd_meta_kernel_bwd_params: null output pointer");
  cudaStream_t st = rd::as_stream(stream);
  if (rd_check_device()) return 1;
  if (B == 0) {
    RD_CUDA(cudaMemsetAsync(grad_w0, 0, sizeof(float) * mk::HID * 3, st));
    RD_CUDA(cudaMemsetAsync(grad_b0, 0, sizeof(float) * mk::HID, st));
    RD_CUDA(cudaMemsetAsync(grad_w1, 0, sizeof(float) * C * mk::HID, st));
    RD_CUDA(cudaMemsetAsync(grad_b1, 0, sizeof(float) * C, st));
    return 0;
  }
  RD_REQUIRE(grad_out && data && coord && w0 && b0 && w1, "rd_meta_kernel_bwd_params: null pointer");
  RD_REQUIRE(workspace && workspace_bytes >= rd_meta_kernel_bwd_workspace_bytes(B, C, H, W),
             "rd_meta_kernel_bwd_params: workspace too small (%zu < %zu)", workspace_bytes,
             rd_meta_kernel_bwd_workspace_bytes(B, C, H, W));
  float* partial = static_cast<float*>(workspace);
  const int nout = C * mk::HID + C + mk::HID * 4;
  if (impl == 0)
    impl = (C == 64 && W % 4 == 0 && ((reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(data)) & 15) == 0) ? 3 : 1;
  if (impl == 3) {
    int nparts = 0;
    if (rd_meta_kernel_bwd_params_ws(grad_out, data, coord, w0, b0, w1, partial, &nparts, B, C, H, W, st)) return 1;
    mk::meta_bwd_param_reduce_kernel<<<(nout + 255) / 256, 256, 0, st>>>(partial, nparts, C, grad_w0, grad_b0, grad_w1,
                                                                         grad_b1);
    rd::count_launch();
    return rd::check_launch("rd_meta_kernel_bwd_params(reduce)");
  }
  const int tiles_w = (W + mk::P_TW - 1) / mk::P_TW;
  const int64_t ntiles = (int64_t)B * H * tiles_w;
  RD_REQUIRE(ntiles <= 0x7fffffffLL, "rd_meta_kernel_bwd_params: too many tiles");
  const size_t smem = sizeof(mk::GpSmem);
  RD_CUDA(cudaFuncSetAttribute(mk::meta_bwd_param_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RD_CUDA(cudaMemsetAsync(partial, 0, (size_t)mk::P_GRID * 2 * nout * sizeof(float), st));
  mk::meta_bwd_param_kernel<<<mk::P_GRID, mk::NT, smem, st>>>(grad_out, data, coord, w0, b0, w1, partial, B, C,
                                                              H, W, tiles_w, (int)ntiles);
  rd::count_launch();
  if (rd::check_launch("rd_meta_kernel_bwd_params")) return 1;
  mk::meta_bwd_param_reduce_kernel<<<(nout + 255) / 256, 256, 0, st>>>(partial, mk::P_GRID * 2, C, grad_w0,
                                                                       grad_b0, grad_w1, grad_b1);
  rd::count_launch();
  return rd::check_launch("rd_meta_kernel_bwd_params(reduce)");
}

int rd_meta_kernel_bwd(const float* grad_out, const float* data, const float* coord, const float* w0,
                       const float* b0, const float* w1, const float* b1, float* grad_data,
                       float* grad_w0, float* grad_b0, float* grad_w1, float* grad_b1, void* workspace,
                       size_t workspace_bytes, int B, int C, int H, int W, int impl,
                       rd_stream_t stream) {
  RD_REQUIRE(grad_data != nullptr, "rd_meta_kernel_bwd: null grad_data");
  if (rd_meta_kernel_bwd_data(grad_out, coord, w0, b0, w1, b1, grad_data, B, C, H, W, impl, stream)) return 1;
  return rd_meta_kernel_bwd_params(grad_out, data, coord, w0, b0, w1, b1, grad_w0, grad_b0, grad_w1, grad_b1,
                                   workspace, workspace_bytes, B, C, H, W, impl, stream);
}

}  // extern "C"
