// Meta-Kernel impl 3, MLP-parameter gradients: persistent, warp-specialised, TMA-fed tcgen05 kernel.
//
// Gradient of MetaKernel.meta_baseline_bias (/root/reference rangedet/symbol/backbone/
// meta_kernel.py:166-240) w.r.t. mlp0/mlp1 weights and biases.  Per (row, 128-pixel) tile and tap k:
//   gw[p][c]   = grad_out[p, c*9+k] * data[p+dk, c]                      (CUDA cores, fp32)
//   hid[p][j]  = relu(W0 . rel + b0)                                     (recomputed, fp32)
//   G1  ghid[p][j] = sum_c gw[p][c] W1[c][j]      tcgen05  M=128 px, N=32, K=64(x hi/lo)  -> TMEM D1
//   G2  gW1[c][j] += sum_p gw[p][c] hid[p][j]     tcgen05  M=128 (c x hi/lo), N=80 (j x hi/lo | 1),
//                   gb1[c] += sum_p gw[p][c]       K=128 px, MN-major views of the SAME smem tiles,
//                                                  accumulated in TMEM across all taps and tiles
//   gW0[j][d] += sum_p relu'(.) ghid[p][j] rel[p][d], gb0 likewise        (fp32 register accumulators)
// bf16 hi+lo splits of gw / hid / W1 keep every product within ~2^-16 of fp32.
// Roles: warp 0 TMA producer (grad_out tap tiles + feature row boxes), warp 1 MMA issuer,
// warps 2-17 builders (four threads per pixel: operand tiles, D1 consumption; they are the
// critical resource -- ncu showed producer and MMA warps spinning on builder-signalled barriers).  One partial result row per CTA goes to the
// workspace; the deterministic reduce kernel of meta_kernel.cu finishes the sum.
#include <stdio.h>
#include <stdlib.h>

#include "../../include/rangedet_b200.h"
#include "rd_common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

namespace mkwp {

constexpr int HID = 32, CCH = 3, C = 64;
constexpr int TW = 128, ROW = TW + 2, DTW = TW + 8;
constexpr int TPP = 4;                        // builder threads per pixel
constexpr int NBUILD = TW * TPP;              // 512: warps 2-17
constexpr int NTHREADS = 64 + NBUILD;         // warp 0 TMA, warp 1 MMA, then the builders
constexpr int NBW = NBUILD / 32;              // builder warps (mbarrier arrival counts)
constexpr int CHUNK = TW * 16;                // one 16-byte chunk over 128 pixel rows (2048 B)
constexpr int GW_BYTES = 16 * CHUNK;          // gw_hi (8 chunks of 8 channels) | gw_lo (8 chunks)
constexpr int H_BYTES = 10 * CHUNK;           // h_hi (4) | h_lo (4) | ones chunk | zero chunk
constexpr int W1T_CHUNK = HID * 16;           // W1^T[j][c]: 32 rows per 16-byte chunk
constexpr int W1T_BYTES = 16 * W1T_CHUNK;     // hi (8 chunks over c) | lo (8 chunks)
constexpr int GO_BYTES = C * TW * 4;          // 32 KB grad_out tap tile
constexpr int DROW_BYTES = C * DTW * 4;       // 34 KB feature row box
constexpr int NS_G = 2, NS_R = 2;
constexpr uint32_t TMEM_COLS = 256;           // D1: 2 x 32 columns at 0/32, D2: 80 columns at 64
constexpr uint32_t D2_COL = 64;
constexpr int BAR_BLD = 1;
constexpr int NOUT = C * HID + C + HID * 4;

struct Smem {
  alignas(1024) float go[NS_G][C * TW];
  alignas(1024) float drow[NS_R][C * DTW];
  alignas(1024) unsigned char gw[GW_BYTES];
  alignas(1024) unsigned char h[H_BYTES];
  alignas(1024) unsigned char w1t[W1T_BYTES];
  alignas(16) float4 w0b[HID];
  alignas(16) float cs[3 * CCH * ROW];
  alignas(8) uint64_t go_full[NS_G], go_empty[NS_G], dr_full[NS_R], dr_empty[NS_R], ops_full, ops_free,
      d1_full[2], d1_empty[2];
  uint32_t tmem_slot;
};

// PROF: diagnostic instantiation (RD_MK_PROF=1).  It must be a template parameter: with 576 threads the
// register cap is 96, and the eight 64-bit cycle counters of a runtime switch spilled (0.57 -> 0.80 ms).
// GOF: storage of grad_out -- 0: (B, C*9, H, W) fp32 (reference op boundary); 1 / 2: haloed NHWC bf16 / fp16 with tap-major
// channels (see meta_kernel_ws.cu): the tap tile is then [128 px][64 ch] x 2 B, 128-byte swizzled, and a builder thread
// fetches its 16 channels with two 16-byte shared loads.
template <int FMT>
__device__ __forceinline__ void unpack8_2b(const uint4& q, float* f) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (FMT == 1) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    } else {
      const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = v.x;
      f[2 * i + 1] = v.y;
    }
  }
}

template <bool PROF, int GOF = 0>
__global__ void __launch_bounds__(NTHREADS, 1)
meta_ws_params_kernel(const __grid_constant__ CUtensorMap tm_go, const __grid_constant__ CUtensorMap tm_data,
                      const float* __restrict__ coord, const float* __restrict__ w0, const float* __restrict__ b0,
                      const float* __restrict__ w1, float* __restrict__ partial, int B, int H, int W, int tiles_w,
                      int ntiles, long long* __restrict__ prof) {
  // prof (diagnostic, RD_MK_PROF=1, else null): per-CTA clock64() sums, [block][16]: 0 producer total,
  // 1 wait go_empty, 2 wait dr_empty | 3 MMA total, 4 wait ops_full, 5 wait d1_empty | 6 builder total,
  // 7 hidden layer, 8 wait go_full/dr_full, 9 products, 10 wait ops_free, 11 operand stores, 12 wait d1_full,
  // 13 D1 consumption, 14 coordinate staging
  auto tick = [&]() -> long long { return PROF ? clock64() : 0ll; };
  long long q0 = 0, q1 = 0, q2 = 0, q3 = 0, q4c = 0, q5 = 0, q6 = 0, q7 = 0;
  const long long t_begin = tick();
  extern __shared__ unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

  if (t == 0) {
    for (int i = 0; i < NS_G; ++i) { tc::mbar_init(&S.go_full[i], 1); tc::mbar_init(&S.go_empty[i], NBW); }
    for (int i = 0; i < NS_R; ++i) { tc::mbar_init(&S.dr_full[i], 1); tc::mbar_init(&S.dr_empty[i], NBW); }
    tc::mbar_init(&S.ops_full, NBW);
    tc::mbar_init(&S.ops_free, 1);
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&S.d1_full[i], 1); tc::mbar_init(&S.d1_empty[i], NBW); }
    tc::fence_mbar_init();
    tma::prefetch_map(&tm_go);
    tma::prefetch_map(&tm_data);
  }
  if (warp == 1) {
    tc::tmem_alloc(&S.tmem_slot, TMEM_COLS);
    tc::tmem_relinquish();
  }
  if (warp >= 2) {  // constants: layer-0 params, W1^T hi/lo in K-major core-matrix layout, ones/zero chunks of H
    const int bt = t - 64;
    for (int j = bt; j < HID; j += NBUILD)
      S.w0b[j] = make_float4(__ldg(w0 + j * 3 + 0), __ldg(w0 + j * 3 + 1), __ldg(w0 + j * 3 + 2), __ldg(b0 + j));
    for (int e = bt; e < HID * 8; e += NBUILD) {  // chunk q (8 channels c = 8q..8q+7) of row j
      const int j = e % HID, q = e / HID;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        float h0, l0, h1, l1;
        tc::split_bf16(__ldg(w1 + (q * 8 + 2 * p) * HID + j), h0, l0);
        tc::split_bf16(__ldg(w1 + (q * 8 + 2 * p + 1) * HID + j), h1, l1);
        hi[p] = tc::pack_bf16x2(h0, h1);
        lo[p] = tc::pack_bf16x2(l0, l1);
      }
      *reinterpret_cast<uint4*>(S.w1t + q * W1T_CHUNK + j * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(S.w1t + (8 + q) * W1T_CHUNK + j * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    if (bt < TW) {
      *reinterpret_cast<uint4*>(S.h + 8 * CHUNK + bt * 16) = make_uint4(tc::pack_bf16x2(1.f, 0.f), 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(S.h + 9 * CHUNK + bt * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    tc::fence_proxy_async_smem();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = S.tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: per tile 3 feature row boxes + 9 grad_out tap tiles, in consumption order =====
    if (lane == 0) {
      uint32_t gg = 0, gr = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int wt = tile % tiles_w, h = (tile / tiles_w) % H, b = tile / (tiles_w * H);
        const int w0px = wt * TW;
        const int bs = w0px >= 4 ? w0px - 4 : 0;
        for (int k = 0; k < 9; ++k, ++gg) {
          if (k % 3 == 0) {
            const int hh = h + k / 3 - 1;
            const uint32_t s = gr % NS_R, ph = (gr / NS_R) & 1;
            const long long ta = tick();
            tc::mbar_wait(&S.dr_empty[s], ph ^ 1);
            q1 += tick() - ta;
            if (hh >= 0 && hh < H) {
              tc::mbar_arrive_expect_tx(&S.dr_full[s], DROW_BYTES);
              tma::load_3d(S.drow[s], &tm_data, &S.dr_full[s], bs, hh, b * C);
            } else {
              tc::mbar_arrive(&S.dr_full[s]);
            }
            ++gr;
          }
          const uint32_t s = gg % NS_G, ph = (gg / NS_G) & 1;
          const long long tb = tick();
          tc::mbar_wait(&S.go_empty[s], ph ^ 1);
          q0 += tick() - tb;
          tc::mbar_arrive_expect_tx(&S.go_full[s], GOF ? GO_BYTES / 2 : GO_BYTES);
          if (GOF) tma::load_4d(S.go[s], &tm_go, &S.go_full[s], k * C, w0px + 1, h + 1, b);
          else tma::load_4d(S.go[s], &tm_go, &S.go_full[s], w0px, h, k, b * C);
        }
      }
      if (PROF) { prof[blockIdx.x * 16 + 0] = tick() - t_begin; prof[blockIdx.x * 16 + 1] = q0; prof[blockIdx.x * 16 + 2] = q1; }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // Converged warp, one elected lane issues; descriptors precomputed (see meta_kernel_ws.cu: the issuing
    // thread's own instruction latency, not the tensor core, bounded the per-MMA descriptor-building loop).
    {
      constexpr uint32_t idesc1 = tc::make_idesc_bf16(128, HID);         // G1: K-major A and B
      constexpr uint32_t idesc2 = tc::make_idesc_bf16(128, 80, 1, 1);    // G2: MN-major A and B
      constexpr uint32_t GC = CHUNK >> 4, WC = W1T_CHUNK >> 4;            // chunk strides in descriptor units
      const uint64_t g1a = tc::make_smem_desc(0, CHUNK, 128, tc::LAYOUT_NONE) | (uint64_t)(tc::smem_u32(S.gw) >> 4);
      const uint64_t g1b = tc::make_smem_desc(0, W1T_CHUNK, 128, tc::LAYOUT_NONE) | (uint64_t)(tc::smem_u32(S.w1t) >> 4);
      const uint64_t g2a = tc::make_smem_desc(0, 128, CHUNK, tc::LAYOUT_NONE) | (uint64_t)(tc::smem_u32(S.gw) >> 4);
      const uint64_t g2b = tc::make_smem_desc(0, 128, CHUNK, tc::LAYOUT_NONE) | (uint64_t)(tc::smem_u32(S.h) >> 4);
      const bool leader = tc::elect_one();
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll 1
        for (int k = 0; k < 9; ++k, ++g) {
          const uint32_t s1 = g & 1, ph1 = (g >> 1) & 1;
          const long long ta = tick();
          tc::mbar_wait(&S.ops_full, g & 1);
          const long long tb = tick();
          tc::mbar_wait(&S.d1_empty[s1], ph1 ^ 1);
          q0 += tb - ta;
          q1 += tick() - tb;
          tc::tc_fence_after();
          if (leader) {
            // G1: D1[px][j] = gw[px][c] . W1T[j][c]^T ; split-bf16 terms (hi,hi) (lo,hi) (hi,lo), 4 K slices each
            const uint32_t d1 = tmem_base + s1 * HID;
            tc::mma_bf16_ss(d1, g1a, g1b, idesc1, 0u);
#pragma unroll
            for (int ks = 1; ks < 4; ++ks) tc::mma_bf16_ss_acc(d1, g1a + 2 * ks * GC, g1b + 2 * ks * WC, idesc1);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) tc::mma_bf16_ss_acc(d1, g1a + (8 + 2 * ks) * GC, g1b + 2 * ks * WC, idesc1);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) tc::mma_bf16_ss_acc(d1, g1a + 2 * ks * GC, g1b + (8 + 2 * ks) * WC, idesc1);
            tc::umma_commit(&S.d1_full[s1]);
            // G2: D2[(c,part)][(j,part)|1] += gw^T . [h | 1]   -- MN-major views: MN block stride = CHUNK
            // (8 channels / hidden units per chunk), K block (8 pixels) stride = 128 B, K step = 16 px = 256 B
            tc::mma_bf16_ss(tmem_base + D2_COL, g2a, g2b, idesc2, g > 0 ? 1u : 0u);
#pragma unroll
            for (int ks = 1; ks < 8; ++ks) tc::mma_bf16_ss_acc(tmem_base + D2_COL, g2a + 16 * ks, g2b + 16 * ks, idesc2);
            tc::umma_commit(&S.ops_free);
          }
          __syncwarp();
        }
      }
      if (PROF && lane == 0) { prof[blockIdx.x * 16 + 3] = tick() - t_begin; prof[blockIdx.x * 16 + 4] = q0; prof[blockIdx.x * 16 + 5] = q1; }
    }
  } else {
    // ===== builders: TPP threads per pixel (TMEM lane); thread `hf` of a pixel owns hidden units
    // [HJ*hf, HJ*hf + HJ) and channels [CPT*hf, CPT*hf + CPT).  They are the critical resource (the producer
    // and MMA warps spin on builder-signalled barriers), and with 2 warps per scheduler they ran at IPC 0.4:
    // four threads per pixel = 4 warps per scheduler to hide the shared-memory / dependency latencies. =====
    const int bt = t - 64;                 // 0..NBUILD-1
    const int hf = (warp - 2) >> 2;        // which part of the per-pixel work
    const int q4 = warp & 3;               // TMEM lane quadrant of this warp
    const int px = q4 * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    constexpr int HJ = HID / TPP;          // hidden units per thread (8)
    constexpr int CPT = C / TPP;           // channels per thread (16)
    constexpr int NCQ = CPT / 8;           // 16-byte operand chunks of gw per thread and part (2)
    constexpr int CS_PER_THREAD = (3 * CCH * ROW + NBUILD - 1) / NBUILD;
    float acc0[HJ][4];  // gW0[j][d], d == 3 -> gb0[j], for j = 16hf + jj
#pragma unroll
    for (int j = 0; j < HJ; ++j)
#pragma unroll
      for (int d = 0; d < 4; ++d) acc0[j][d] = 0.f;
    uint32_t g = 0, gr = 0;
    bool have_prev = false;
    uint32_t prev_mask = 0, prev_g = 0;
    float prev_rel[3] = {0.f, 0.f, 0.f};

    auto consume_d1 = [&](uint32_t gp, uint32_t mask, const float* rel) {
      const uint32_t s1 = gp & 1, ph1 = (gp >> 1) & 1;
      const long long tw = tick();
      tc::mbar_wait(&S.d1_full[s1], ph1);
      const long long tw2 = tick();
      q5 += tw2 - tw;
      __syncwarp();
      tc::tc_fence_after();
      float v[HJ];
      tc::tmem_ld_x8(tmem_base + lane_sel + s1 * HID + hf * HJ, v);
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&S.d1_empty[s1]);
#pragma unroll
      for (int j = 0; j < HJ; ++j) {
        const float gh = ((mask >> j) & 1u) ? v[j] : 0.f;
        acc0[j][0] = fmaf(gh, rel[0], acc0[j][0]);
        acc0[j][1] = fmaf(gh, rel[1], acc0[j][1]);
        acc0[j][2] = fmaf(gh, rel[2], acc0[j][2]);
        acc0[j][3] += gh;
      }
      q6 += tick() - tw2;
    };
    // coordinate tile of `tile` -> registers (software prefetch: issued one tile ahead)
    auto load_coords = [&](int tile, float (&cp)[CS_PER_THREAD]) {
      const int wt = tile % tiles_w, h = (tile / tiles_w) % H, b = tile / (tiles_w * H);
      const int w0px = wt * TW;
#pragma unroll
      for (int i = 0; i < CS_PER_THREAD; ++i) {
        const int e = bt + i * NBUILD;
        float v = 0.f;
        if (e < 3 * CCH * ROW) {
          const int col = e % ROW, d = (e / ROW) % CCH, r = e / (ROW * CCH);
          const int hh = h + r - 1, ww = w0px + col - 1;
          if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(coord + (((int64_t)b * CCH + d) * H + hh) * W + ww);
        }
        cp[i] = v;
      }
    };
    float cpre[CS_PER_THREAD];
    if ((int)blockIdx.x < ntiles) load_coords(blockIdx.x, cpre);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int wt = tile % tiles_w, h = (tile / tiles_w) % H;
      const int w0px = wt * TW;
      const int bs = w0px >= 4 ? w0px - 4 : 0;
      const long long ts0 = tick();
      tma::named_bar_sync(BAR_BLD, NBUILD);  // everyone finished reading the previous coordinate tile
#pragma unroll
      for (int i = 0; i < CS_PER_THREAD; ++i) {
        const int e = bt + i * NBUILD;
        if (e < 3 * CCH * ROW) S.cs[e] = cpre[i];
      }
      tma::named_bar_sync(BAR_BLD, NBUILD);
      if (tile + (int)gridDim.x < ntiles) load_coords(tile + gridDim.x, cpre);  // in flight during this tile
      q7 += tick() - ts0;
      const float c0 = S.cs[(1 * CCH + 0) * ROW + px + 1];
      const float c1 = S.cs[(1 * CCH + 1) * ROW + px + 1];
      const float c2 = S.cs[(1 * CCH + 2) * ROW + px + 1];
      const bool px_ok = (w0px + px) < W;  // tile overhang: these pixels contribute nothing

      for (int k = 0; k < 9; ++k, ++g) {
        const int dy = k / 3 - 1, dx = k % 3 - 1;
        const long long b0t = tick();
        // ---- this thread's hidden units of (pixel, tap), kept packed in registers
        const int ccol = px + 1 + dx, r = dy + 1;
        const float r0 = S.cs[(r * CCH + 0) * ROW + ccol] - c0;
        const float r1 = S.cs[(r * CCH + 1) * ROW + ccol] - c1;
        const float r2 = S.cs[(r * CCH + 2) * ROW + ccol] - c2;
        uint32_t hhi[HJ / 2], hlo[HJ / 2], mask = 0;
#pragma unroll
        for (int jp = 0; jp < HJ / 2; ++jp) {
          float hv[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float4 wv = S.w0b[hf * HJ + jp * 2 + u];
            float z = wv.w;
            z = fmaf(wv.x, r0, z);
            z = fmaf(wv.y, r1, z);
            z = fmaf(wv.z, r2, z);
            hv[u] = fmaxf(z, 0.f);
            mask |= (hv[u] > 0.f ? 1u : 0u) << (jp * 2 + u);
          }
          tc::split_pack_bf16x2(hv[0], hv[1], hhi[jp], hlo[jp]);
        }
        // ---- inputs of this tap
        const uint32_t sg = g % NS_G, phg = (g / NS_G) & 1;
        const uint32_t sr = gr % NS_R, phr = (gr / NS_R) & 1;
        const long long b1t = tick();
        tc::mbar_wait(&S.go_full[sg], phg);
        if (dx == -1) tc::mbar_wait(&S.dr_full[sr], phr);
        const long long b2t = tick();
        const int col = w0px + px + dx - bs;
        const bool ok = px_ok && (h + dy >= 0) && (h + dy < H) && col >= 0;
        const float* gt = S.go[sg] + px;
        const float* dt = S.drow[sr] + (ok ? col : 0);
        // products first (registers), so the shared-memory loads are not serialised behind the stores
        uint32_t ghi[NCQ][4], glo[NCQ][4];
#pragma unroll
        for (int cq = 0; cq < NCQ; ++cq) {  // 8 channels per 16-byte chunk; this thread: chunks NCQ*hf ..
          float gv[8], dv[8];
          if (GOF) {
            const uint4 q = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(S.go[sg]) + px * 128 +
                                                            (((hf * NCQ + cq) ^ (px & 7)) << 4));
            unpack8_2b<GOF>(q, gv);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = (hf * NCQ + cq) * 8 + i;
            if (!GOF) gv[i] = gt[c * TW];
            dv[i] = ok ? dt[c * DTW] : 0.f;
          }
#pragma unroll
          for (int p = 0; p < 4; ++p)
            tc::split_pack_bf16x2(gv[2 * p] * dv[2 * p], gv[2 * p + 1] * dv[2 * p + 1], ghi[cq][p], glo[cq][p]);
        }
        __syncwarp();
        if (lane == 0) {  // input tiles fully read by this warp
          tc::mbar_arrive(&S.go_empty[sg]);
          if (dx == 1) tc::mbar_arrive(&S.dr_empty[sr]);
        }
        if (dx == 1) ++gr;
        // ---- operand tiles may be overwritten once the previous tap's MMAs have read them
        const long long b3t = tick();
        tc::mbar_wait(&S.ops_free, (g & 1) ^ 1);
        const long long b4t = tick();
#pragma unroll
        for (int q = 0; q < HJ / 8; ++q) {  // h chunks of 8 hidden units: hi at [0,4), lo at [4,8)
          *reinterpret_cast<uint4*>(S.h + (hf * (HJ / 8) + q) * CHUNK + px * 16) =
              make_uint4(hhi[q * 4 + 0], hhi[q * 4 + 1], hhi[q * 4 + 2], hhi[q * 4 + 3]);
          *reinterpret_cast<uint4*>(S.h + (4 + hf * (HJ / 8) + q) * CHUNK + px * 16) =
              make_uint4(hlo[q * 4 + 0], hlo[q * 4 + 1], hlo[q * 4 + 2], hlo[q * 4 + 3]);
        }
#pragma unroll
        for (int cq = 0; cq < NCQ; ++cq) {  // gw chunks of 8 channels: hi at [0,8), lo at [8,16)
          *reinterpret_cast<uint4*>(S.gw + (hf * NCQ + cq) * CHUNK + px * 16) =
              make_uint4(ghi[cq][0], ghi[cq][1], ghi[cq][2], ghi[cq][3]);
          *reinterpret_cast<uint4*>(S.gw + (8 + hf * NCQ + cq) * CHUNK + px * 16) =
              make_uint4(glo[cq][0], glo[cq][1], glo[cq][2], glo[cq][3]);
        }
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&S.ops_full);
        {
          const long long b5t = tick();
          q0 += b1t - b0t; q1 += b2t - b1t; q2 += b3t - b2t; q3 += b4t - b3t; q4c += b5t - b4t;
        }
        // ---- consume the previous tap's D1 while this tap's MMAs run
        if (have_prev) consume_d1(prev_g, prev_mask, prev_rel);
        have_prev = true;
        prev_g = g;
        prev_mask = ok ? mask : 0u;  // (gw is 0 for !ok pixels, so ghid is 0 anyway)
        prev_rel[0] = r0; prev_rel[1] = r1; prev_rel[2] = r2;
      }
    }
    if (have_prev) consume_d1(prev_g, prev_mask, prev_rel);
    const uint32_t g_total = g;
    if (PROF && bt == 0) {
      long long* pr = prof + blockIdx.x * 16;
      pr[6] = tick() - t_begin; pr[7] = q0; pr[8] = q1; pr[9] = q2; pr[10] = q3; pr[11] = q4c; pr[12] = q5; pr[13] = q6; pr[14] = q7;
    }

    // ---- CTA result: gW0/gb0 from registers (cross-thread sum through smem), gW1/gb1 from TMEM D2
    tma::named_bar_sync(BAR_BLD, NBUILD);              // all builders finished with go / drow tiles
    float* red = reinterpret_cast<float*>(S.go);       // 128 x 128 floats = 64 KB = the two go stages
#pragma unroll
    for (int j = 0; j < HJ; ++j)
#pragma unroll
      for (int d = 0; d < 4; ++d) red[((hf * HJ + j) * 4 + d) * 128 + px] = acc0[j][d];
    tma::named_bar_sync(BAR_BLD, NBUILD);
    float* prow = partial + (int64_t)blockIdx.x * NOUT;
    if (bt < 128) {
      float s = 0.f;
      for (int i = 0; i < 128; ++i) s += red[bt * 128 + ((i + bt) & 127)];
      prow[C * HID + C + bt] = s;  // (j,d) -> j*4+d, as the reduce kernel expects
    }
    tma::named_bar_sync(BAR_BLD, NBUILD);
    if (hf == 0) {  // warps 2-5 cover the 128 TMEM lanes once
      if (g_total > 0) {
        tc::mbar_wait(&S.ops_free, (g_total - 1) & 1);   // last G2 complete
        __syncwarp();
        tc::tc_fence_after();
        float v[32], u[32], wv[16];
        tc::tmem_ld_x32(tmem_base + lane_sel + D2_COL, v);        // x h_hi
        tc::tmem_ld_x32(tmem_base + lane_sel + D2_COL + 32, u);   // x h_lo
        tc::tmem_ld_x16(tmem_base + lane_sel + D2_COL + 64, wv);  // x [1, 0...]
#pragma unroll
        for (int j = 0; j < 32; ++j) red[px * 33 + j] = v[j] + u[j];
        red[px * 33 + 32] = wv[0];
      } else {
        for (int j = 0; j < 33; ++j) red[px * 33 + j] = 0.f;
      }
    }
    tma::named_bar_sync(BAR_BLD, NBUILD);
    if (bt < C) {
      const float* a = red + bt * 33;
      const float* bb = red + (bt + C) * 33;   // gw_lo rows
      for (int j = 0; j < HID; ++j) prow[bt * HID + j] = a[j] + bb[j];
      prow[C * HID + bt] = a[32] + bb[32];
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace mkwp

// partial must hold gridDim rows of NOUT floats; returns the number of rows written in *nparts
static int params_launch(const void* grad_out, int gof, const float* data, const float* coord, const float* w0,
                         const float* b0, const float* w1, float* partial, int* nparts, int B, int C, int H, int W,
                         cudaStream_t stream);

int rd_meta_kernel_bwd_params_ws(const float* grad_out, const float* data, const float* coord, const float* w0,
                                 const float* b0, const float* w1, float* partial, int* nparts, int B, int C, int H,
                                 int W, cudaStream_t stream) {
  return params_launch(grad_out, 0, data, coord, w0, b0, w1, partial, nparts, B, C, H, W, stream);
}
// grad_out as the haloed NHWC tap-major gradient (gof 1: bf16, 2: fp16)
int rd_meta_kernel_bwd_params_ws_nhwc(const void* grad_out_pad, int gof, const float* data, const float* coord,
                                      const float* w0, const float* b0, const float* w1, float* partial, int* nparts, int B,
                                      int C, int H, int W, cudaStream_t stream) {
  return params_launch(grad_out_pad, gof, data, coord, w0, b0, w1, partial, nparts, B, C, H, W, stream);
}

static int params_launch(const void* grad_out, int gof, const float* data, const float* coord, const float* w0,
                         const float* b0, const float* w1, float* partial, int* nparts, int B, int C, int H, int W,
                         cudaStream_t stream) {
  using namespace mkwp;
  RD_REQUIRE(C == mkwp::C, "Meta-Kernel impl 3 (TMA/tcgen05) is specialised for C == 64 (got %d)", C);
  RD_REQUIRE(W % 4 == 0, "Meta-Kernel impl 3 needs W %% 4 == 0; W=%d", W);
  const int tiles_w = (W + TW - 1) / TW;
  const int64_t ntiles = (int64_t)B * H * tiles_w;
  RD_REQUIRE(ntiles <= 0x7fffffffLL, "rd_meta_kernel_bwd_params: too many tiles");
  CUtensorMap tm_go, tm_data;
  const uint64_t plane = (uint64_t)H * W * 4;
  const uint64_t d3[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)B * C};
  const uint64_t s3[2] = {(uint64_t)W * 4, plane};
  const uint32_t b3in[3] = {(uint32_t)DTW, 1u, (uint32_t)C};
  const uint64_t d4[4] = {(uint64_t)W, (uint64_t)H, 9u, (uint64_t)B * C};
  const uint64_t s4[3] = {(uint64_t)W * 4, plane, plane * 9};
  const uint32_t b4[4] = {(uint32_t)TW, 1u, 1u, (uint32_t)C};
  if (gof) {   // whole haloed NHWC tensor (9C, W+2, H+2, B): pixel (h, w) sits at (w+1, h+1)
    const uint64_t Wp = (uint64_t)W + 2, Hp = (uint64_t)H + 2, CO = 9u * mkwp::C;
    const uint64_t dn[4] = {CO, Wp, Hp, (uint64_t)B};
    const uint64_t sn[3] = {CO * 2, Wp * CO * 2, Hp * Wp * CO * 2};
    const uint32_t bn[4] = {(uint32_t)mkwp::C, (uint32_t)TW, 1u, 1u};
    if (tma::make_map(&tm_go, gof == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, grad_out, 4, dn, sn,
                      bn, CU_TENSOR_MAP_SWIZZLE_128B))
      return 1;
  } else if (tma::make_map(&tm_go, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, grad_out, 4, d4, s4, b4, CU_TENSOR_MAP_SWIZZLE_NONE)) {
    return 1;
  }
  if (tma::make_map(&tm_data, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, data, 3, d3, s3, b3in, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  int dev = 0, sms = 0;
  RD_CUDA(cudaGetDevice(&dev));
  RD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = sizeof(Smem) + 1024;
  const int64_t grid = ntiles < sms ? ntiles : sms;
  RD_CUDA(rd::smem_optin(meta_ws_params_kernel<false>, smem));
  RD_CUDA(rd::smem_optin(meta_ws_params_kernel<true>, smem));
  RD_CUDA(rd::smem_optin(meta_ws_params_kernel<false, 1>, smem));
  RD_CUDA(rd::smem_optin(meta_ws_params_kernel<false, 2>, smem));
  static const bool want_prof = [] { const char* e = getenv("RD_MK_PROF"); return e && e[0] == '1'; }();
  long long* d_prof = nullptr;
  if (want_prof) {
    static long long* buf = nullptr;
    if (!buf) RD_CUDA(cudaMalloc(&buf, 1024 * 16 * sizeof(long long)));
    RD_CUDA(cudaMemsetAsync(buf, 0, 1024 * 16 * sizeof(long long), stream));
    d_prof = buf;
  }
  if (gof == 1)
    meta_ws_params_kernel<false, 1><<<(unsigned)grid, NTHREADS, smem, stream>>>(tm_go, tm_data, coord, w0, b0, w1, partial, B,
                                                                                 H, W, tiles_w, (int)ntiles, nullptr);
  else if (gof == 2)
    meta_ws_params_kernel<false, 2><<<(unsigned)grid, NTHREADS, smem, stream>>>(tm_go, tm_data, coord, w0, b0, w1, partial, B,
                                                                                 H, W, tiles_w, (int)ntiles, nullptr);
  else if (want_prof)
    meta_ws_params_kernel<true><<<(unsigned)grid, NTHREADS, smem, stream>>>(tm_go, tm_data, coord, w0, b0, w1, partial, B, H,
                                                                             W, tiles_w, (int)ntiles, d_prof);
  else
    meta_ws_params_kernel<false><<<(unsigned)grid, NTHREADS, smem, stream>>>(tm_go, tm_data, coord, w0, b0, w1, partial, B,
                                                                              H, W, tiles_w, (int)ntiles, nullptr);
  rd::count_launch();
  if (want_prof && !gof) {  // diagnostic: synchronous, mean cycles per TAP and role
    RD_CUDA(cudaStreamSynchronize(stream));
    static long long hbuf[1024 * 16];
    RD_CUDA(cudaMemcpy(hbuf, d_prof, sizeof(long long) * 16 * grid, cudaMemcpyDeviceToHost));
    double m[16] = {0};
    for (int b = 0; b < grid; ++b)
      for (int k = 0; k < 16; ++k) m[k] += (double)hbuf[b * 16 + k] / grid;
    const double tp = 9.0 * (double)ntiles / grid;
    fprintf(stderr,
            "[rd_meta prof bwd_params] taps/cta=%.0f | per tap: producer %.0f (wait go_empty %.0f, dr_empty %.0f) | mma %.0f (wait "
            "ops_full %.0f, d1_empty %.0f) | builder %.0f (hidden %.0f, wait inputs %.0f, products %.0f, wait ops_free %.0f, "
            "stores %.0f, wait d1_full %.0f, D1 use %.0f, coord staging %.0f)\n",
            tp, m[0] / tp, m[1] / tp, m[2] / tp, m[3] / tp, m[4] / tp, m[5] / tp, m[6] / tp, m[7] / tp, m[8] / tp, m[9] / tp,
            m[10] / tp, m[11] / tp, m[12] / tp, m[13] / tp, m[14] / tp);
  }
  *nparts = (int)grid;
  return rd::check_launch("rd_meta_kernel_bwd_params(impl 3)");
}
