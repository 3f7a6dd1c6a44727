// Meta-Kernel impl 3: persistent, warp-specialised, TMA-fed tcgen05 kernel (forward and grad_data).
//
// Replaces MetaKernel.meta_baseline_bias (/root/reference rangedet/symbol/backbone/
// meta_kernel.py:166-240) and its data-gradient.  One CTA per SM walks (row, 128-pixel) tiles; per
// tile 9 taps flow through four mbarrier-synchronised rings:
//
//   warp 0      TMA producer    feature rows (fwd: one [64 ch][136 px] box per neighbour ROW, shared by
//                               its three dx taps) / grad_out tap planes (bwd) -> smem.  TMA tile mode
//                               on this part faults on negative or non-16-byte-aligned coordinates
//                               (probed: tests/test_gpu_parity.py::test_tma_probe), so boxes start at
//                               max(w0-4, 0); rows outside the image are not loaded (treated as 0)
//                               and the right border is TMA's zero fill
//   warps 2-5   hidden layer    rel xyz -> relu(W0 rel + b0) fp32 -> bf16 hi|lo, K-major core-matrix
//                               layout (A operand ring)
//   warp 1      MMA issuer      9 x tcgen05.mma M128 N64 K16: [h_hi|h_lo] x [W1_hi|W1_lo]^T (+ bias
//                               K-slice) -> fp32 accumulator ring in TMEM (4 x 64 columns)
//   warps 6-9   epilogue        tcgen05.ld -> x tap tile (smem, conflict-free) ->
//                                 fwd: product tile -> smem -> TMA store (4-D map over (W,H,9,B*C))
//                                 bwd: accumulate over the 9 taps in registers -> smem -> TMA store
// HBM traffic = algorithmic (2572 B/pixel either direction); every global access is a TMA bulk
// transfer, so bytes in flight per SM are bounded by the rings (96 KB loads + 64 KB stores), not by
// per-thread load/store queues.
#include <stdlib.h>

#include "../../include/rangedet_b200.h"
#include "rd_common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

namespace mkws {

constexpr int HID = 32, CCH = 3, C = 64;
constexpr int TW = 128, ROW = TW + 2;
constexpr int EPI_WARPS = 8;                 // two per TMEM lane quadrant: each takes 32 of the 64 channel columns
constexpr int NTHREADS = (6 + EPI_WARPS) * 32;   // warp 0 TMA, warp 1 MMA, warps 2-5 hidden layer, warps 6-13 epilogue
constexpr int A_CHUNK = TW * 16;             // one 16-byte K-chunk over 128 rows
constexpr int A_BYTES = 8 * A_CHUNK;         // h_hi (4 chunks) | h_lo (4 chunks)
constexpr int B_CHUNK = C * 16;
constexpr int B_BYTES = 10 * B_CHUNK;        // W1_hi | W1_lo | bias chunk | zero chunk
constexpr int ONES_BYTES = 2 * A_CHUNK;
constexpr int DTW = TW + 8;                  // input box width: pixels [w0-4, w0+132)
constexpr int IN_BYTES = C * DTW * 4;        // 34 KB input box
constexpr int TILE_BYTES = C * TW * 4;       // 32 KB output tile
constexpr int NS_D = 3, NS_O = 2, NS_A = 2, NS_T = 4;
constexpr uint32_t TMEM_COLS = NS_T * C;     // 256
constexpr int BAR_HID = 1, BAR_EPI = 2;      // named barriers (0 = __syncthreads)

// eight 2-byte floats (FMT 1: bf16, 2: fp16) -> fp32
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

struct Smem {
  alignas(1024) float dtile[NS_D][C * DTW];
  alignas(1024) float otile[NS_O][C * TW];
  alignas(1024) unsigned char a[NS_A][A_BYTES];
  alignas(1024) unsigned char bw[B_BYTES];
  alignas(1024) unsigned char ones[ONES_BYTES];
  alignas(16) float4 w0b[HID];
  alignas(16) float cs[3 * CCH * ROW];
  alignas(8) uint64_t d_full[NS_D], d_empty[NS_D], a_full[NS_A], a_empty[NS_A], t_full[NS_T], t_empty[NS_T];
  uint32_t tmem_slot;
};

// MODE 0: forward (tap tile = features, output = 9 product planes per channel, NCHW fp32)
// MODE 1: grad_data (tap tile = grad_out plane 8-k at the mirrored pixel, output = sum over taps)
// MODE 2: forward fused with the point_wise_mlp BatchNorm + ReLU of meta_kernel_conv
//         (dla_backbone.py:92-94): y = relu(out * scale + shift) written as haloed NHWC bf16 with
//         TAP-MAJOR channels (k*64 + c) -- the aggregation 1x1 conv's weight is permuted to match --
//         so each tap is one 64-channel x 128-pixel 128B-swizzled TMA store.
//
// GOF (MODE 1 only): storage of grad_out.  0: (B, C*9, H, W) fp32, channel c*9+k -- the reference op boundary
// (meta_kernel.py:232-239).  1 / 2: haloed NHWC bf16 / fp16 with TAP-MAJOR channels (k*64 + c), which is what the
// BatchNorm backward of the training graph writes: the tap box is then [136 px][64 ch] x 2 B with the 128-byte swizzle
// (half the bytes, and eight 16-byte shared loads per tap and thread instead of 64 scalar ones).
template <int MODE, bool PROF, int GOF = 0>
__global__ void __launch_bounds__(NTHREADS, 1)
meta_ws_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
               const float* __restrict__ coord, const float* __restrict__ w0, const float* __restrict__ b0,
               const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ ep_scale,
               const float* __restrict__ ep_shift, int ep_relu, int B, int H, int W, int tiles_w, int ntiles,
               long long* __restrict__ prof) {
  // prof (diagnostic, RD_MK_PROF=1, else null): per-CTA clock64() sums, [block][16]:
  // 0 producer total, 1 wait d_empty | 2 builders total, 3 wait a_empty, 4 coordinate staging |
  // 5 MMA total, 6 wait a_full, 7 wait t_empty | 8 epilogue total, 9 wait t_full, 10 wait d_full,
  // 11 wait store-read + barrier, 12 body, 13 second barrier + store issue
  auto tick = [&]() -> long long { return PROF ? clock64() : 0ll; };
  long long pa = 0, pb = 0, pc = 0, pd = 0, pe = 0;
  const long long t_begin = tick();
  extern __shared__ unsigned char smem_raw[];
  // TMA destinations need 128-byte (we use 1024) alignment; the dynamic window is only guaranteed 16
  Smem& S = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

  // ---- setup ---------------------------------------------------------------------------------
  if (t == 0) {
    for (int i = 0; i < NS_D; ++i) { tc::mbar_init(&S.d_full[i], 1); tc::mbar_init(&S.d_empty[i], EPI_WARPS); }
    for (int i = 0; i < NS_A; ++i) { tc::mbar_init(&S.a_full[i], 4); tc::mbar_init(&S.a_empty[i], 1); }
    for (int i = 0; i < NS_T; ++i) { tc::mbar_init(&S.t_full[i], 1); tc::mbar_init(&S.t_empty[i], EPI_WARPS); }
    tc::fence_mbar_init();
    tma::prefetch_map(&tm_in);
    tma::prefetch_map(&tm_out);
  }
  if (warp == 1) {
    tc::tmem_alloc(&S.tmem_slot, TMEM_COLS);
    tc::tmem_relinquish();
  }
  if (warp >= 2 && warp < 6) {  // constant operands: W1 hi/lo, bias slice, ones slice, layer-0 params
    const int ht = t - 64;
    for (int j = ht; j < HID; j += 128)
      S.w0b[j] = make_float4(__ldg(w0 + j * 3 + 0), __ldg(w0 + j * 3 + 1), __ldg(w0 + j * 3 + 2), __ldg(b0 + j));
    for (int e = ht; e < C * 4; e += 128) {
      const int c = e % C, q = e / C;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        float h0, l0, h1, l1;
        tc::split_bf16(__ldg(w1 + c * HID + q * 8 + 2 * p), h0, l0);
        tc::split_bf16(__ldg(w1 + c * HID + q * 8 + 2 * p + 1), h1, l1);
        hi[p] = tc::pack_bf16x2(h0, h1);
        lo[p] = tc::pack_bf16x2(l0, l1);
      }
      *reinterpret_cast<uint4*>(S.bw + q * B_CHUNK + c * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(S.bw + (4 + q) * B_CHUNK + c * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    for (int c = ht; c < C; c += 128) {
      float h, l;
      tc::split_bf16(__ldg(b1 + c), h, l);
      *reinterpret_cast<uint4*>(S.bw + 8 * B_CHUNK + c * 16) = make_uint4(tc::pack_bf16x2(h, l), 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(S.bw + 9 * B_CHUNK + c * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    *reinterpret_cast<uint4*>(S.ones + ht * 16) = make_uint4(tc::pack_bf16x2(1.f, 1.f), 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(S.ones + A_CHUNK + ht * 16) = make_uint4(0u, 0u, 0u, 0u);
    tc::fence_proxy_async_smem();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = S.tmem_slot;

  // ---- roles ---------------------------------------------------------------------------------
  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int wt = tile % tiles_w, h = (tile / tiles_w) % H, b = tile / (tiles_w * H);
        const int w0px = wt * TW;
        const int bs = w0px >= 4 ? w0px - 4 : 0;  // aligned, non-negative box start
        constexpr int UNITS = MODE != 1 ? 3 : 9;  // fwd: one box per neighbour row; bwd: one per tap
        for (int u = 0; u < UNITS; ++u, ++g) {
          const int dy = MODE != 1 ? u - 1 : u / 3 - 1;
          const uint32_t s = g % NS_D, ph = (g / NS_D) & 1;
          const long long ta = tick();
          tc::mbar_wait(&S.d_empty[s], ph ^ 1);
          pa += tick() - ta;
          const int hh = h + dy;
          if (hh >= 0 && hh < H) {
            tc::mbar_arrive_expect_tx(&S.d_full[s], GOF ? IN_BYTES / 2 : IN_BYTES);
            if (MODE != 1) tma::load_3d(S.dtile[s], &tm_in, &S.d_full[s], bs, hh, b * C);
            else if (GOF) tma::load_4d(S.dtile[s], &tm_in, &S.d_full[s], (8 - u) * C, bs + 1, hh + 1, b);
            else tma::load_4d(S.dtile[s], &tm_in, &S.d_full[s], bs, hh, 8 - u, b * C);
          } else {
            tc::mbar_arrive(&S.d_full[s]);  // row outside the image: nothing to load, consumers use 0
          }
        }
      }
      if (PROF) { prof[blockIdx.x * 16 + 0] = tick() - t_begin; prof[blockIdx.x * 16 + 1] = pa; }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // The warp runs converged and one elected lane issues: measured with scripts/mma_bench2.cu, a lean
    // loop sustains the tensor-core floor (48 cycles per M128 N64 K16 MMA with both operands in shared
    // memory), while per-MMA descriptor rebuilding under `if (lane == 0)` costs 130+ cycles of the issuing
    // thread.  Descriptors differ only in their start-address field: one 64-bit add each.
    // Split-bf16 product x.w ~ xh.wh + xl.wh + xh.wl ; the xl.wl term (2^-16 relative) is dropped.
    {
      constexpr uint32_t idesc = tc::make_idesc_bf16(128, C);
      const uint64_t a_hi = tc::make_smem_desc(0, A_CHUNK, 128, tc::LAYOUT_NONE);
      const uint64_t b_hi = tc::make_smem_desc(0, B_CHUNK, 128, tc::LAYOUT_NONE);
      const uint64_t bdesc = b_hi | (uint64_t)(tc::smem_u32(S.bw) >> 4);
      const uint64_t odesc = a_hi | (uint64_t)(tc::smem_u32(S.ones) >> 4);
      constexpr uint32_t AC = A_CHUNK >> 4, BC = B_CHUNK >> 4;  // chunk strides in descriptor address units
      const bool leader = tc::elect_one();
      uint32_t sa = 0, pha = 0, st = 0, pht = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll 1
        for (int k = 0; k < 9; ++k) {
          const long long ta = tick();
          tc::mbar_wait(&S.a_full[sa], pha);
          const long long tb = tick();
          tc::mbar_wait(&S.t_empty[st], pht ^ 1);
          pa += tb - ta;
          pb += tick() - tb;
          tc::tc_fence_after();
          if (leader) {
            const uint32_t d_tmem = tmem_base + st * C;
            const uint64_t adesc = a_hi | (uint64_t)(tc::smem_u32(S.a[sa]) >> 4);
            // (A part, B part): (hi,hi) (lo,hi) (hi,lo); two K=16 slices each; then the bias slice
            tc::mma_bf16_ss(d_tmem, adesc, bdesc, idesc, 0u);
            tc::mma_bf16_ss_acc(d_tmem, adesc + 2 * AC, bdesc + 2 * BC, idesc);
            tc::mma_bf16_ss_acc(d_tmem, adesc + 4 * AC, bdesc, idesc);
            tc::mma_bf16_ss_acc(d_tmem, adesc + 6 * AC, bdesc + 2 * BC, idesc);
            tc::mma_bf16_ss_acc(d_tmem, adesc, bdesc + 4 * BC, idesc);
            tc::mma_bf16_ss_acc(d_tmem, adesc + 2 * AC, bdesc + 6 * BC, idesc);
            tc::mma_bf16_ss_acc(d_tmem, odesc, bdesc + 8 * BC, idesc);
            tc::umma_commit(&S.a_empty[sa]);  // A slot reusable once these MMAs have read it
            tc::umma_commit(&S.t_full[st]);   // accumulator ready
          }
          __syncwarp();
          if (++sa == NS_A) { sa = 0; pha ^= 1; }
          if (++st == NS_T) { st = 0; pht ^= 1; }
        }
      }
      if (PROF && lane == 0) {
        prof[blockIdx.x * 16 + 5] = tick() - t_begin;
        prof[blockIdx.x * 16 + 6] = pa;
        prof[blockIdx.x * 16 + 7] = pb;
      }
    }
  } else if (warp < 6) {
    // ===== hidden-layer producers (128 threads, thread = pixel) =====
    const int ht = t - 64;
    uint32_t g = 0;
    // Coordinate tile (3 rows x 3 channels x 130 columns) of a tile -> registers.  Issued one tile AHEAD:
    // the per-role cycle profile (RD_MK_PROF) showed 10.9k of the builders' 25.5k cycles per tile spent in
    // this staging step (exposed global-load latency between two barriers), and the builders are the
    // critical path (the MMA warp waited 80 % of the time for operand tiles).
    constexpr int CS_PER_THREAD = (3 * CCH * ROW + 127) / 128;
    auto load_coords = [&](int tile, float (&cp)[CS_PER_THREAD]) {
      const int wt = tile % tiles_w, h = (tile / tiles_w) % H, b = tile / (tiles_w * H);
      const int w0px = wt * TW;
#pragma unroll
      for (int i = 0; i < CS_PER_THREAD; ++i) {
        const int e = ht + i * 128;
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
      const long long tc0 = tick();
      tma::named_bar_sync(BAR_HID, 128);  // everyone finished reading the previous coordinate tile
#pragma unroll
      for (int i = 0; i < CS_PER_THREAD; ++i) {
        const int e = ht + i * 128;
        if (e < 3 * CCH * ROW) S.cs[e] = cpre[i];  // e == (r * CCH + d) * ROW + col
      }
      tma::named_bar_sync(BAR_HID, 128);
      if (tile + (int)gridDim.x < ntiles) load_coords(tile + gridDim.x, cpre);  // in flight during this tile
      pb += tick() - tc0;
      const float c0 = S.cs[(1 * CCH + 0) * ROW + ht + 1];
      const float c1 = S.cs[(1 * CCH + 1) * ROW + ht + 1];
      const float c2 = S.cs[(1 * CCH + 2) * ROW + ht + 1];
      for (int k = 0; k < 9; ++k, ++g) {
        const int dy = k / 3 - 1, dx = k % 3 - 1;
        const int col = ht + 1 + dx, r = dy + 1;
        float r0 = S.cs[(r * CCH + 0) * ROW + col] - c0;
        float r1 = S.cs[(r * CCH + 1) * ROW + col] - c1;
        float r2 = S.cs[(r * CCH + 2) * ROW + col] - c2;
        if (MODE == 1) { r0 = -r0; r1 = -r1; r2 = -r2; }  // rel seen from the mirrored pixel
        const uint32_t sa = g % NS_A, pha = (g / NS_A) & 1;
        const long long ta = tick();
        tc::mbar_wait(&S.a_empty[sa], pha ^ 1);
        pa += tick() - ta;
        unsigned char* abuf = S.a[sa];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            float hv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const float4 wv = S.w0b[q * 8 + 2 * p + u];
              float z = wv.w;
              z = fmaf(wv.x, r0, z);
              z = fmaf(wv.y, r1, z);
              z = fmaf(wv.z, r2, z);
              hv[u] = fmaxf(z, 0.f);
            }
            tc::split_pack_bf16x2(hv[0], hv[1], hi[p], lo[p]);
          }
          *reinterpret_cast<uint4*>(abuf + q * A_CHUNK + ht * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(abuf + (4 + q) * A_CHUNK + ht * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&S.a_full[sa]);
      }
    }
    if (PROF && ht == 0) {
      prof[blockIdx.x * 16 + 2] = tick() - t_begin;
      prof[blockIdx.x * 16 + 3] = pa;
      prof[blockIdx.x * 16 + 4] = pb;
    }
  } else {
    // ===== epilogue (256 threads: thread = TMEM lane = pixel, two warps per lane quadrant, 32 channel columns each) =====
    // (Four warps -- one per scheduler -- spent 9.7 k of 17 k cycles per tile in this body with nothing to switch to while
    //  a TMEM load or a shared-memory round trip was in flight; the MMA warp waited 8.3 k on t_empty.)
    const int q4 = warp & 3;                 // TMEM lane quadrant this warp may read
    const int px = q4 * 32 + lane;
    const int half = (warp - 6) >> 2;        // channel columns [32 half, 32 half + 32)
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const bool leader = (warp == 6 && lane == 0);
    constexpr int EPI_THREADS = EPI_WARPS * 32;
    uint32_t g = 0, n_store = 0;
    float gd[MODE == 1 ? C / 2 : 1];
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int wt = tile % tiles_w, h = (tile / tiles_w) % H, b = tile / (tiles_w * H);
      const int w0px = wt * TW;
      if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < C / 2; ++i) gd[i] = 0.f;
      }
      const int bs = w0px >= 4 ? w0px - 4 : 0;
      for (int k = 0; k < 9; ++k, ++g) {
        const int dy = k / 3 - 1, dx = k % 3 - 1;
        const uint32_t st = g % NS_T, pht = (g / NS_T) & 1;
        const uint32_t gu = MODE != 1 ? g / 3 : g;  // input unit (row box / tap box) this tap reads
        const uint32_t sd = gu % NS_D, phd = (gu / NS_D) & 1;
        const long long e0 = tick();
        tc::mbar_wait(&S.t_full[st], pht);
        const long long e1 = tick();
        if (MODE == 1 || dx == -1) tc::mbar_wait(&S.d_full[sd], phd);
        __syncwarp();
        tc::tc_fence_after();
        const long long e2 = tick();
        pa += e1 - e0;
        pb += e2 - e1;
        // pixel read by this thread for this tap, as a column of the input box (or: outside -> 0)
        const int col = w0px + px + dx - bs;
        const bool ok = (h + dy >= 0) && (h + dy < H) && col >= 0;
        const float* dt = S.dtile[sd] + (ok ? col : 0);
        const bool last_use = MODE == 1 || dx == 1;
        if (MODE == 2) {
          const uint32_t so = n_store % NS_O;
          if (leader) tma::store_wait_read<NS_O - 1>();
          tma::named_bar_sync(BAR_EPI, EPI_THREADS);
          unsigned char* ot = reinterpret_cast<unsigned char*>(S.otile[so]);  // [128 px][64 ch] bf16, 128B swizzle
          const float* esc = ep_scale + k * C;
          const float* esh = ep_shift + k * C;
          {
            float d[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) d[i] = ok ? dt[(half * 32 + i) * DTW] : 0.f;
            float v[32];
            tc::tmem_ld_x32(tmem_base + lane_sel + st * C + half * 32, v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int c = half * 32 + j * 8 + 2 * e;
                float a = fmaf(d[j * 8 + 2 * e] * v[j * 8 + 2 * e], __ldg(esc + c), __ldg(esh + c));
                float bb = fmaf(d[j * 8 + 2 * e + 1] * v[j * 8 + 2 * e + 1], __ldg(esc + c + 1), __ldg(esh + c + 1));
                if (ep_relu & 1) { a = fmaxf(a, 0.f); bb = fmaxf(bb, 0.f); }
                pk[e] = (ep_relu & 2) ? tc::pack_f16x2(a, bb) : tc::pack_bf16x2(a, bb);   // bit 1: fp16 storage
              }
              const int chunk = half * 4 + j;
              *reinterpret_cast<uint4*>(ot + px * 128 + ((chunk ^ (px & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
          tc::tc_fence_before();
          tc::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tc::mbar_arrive(&S.t_empty[st]);
            if (last_use) tc::mbar_arrive(&S.d_empty[sd]);
          }
          tma::named_bar_sync(BAR_EPI, EPI_THREADS);
          if (leader) {
            tma::store_4d(&tm_out, ot, k * C, w0px, h, b);
            tma::store_commit();
          }
          ++n_store;
        } else if (MODE == 0) {
          const uint32_t so = n_store % NS_O;
          if (leader) tma::store_wait_read<NS_O - 1>();  // the store that last used otile[so] has read it
          tma::named_bar_sync(BAR_EPI, EPI_THREADS);
          const long long e3 = tick();
          pc += e3 - e2;
          float* ot = S.otile[so];
          {
            // all 32 tap-tile loads first: the stores below may alias them as far as the compiler
            // knows, so interleaving would expose the full shared-memory latency per element
            float d[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) d[i] = ok ? dt[(half * 32 + i) * DTW] : 0.f;
            float v[32];
            tc::tmem_ld_x32(tmem_base + lane_sel + st * C + half * 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) ot[(half * 32 + i) * TW + px] = d[i] * v[i];
          }
          tc::tc_fence_before();
          tc::fence_proxy_async_smem();
          __syncwarp();
          const long long e4 = tick();
          pd += e4 - e3;
          if (lane == 0) {
            tc::mbar_arrive(&S.t_empty[st]);
            if (last_use) tc::mbar_arrive(&S.d_empty[sd]);
          }
          tma::named_bar_sync(BAR_EPI, EPI_THREADS);
          if (leader) {
            tma::store_4d(&tm_out, ot, w0px, h, k, b * C);
            tma::store_commit();
          }
          pe += tick() - e4;
          ++n_store;
        } else {
          {
            float d[32];
            if (GOF) {   // [px][64 ch] 2-byte elements, 16-byte chunk j of row r stored at chunk j ^ (r & 7)
              const int r = ok ? col : 0;
              const unsigned char* rowp = reinterpret_cast<const unsigned char*>(S.dtile[sd]) + r * 128;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 q = *reinterpret_cast<const uint4*>(rowp + (((half * 4 + j) ^ (r & 7)) << 4));
                unpack8_2b<GOF>(q, d + j * 8);
              }
              if (!ok) {
#pragma unroll
                for (int i = 0; i < 32; ++i) d[i] = 0.f;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) d[i] = ok ? dt[(half * 32 + i) * DTW] : 0.f;
            }
            float v[32];
            tc::tmem_ld_x32(tmem_base + lane_sel + st * C + half * 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) gd[i] = fmaf(d[i], v[i], gd[i]);
          }
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            tc::mbar_arrive(&S.t_empty[st]);
            tc::mbar_arrive(&S.d_empty[sd]);
          }
        }
      }
      if (MODE == 1) {  // one output tile per image tile
        const uint32_t so = n_store % NS_O;
        if (leader) tma::store_wait_read<NS_O - 1>();
        tma::named_bar_sync(BAR_EPI, EPI_THREADS);
        float* ot = S.otile[so];
#pragma unroll
        for (int c = 0; c < C / 2; ++c) ot[(half * 32 + c) * TW + px] = gd[c];
        tc::fence_proxy_async_smem();
        tma::named_bar_sync(BAR_EPI, EPI_THREADS);
        if (leader) {
          tma::store_3d(&tm_out, ot, w0px, h, b * C);
          tma::store_commit();
        }
        ++n_store;
      }
    }
    if (leader) tma::store_wait_all<0>();
    if (PROF && leader) {
      prof[blockIdx.x * 16 + 8] = tick() - t_begin;
      prof[blockIdx.x * 16 + 9] = pa;
      prof[blockIdx.x * 16 + 10] = pb;
      prof[blockIdx.x * 16 + 11] = pc;
      prof[blockIdx.x * 16 + 12] = pd;
      prof[blockIdx.x * 16 + 13] = pe;
    }
  }

  // ---- teardown ------------------------------------------------------------------------------
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

inline int launch(int mode, const void* tap_src_v, void* dst, const float* coord, const float* w0, const float* b0,
                  const float* w1, const float* b1, const float* ep_scale, const float* ep_shift, int ep_relu, int B,
                  int Cc, int H, int W, cudaStream_t stream, int gof = 0) {
  const float* tap_src = static_cast<const float*>(tap_src_v);
  RD_REQUIRE(Cc == C, "Meta-Kernel impl 3 (TMA/tcgen05) is specialised for C == 64 (got %d)", Cc);
  RD_REQUIRE(W % 4 == 0, "Meta-Kernel impl 3 needs W %% 4 == 0 (TMA row stride must be a multiple of 16 B); W=%d", W);
  RD_REQUIRE((reinterpret_cast<uintptr_t>(tap_src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
             "Meta-Kernel impl 3 needs 16-byte aligned tensors");
  const int tiles_w = (W + TW - 1) / TW;
  const int64_t ntiles = (int64_t)B * H * tiles_w;
  RD_REQUIRE(ntiles <= 0x7fffffffLL, "rd_meta_kernel: too many tiles");
  CUtensorMap tm_in, tm_out;
  const uint64_t plane = (uint64_t)H * W * 4;
  // features / grad_data viewed as (W, H, B*C); out / grad_out viewed as (W, H, 9, B*C)
  const uint64_t d3[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)B * C};
  const uint64_t s3[2] = {(uint64_t)W * 4, plane};
  const uint32_t b3[3] = {(uint32_t)TW, 1u, (uint32_t)C};       // output tile
  const uint32_t b3in[3] = {(uint32_t)DTW, 1u, (uint32_t)C};   // input row box
  const uint64_t d4[4] = {(uint64_t)W, (uint64_t)H, 9u, (uint64_t)B * C};
  const uint64_t s4[3] = {(uint64_t)W * 4, plane, plane * 9};
  const uint32_t b4[4] = {(uint32_t)TW, 1u, 1u, (uint32_t)C};
  const uint32_t b4in[4] = {(uint32_t)DTW, 1u, 1u, (uint32_t)C};
  if (mode == 0) {
    if (tma::make_map(&tm_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, tap_src, 3, d3, s3, b3in, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
    if (tma::make_map(&tm_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, dst, 4, d4, s4, b4, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  } else if (mode == 1) {
    if (gof) {   // whole haloed NHWC tensor (9C, W+2, H+2, B): pixel (h, w) sits at (w+1, h+1); the halo is zero
      const uint64_t Wp = (uint64_t)W + 2, Hp = (uint64_t)H + 2, CO = 9u * C;
      const uint64_t dn[4] = {CO, Wp, Hp, (uint64_t)B};
      const uint64_t sn[3] = {CO * 2, Wp * CO * 2, Hp * Wp * CO * 2};
      const uint32_t bn[4] = {(uint32_t)C, (uint32_t)DTW, 1u, 1u};
      if (tma::make_map(&tm_in, gof == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, tap_src_v, 4, dn,
                        sn, bn, CU_TENSOR_MAP_SWIZZLE_128B))
        return 1;
    } else if (tma::make_map(&tm_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, tap_src, 4, d4, s4, b4in, CU_TENSOR_MAP_SWIZZLE_NONE)) {
      return 1;
    }
    if (tma::make_map(&tm_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, dst, 3, d3, s3, b3, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  } else {
    if (tma::make_map(&tm_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, tap_src, 3, d3, s3, b3in, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
    // interior view of the haloed NHWC bf16 output (9C, W, H, B)
    const uint64_t Wp = (uint64_t)W + 2, Hp = (uint64_t)H + 2, CO = 9u * C;
    const uint64_t dn[4] = {CO, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t sn[3] = {CO * 2, Wp * CO * 2, Hp * Wp * CO * 2};
    const uint32_t bn[4] = {(uint32_t)C, (uint32_t)TW, 1u, 1u};
    const char* y_int = static_cast<const char*>(dst) + (Wp + 1) * CO * 2;
    if (tma::make_map(&tm_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, y_int, 4, dn, sn, bn, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  int dev = 0, sms = 0;
  RD_CUDA(cudaGetDevice(&dev));
  RD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = sizeof(Smem) + 1024;
  const int64_t grid = ntiles < sms ? ntiles : sms;
  static const bool want_prof = [] { const char* e = getenv("RD_MK_PROF"); return e && e[0] == '1'; }();
  long long* d_prof = nullptr;
  if (want_prof) {
    static long long* buf = nullptr;
    if (!buf) RD_CUDA(cudaMalloc(&buf, 1024 * 16 * sizeof(long long)));
    RD_CUDA(cudaMemsetAsync(buf, 0, 1024 * 16 * sizeof(long long), stream));
    d_prof = buf;
  }
  RD_CUDA(rd::smem_optin(meta_ws_kernel<0, false>, smem));
  RD_CUDA(rd::smem_optin(meta_ws_kernel<1, false>, smem));
  RD_CUDA(rd::smem_optin(meta_ws_kernel<2, false>, smem));
  RD_CUDA(rd::smem_optin(meta_ws_kernel<0, true>, smem));
  RD_CUDA(rd::smem_optin(meta_ws_kernel<1, true>, smem));
  RD_CUDA(rd::smem_optin(meta_ws_kernel<1, false, 1>, smem));
  RD_CUDA(rd::smem_optin(meta_ws_kernel<1, false, 2>, smem));
#define RD_MKWS_LAUNCH(M, PR, SC, SH, RELU)                                                                              \
  meta_ws_kernel<M, PR><<<(unsigned)grid, NTHREADS, smem, stream>>>(tm_in, tm_out, coord, w0, b0, w1, b1, SC, SH, RELU, B, H, \
                                                                    W, tiles_w, (int)ntiles, d_prof)
  if (mode == 0) {
    if (want_prof) RD_MKWS_LAUNCH(0, true, nullptr, nullptr, 0);
    else RD_MKWS_LAUNCH(0, false, nullptr, nullptr, 0);
  } else if (mode == 2) {
    RD_MKWS_LAUNCH(2, false, ep_scale, ep_shift, ep_relu);
  } else if (gof) {
    if (gof == 1)
      meta_ws_kernel<1, false, 1><<<(unsigned)grid, NTHREADS, smem, stream>>>(tm_in, tm_out, coord, w0, b0, w1, b1, nullptr, nullptr,
                                                                              0, B, H, W, tiles_w, (int)ntiles, d_prof);
    else
      meta_ws_kernel<1, false, 2><<<(unsigned)grid, NTHREADS, smem, stream>>>(tm_in, tm_out, coord, w0, b0, w1, b1, nullptr, nullptr,
                                                                              0, B, H, W, tiles_w, (int)ntiles, d_prof);
  } else {
    if (want_prof) RD_MKWS_LAUNCH(1, true, nullptr, nullptr, 0);
    else RD_MKWS_LAUNCH(1, false, nullptr, nullptr, 0);
  }
#undef RD_MKWS_LAUNCH
  rd::count_launch();
  if (want_prof && mode != 2 && !gof) {  // diagnostic: synchronous, prints mean cycles per tile and role
    RD_CUDA(cudaStreamSynchronize(stream));
    static long long hbuf[1024 * 16];
    RD_CUDA(cudaMemcpy(hbuf, d_prof, sizeof(long long) * 16 * grid, cudaMemcpyDeviceToHost));
    double m[16] = {0};
    for (int b = 0; b < grid; ++b)
      for (int k = 0; k < 16; ++k) m[k] += (double)hbuf[b * 16 + k] / grid;
    const double tl = (double)ntiles / grid;
    fprintf(stderr,
            "[rd_meta prof mode %d] tiles/cta=%.1f | per tile: producer %.0f (wait d_empty %.0f) | builders %.0f (wait a_empty "
            "%.0f, coord staging %.0f) | mma %.0f (wait a_full %.0f, wait t_empty %.0f) | epi %.0f (wait t_full %.0f, wait d_full "
            "%.0f, wait store+bar %.0f, body %.0f, bar+store %.0f)\n",
            mode, tl, m[0] / tl, m[1] / tl, m[2] / tl, m[3] / tl, m[4] / tl, m[5] / tl, m[6] / tl, m[7] / tl, m[8] / tl,
            m[9] / tl, m[10] / tl, m[11] / tl, m[12] / tl, m[13] / tl);
  }
  return rd::check_launch(mode == 1 ? "rd_meta_kernel_bwd_data(impl 3)" : "rd_meta_kernel_fwd(impl 3)");
}

}  // namespace mkws

int rd_meta_kernel_fwd_ws(const float* data, const float* coord, const float* w0, const float* b0, const float* w1,
                          const float* b1, float* out, int B, int C, int H, int W, cudaStream_t stream) {
  return mkws::launch(0, data, out, coord, w0, b0, w1, b1, nullptr, nullptr, 0, B, C, H, W, stream);
}
int rd_meta_kernel_bwd_data_ws(const float* grad_out, const float* coord, const float* w0, const float* b0,
                               const float* w1, const float* b1, float* grad_data, int B, int C, int H, int W,
                               cudaStream_t stream) {
  return mkws::launch(1, grad_out, grad_data, coord, w0, b0, w1, b1, nullptr, nullptr, 0, B, C, H, W, stream);
}

// grad_data from the haloed NHWC tap-major gradient (gof 1: bf16, 2: fp16)
int rd_meta_kernel_bwd_data_ws_nhwc(const void* grad_out_pad, int gof, const float* coord, const float* w0, const float* b0,
                                    const float* w1, const float* b1, float* grad_data, int B, int C, int H, int W,
                                    cudaStream_t stream) {
  return mkws::launch(1, grad_out_pad, grad_data, coord, w0, b0, w1, b1, nullptr, nullptr, 0, B, C, H, W, stream, gof);
}

extern "C" int rd_meta_kernel_fwd_nhwc_bf16(const float* data, const float* coord, const float* w0, const float* b0,
                                            const float* w1, const float* b1, const float* scale, const float* shift,
                                            int relu, void* y_pad, int B, int C, int H, int W, rd_stream_t stream) {
  RD_REQUIRE(data && coord && w0 && b0 && w1 && b1 && scale && shift && y_pad, "rd_meta_kernel_fwd_nhwc_bf16: null pointer");
  RD_REQUIRE(B > 0 && H > 0 && W > 0, "rd_meta_kernel_fwd_nhwc_bf16: bad shape");
  if (rd_check_device()) return 1;
  return mkws::launch(2, data, y_pad, coord, w0, b0, w1, b1, scale, shift, relu ? 1 : 0, B, C, H, W, rd::as_stream(stream));
}

// Same kernel, output stored as fp16 (the reference's training storage type, config:35): bit 1 of the epilogue flag.
extern "C" int rd_meta_kernel_fwd_nhwc_f16(const float* data, const float* coord, const float* w0, const float* b0,
                                           const float* w1, const float* b1, const float* scale, const float* shift,
                                           int relu, void* y_pad, int B, int C, int H, int W, rd_stream_t stream) {
  RD_REQUIRE(data && coord && w0 && b0 && w1 && b1 && scale && shift && y_pad, "rd_meta_kernel_fwd_nhwc_f16: null pointer");
  RD_REQUIRE(B > 0 && H > 0 && W > 0, "rd_meta_kernel_fwd_nhwc_f16: bad shape");
  if (rd_check_device()) return 1;
  return mkws::launch(2, data, y_pad, coord, w0, b0, w1, b1, scale, shift, (relu ? 1 : 0) | 2, B, C, H, W, rd::as_stream(stream));
}
