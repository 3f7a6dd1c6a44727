// Implicit-GEMM 2-D convolution (3x3 pad 1 or 1x1, stride 1) on tcgen05 tensor cores, NHWC bf16.
//
// Replaces the dense contractions of the DLA backbone / RPN head that the reference delegates to
// mx.sym.Convolution -> cuDNN (/root/reference mxnext/simple.py:123-158; call sites
// rangedet/symbol/backbone/dla_backbone.py:23-50,95 and rangedet/symbol/head/builder.py:198-266),
// with the following BatchNorm (inference form: per-channel scale/shift), ReLU and residual add
// (dla_backbone.py:31-56) folded into the epilogue.
//
// GEMM view per output tile:  M = 128 consecutive pixels of one image row, N = Cout (64 | 128),
// K = taps x Cin.  Activations live in HBM as zero-haloed NHWC bf16 [N][H+2][W+2][C] so that every
// tap is a plain TMA box (this part's TMA faults on negative coordinates); a box is 64 channels
// (128 B, one 128B-swizzle atom) x 128 pixels = the K-major A operand of four K=16 MMAs.  Weights
// are pre-packed [tap][Cout][Cin] bf16; when all taps fit (<= 80 KB) they stay resident in shared
// memory for the whole persistent CTA, otherwise they stream with the activations.
//   warp 0  TMA producer      warp 1  MMA issuer (tcgen05.mma M128 x Cout x K16, fp32 in TMEM,
//   warps 2-5 epilogue         two accumulators so tile i+1's MMAs overlap tile i's epilogue)
// Epilogue: tcgen05.ld -> x scale[c] + shift[c] (+ residual) -> ReLU -> bf16 -> 128B-swizzled
// staging tile -> TMA store into the interior view of the haloed output.
#include <cuda_bf16.h>

#include "../../include/rangedet_b200.h"
#include "rd_common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

namespace conv {

constexpr int TM = 128;              // pixels per tile
constexpr int KC = 64;               // channels per TMA box / swizzle atom
constexpr int A_BYTES = TM * KC * 2; // 16 KB
constexpr int NTHREADS = 192;
constexpr int BAR_EPI = 1;

struct Params {
  int N, H, W, Cin, Cout, taps;      // taps = 9 (3x3, pad 1) or 1 (1x1)
  int tiles_w, ntiles;
  int relu, has_residual;
  int b_resident, nstages;
  int a_off, b_off, o_off, misc_off; // byte offsets into the 1024-aligned dynamic shared memory
  int stage_bytes;
};

__global__ void __launch_bounds__(NTHREADS, 1)
conv_fprop_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                  const __grid_constant__ CUtensorMap tm_y, const float* __restrict__ scale,
                  const float* __restrict__ shift, const __nv_bfloat16* __restrict__ residual,
                  int64_t res_row_stride, int64_t res_img_stride, const Params P) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int kh = P.Cin / KC;                    // K boxes per tap
  const int nh = P.Cout / KC;                   // output halves
  const int b_tile = P.Cout * KC * 2;           // one (tap, k-half) weight tile
  unsigned char* sA = base + P.a_off;           // stages: [A boxes | (B tiles if streaming)]
  unsigned char* sB = base + P.b_off;           // resident weights
  unsigned char* sO = base + P.o_off;           // epilogue staging, nh x 16 KB
  uint64_t* full = reinterpret_cast<uint64_t*>(base + P.misc_off);
  uint64_t* empty = full + 8;
  uint64_t* t_full = empty + 8;
  uint64_t* t_empty = t_full + 2;
  uint64_t* w_full = t_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  float* s_scale = reinterpret_cast<float*>(tmem_slot + 2);
  float* s_shift = s_scale + 128;

  if (t == 0) {
    for (int i = 0; i < P.nstages; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&t_full[i], 1); tc::mbar_init(&t_empty[i], 4); }
    tc::mbar_init(w_full, 1);
    tc::fence_mbar_init();
    tma::prefetch_map(&tm_x);
    tma::prefetch_map(&tm_w);
    tma::prefetch_map(&tm_y);
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 256);
    tc::tmem_relinquish();
  }
  for (int c = t; c < P.Cout; c += NTHREADS) {
    s_scale[c] = scale ? __ldg(scale + c) : 1.f;
    s_shift[c] = shift ? __ldg(shift + c) : 0.f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int units = P.taps * kh;  // pipeline stages consumed per tile

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      if (P.b_resident) {  // all weight tiles once
        tc::mbar_arrive_expect_tx(w_full, (uint32_t)(units * b_tile));
        for (int tap = 0; tap < P.taps; ++tap)
          for (int q = 0; q < kh; ++q)
            tma::load_3d(sB + (tap * kh + q) * b_tile, &tm_w, w_full, q * KC, 0, tap);
      }
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        const int wt = tile % P.tiles_w, h = (tile / P.tiles_w) % P.H, n = tile / (P.tiles_w * P.H);
        const int w0 = wt * TM;
        for (int tap = 0; tap < P.taps; ++tap) {
          const int dy = P.taps == 9 ? tap / 3 : 1, dx = P.taps == 9 ? tap % 3 : 1;  // offsets into the halo frame
          for (int q = 0; q < kh; ++q, ++g) {
            const uint32_t s = g % P.nstages, ph = (g / P.nstages) & 1;
            tc::mbar_wait(&empty[s], ph ^ 1);
            unsigned char* st = sA + s * P.stage_bytes;
            tc::mbar_arrive_expect_tx(&full[s], (uint32_t)(A_BYTES + (P.b_resident ? 0 : b_tile)));
            tma::load_4d(st, &tm_x, &full[s], q * KC, w0 + dx, h + dy, n);
            if (!P.b_resident) tma::load_3d(st + A_BYTES, &tm_w, &full[s], q * KC, 0, tap);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc_bf16(TM, P.Cout);
      if (P.b_resident) tc::mbar_wait(w_full, 0);
      uint32_t g = 0, it = 0;
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1, pha = (it >> 1) & 1;
        tc::mbar_wait(&t_empty[acc], pha ^ 1);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 128;
        for (int u = 0; u < units; ++u, ++g) {
          const uint32_t s = g % P.nstages, ph = (g / P.nstages) & 1;
          tc::mbar_wait(&full[s], ph);
          tc::tc_fence_after();
          const uint32_t a_addr = tc::smem_u32(sA + s * P.stage_bytes);
          const uint32_t b_addr = P.b_resident ? tc::smem_u32(sB + u * b_tile) : a_addr + A_BYTES;
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks) {
            // 128B-swizzled K-major tiles: 8-row groups 1024 B apart, K advances 32 B inside the atom
            const uint64_t ad = tc::make_smem_desc(a_addr + ks * 32, 0, 1024, tc::LAYOUT_SW128);
            const uint64_t bd = tc::make_smem_desc(b_addr + ks * 32, 0, 1024, tc::LAYOUT_SW128);
            tc::mma_bf16_ss(d_tmem, ad, bd, idesc, (u > 0 || ks > 0) ? 1u : 0u);
          }
          tc::umma_commit(&empty[s]);
        }
        tc::umma_commit(&t_full[acc]);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: thread = pixel row = TMEM lane =====
    const int q4 = warp & 3;
    const int px = q4 * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const bool leader = (warp == 2 && lane == 0);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x, ++it) {
      const int wt = tile % P.tiles_w, h = (tile / P.tiles_w) % P.H, n = tile / (P.tiles_w * P.H);
      const int w0 = wt * TM;
      const uint32_t acc = it & 1, pha = (it >> 1) & 1;
      tc::mbar_wait(&t_full[acc], pha);
      __syncwarp();
      tc::tc_fence_after();
      if (leader) tma::store_wait_read<0>();  // previous tile's stores have read the staging buffer
      tma::named_bar_sync(BAR_EPI, 128);
      const bool in_img = (w0 + px) < P.W;
      const __nv_bfloat16* rrow =
          P.has_residual ? residual + (int64_t)n * res_img_stride + (int64_t)h * res_row_stride + (int64_t)(w0 + px) * P.Cout
                         : nullptr;
      for (int c0 = 0; c0 < P.Cout; c0 += 32) {
        float v[32];
        tc::tmem_ld_x32(tmem_base + lane_sel + acc * 128 + c0, v);
        uint4 rv[4];
        if (P.has_residual && in_img) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rv[j] = __ldg(reinterpret_cast<const uint4*>(rrow + c0) + j);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // 8 channels -> one 16-byte chunk
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = c0 + j * 8 + 2 * e;
            float a = fmaf(v[j * 8 + 2 * e], s_scale[c], s_shift[c]);
            float b = fmaf(v[j * 8 + 2 * e + 1], s_scale[c + 1], s_shift[c + 1]);
            if (P.has_residual && in_img) {
              const uint32_t w = reinterpret_cast<const uint32_t*>(&rv[j])[e];
              const __nv_bfloat162 r2 = *reinterpret_cast<const __nv_bfloat162*>(&w);
              a += __bfloat162float(r2.x);
              b += __bfloat162float(r2.y);
            }
            if (P.relu) {
              a = fmaxf(a, 0.f);
              b = fmaxf(b, 0.f);
            }
            pk[e] = tc::pack_bf16x2(a, b);
          }
          const int half = (c0 + j * 8) / KC;
          const int chunk = ((c0 + j * 8) % KC) / 8;
          // 128B swizzle: 16-byte chunk index XOR (row % 8)
          *reinterpret_cast<uint4*>(sO + half * A_BYTES + px * 128 + ((chunk ^ (px & 7)) << 4)) =
              make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      tc::tc_fence_before();
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&t_empty[acc]);
      tma::named_bar_sync(BAR_EPI, 128);
      if (leader) {
        for (int hf = 0; hf < nh; ++hf) tma::store_4d(&tm_y, sO + hf * A_BYTES, hf * KC, w0, h, n);
        tma::store_commit();
      }
    }
    if (leader) tma::store_wait_all<0>();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 256);
}

}  // namespace conv

extern "C" int rd_conv2d_nhwc_bf16(const void* x_pad, const void* w_packed, const float* scale, const float* shift,
                                   const void* residual_pad, void* y_pad, int N, int H, int W, int Cin, int Cout,
                                   int ksize, int relu, rd_stream_t stream) {
  using namespace conv;
  RD_REQUIRE(x_pad && w_packed && y_pad, "rd_conv2d_nhwc_bf16: null pointer");
  RD_REQUIRE(ksize == 3 || ksize == 1, "rd_conv2d_nhwc_bf16: kernel size must be 3 (pad 1) or 1 (got %d)", ksize);
  RD_REQUIRE((Cin == 64 || Cin == 128) && (Cout == 64 || Cout == 128),
             "rd_conv2d_nhwc_bf16: Cin and Cout must be 64 or 128 (got %d, %d); pad the channels", Cin, Cout);
  RD_REQUIRE(N > 0 && H > 0 && W > 0, "rd_conv2d_nhwc_bf16: bad shape");
  if (rd_check_device()) return 1;
  Params P;
  P.N = N; P.H = H; P.W = W; P.Cin = Cin; P.Cout = Cout; P.taps = ksize * ksize;
  P.tiles_w = (W + TM - 1) / TM;
  const int64_t ntiles = (int64_t)N * H * P.tiles_w;
  RD_REQUIRE(ntiles <= 0x7fffffffLL, "rd_conv2d_nhwc_bf16: too many tiles");
  P.ntiles = (int)ntiles;
  P.relu = relu ? 1 : 0;
  P.has_residual = residual_pad ? 1 : 0;
  const int kh = Cin / KC, nh = Cout / KC;
  const int b_tile = Cout * KC * 2;
  const int w_bytes = P.taps * kh * b_tile;
  P.b_resident = w_bytes <= 80 * 1024 ? 1 : 0;
  P.stage_bytes = A_BYTES + (P.b_resident ? 0 : b_tile);
  const int o_bytes = nh * A_BYTES;
  const int misc = 2048;
  const int budget = 220 * 1024 - o_bytes - misc - (P.b_resident ? w_bytes : 0);
  int ns = budget / P.stage_bytes;
  if (ns > 8) ns = 8;
  RD_REQUIRE(ns >= 2, "rd_conv2d_nhwc_bf16: shared memory budget too small");
  P.nstages = ns;
  P.a_off = 0;
  P.b_off = ns * P.stage_bytes;
  P.o_off = P.b_off + (P.b_resident ? w_bytes : 0);
  P.misc_off = P.o_off + o_bytes;
  const size_t smem = (size_t)P.misc_off + misc + 1024;

  CUtensorMap tm_x, tm_w, tm_y;
  const uint64_t Wp = (uint64_t)W + 2, Hp = (uint64_t)H + 2;
  {  // haloed input (C, W+2, H+2, N)
    const uint64_t d[4] = {(uint64_t)Cin, Wp, Hp, (uint64_t)N};
    const uint64_t s[3] = {(uint64_t)Cin * 2, Wp * Cin * 2, Hp * Wp * Cin * 2};
    const uint32_t b[4] = {(uint32_t)KC, (uint32_t)TM, 1u, 1u};
    if (tma::make_map(&tm_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, x_pad, 4, d, s, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {  // packed weights (Cin, Cout, taps)
    const uint64_t d[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)P.taps};
    const uint64_t s[2] = {(uint64_t)Cin * 2, (uint64_t)Cin * Cout * 2};
    const uint32_t b[3] = {(uint32_t)KC, (uint32_t)Cout, 1u};
    if (tma::make_map(&tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w_packed, 3, d, s, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {  // interior view of the haloed output (C, W, H, N): stores are clipped at W, never touch the halo
    const uint64_t d[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    const uint64_t s[3] = {(uint64_t)Cout * 2, Wp * Cout * 2, Hp * Wp * Cout * 2};
    const uint32_t b[4] = {(uint32_t)KC, (uint32_t)TM, 1u, 1u};
    const char* y_int = static_cast<const char*>(y_pad) + (Wp + 1) * Cout * 2;
    if (tma::make_map(&tm_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, y_int, 4, d, s, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  const __nv_bfloat16* res = nullptr;
  if (residual_pad) res = static_cast<const __nv_bfloat16*>(residual_pad) + (Wp + 1) * Cout;  // interior origin
  int dev = 0, sms = 0;
  RD_CUDA(cudaGetDevice(&dev));
  RD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static size_t smem_set = 0;
  if (smem > smem_set) {
    RD_CUDA(cudaFuncSetAttribute(conv_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  const int grid = P.ntiles < sms ? P.ntiles : sms;
  conv_fprop_kernel<<<grid, NTHREADS, smem, rd::as_stream(stream)>>>(tm_x, tm_w, tm_y, scale, shift, res,
                                                                      (int64_t)Wp * Cout, (int64_t)Hp * Wp * Cout, P);
  rd::count_launch();
  return rd::check_launch("rd_conv2d_nhwc_bf16");
}
