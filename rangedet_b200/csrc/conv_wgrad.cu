// Weight gradient of the NHWC bf16 convolution family on tcgen05 tensor cores (sm_100a).
//
// Backward of mx.sym.Convolution / mx.sym.Deconvolution w.r.t. the weight, which the reference gets
// from MXNet/cuDNN for every layer of the DLA backbone and RPN head (/root/reference
// mxnext/simple.py:123-158,545-580; call sites rangedet/symbol/backbone/dla_backbone.py:17-127,
// rangedet/symbol/head/builder.py:198-266).  One generic contraction over pixels:
//
//     G[tap][a][b] = sum over (n, h, w) of  A[n, h, w][a] * B[n, h + dy, w*stride + dx][b]
//
// A (CA = 64 or 128 channels, e.g. the gradient dz of the conv output) and B (CB channels, e.g. the
// layer input x) are zero-haloed NHWC bf16, G is fp32 [taps][CA][CB] -- exactly the packed layout of the
// forward weights ([tap][Cout][Cin]) when A = dz, B = x.  For a transposed convolution the roles swap
// (A = layer input, B = dz in its phase-grouped view), see rangedet_b200/train.py.
//
// GEMM view: the contraction dimension K is the PIXEL index, which is the outer dimension of both NHWC
// operands, so both MMA operands are MN-major: a TMA box [128 px][64 ch] with the 128B swizzle is
// exactly the canonical MN-major SW128 tile (8 chunk x 8 row atoms, atoms of 64 channels LBO apart,
// 8-pixel groups SBO = 1024 B apart).  M = CA (rows 64..127 are don't-care when CA = 64), N = a CB
// chunk, K = 16 pixels per tcgen05.mma, 8 MMAs per 128-pixel tile and tap.  The three dx taps of one
// dy are row-shifted views (+128 B per pixel) of ONE 130-pixel strip of B, each with its own TMEM
// accumulator.  A CTA = (job, split): a job is one dy row of taps x one CB chunk (<= 512 TMEM columns),
// a split is a contiguous range of pixel tiles; accumulators stay in TMEM for the whole CTA, partial
// results go to the workspace and a second kernel adds the splits in fixed order (deterministic).
//   warp 0 TMA producer | warp 1 MMA issuer | warps 2-5 final TMEM -> workspace
// RD_WGRAD_NOSWZ=1 selects the no-swizzle canonical layout instead (8-channel TMA boxes; the layout
// tests/test_gpu_parity.py::test_tc_probe validates) -- a diagnostic cross-check of the SW128 path.
#include <stdlib.h>
#include <string.h>

#include "../../include/rangedet_b200.h"
#include "act_type.cuh"
#include "rd_common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

namespace RD_ACT_NS(wg) {

constexpr int TM = 128;        // pixels per K tile
constexpr int NTHREADS = 192;
constexpr int MAX_JOBS = 48, MAX_STAGES = 6;

struct Job {
  int dy, dx0, ndx, cb0, tap0;
};

struct Plan {
  int N, H, W;                 // A resolution
  int CA, CB, CBJ;
  int ksize, stride_w, ntaps;
  int njobs, nsplit, tiles_w, ntiles;
  int swz;                     // 1: SW128 tiles (64-channel boxes), 0: no-swizzle (8-channel boxes)
  int ch_unit;                 // channels per TMA box (64 or 8)
  int a_units, b_units;        // TMA boxes per stage
  int a_unit_bytes, b_unit_bytes, b_box_px;
  int a_bytes, b_bytes;        // bytes landed per stage (mbarrier transaction count)
  int stage_bytes, nstages;
  int tmem_cols;
  Job jobs[MAX_JOBS];
};

struct Misc {
  uint64_t full[MAX_STAGES], empty[MAX_STAGES], done;
  uint32_t tmem_slot, pad;
};

__global__ void __launch_bounds__(NTHREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
             float* __restrict__ partial, const __grid_constant__ Plan P) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  Misc& M = *reinterpret_cast<Misc*>(base + P.nstages * P.stage_bytes);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) rd::pdl_trigger();
  const int job_id = blockIdx.x / P.nsplit, split = blockIdx.x % P.nsplit;
  const Job J = P.jobs[job_id];
  const int t_begin = (int)((int64_t)P.ntiles * split / P.nsplit);
  const int t_end = (int)((int64_t)P.ntiles * (split + 1) / P.nsplit);

  if (t == 0) {
    for (int i = 0; i < MAX_STAGES; ++i) { tc::mbar_init(&M.full[i], 1); tc::mbar_init(&M.empty[i], 1); }
    tc::mbar_init(&M.done, 1);
    tc::fence_mbar_init();
    tma::prefetch_map(&tm_a);
    tma::prefetch_map(&tm_b);
  }
  if (warp == 1) {
    tc::tmem_alloc(&M.tmem_slot, (uint32_t)P.tmem_cols);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = M.tmem_slot;
  rd::pdl_wait();   // prologue done: only now are the predecessors' tensors (and the partial buffer) touched

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t s = 0, par = 1;
      for (int tl = t_begin; tl < t_end; ++tl) {
        const int w0 = (tl % P.tiles_w) * TM, h = (tl / P.tiles_w) % P.H, n = tl / (P.tiles_w * P.H);
        tc::mbar_wait(&M.empty[s], par);
        tc::mbar_arrive_expect_tx(&M.full[s], (uint32_t)(P.a_bytes + P.b_bytes));
        unsigned char* sa = base + s * P.stage_bytes;
        unsigned char* sb = sa + P.a_units * P.a_unit_bytes;
        for (int u = 0; u < P.a_units; ++u)
          tma::load_4d(sa + u * P.a_unit_bytes, &tm_a, &M.full[s], u * P.ch_unit, w0 + 1, h + 1, n);
        for (int u = 0; u < P.b_units; ++u)
          tma::load_4d(sb + u * P.b_unit_bytes, &tm_b, &M.full[s], J.cb0 + u * P.ch_unit, w0 * P.stride_w + J.dx0,
                       h + J.dy, n);
        if (++s == (uint32_t)P.nstages) { s = 0; par ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer: converged warp, one elected lane issues =====
    const uint32_t idesc = tc::make_idesc_f16kind(128, P.CBJ, RD_ACT_MMA_FMT, 1, 1);  // both operands MN-major
    // MN-major descriptors.  SW128: LBO = stride between 64-channel atoms, SBO = 1024 (8 pixel rows);
    // no swizzle: LBO = 128 (8 pixels x 16 B), SBO = stride between 8-channel chunks.
    const uint64_t a_hi = P.swz ? tc::make_smem_desc(0, (uint32_t)P.a_unit_bytes, 1024, tc::LAYOUT_SW128)
                                : tc::make_smem_desc(0, 128, (uint32_t)P.a_unit_bytes, tc::LAYOUT_NONE);
    const uint64_t b_hi = P.swz ? tc::make_smem_desc(0, (uint32_t)P.b_unit_bytes, 1024, tc::LAYOUT_SW128)
                                : tc::make_smem_desc(0, 128, (uint32_t)P.b_unit_bytes, tc::LAYOUT_NONE);
    const uint32_t kstep = P.swz ? (2048u >> 4) : (256u >> 4);  // 16 pixels, in 16-byte address units
    const uint32_t dxstep = P.swz ? (128u >> 4) : 1u;           // one pixel
    const bool leader = tc::elect_one();
    uint32_t s = 0, par = 0;
    for (int tl = t_begin; tl < t_end; ++tl) {
      tc::mbar_wait(&M.full[s], par);
      tc::tc_fence_after();
      if (leader) {
        const uint32_t a_lo = tc::smem_u32(base + s * P.stage_bytes) >> 4;
        const uint32_t b_lo = a_lo + (uint32_t)((P.a_units * P.a_unit_bytes) >> 4);
        for (int d = 0; d < J.ndx; ++d) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(d * P.CBJ);
#pragma unroll
          for (uint32_t ks = 0; ks < 8; ++ks) {
            const uint64_t ad = a_hi | (uint64_t)((a_lo + ks * kstep) & 0x3FFF);
            const uint64_t bd = b_hi | (uint64_t)((b_lo + (uint32_t)d * dxstep + ks * kstep) & 0x3FFF);
            tc::mma_bf16_ss(d_tmem, ad, bd, idesc, (tl == t_begin && ks == 0) ? 0u : 1u);
          }
        }
        tc::umma_commit(&M.empty[s]);
      }
      __syncwarp();
      if (++s == (uint32_t)P.nstages) { s = 0; par ^= 1; }
    }
    if (leader) tc::umma_commit(&M.done);
    __syncwarp();
  } else {
    // ===== final epilogue: TMEM lane = A channel, columns = (dx, B channel) =====
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    tc::mbar_wait(&M.done, 0);
    __syncwarp();
    tc::tc_fence_after();
    if (q4 * 32 < P.CA) {  // whole warps only: tcgen05.ld is warp-collective
      const uint32_t t_acc = tmem_base + ((uint32_t)(q4 * 32) << 16);
      for (int d = 0; d < J.ndx; ++d) {
        float* dst = partial + (((int64_t)split * P.ntaps + (J.tap0 + d)) * P.CA + row) * P.CB + J.cb0;
        for (int c0 = 0; c0 < P.CBJ; c0 += 32) {
          float v[32];
          if (t_end > t_begin) {
            tc::tmem_ld_x32(t_acc + (uint32_t)(d * P.CBJ + c0), v);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
            reinterpret_cast<float4*>(dst + c0)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// out[i] = sum over splits of partial[s][i], fixed order.  A warp covers 4 consecutive float4 outputs x 8 split groups
// (lane = group * 4 + element): every group streams its splits (g, g + 8, ...) with 64-byte contiguous reads, so a thread
// has nsplit / 8 independent loads in flight instead of walking all ~49 splits on its own (the first version spent
// 15 us per layer, latency-bound, 1.4 ms per training step); the groups are then combined by a fixed shuffle tree:
// deterministic.
__global__ void __launch_bounds__(256) reduce_kernel(const float4* __restrict__ partial, float4* __restrict__ out, int n4, int nsplit) {
  if (threadIdx.x == 0) rd::pdl_trigger();
  rd::pdl_wait();
  const int lane = threadIdx.x & 31, el = lane & 3, g = lane >> 2;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  for (int base = (((int)blockIdx.x * (int)blockDim.x + (int)threadIdx.x) >> 5) * 4; base < n4; base += warps_total * 4) {
    const int i = base + el;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) {
#pragma unroll 4
      for (int s = g; s < nsplit; s += 8) {
        const float4 b = partial[(int64_t)s * n4 + i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
    }
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
      a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
      a.z += __shfl_xor_sync(0xffffffffu, a.z, o);
      a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
    }
    if (g == 0 && i < n4) out[i] = a;
  }
}

static int make_plan(Plan& P, int N, int H, int W, int CA, int CB, int ksize, int stride_w) {
  memset(&P, 0, sizeof(P));
  RD_REQUIRE(N > 0 && H > 0 && W > 0, "rd_conv2d_wgrad: bad shape");
  RD_REQUIRE(CA == 64 || CA == 128, "rd_conv2d_wgrad: CA must be 64 or 128 (got %d)", CA);
  RD_REQUIRE(CB >= 64 && CB % 64 == 0 && CB <= 1024, "rd_conv2d_wgrad: CB must be a multiple of 64, <= 1024 (got %d)", CB);
  RD_REQUIRE(ksize == 3 || ksize == 1, "rd_conv2d_wgrad: kernel size must be 3 or 1 (got %d)", ksize);
  RD_REQUIRE(stride_w == 1 || stride_w == 2, "rd_conv2d_wgrad: W stride must be 1 or 2 (got %d)", stride_w);
  P.N = N; P.H = H; P.W = W; P.CA = CA; P.CB = CB; P.ksize = ksize; P.stride_w = stride_w;
  P.ntaps = ksize * ksize;
  const char* e = getenv("RD_WGRAD_NOSWZ");
  P.swz = (e && e[0] == '1') ? 0 : 1;
  P.ch_unit = P.swz ? 64 : 8;
  const bool strip = (ksize == 3 && stride_w == 1);
  const int ndx = strip ? 3 : 1;
  const int cap = 512 / ndx;  // TMEM columns per tap
  P.CBJ = (CB % 256 == 0 && cap >= 256) ? 256 : (CB % 128 == 0 && cap >= 128) ? 128 : 64;
  int nj = 0;
  for (int dy = 0; dy < ksize; ++dy)
    for (int dx = 0; dx < ksize; dx += ndx)
      for (int cb0 = 0; cb0 < CB; cb0 += P.CBJ) {
        RD_REQUIRE(nj < MAX_JOBS, "rd_conv2d_wgrad: too many jobs");
        Job& J = P.jobs[nj++];
        J.dy = ksize == 3 ? dy : 1;
        J.dx0 = ksize == 3 ? dx : 1;
        J.ndx = ndx;
        J.cb0 = cb0;
        J.tap0 = dy * ksize + dx;
      }
  P.njobs = nj;
  P.tiles_w = (W + TM - 1) / TM;
  const int64_t ntiles = (int64_t)N * H * P.tiles_w;
  RD_REQUIRE(ntiles <= 0x7fffffffLL, "rd_conv2d_wgrad: too many tiles");
  P.ntiles = (int)ntiles;
  int nsplit = 148 / nj;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > P.ntiles) nsplit = P.ntiles;
  P.nsplit = nsplit;
  P.a_units = CA / P.ch_unit;
  P.b_units = P.CBJ / P.ch_unit;
  const int b_px = strip ? TM + 2 : TM;       // pixels landed per B box
  P.b_box_px = strip ? TM + 2 : TM * stride_w;  // box extent in tensor elements (element stride = stride_w)
  if (P.swz) {
    P.a_unit_bytes = TM * 128;
    P.b_unit_bytes = ((b_px * 128 + 1023) / 1024) * 1024;
  } else {
    P.a_unit_bytes = TM * 16;
    P.b_unit_bytes = ((b_px * 16 + 127) / 128) * 128;
  }
  P.a_bytes = P.a_units * TM * P.ch_unit * 2;
  P.b_bytes = P.b_units * b_px * P.ch_unit * 2;
  P.stage_bytes = ((P.a_units * P.a_unit_bytes + P.b_units * P.b_unit_bytes + 1023) / 1024) * 1024;
  {
    static int cap_kb = -1;   // RD_WGRAD_SMEM_KB: cap of the operand ring (tuning: co-residency with the BatchNorm passes)
    if (cap_kb < 0) {
      const char* e = getenv("RD_WGRAD_SMEM_KB");
      cap_kb = e ? atoi(e) : 0;
      if (cap_kb < 32 || cap_kb > 220) cap_kb = 220;
    }
    P.nstages = (cap_kb * 1024) / P.stage_bytes;
    if (P.nstages < 2) P.nstages = (220 * 1024) / P.stage_bytes;   // a cap below two stages is ignored
  }
  if (P.nstages > MAX_STAGES) P.nstages = MAX_STAGES;
  RD_REQUIRE(P.nstages >= 2, "rd_conv2d_wgrad: stage of %d bytes does not fit twice in shared memory", P.stage_bytes);
  int cols = 32;
  while (cols < ndx * P.CBJ) cols <<= 1;
  P.tmem_cols = cols;
  return 0;
}

}  // namespace wg_<storage type>
namespace wg = RD_ACT_NS(wg);

extern "C" {

#ifndef RD_ACT_F16   // storage-type independent: defined once, by the bf16 pass
size_t rd_conv2d_wgrad_workspace_bytes(int N, int H, int W, int CA, int CB, int ksize, int stride_w) {
  wg::Plan P;
  if (wg::make_plan(P, N, H, W, CA, CB, ksize, stride_w)) return 0;
  return (size_t)P.nsplit * P.ntaps * CA * CB * sizeof(float);
}
#endif

int RD_ACT_FN(rd_conv2d_wgrad_nhwc_, )(const void* a_pad, const void* b_pad, float* g, int N, int H, int W, int CA, int CB,
                              int ksize, int stride_w, void* workspace, size_t workspace_bytes, rd_stream_t stream) {
  wg::Plan P;
  if (wg::make_plan(P, N, H, W, CA, CB, ksize, stride_w)) return 1;
  RD_REQUIRE(a_pad && b_pad && g && workspace, "rd_conv2d_wgrad: null pointer");
  RD_REQUIRE(workspace_bytes >= (size_t)P.nsplit * P.ntaps * CA * CB * sizeof(float), "rd_conv2d_wgrad: workspace too small");
  if (rd_check_device()) return 1;
  cudaStream_t s = rd::as_stream(stream);
  const uint64_t Hp = (uint64_t)H + 2, Wa = (uint64_t)W + 2, Wb = (uint64_t)W * stride_w + 2;
  const CUtensorMapSwizzle sw = P.swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUtensorMap tm_a, tm_b;
  {
    const uint64_t d[4] = {(uint64_t)CA, Wa, Hp, (uint64_t)N};
    const uint64_t st[3] = {(uint64_t)CA * 2, Wa * CA * 2, Hp * Wa * CA * 2};
    const uint32_t b[4] = {(uint32_t)P.ch_unit, (uint32_t)wg::TM, 1u, 1u};
    if (tma::make_map(&tm_a, RD_ACT_TMA_TYPE, a_pad, 4, d, st, b, sw)) return 1;
  }
  {
    const uint64_t d[4] = {(uint64_t)CB, Wb, Hp, (uint64_t)N};
    const uint64_t st[3] = {(uint64_t)CB * 2, Wb * CB * 2, Hp * Wb * CB * 2};
    const uint32_t b[4] = {(uint32_t)P.ch_unit, (uint32_t)P.b_box_px, 1u, 1u};
    const uint32_t es[4] = {1u, (uint32_t)stride_w, 1u, 1u};
    if (tma::make_map_es(&tm_b, RD_ACT_TMA_TYPE, b_pad, 4, d, st, b, es, sw)) return 1;
  }
  const size_t smem = (size_t)P.nstages * P.stage_bytes + sizeof(wg::Misc) + 1024;
  RD_REQUIRE(smem <= 227 * 1024, "rd_conv2d_wgrad: shared memory layout exceeds 227 KB (%zu)", smem);
  RD_CUDA(rd::smem_optin(wg::wgrad_kernel, smem));
  float* partial = static_cast<float*>(workspace);
  RD_CUDA(rd::launch(wg::wgrad_kernel, dim3(P.njobs * P.nsplit), dim3(wg::NTHREADS), smem, s, tm_a, tm_b, partial, P));
  const int n4 = P.ntaps * CA * CB / 4;
  {
    const int64_t warps = ((int64_t)n4 + 3) / 4;                 // one warp per 4 outputs
    int64_t blocks = (warps + 7) / 8;                            // 8 warps per block
    if (blocks > 148 * 8) blocks = 148 * 8;
    RD_CUDA(rd::launch(wg::reduce_kernel, dim3((unsigned)blocks), dim3(256), 0, s, reinterpret_cast<const float4*>(partial),
                       reinterpret_cast<float4*>(g), n4, (int)P.nsplit));
  }
  rd::count_launch(2);
  return rd::check_launch("rd_conv2d_wgrad");
}

}  // extern "C"
