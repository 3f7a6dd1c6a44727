// Implicit-GEMM 2-D convolution family on tcgen05 tensor cores, NHWC bf16 (sm_100a).
//
// Replaces the dense contractions of the DLA backbone / RPN head that the reference delegates to
// MXNet/cuDNN: mx.sym.Convolution (/root/reference mxnext/simple.py:123-158) and
// mx.sym.Deconvolution (mxnext/simple.py:545-580); call sites rangedet/symbol/backbone/
// dla_backbone.py:23-56 (3x3, stride (1,1) and (1,2), 1x1 projections), :95 (1x1 576->64
// aggregation conv after the Meta-Kernel), :120-127 (deconv (3,8)/(1,4)/(1,2) and (3,4)/(1,2)/(1,1))
// and rangedet/symbol/head/builder.py:198-266.  The following inference-form BatchNorm (per-channel
// scale/shift), ReLU and residual add are folded into the epilogue.
//
// GEMM view per tile: M = 128 consecutive output-resolution columns of one image row, N = Cout,
// K = taps x Cin.  Activations are zero-haloed NHWC bf16 [N][H+2][W+2][C], so every tap is a plain
// TMA box with non-negative coordinates (this part's TMA faults on negative ones): 64 channels
// (128 B = one 128B-swizzle atom) x 128 pixels = the K-major A operand of four K=16 MMAs; a W-stride
// of 2 is the tensor map's element stride.  A transposed convolution with W-stride S is S phase
// convolutions that share input tiles: each phase gets its own TMEM accumulator.
// The per-tile work is a small "tap program" (Params::loads): for each input tile offset, the list
// of (accumulator, weight tap) pairs that consume it.  Weights are pre-packed [tap][Cout][Cin]; when
// they all fit in shared memory they stay resident for the whole persistent CTA, otherwise each
// use streams its 64-channel weight tile through the same ring as the activations.
//   warp 0  TMA producer      warp 1  MMA issuer (tcgen05.mma M128 x Cout x K16, fp32 in TMEM)
//   warps 2-5 epilogue: tcgen05.ld -> x scale + shift (+ residual) -> ReLU -> bf16 ->
//       conv: 128B-swizzled staging tile -> TMA store into the interior view of the haloed output
//       deconv: each thread owns S consecutive output pixels -> direct 16-byte stores
#include <stdlib.h>
#include <string.h>

#include "../../include/rangedet_b200.h"
#include "act_type.cuh"
#include "rd_common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

// conv_t.cu: 3x3 / stride 1 / Cout = 128 in the transposed GEMM orientation (M = Cout, N = 256 pixels)
int RD_ACT_FN(rd_convt_run_, )(const void* x_pad, const void* w_packed, const float* scale, const float* shift,
                               const void* residual_pad, void* y_pad, int N, int H, int W, int Cin, int Cout, int relu,
                               cudaStream_t stream, float* stats, int* stats_slots, const void* bn_z = nullptr,
                               const float* bn_coef = nullptr, int bn_mask = 0);

namespace RD_ACT_NS(conv) {

constexpr int TM = 128;               // GEMM rows per tile
constexpr int KC = 64;                // channels per TMA box / swizzle atom
constexpr int SLOT = TM * KC * 2;     // 16 KB ring slot (A tile, or a 128-row weight tile)
constexpr int NTHREADS = 384;
constexpr int BAR_EPI = 1;
constexpr int MAX_LOADS = 9, MAX_USES = 4;
constexpr uint32_t USE_NEWLOAD = 1, USE_LASTOFLOAD = 2, USE_FIRST = 4;  // flags of a per-use record
constexpr int STATS_STRIDE = 1184;    // = bn::MAX_BLOCKS: slots per channel row of the batch-statistics partials

struct Load {
  int8_t dy, dx;            // offsets in the haloed input frame (already >= 0)
  uint8_t nuse;
  uint8_t acc[MAX_USES];    // accumulator (phase) fed by this input tile
  uint8_t tap[MAX_USES];    // packed weight tap
  uint8_t off[MAX_USES];    // strip mode: first pixel row of the loaded strip this use starts at (dx)
};

struct SLoad {  // shared-memory form of Load: the per-use bytes packed into words
  uint32_t nuse, acc, tap, off;
  int dy, dx;
};

constexpr int MAX_PIPES = 2, MAX_SA = 4, MAX_SB = 8;

struct Params {
  int N, H, W_out_tiles;    // W_out_tiles: GEMM-row extent along W (output cols for conv, input cols for deconv)
  int Cin, Cout;
  int tiles_w, ntiles;
  int in_stride_w;          // input pixels advanced per GEMM row (1, or 2 for W-strided conv)
  int nacc, nloads, ntaps, nuses_total;
  int relu, has_residual, res_after_relu;
  int deconv_s;             // 0: convolution (TMA-store epilogue); S>0: transposed conv, phase count S
  int b_resident, acc_bufs;
  int npipes, nsa, nsb;     // pipelines, activation ring stages per pipeline, weight ring stages (shared)
  int a_off[MAX_PIPES], o_off[MAX_PIPES], b_off, w_off, misc_off;
  int slot_bytes, a_bytes;  // activation ring slot size, bytes of one activation load
  int64_t y_row, y_img;     // element strides of the haloed output (deconv / residual addressing)
  long long* prof;          // diagnostic builds only (RD_CONV_PROF=1), else null
  Load loads[MAX_LOADS];
};

struct Misc {  // barriers and small tables, at Params::misc_off of the dynamic shared memory
  uint64_t a_full[MAX_PIPES][MAX_SA], a_empty[MAX_PIPES][MAX_SA];
  uint64_t b_full[MAX_SB], b_empty[MAX_SB];
  uint64_t t_full[MAX_PIPES][2], t_empty[MAX_PIPES][2];
  uint64_t w_full;
  uint32_t tmem_slot, pad;
  float scale[128], shift[128];
  uint4 uses[MAX_LOADS * MAX_USES];  // flat per-use records for the MMA warps (one K-half)
  SLoad loads[MAX_LOADS];            // tap program for the producer
};
constexpr int MISC_BYTES = 3072;
static_assert(sizeof(Misc) <= MISC_BYTES, "Misc grew past its reservation");

// Two complete pipelines per CTA -- {MMA warp, four epilogue warps, TMEM accumulators, activation ring,
// output staging tile} each -- fed by one TMA producer warp and sharing the weights (resident, or
// streamed through ONE ring whose slots are released when both pipelines have used them: every weight
// tile fetched from L2 serves two GEMM tiles).  Why two: one issuing thread sustains at best one MMA per
// ~100-130 cycles in this loop (waits, commits, descriptor moves), the tensor core needs 48 (N=64) /
// 64 (N=128); two issuers interleave on the same tensor core.
//   warp 0 producer | warp 1 MMA pipe 0 | warps 2-5 epilogue pipe 0 | warp 6 MMA pipe 1 | warp 7 idle |
//   warps 8-11 epilogue pipe 1.   A super-tile is npipes consecutive tiles; pipe p takes tile st*npipes+p.
// PROF: diagnostic build (RD_CONV_PROF=1) that accumulates clock64() cycles per role into P.prof
// [block][16]: 0 producer total, 1 producer wait(a_empty), 2 producer wait(b_empty); 4 MMA(pipe 0) total,
// 5 wait(a_full), 6 wait(t_empty), 7 wait(b_full); 8 epilogue(pipe 0) total, 9 wait(t_full),
// 10 wait(store read)+barrier, 11 TMEM->smem body.
template <bool PROF>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
            const __grid_constant__ CUtensorMap tm_y, const float* __restrict__ scale,
            const float* __restrict__ shift, const act_t* __restrict__ residual,
            act_t* __restrict__ y_interior, float* __restrict__ stats, const __grid_constant__ Params P) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) rd::pdl_trigger();   // the next kernel may take SMs as this grid's CTAs leave them
  auto tick = [&]() -> long long { return PROF ? clock64() : 0ll; };
  long long pc1 = 0, pc2 = 0, pc3 = 0;
  const long long t_begin = tick();
  const int kh = P.Cin / KC, nh = P.Cout / KC;
  const int b_tile = P.Cout * KC * 2;
  const int npipes = P.npipes;
  const int nsuper = (P.ntiles + npipes - 1) / npipes;
  unsigned char* sW = base + P.w_off;
  unsigned char* ringB = base + P.b_off;
  Misc& M = *reinterpret_cast<Misc*>(base + P.misc_off);

  if (t < P.nloads) {
    SLoad L;
    L.nuse = P.loads[t].nuse;
    L.acc = L.tap = L.off = 0;
    for (int u = 0; u < MAX_USES; ++u) {
      L.acc |= (uint32_t)P.loads[t].acc[u] << (8 * u);
      L.tap |= (uint32_t)P.loads[t].tap[u] << (8 * u);
      L.off |= (uint32_t)P.loads[t].off[u] << (8 * u);
    }
    L.dy = P.loads[t].dy;
    L.dx = P.loads[t].dx;
    M.loads[t] = L;
  }
  if (t == 32) {
    int i = 0;
    uint32_t started = 0;
    for (int l = 0; l < P.nloads; ++l)
      for (int u = 0; u < P.loads[l].nuse; ++u, ++i) {
        const uint32_t acc = P.loads[l].acc[u], tap = P.loads[l].tap[u], off = P.loads[l].off[u];
        uint4 r;
        r.x = off * 8;  // strip mode: the view starts `off` pixel rows (128 B = 8 address units) into the tile
        r.y = (tc::smem_u32(sW) >> 4) + tap * (uint32_t)kh * (uint32_t)(b_tile >> 4);  // resident weights
        r.z = acc * (uint32_t)P.Cout;
        r.w = (u == 0 ? USE_NEWLOAD : 0u) | (u == P.loads[l].nuse - 1 ? USE_LASTOFLOAD : 0u) |
              ((started >> acc) & 1u ? 0u : USE_FIRST);
        started |= 1u << acc;
        M.uses[i] = r;
      }
  }
  if (t == 0) {
    for (int p = 0; p < MAX_PIPES; ++p) {
      for (int i = 0; i < MAX_SA; ++i) { tc::mbar_init(&M.a_full[p][i], 1); tc::mbar_init(&M.a_empty[p][i], 1); }
      for (int i = 0; i < 2; ++i) { tc::mbar_init(&M.t_full[p][i], 1); tc::mbar_init(&M.t_empty[p][i], 4); }
    }
    for (int i = 0; i < MAX_SB; ++i) { tc::mbar_init(&M.b_full[i], 1); tc::mbar_init(&M.b_empty[i], (uint32_t)npipes); }
    tc::mbar_init(&M.w_full, 1);
    tc::fence_mbar_init();
    tma::prefetch_map(&tm_x);
    tma::prefetch_map(&tm_w);
    if (!P.deconv_s) tma::prefetch_map(&tm_y);
  }
  if (warp == 1) {
    tc::tmem_alloc(&M.tmem_slot, 512);
    tc::tmem_relinquish();
  }
  rd::pdl_wait();   // everything above touched only shared memory / TMEM / kernel parameters
  for (int c = t; c < P.Cout; c += NTHREADS) {
    M.scale[c] = scale ? __ldg(scale + c) : 1.f;
    M.shift[c] = shift ? __ldg(shift + c) : 0.f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = M.tmem_slot;
  const uint32_t acc_stride = (uint32_t)P.Cout;             // columns per accumulator
  const uint32_t buf_stride = (uint32_t)(P.nacc * P.Cout);  // columns per accumulator set

  if (warp == 0) {
    // ===== TMA producer (both pipelines + the shared weight ring) =====
    if (lane == 0) {
      if (P.b_resident) {
        tc::mbar_arrive_expect_tx(&M.w_full, (uint32_t)(P.ntaps * kh * b_tile));
        for (int tap = 0; tap < P.ntaps; ++tap)
          for (int q = 0; q < kh; ++q) tma::load_3d(sW + (tap * kh + q) * b_tile, &tm_w, &M.w_full, q * KC, 0, tap);
      }
      uint32_t sa0 = 0, pa0 = 1, sa1 = 0, pa1 = 1, sb = 0, pb = 1;  // ring positions, parity of the `empty` wait
      const uint32_t nsa = (uint32_t)P.nsa, nsb = (uint32_t)P.nsb;
      const int nloads = P.nloads, resident = P.b_resident;
      unsigned char* ringA0 = base + P.a_off[0];
      unsigned char* ringA1 = base + P.a_off[1];
      for (int st = blockIdx.x; st < nsuper; st += gridDim.x) {
        int tl0 = st * npipes, tl1 = tl0 + 1;
        if (tl1 >= P.ntiles) tl1 = P.ntiles - 1;  // odd tail: pipe 1 recomputes the last tile, its store is skipped
        const int wx0 = (tl0 % P.tiles_w) * TM * P.in_stride_w, h0 = (tl0 / P.tiles_w) % P.H, n0 = tl0 / (P.tiles_w * P.H);
        const int wx1 = (tl1 % P.tiles_w) * TM * P.in_stride_w, h1 = (tl1 / P.tiles_w) % P.H, n1 = tl1 / (P.tiles_w * P.H);
        for (int q = 0; q < kh; ++q)
          for (int l = 0; l < nloads; ++l) {
            const SLoad L = M.loads[l];
            {
              const long long ta = tick();
              tc::mbar_wait(&M.a_empty[0][sa0], pa0);
              pc1 += tick() - ta;
              tc::mbar_arrive_expect_tx(&M.a_full[0][sa0], (uint32_t)P.a_bytes);
              tma::load_4d(ringA0 + sa0 * P.slot_bytes, &tm_x, &M.a_full[0][sa0], q * KC, wx0 + L.dx, h0 + L.dy, n0);
              if (++sa0 == nsa) { sa0 = 0; pa0 ^= 1; }
            }
            if (npipes == 2) {
              const long long ta = tick();
              tc::mbar_wait(&M.a_empty[1][sa1], pa1);
              pc1 += tick() - ta;
              tc::mbar_arrive_expect_tx(&M.a_full[1][sa1], (uint32_t)P.a_bytes);
              tma::load_4d(ringA1 + sa1 * P.slot_bytes, &tm_x, &M.a_full[1][sa1], q * KC, wx1 + L.dx, h1 + L.dy, n1);
              if (++sa1 == nsa) { sa1 = 0; pa1 ^= 1; }
            }
            if (!resident)
              for (int u = 0; u < (int)L.nuse; ++u) {
                const long long tb = tick();
                tc::mbar_wait(&M.b_empty[sb], pb);
                pc2 += tick() - tb;
                tc::mbar_arrive_expect_tx(&M.b_full[sb], (uint32_t)b_tile);
                tma::load_3d(ringB + sb * b_tile, &tm_w, &M.b_full[sb], q * KC, 0, (int)((L.tap >> (8 * u)) & 0xff));
                if (++sb == nsb) { sb = 0; pb ^= 1; }
              }
          }
      }
      if (PROF) {
        P.prof[blockIdx.x * 16 + 0] = tick() - t_begin;
        P.prof[blockIdx.x * 16 + 1] = pc1;
        P.prof[blockIdx.x * 16 + 2] = pc2;
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 6) {
    // ===== MMA issuer of pipeline p =====
    // The whole warp runs this loop CONVERGED and one elected lane issues.  Measured (scripts/mma_bench*.cu):
    // tcgen05.mma itself sustains its floor (48 cycles at N=64 -- shared-memory operand reads --, 64 at
    // N=128), but a loop that rebuilds descriptors, unpacks the tap program and polls per MMA spends
    // 130-240 cycles of the issuing thread per MMA.  So: one 16-byte record per use, read as a broadcast,
    // descriptors that differ only in their start-address field, four back-to-back MMAs per use.
    const int p = warp == 1 ? 0 : 1;
    if (p < npipes) {
      const uint32_t idesc = tc::make_idesc_f16kind(TM, P.Cout, RD_ACT_MMA_FMT);
      const uint64_t desc_hi = tc::make_smem_desc(0, 0, 1024, tc::LAYOUT_SW128);  // everything but the address
      const uint32_t ringA_lo = tc::smem_u32(base + P.a_off[p]) >> 4, ringB_lo = tc::smem_u32(ringB) >> 4;
      const uint32_t slot_lo = (uint32_t)P.slot_bytes >> 4, btile_lo = (uint32_t)b_tile >> 4;
      const uint32_t nsa = (uint32_t)P.nsa, nsb = (uint32_t)P.nsb;
      const int nuses = P.nuses_total, resident = P.b_resident;
      const bool leader = tc::elect_one();
      if (resident) tc::mbar_wait(&M.w_full, 0);
      uint32_t sa = 0, pha = 0, sb = 0, phb = 0;  // ring positions of the next units, phase parities
      uint32_t it = 0;
      for (int st = blockIdx.x; st < nsuper; st += gridDim.x, ++it) {
        const uint32_t buf = P.acc_bufs == 2 ? (it & 1) : 0;
        const uint32_t use_n = P.acc_bufs == 2 ? (it >> 1) : it;  // how many times this buffer was used before
        const long long te = tick();
        tc::mbar_wait(&M.t_empty[p][buf], (use_n & 1) ^ 1);
        pc2 += tick() - te;
        const uint32_t d_base = tmem_base + ((uint32_t)p * P.acc_bufs + buf) * buf_stride;
        for (int q = 0; q < kh; ++q) {
          uint32_t a_lo = 0, cur_sa = 0;
          const uint32_t bq = (uint32_t)q * btile_lo;
#pragma unroll 1
          for (int i = 0; i < nuses; ++i) {
            const uint4 r = M.uses[i];  // x: A offset (16-B units), y: resident B address, z: TMEM column, w: flags
            if (r.w & USE_NEWLOAD) {
              cur_sa = sa;
              const long long tf = tick();
              tc::mbar_wait(&M.a_full[p][sa], pha);
              pc1 += tick() - tf;
              a_lo = ringA_lo + sa * slot_lo;
              if (++sa == nsa) { sa = 0; pha ^= 1; }
            }
            uint32_t b_lo = r.y + bq, cur_sb = 0;
            if (!resident) {
              cur_sb = sb;
              const long long tf = tick();
              tc::mbar_wait(&M.b_full[sb], phb);
              pc3 += tick() - tf;
              b_lo = ringB_lo + sb * btile_lo;
              if (++sb == nsb) { sb = 0; phb ^= 1; }
            }
            tc::tc_fence_after();
            if (leader) {
              // 128B-swizzled K-major tiles: K advances 32 B (2 address units) inside the swizzle atom
              const uint64_t ad = desc_hi | (uint64_t)((a_lo + r.x) & 0x3FFF);
              const uint64_t bd = desc_hi | (uint64_t)(b_lo & 0x3FFF);
              const uint32_t d_tmem = d_base + r.z;
              tc::mma_bf16_ss(d_tmem, ad, bd, idesc, ((r.w & USE_FIRST) && q == 0) ? 0u : 1u);
              tc::mma_bf16_ss_acc(d_tmem, ad + 2, bd + 2, idesc);
              tc::mma_bf16_ss_acc(d_tmem, ad + 4, bd + 4, idesc);
              tc::mma_bf16_ss_acc(d_tmem, ad + 6, bd + 6, idesc);
              if (!resident) tc::umma_commit(&M.b_empty[cur_sb]);
              if (r.w & USE_LASTOFLOAD) tc::umma_commit(&M.a_empty[p][cur_sa]);
            }
            __syncwarp();
          }
        }
        if (leader) tc::umma_commit(&M.t_full[p][buf]);
        __syncwarp();
      }
      if (PROF && lane == 0 && p == 0) {
        P.prof[blockIdx.x * 16 + 4] = tick() - t_begin;
        P.prof[blockIdx.x * 16 + 5] = pc1;
        P.prof[blockIdx.x * 16 + 6] = pc2;
        P.prof[blockIdx.x * 16 + 7] = pc3;
      }
    }
  } else if (warp != 7) {
    // ===== epilogue of pipeline p: thread = GEMM row = TMEM lane =====
    const int p = warp >= 8 ? 1 : 0;
    if (p < npipes) {
      const int q4 = warp & 3;  // TMEM lane quadrant this warp may read
      const int px = q4 * 32 + lane;
      const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
      const bool leader = ((warp == 2 || warp == 8) && lane == 0);
      const int bar_id = BAR_EPI + p;
      unsigned char* sO = base + P.o_off[p];
      // Fused batch statistics (training-mode BatchNorm, mxnext/complicate.py:32-43): after the tile is staged, the 128
      // epilogue threads re-map to (channel pair, row group) and add the column sums of the STORED values (sum z,
      // sum z^2) into four registers that live across all tiles of this CTA; one partial per (CTA, pipe, row group)
      // at the end.  Saves the separate full read of z by rd_bn_train_stats.
      const int e128 = q4 * 32 + lane;
      const int ncp = P.Cout >> 1;                 // channel pairs: 64 or 32
      const int st_cp = e128 & (ncp - 1), st_rg = e128 / ncp, st_rows = ncp;   // rows per group = 128 / (128 / ncp)
      const uint32_t st_col = (uint32_t)(((2 * st_cp) / KC) * SLOT + (((2 * st_cp) % 8) * 2));
      const uint32_t st_chunk = (uint32_t)(((2 * st_cp) % KC) / 8);
      float st_s0 = 0.f, st_s1 = 0.f, st_q0 = 0.f, st_q1 = 0.f;
      uint32_t it = 0;
      for (int st = blockIdx.x; st < nsuper; st += gridDim.x, ++it) {
        const int tile = st * npipes + p;
        const bool real = tile < P.ntiles;  // odd tail: the duplicate tile of pipe 1 is computed but not stored
        const int tl = real ? tile : P.ntiles - 1;
        const int wt = tl % P.tiles_w, h = (tl / P.tiles_w) % P.H, n = tl / (P.tiles_w * P.H);
        const int w0 = wt * TM;
        const uint32_t buf = P.acc_bufs == 2 ? (it & 1) : 0;
        const uint32_t use_n = P.acc_bufs == 2 ? (it >> 1) : it;
        const long long e0 = tick();
        tc::mbar_wait(&M.t_full[p][buf], use_n & 1);
        __syncwarp();
        tc::tc_fence_after();
        const long long e1 = tick();
        const bool in_img = real && (w0 + px) < P.W_out_tiles;
        if (!P.deconv_s) {
          if (leader) tma::store_wait_read<0>();  // previous tile's stores have read the staging buffer
          tma::named_bar_sync(bar_id, 128);
        }
        const long long e2 = tick();
        pc1 += e1 - e0;
        pc2 += e2 - e1;
        const uint32_t t_acc = tmem_base + lane_sel + ((uint32_t)p * P.acc_bufs + buf) * buf_stride;
        const int nphase = P.deconv_s ? P.deconv_s : 1;
        for (int ph = 0; ph < nphase; ++ph) {
          // output pixel of this thread for this phase (interior coordinates)
          const int64_t opix = P.deconv_s ? (int64_t)(w0 + px) * P.deconv_s + ph : (int64_t)(w0 + px);
          const int64_t ooff = (int64_t)n * P.y_img + (int64_t)h * P.y_row + opix * P.Cout;
          for (int c0 = 0; c0 < P.Cout; c0 += 32) {
            float v[32];
            tc::tmem_ld_x32(t_acc + ph * acc_stride + c0, v);
            uint4 rv[4];
            if (P.has_residual && in_img) {
#pragma unroll
              for (int j = 0; j < 4; ++j) rv[j] = __ldg(reinterpret_cast<const uint4*>(residual + ooff + c0) + j);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // 8 channels -> one 16-byte chunk
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int c = c0 + j * 8 + 2 * e;
                float a = fmaf(v[j * 8 + 2 * e], M.scale[c], M.shift[c]);
                float b = fmaf(v[j * 8 + 2 * e + 1], M.scale[c + 1], M.shift[c + 1]);
                float ra = 0.f, rb = 0.f;
                if (P.has_residual && in_img) {
                  const uint32_t w = reinterpret_cast<const uint32_t*>(&rv[j])[e];
                  act::unpack2(w, ra, rb);
                }
                if (!P.res_after_relu) { a += ra; b += rb; }
                if (P.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
                if (P.res_after_relu) { a += ra; b += rb; }
                pk[e] = act::pack2(a, b);
              }
              if (P.deconv_s) {
                if (in_img) *reinterpret_cast<uint4*>(y_interior + ooff + c0 + j * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              } else {
                // staging tile + TMA store; direct 16-byte global stores were measured slower (epilogue body
                // 5267 -> 9010 cycles per 128->128 tile: every warp-wide store touches 32 lines)
                const int half = (c0 + j * 8) / KC, chunk = ((c0 + j * 8) % KC) / 8;
                // 128B swizzle: 16-byte chunk index XOR (row % 8)
                *reinterpret_cast<uint4*>(sO + half * SLOT + px * 128 + ((chunk ^ (px & 7)) << 4)) =
                    make_uint4(pk[0], pk[1], pk[2], pk[3]);
              }
            }
          }
        }
        tc::tc_fence_before();
        if (!P.deconv_s) tc::fence_proxy_async_smem();
        __syncwarp();
        pc3 += tick() - e2;
        if (lane == 0) tc::mbar_arrive(&M.t_empty[p][buf]);
        if (!P.deconv_s) {
          tma::named_bar_sync(bar_id, 128);
          if (leader && real) {
            for (int hf = 0; hf < nh; ++hf) tma::store_4d(&tm_y, sO + hf * SLOT, hf * KC, w0, h, n);
            tma::store_commit();
          }
          if (stats != nullptr && real) {
            const int nvalid = min(TM, P.W_out_tiles - w0);
            const int r1 = min((st_rg + 1) * st_rows, nvalid);
#pragma unroll 4
            for (int r = st_rg * st_rows; r < r1; ++r) {
              const uint32_t wv = *reinterpret_cast<const uint32_t*>(sO + st_col + r * 128 + ((st_chunk ^ (uint32_t)(r & 7)) << 4));
              float f0, f1;
              act::unpack2(wv, f0, f1);
              st_s0 += f0;
              st_s1 += f1;
              st_q0 = fmaf(f0, f0, st_q0);
              st_q1 = fmaf(f1, f1, st_q1);
            }
          }
        }
      }
      if (!P.deconv_s && leader) tma::store_wait_all<0>();
      if (stats != nullptr && !P.deconv_s) {
        // this pipeline's (row group x channel) partials -> its (now idle) staging tile; the CTA adds the pipelines and
        // row groups in fixed order after the final barrier and writes ONE slot per channel (the finalize kernel then
        // walks 148 slots per channel instead of 1184)
        tma::named_bar_sync(bar_id, 128);   // the leader's TMA stores have finished reading the staging tile
        float* red = reinterpret_cast<float*>(sO) + st_rg * 2 * P.Cout;
        const int c = 2 * st_cp;
        red[c] = st_s0;
        red[c + 1] = st_s1;
        red[P.Cout + c] = st_q0;
        red[P.Cout + c + 1] = st_q1;
      }
      if (PROF && leader && p == 0) {
        P.prof[blockIdx.x * 16 + 8] = tick() - t_begin;
        P.prof[blockIdx.x * 16 + 9] = pc1;
        P.prof[blockIdx.x * 16 + 10] = pc2;
        P.prof[blockIdx.x * 16 + 11] = pc3;
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
  if (stats != nullptr && !P.deconv_s) {
    // partial[(which * Cout + channel) * STATS_STRIDE + slot], slot = CTA: the layout bn::fwd_finalize_kernel reads
    const int nrg = TM / (P.Cout >> 1);
    for (int v = t; v < 2 * P.Cout; v += NTHREADS) {
      float a = 0.f;
      for (int p = 0; p < npipes; ++p)
        for (int rg = 0; rg < nrg; ++rg) a += reinterpret_cast<const float*>(base + P.o_off[p])[rg * 2 * P.Cout + v];
      stats[(int64_t)v * STATS_STRIDE + blockIdx.x] = a;
    }
  }
}

// mode: 0 = conv ksize x ksize (pad ksize/2), W-stride stride_w in {1,2}; 1 = deconv (3,8)/(1,4)/(1,2);
//       2 = deconv (3,4)/(1,2)/(1,1); 3 = deconv (3,3)/(1,2)/(1,1) with output width 2*W (the data gradient
//       of a 3x3 W-stride-2 convolution)
static int run(int mode, const void* x_pad, const void* w_packed, const float* scale, const float* shift,
               const void* residual_pad, void* y_pad, int N, int H, int W_in, int Cin, int Cout, int ksize,
               int stride_w, int relu, int res_after_relu, cudaStream_t stream, int y_ctotal = 0, int y_coff = 0,
               float* stats = nullptr, int* stats_slots = nullptr) {
  RD_REQUIRE(x_pad && w_packed && y_pad, "rd_conv: null pointer");
  // channel-slice output: write channels [y_coff, y_coff + Cout) of a y tensor with y_ctotal channels
  if (y_ctotal == 0) y_ctotal = Cout;
  RD_REQUIRE(y_ctotal >= Cout && y_coff >= 0 && y_coff + Cout <= y_ctotal && y_ctotal % 8 == 0 && y_coff % 8 == 0,
             "rd_conv: bad output channel slice (%d + %d of %d)", y_coff, Cout, y_ctotal);
  RD_REQUIRE(y_ctotal == Cout || (mode == 0 && !residual_pad), "rd_conv: a channel-slice output needs a plain conv without residual");
  RD_REQUIRE(Cin >= 64 && Cin % 64 == 0 && Cin <= 1024, "rd_conv: Cin must be a multiple of 64 (got %d); pad the channels", Cin);
  RD_REQUIRE(Cout == 64 || Cout == 128, "rd_conv: Cout must be 64 or 128 (got %d); pad the channels", Cout);
  RD_REQUIRE(N > 0 && H > 0 && W_in > 0, "rd_conv: bad shape");
  if (rd_check_device()) return 1;
  {
    // 3x3 stride-1 convolutions with 128 output channels (head towers, res2 / res3a / res3 / agg2 and their data
    // gradients) take the transposed orientation of conv_t.cu: 96 instead of 128 B/clk of shared-memory operand reads per
    // MMA, tiles over the flattened pixel grid.  RD_CONV_T=0 keeps them here (A/B timing, cross-check).
    // Cout == 64 also runs there (upper half of the weight tile zero-filled by TMA), but half of every M128 MMA is then
    // wasted and it measured SLOWER than the pixel-major kernel (64->64 @2x64x2656: 47.8 vs 35.6 us): opt-in only.
    static const int t64 = [] { const char* e = getenv("RD_CONV_T64"); return (e && e[0] == '1') ? 1 : 0; }();
    if (rd::conv_t_enabled() && mode == 0 && ksize == 3 && stride_w == 1 && (Cout == 128 || (Cout == 64 && t64)) &&
        y_ctotal == Cout && !res_after_relu)
      return RD_ACT_FN(rd_convt_run_, )(x_pad, w_packed, scale, shift, residual_pad, y_pad, N, H, W_in, Cin, Cout, relu, stream,
                                        stats, stats_slots);
  }
  Params P;
  memset(&P, 0, sizeof(P));
  P.N = N; P.H = H; P.Cin = Cin; P.Cout = Cout;
  P.relu = relu ? 1 : 0;
  P.has_residual = residual_pad ? 1 : 0;
  P.res_after_relu = res_after_relu ? 1 : 0;
  int W_out;  // output width in pixels
  bool strip = false;
  if (mode == 0) {
    RD_REQUIRE(ksize == 3 || ksize == 1, "rd_conv2d: kernel size must be 3 (pad 1) or 1 (got %d)", ksize);
    RD_REQUIRE(stride_w == 1 || stride_w == 2, "rd_conv2d: W stride must be 1 or 2 (got %d)", stride_w);
    RD_REQUIRE(stride_w == 1 || W_in % 2 == 0, "rd_conv2d: W must be even for W-stride 2");
    W_out = W_in / stride_w;
    P.W_out_tiles = W_out;
    P.in_stride_w = stride_w;
    P.nacc = 1;
    P.deconv_s = 0;
    P.ntaps = ksize * ksize;
    const char* es = getenv("RD_CONV_STRIP");
    // Strip mode (default): one 130-pixel row strip per dy serves the three dx taps as row-shifted views,
    // 3 activation loads per K-half instead of 9.  Measured 0.086 vs 0.095 ms (64->64 @ 4x64x2656) and
    // 0.202 vs 0.258 ms (128->128) once the MMA issue loop stopped being the bound.  RD_CONV_STRIP=0
    // selects nine aligned loads (diagnostic).
    strip = (ksize == 3 && stride_w == 1 && !(es && es[0] == '0'));
    if (strip) {
      // one 130-pixel row strip per dy serves the three dx taps as row-shifted views of the same tile
      P.nloads = 3;
      for (int dy = 0; dy < 3; ++dy) {
        Load& L = P.loads[dy];
        L.dy = (int8_t)dy; L.dx = 0; L.nuse = 3;
        for (int dx = 0; dx < 3; ++dx) { L.acc[dx] = 0; L.tap[dx] = (uint8_t)(dy * 3 + dx); L.off[dx] = (uint8_t)dx; }
      }
    } else {
      P.nloads = P.ntaps;
      for (int tap = 0; tap < P.ntaps; ++tap) {
        Load& L = P.loads[tap];
        L.dy = (int8_t)(ksize == 3 ? tap / 3 : 1);
        L.dx = (int8_t)(ksize == 3 ? tap % 3 : 1);
        L.nuse = 1; L.acc[0] = 0; L.tap[0] = (uint8_t)tap;
      }
    }
  } else {
    // out[oh, ow] += x[ih, iw] w[ky, kx]  with  oh = ih - 1 + ky,  ow = iw*S - pad + kx
    const int S = mode == 1 ? 4 : 2, KW = mode == 1 ? 8 : (mode == 2 ? 4 : 3), pad = mode == 1 ? 2 : 1;
    RD_REQUIRE(S * Cout <= 512, "rd_deconv: S*Cout accumulators exceed TMEM");
    W_out = W_in * S;
    P.W_out_tiles = W_in;
    P.in_stride_w = 1;
    P.nacc = S;
    P.deconv_s = S;
    P.ntaps = 3 * KW;
    // distinct input tiles: dih in {-1,0,1} x diw in {-1,0,1}
    int nl = 0;
    for (int dih = -1; dih <= 1; ++dih)
      for (int diw = -1; diw <= 1; ++diw) {
        Load L;
        memset(&L, 0, sizeof(L));
        L.dy = (int8_t)(dih + 1);
        L.dx = (int8_t)(diw + 1);
        const int ky = 1 - dih;  // ih = oh + 1 - ky
        for (int ph = 0; ph < S; ++ph)
          for (int kx = 0; kx < KW; ++kx) {
            // ow = j*S + ph = iw*S - pad + kx  with iw = j + diw  ->  kx = ph + pad - diw*S
            if (kx == ph + pad - diw * S) {
              RD_REQUIRE(L.nuse < MAX_USES, "rd_deconv: tap program overflow");
              L.acc[L.nuse] = (uint8_t)ph;
              L.tap[L.nuse] = (uint8_t)(ky * KW + kx);
              L.nuse++;
            }
          }
        if (L.nuse) P.loads[nl++] = L;
      }
    P.nloads = nl;
  }
  P.tiles_w = (P.W_out_tiles + TM - 1) / TM;
  P.nuses_total = 0;
  for (int l = 0; l < P.nloads; ++l) P.nuses_total += P.loads[l].nuse;
  const int64_t ntiles = (int64_t)N * H * P.tiles_w;
  RD_REQUIRE(ntiles <= 0x7fffffffLL, "rd_conv: too many tiles");
  P.ntiles = (int)ntiles;
  const int kh = Cin / KC, nh = Cout / KC;
  const int b_tile = Cout * KC * 2;
  const int w_bytes = P.ntaps * kh * b_tile;
  const int box_px = strip ? TM + 2 : TM * (mode == 0 ? stride_w : 1);
  P.a_bytes = (strip ? TM + 2 : TM) * KC * 2;
  P.slot_bytes = strip ? 17 * 1024 : SLOT;
  // Pipelines: two whenever both accumulator sets fit in the 512 TMEM columns (RD_CONV_PIPES=1 forces one).
  {
    const char* ep = getenv("RD_CONV_PIPES");
    P.npipes = (2 * P.nacc * Cout <= 512 && !(ep && ep[0] == '1') && P.ntiles >= 2) ? 2 : 1;
  }
  P.acc_bufs = (2 * P.npipes * P.nacc * Cout <= 512) ? 2 : 1;
  // Shared memory: [activation rings][weight ring | resident weights][output staging tiles][Misc]
  const int o_bytes = P.deconv_s ? 0 : nh * SLOT;  // per pipeline
  const int budget = 222 * 1024 - MISC_BYTES - P.npipes * o_bytes;
  P.b_resident = (w_bytes <= 96 * 1024 && w_bytes <= budget - P.npipes * 2 * P.slot_bytes) ? 1 : 0;
  if (P.b_resident) {
    P.nsb = 0;
    P.nsa = (budget - w_bytes) / (P.npipes * P.slot_bytes);
  } else {
    // weights stream: two activation stages per pipeline, the rest of the budget is the weight ring
    // (measured 0.1728 ms with 2 + 5 stages vs 0.1767 ms with 3 + 3 for 128->128 @ 4x64x2656)
    P.nsa = 2;
    P.nsb = (budget - P.npipes * P.nsa * P.slot_bytes) / b_tile;
  }
  if (!P.b_resident) {  // diagnostic override of the split between activation and weight stages
    const char* ea = getenv("RD_CONV_NSA");
    if (ea && ea[0] >= '2' && ea[0] <= '4') {
      P.nsa = ea[0] - '0';
      P.nsb = (budget - P.npipes * P.nsa * P.slot_bytes) / b_tile;
    }
  }
  if (P.nsa > MAX_SA) P.nsa = MAX_SA;
  if (P.nsb > MAX_SB) P.nsb = MAX_SB;
  RD_REQUIRE(P.nsa >= 2 && (P.b_resident || P.nsb >= 2), "rd_conv: shared memory budget too small (nsa %d, nsb %d)", P.nsa, P.nsb);
  {
    int off = 0;
    for (int p = 0; p < MAX_PIPES; ++p) { P.a_off[p] = off; if (p < P.npipes) off += P.nsa * P.slot_bytes; }
    P.b_off = P.w_off = off;
    off += P.b_resident ? w_bytes : P.nsb * b_tile;
    for (int p = 0; p < MAX_PIPES; ++p) { P.o_off[p] = off; if (p < P.npipes) off += o_bytes; }
    P.misc_off = off;
  }
  const size_t smem = (size_t)P.misc_off + MISC_BYTES + 1024;
  RD_REQUIRE(smem <= 227 * 1024, "rd_conv: shared memory layout exceeds 227 KB (%zu)", smem);
  const uint64_t Wp_in = (uint64_t)W_in + 2, Hp = (uint64_t)H + 2, Wp_out = (uint64_t)W_out + 2;
  P.y_row = (int64_t)Wp_out * Cout;
  P.y_img = (int64_t)Hp * Wp_out * Cout;

  CUtensorMap tm_x, tm_w, tm_y;
  {  // haloed input (C, W+2, H+2, N); a W-stride of 2 is the element stride of dimension 1
    const uint64_t d[4] = {(uint64_t)Cin, Wp_in, Hp, (uint64_t)N};
    const uint64_t s[3] = {(uint64_t)Cin * 2, Wp_in * Cin * 2, Hp * Wp_in * Cin * 2};
    const uint32_t b[4] = {(uint32_t)KC, (uint32_t)box_px, 1u, 1u};
    const uint32_t es[4] = {1u, (uint32_t)P.in_stride_w, 1u, 1u};
    if (tma::make_map_es(&tm_x, RD_ACT_TMA_TYPE, x_pad, 4, d, s, b, es, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {  // packed weights (Cin, Cout, taps)
    const uint64_t d[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)P.ntaps};
    const uint64_t s[2] = {(uint64_t)Cin * 2, (uint64_t)Cin * Cout * 2};
    const uint32_t b[3] = {(uint32_t)KC, (uint32_t)Cout, 1u};
    if (tma::make_map(&tm_w, RD_ACT_TMA_TYPE, w_packed, 3, d, s, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  char* y_int = static_cast<char*>(y_pad) + ((Wp_out + 1) * y_ctotal + y_coff) * 2;  // interior origin of the haloed output
  {  // interior view (C, W_out, H, N): stores are clipped at W_out, never touch the halo
    const uint64_t d[4] = {(uint64_t)Cout, (uint64_t)W_out, (uint64_t)H, (uint64_t)N};
    const uint64_t s[3] = {(uint64_t)y_ctotal * 2, Wp_out * y_ctotal * 2, Hp * Wp_out * y_ctotal * 2};
    const uint32_t b[4] = {(uint32_t)KC, (uint32_t)TM, 1u, 1u};
    if (tma::make_map(&tm_y, RD_ACT_TMA_TYPE, y_int, 4, d, s, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  const act_t* res = nullptr;
  if (residual_pad) res = static_cast<const act_t*>(residual_pad) + (Wp_out + 1) * Cout;
  int dev = 0, sms = 0;
  RD_CUDA(cudaGetDevice(&dev));
  RD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  RD_CUDA(rd::smem_optin(conv_kernel<false>, smem));
  RD_CUDA(rd::smem_optin(conv_kernel<true>, smem));
  const int nsuper = (P.ntiles + P.npipes - 1) / P.npipes;
  const int grid = nsuper < sms ? nsuper : sms;
  if (stats) {
    RD_REQUIRE(mode == 0 && y_ctotal == Cout, "rd_conv2d stats: only the plain convolution epilogue accumulates batch statistics");
    const int slots = grid;   // one partial per CTA and channel
    RD_REQUIRE(slots <= STATS_STRIDE, "rd_conv2d stats: %d partial slots exceed %d", slots, STATS_STRIDE);
    if (stats_slots) *stats_slots = slots;
  }
  static const bool prof = [] { const char* e = getenv("RD_CONV_PROF"); return e && e[0] == '1'; }();
  if (prof) {  // diagnostic: per-role cycle counters, synchronous, printed to stderr
    static long long* d_prof = nullptr;
    if (!d_prof) RD_CUDA(cudaMalloc(&d_prof, 1024 * 16 * sizeof(long long)));
    RD_CUDA(cudaMemsetAsync(d_prof, 0, 1024 * 16 * sizeof(long long), stream));
    P.prof = d_prof;
    RD_CUDA(rd::launch(conv_kernel<true>, dim3(grid), dim3(NTHREADS), smem, stream, tm_x, tm_w, tm_y, scale, shift, res,
                       reinterpret_cast<act_t*>(y_int), stats, P));
    RD_CUDA(cudaStreamSynchronize(stream));
    static long long h[1024 * 16];
    RD_CUDA(cudaMemcpy(h, d_prof, sizeof(long long) * 16 * grid, cudaMemcpyDeviceToHost));
    double m[16] = {0};
    for (int b = 0; b < grid; ++b)
      for (int k = 0; k < 16; ++k) m[k] += (double)h[b * 16 + k] / grid;
    const double tiles = (double)((P.ntiles + P.npipes - 1) / P.npipes) / grid;
    fprintf(stderr,
            "[rd_conv prof] Cin=%d Cout=%d pipes=%d nsa=%d nsb=%d resident=%d super-tiles/cta=%.1f | per super-tile: producer %.0f "
            "(wait a_empty %.0f, b_empty %.0f) | mma0 %.0f (wait a_full %.0f, b_full %.0f, t_empty %.0f) | epi0 %.0f (wait t_full "
            "%.0f, wait store+bar %.0f, body %.0f)\n",
            P.Cin, P.Cout, P.npipes, P.nsa, P.nsb, P.b_resident, tiles, m[0] / tiles, m[1] / tiles, m[2] / tiles, m[4] / tiles,
            m[5] / tiles, m[7] / tiles, m[6] / tiles, m[8] / tiles, m[9] / tiles, m[10] / tiles, m[11] / tiles);
  } else {
    RD_CUDA(rd::launch(conv_kernel<false>, dim3(grid), dim3(NTHREADS), smem, stream, tm_x, tm_w, tm_y, scale, shift, res,
                       reinterpret_cast<act_t*>(y_int), stats, P));
  }
  rd::count_launch();
  return rd::check_launch("rd_conv");
}

}  // namespace conv_<storage type>
namespace conv = RD_ACT_NS(conv);

extern "C" {

int RD_ACT_FN(rd_conv2d_nhwc_, )(const void* x_pad, const void* w_packed, const float* scale, const float* shift,
                        const void* residual_pad, void* y_pad, int N, int H, int W, int Cin, int Cout, int ksize,
                        int stride_w, int relu, rd_stream_t stream) {
  return conv::run(0, x_pad, w_packed, scale, shift, residual_pad, y_pad, N, H, W, Cin, Cout, ksize, stride_w, relu, 0,
                   rd::as_stream(stream));
}

int RD_ACT_FN(rd_conv2d_nhwc_, _stats)(const void* x_pad, const void* w_packed, void* y_pad, int N, int H, int W, int Cin, int Cout,
                                       int ksize, int stride_w, float* stats_partial, size_t stats_bytes, int* stats_slots,
                                       rd_stream_t stream) {
  RD_REQUIRE(stats_partial && stats_slots, "rd_conv2d_nhwc_stats: null statistics buffer");
  RD_REQUIRE(stats_bytes >= (size_t)conv::STATS_STRIDE * 2 * (size_t)Cout * sizeof(float),
             "rd_conv2d_nhwc_stats: statistics workspace too small (rd_bn_workspace_bytes(Cout))");
  return conv::run(0, x_pad, w_packed, nullptr, nullptr, nullptr, y_pad, N, H, W, Cin, Cout, ksize, stride_w, 0, 0,
                   rd::as_stream(stream), 0, 0, stats_partial, stats_slots);
}

// Data-gradient convolution that also accumulates the BatchNorm-backward sums of the layer below (see the header).
int RD_ACT_FN(rd_conv2d_nhwc_, _bwdstats)(const void* x_pad, const void* w_packed, void* y_pad, const void* bn_z_pad,
                                          const float* bn_coef, int bn_mask_mode, int N, int H, int W, int Cin, int Cout,
                                          float* sums_partial, size_t sums_bytes, int* sums_slots, rd_stream_t stream) {
  RD_REQUIRE(x_pad && w_packed && y_pad && bn_z_pad && bn_coef && sums_partial && sums_slots, "rd_conv2d_nhwc_bwdstats: null pointer");
  RD_REQUIRE(Cout == 128 && Cin >= 64 && Cin % 64 == 0 && Cin <= 1024,
             "rd_conv2d_nhwc_bwdstats: 3x3 / stride 1 with 128 output channels only (got %d -> %d)", Cin, Cout);
  RD_REQUIRE(bn_mask_mode == 0 || bn_mask_mode == 2, "rd_conv2d_nhwc_bwdstats: mask_mode must be 0 or 2");
  RD_REQUIRE(N > 0 && H > 0 && W > 0, "rd_conv2d_nhwc_bwdstats: bad shape");
  RD_REQUIRE(sums_bytes >= (size_t)conv::STATS_STRIDE * 2 * (size_t)Cout * sizeof(float),
             "rd_conv2d_nhwc_bwdstats: sums workspace too small (rd_bn_workspace_bytes(Cout))");
  if (rd_check_device()) return 1;
  return RD_ACT_FN(rd_convt_run_, )(x_pad, w_packed, nullptr, nullptr, nullptr, y_pad, N, H, W, Cin, Cout, 0, rd::as_stream(stream),
                                    sums_partial, sums_slots, bn_z_pad, bn_coef, bn_mask_mode);
}

int RD_ACT_FN(rd_conv2d_nhwc_, _slice)(const void* x_pad, const void* w_packed, const float* scale, const float* shift,
                              void* y_pad, int N, int H, int W, int Cin, int Cout, int ksize, int stride_w, int relu,
                              int y_ctotal, int y_coff, rd_stream_t stream) {
  return conv::run(0, x_pad, w_packed, scale, shift, nullptr, y_pad, N, H, W, Cin, Cout, ksize, stride_w, relu, 0,
                   rd::as_stream(stream), y_ctotal, y_coff);
}

int RD_ACT_FN(rd_deconv2d_nhwc_, )(const void* x_pad, const void* w_packed, const float* scale, const float* shift,
                          const void* residual_pad, void* y_pad, int N, int H, int W, int Cin, int Cout, int kw,
                          int relu, rd_stream_t stream) {
  RD_REQUIRE(kw == 8 || kw == 4 || kw == 3,
             "rd_deconv2d_nhwc_bf16: supported kernels are (3,8)/(1,4)/(1,2), (3,4)/(1,2)/(1,1) and (3,3)/(1,2)/(1,1)");
  return conv::run(kw == 8 ? 1 : (kw == 4 ? 2 : 3), x_pad, w_packed, scale, shift, residual_pad, y_pad, N, H, W, Cin, Cout, 3, 1, relu, 1,
                   rd::as_stream(stream));
}

}  // extern "C"
