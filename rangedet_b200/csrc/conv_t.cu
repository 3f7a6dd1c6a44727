// 3x3 stride-1 convolution with 128 output channels, "transposed" GEMM orientation (sm_100a, tcgen05).
//
// Same operation and tensor layouts as the 3x3 / stride-1 case of conv_tc.cu (mx.sym.Convolution of /root/reference
// mxnext/simple.py:123-158 at the 128-channel call sites: rangedet/symbol/backbone/dla_backbone.py:23-56 for res2 / res3a /
// res3 / agg2, rangedet/symbol/head/builder.py:198-246 for the 24 head-tower layers = 57 % of the model's FLOPs), but the
// MMA operands swap roles:
//
//     D[Cout = 128 TMEM lanes][256 pixel columns]  +=  W_tap[128 x 64]  .  X_tap[256 px x 64]^T        (M128 N256 K16)
//
// Why: with both operands in shared memory an M128 x N x K16 MMA reads (128 + N) * 32 B.  In conv_tc.cu (M = pixels,
// N = Cout = 128) that is 8 KB per 64 tensor-cycles = 128 B/clk, the whole shared-memory bandwidth of the SM -- before
// the TMA writes and the epilogue's staging traffic; ncu: tensor pipe 61-64 %.  Here N = 256 pixels: 12 KB per 128
// cycles = 96 B/clk.  (A cta_group::2 pair would reach the same ratio with two SMs and cluster plumbing.)
//
// Tiles run over the FLATTENED haloed pixel grid [N*(H+2)*(W+2)] instead of per image row: a tile is 256 consecutive
// flat pixels, a tap (dy, dx) is the constant offset (dy-1)*(W+2) + (dx-1), so every width packs tiles densely
// (W = 664: 666 half-tiles' worth instead of 768; W = 166: 173 instead of 256).  Outputs that fall on halo pixels are
// forced to zero in the epilogue, which is what the halo holds anyway; the first tile starts at the first interior pixel,
// so no TMA coordinate is negative.
//   warp 0    TMA producer: per K-half and dy ONE 258-pixel strip (a 256-pixel box + an 8-pixel box; the three dx taps
//             are row-shifted views of it), and the 128 x 64 weight tile of every tap, through two rings
//   warp 1    MMA issuer (converged warp, elected lane): 4 x tcgen05.mma M128 N256 K16 per tap and K-half, fp32
//             accumulators double-buffered in all 512 TMEM columns
//   warps 2-17 epilogue (four per TMEM lane quadrant, the tile's 32-column chunks dealt among them): thread = TMEM lane = OUTPUT
//             CHANNEL; scale / shift / residual / ReLU, halo mask, storage rounding in pixel pairs, lane pairs swap halves so
//             that 4-byte words [pixel][c, c+1] go into a [pixel][channel] 128B-swizzled staging tile -> two TMA stores;
//             the BatchNorm batch statistics of the stored values are plain per-thread sums here.
#include <stdlib.h>
#include <string.h>

#include "../../include/rangedet_b200.h"
#include "act_type.cuh"
#include "rd_common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

namespace RD_ACT_NS(convt) {

constexpr int TN = 256;                       // largest tile (pixels = MMA N); Params::tn (128..256, multiple of 32) is what a launch uses
constexpr int CO = 128;                       // output channels = MMA M
constexpr int KC = 64;                        // channels per K-half (one 128B swizzle atom)
constexpr int STRIP_ROWS = 264;               // 258 pixels used; 256-pixel box + 8-pixel box
constexpr int STRIP_BYTES = STRIP_ROWS * 128; // 33792
constexpr int WT_BYTES = CO * 128;            // 16384: one tap, one K-half
constexpr int HALF_BYTES = TN * 128;          // 32768: staging tile of 64 channels
constexpr int MAX_SA = 3, MAX_SB = 5;        // ring depths are chosen on the host (Params::nsa / nsb)
constexpr int EPI_WARPS = 16;                // four per TMEM lane quadrant
constexpr int NTHREADS = (3 + EPI_WARPS) * 32;   // warp 0 producer, warp 1 MMA, warps 2-17 epilogue, warp 18 staging DMA
constexpr int STATS_STRIDE = 1184;            // = bn::MAX_BLOCKS

struct Params {
  int kh;                 // Cin / 64
  int Wp, Hp;             // haloed width / height
  int P_total;            // N * Hp * Wp flat pixels
  int p_first;            // first interior pixel = Wp + 1
  int ntiles;
  int relu, has_res;
  int cout;               // 128, or 64: the weight tile's rows 64..127 are then TMA zero fill and TMEM lanes 64..127 idle
  int nsa, nsb;           // strip / weight ring stages
  int tn;                 // pixels per tile (MMA N): chosen per shape so that the last wave of tiles is full (see pick_tn)
  int bstat;              // BatchNorm-backward sums of the layer BELOW in the epilogue (see rd_conv2d_nhwc_*_bwdstats):
                          // 0 off, 1 g = dy, 2 g = dy where z*a + b > 0 (the ReLU mask recomputed from z)
  int off_w, off_o, off_misc;
};

struct Misc {
  uint64_t a_full[MAX_SA], a_empty[MAX_SA], b_full[MAX_SB], b_empty[MAX_SB], t_full[2], t_empty[2];
  uint64_t r_full[2];     // staging region (first / second part of the tile's pixels) writable: previous store has read it and,
                          // with a residual / z tile, that tile has landed
  uint64_t s_done[2];     // staging region complete: all epilogue warps have written their part
  uint32_t tmem_slot, pad;
  uint32_t mask[2][8];    // halo bits of the tile's pixels, one word per 32-column chunk, double-buffered by tile parity: written
                          // by the DMA warp before it hands region A over (mbarrier release / acquire orders it)
  float scale[CO], shift[CO];
};

// PROF: diagnostic build (RD_CONVT_PROF=1): clock64() cycles per role into prof[block][8]: 0 MMA total, 1 wait(a_full),
// 2 wait(b_full), 3 wait(t_empty); 4 epilogue total, 5 wait(t_full), 6 wait(store read) + mask + barrier, 7 body.
template <bool PROF>
__global__ void __launch_bounds__(NTHREADS, 1)
convt_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_x2,
             const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_yA,
             const __grid_constant__ CUtensorMap tm_yB, const __grid_constant__ CUtensorMap tm_rA,
             const __grid_constant__ CUtensorMap tm_rB, const float* __restrict__ scale, const float* __restrict__ shift,
             float* __restrict__ stats, const float* __restrict__ bcoef, long long* __restrict__ prof,
             const __grid_constant__ Params P) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  auto tick = [&]() -> long long { return PROF ? clock64() : 0ll; };
  const long long t_begin = tick();
  long long pc1 = 0, pc2 = 0, pc3 = 0;
  if (t == 0) rd::pdl_trigger();
  unsigned char* strips = base;
  unsigned char* wts = base + P.off_w;
  unsigned char* sO = base + P.off_o;
  Misc& M = *reinterpret_cast<Misc*>(base + P.off_misc);
  const uint32_t NSA = (uint32_t)P.nsa, NSB = (uint32_t)P.nsb;

  if (t == 0) {
    for (int i = 0; i < MAX_SA; ++i) { tc::mbar_init(&M.a_full[i], 1); tc::mbar_init(&M.a_empty[i], 1); }
    for (int i = 0; i < MAX_SB; ++i) { tc::mbar_init(&M.b_full[i], 1); tc::mbar_init(&M.b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&M.t_full[i], 1); tc::mbar_init(&M.t_empty[i], EPI_WARPS); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&M.r_full[i], 1); tc::mbar_init(&M.s_done[i], EPI_WARPS); }
    tc::fence_mbar_init();
    tma::prefetch_map(&tm_x);
    tma::prefetch_map(&tm_x2);
    tma::prefetch_map(&tm_w);
    tma::prefetch_map(&tm_yA);
    tma::prefetch_map(&tm_yB);
  }
  if (warp == 1) {
    tc::tmem_alloc(&M.tmem_slot, 512);
    tc::tmem_relinquish();
  }
  rd::pdl_wait();   // everything above touched only shared memory / TMEM / kernel parameters
  for (int c = t; c < P.cout; c += NTHREADS) {
    M.scale[c] = scale ? __ldg(scale + c) : 1.f;
    M.shift[c] = shift ? __ldg(shift + c) : 0.f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = M.tmem_slot;
  const int kh = P.kh;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t sa = 0, pa = 1, sb = 0, pb = 1;   // ring positions, parity of the `empty` wait
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        const int p0 = P.p_first + tile * P.tn;
        for (int q = 0; q < kh; ++q)
          for (int dy = 0; dy < 3; ++dy) {
            const int ps = p0 + (dy - 1) * P.Wp - 1;   // >= 0: p0 >= Wp + 1
            tc::mbar_wait(&M.a_empty[sa], pa);
            tc::mbar_arrive_expect_tx(&M.a_full[sa], (uint32_t)(P.tn + 8) * 128u);
            tma::load_2d(strips + sa * STRIP_BYTES, &tm_x, &M.a_full[sa], q * KC, ps);
            tma::load_2d(strips + sa * STRIP_BYTES + P.tn * 128, &tm_x2, &M.a_full[sa], q * KC, ps + P.tn);
            if (++sa == NSA) { sa = 0; pa ^= 1; }
            for (int dx = 0; dx < 3; ++dx) {
              tc::mbar_wait(&M.b_empty[sb], pb);
              tc::mbar_arrive_expect_tx(&M.b_full[sb], (uint32_t)WT_BYTES);
              tma::load_3d(wts + sb * WT_BYTES, &tm_w, &M.b_full[sb], q * KC, 0, dy * 3 + dx);
              if (++sb == NSB) { sb = 0; pb ^= 1; }
            }
          }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer: converged warp, one elected lane issues =====
    const uint32_t idesc = tc::make_idesc_f16kind(CO, P.tn, RD_ACT_MMA_FMT);
    const uint64_t desc_hi = tc::make_smem_desc(0, 0, 1024, tc::LAYOUT_SW128);  // everything but the address
    const uint32_t strip_lo = tc::smem_u32(strips) >> 4, wt_lo = tc::smem_u32(wts) >> 4;
    const bool leader = tc::elect_one();
    uint32_t sa = 0, pha = 0, sb = 0, phb = 0, it = 0;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1;
      { const long long a0 = tick(); tc::mbar_wait(&M.t_empty[buf], ((it >> 1) & 1) ^ 1); pc3 += tick() - a0; }
      const uint32_t d_tmem = tmem_base + buf * (uint32_t)TN;
      bool first = true;
      for (int q = 0; q < kh; ++q)
        for (int dy = 0; dy < 3; ++dy) {
          { const long long a0 = tick(); tc::mbar_wait(&M.a_full[sa], pha); pc1 += tick() - a0; }
          const uint32_t x_lo = strip_lo + sa * (uint32_t)(STRIP_BYTES >> 4);
#pragma unroll 1
          for (int dx = 0; dx < 3; ++dx) {
            { const long long a0 = tick(); tc::mbar_wait(&M.b_full[sb], phb); pc2 += tick() - a0; }
            tc::tc_fence_after();
            if (leader) {
              // A = weight tile (128 rows), B = strip rows [dx, dx + 256): one pixel = one 128-byte row = 8 address units
              const uint64_t wd = desc_hi | (uint64_t)((wt_lo + sb * (uint32_t)(WT_BYTES >> 4)) & 0x3FFF);
              const uint64_t xd = desc_hi | (uint64_t)((x_lo + (uint32_t)dx * 8u) & 0x3FFF);
              tc::mma_bf16_ss(d_tmem, wd, xd, idesc, first ? 0u : 1u);
              tc::mma_bf16_ss_acc(d_tmem, wd + 2, xd + 2, idesc);
              tc::mma_bf16_ss_acc(d_tmem, wd + 4, xd + 4, idesc);
              tc::mma_bf16_ss_acc(d_tmem, wd + 6, xd + 6, idesc);
              tc::umma_commit(&M.b_empty[sb]);
              if (dx == 2) tc::umma_commit(&M.a_empty[sa]);
            }
            first = false;
            __syncwarp();
            if (++sb == NSB) { sb = 0; phb ^= 1; }
          }
          if (++sa == NSA) { sa = 0; pha ^= 1; }
        }
      if (leader) tc::umma_commit(&M.t_full[buf]);
      __syncwarp();
    }
    if (PROF && lane == 0) {
      prof[blockIdx.x * 8 + 0] = tick() - t_begin;
      prof[blockIdx.x * 8 + 1] = pc1;
      prof[blockIdx.x * 8 + 2] = pc2;
      prof[blockIdx.x * 8 + 3] = pc3;
    }
  } else if (warp < 2 + EPI_WARPS) {
    // ===== epilogue: thread = TMEM lane = output channel; SIXTEEN warps, four per TMEM lane quadrant =====
    // (The body is a chain of TMEM load -> shared-memory read -> convert -> shuffle -> shared-memory write latencies: four
    //  warps writing 2-byte elements spent 14.4 k cycles per tile against 9.2 k of MMA time, eight warps 8.2 k (plain) /
    //  10.7 k (residual) / 12.8 k (BN-backward sums), sixteen 5.1 / 6.9 / 8.7 k -- every variant is now behind the MMA
    //  warp's 11 k.  The warps of a quadrant take the 32-column chunks round-robin, values are rounded in pairs, and lane
    //  pairs (c, c+1) swap halves so that every store is a 4-byte word
    //  [pixel][c, c+1] -- 16 lanes cover 64 contiguous bytes of a row, the 128B swizzle keeps the two rows of a warp-wide
    //  store in different banks.)
    // The tile's pixels are handled as TWO staging regions (first / second half of the 32-column chunks).  The DMA warp
    // stores region A and reloads it for the next tile (residual / z tile) while these warps work on region B, and vice
    // versa: no epilogue warp ever waits for a TMA store to drain or for a residual load issued just now -- with a single
    // region those two waits were 2.3 k (plain) / 5.3 k (residual) of ~12-16 k cycles per tile (RD_CONVT_PROF=1).
    const int ew = warp - 2;                      // 0..EPI_WARPS-1
    const int q4 = warp & 3;                      // TMEM lane quadrant this warp may read
    const int part = ew >> 2;                     // which of the quadrant's warps: chunks are dealt round-robin among them
    constexpr int NPART = EPI_WARPS / 4;
    const int nch = P.tn >> 5;                    // 32-column chunks of the tile (4..8)
    const int nA = (nch + 1) >> 1;                // chunks of region A
    const int c = q4 * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const bool live = c < P.cout;                 // whole warps: q4 >= 2 idles when Cout == 64 (it still arrives on every barrier)
    const float sc = live ? M.scale[c] : 0.f, sh = live ? M.shift[c] : 0.f;
    // BatchNorm below (bstat): a | b | mean rows of its coefficient block
    const float za = (P.bstat && live) ? __ldg(bcoef + c) : 0.f, zb = (P.bstat && live) ? __ldg(bcoef + P.cout + c) : 0.f;
    const float zm = (P.bstat && live) ? __ldg(bcoef + 2 * P.cout + c) : 0.f;
    const uint32_t odd = (uint32_t)(lane & 1);
    const uint32_t sel = odd ? 0x3276u : 0x5410u; // odd: (partner.hi, mine.hi) = channels (c-1, c) of pixel j+1; even: (mine.lo, partner.lo)
    unsigned char* my_row = sO + (c >> 6) * HALF_BYTES + odd * 128 + (c & 6) * 2;   // word (c & ~1) of row j + odd
    unsigned char* my_col = sO + (c >> 6) * HALF_BYTES + (c & 7) * 2;               // 2-byte element of channel c (residual reads)
    const uint32_t my_chunk = (uint32_t)((c & 63) >> 3);
    float s_sum = 0.f, s_sq = 0.f;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1;
      const uint32_t t_acc = tmem_base + lane_sel + buf * (uint32_t)TN;
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
        const int s0 = sub ? nA : 0, n = sub ? nch - nA : nA;
        const int lo = s0 + part, hi = s0 + n;
        const long long e0 = tick();
        tc::mbar_wait(&M.r_full[sub], it & 1);
        const long long e1 = tick();
        pc2 += e1 - e0;
        if (sub == 0) {
          tc::mbar_wait(&M.t_full[buf], (it >> 1) & 1);
          __syncwarp();
          tc::tc_fence_after();
          pc1 += tick() - e1;
        }
        const long long e2 = tick();
#pragma unroll 1
        for (int ch = lo; ch < (live ? hi : lo); ch += NPART) {
          const uint32_t mbits = M.mask[it & 1][ch];
          float v[32];
          tc::tmem_ld_x32(t_acc + (uint32_t)(ch * 32), v);
          const int px0 = ch * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float a0 = fmaf(v[j], sc, sh), a1 = fmaf(v[j + 1], sc, sh);
            float z0 = 0.f, z1 = 0.f;
            if (P.has_res || P.bstat) {
              z0 = act::to_float(*reinterpret_cast<const act_t*>(my_col + (px0 + j) * 128 + ((my_chunk ^ (uint32_t)(j & 7)) << 4)));
              z1 = act::to_float(*reinterpret_cast<const act_t*>(my_col + (px0 + j + 1) * 128 + ((my_chunk ^ (uint32_t)((j + 1) & 7)) << 4)));
            }
            if (P.has_res) { a0 += z0; a1 += z1; }
            if (P.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
            if ((mbits >> j) & 1u) a0 = 0.f;        // halo pixels: keep them zero
            if ((mbits >> (j + 1)) & 1u) a1 = 0.f;
            const uint32_t mine = act::pack2(a0, a1);                      // channel c: (pixel j, pixel j + 1)
            float f0, f1;
            act::unpack2(mine, f0, f1);
            if (P.bstat) {   // S1 = sum g, S2 = sum g (z - mean) of the stored gradient (what bn::s_bwd_reduce_kernel would read back)
              if (P.bstat == 2) {
                f0 = fmaf(z0, za, zb) > 0.f ? f0 : 0.f;
                f1 = fmaf(z1, za, zb) > 0.f ? f1 : 0.f;
              }
              s_sum += f0 + f1;
              s_sq = fmaf(f0, z0 - zm, fmaf(f1, z1 - zm, s_sq));
            } else {
              s_sum += f0 + f1;
              s_sq = fmaf(f0, f0, fmaf(f1, f1, s_sq));
            }
            const uint32_t theirs = __shfl_xor_sync(0xffffffffu, mine, 1);
            const uint32_t word = __byte_perm(mine, theirs, sel);          // channels (c & ~1, +1) of pixel j + odd
            // all lanes of a pair have passed their residual reads of both rows before either writes (shuffle above)
            *reinterpret_cast<uint32_t*>(my_row + (px0 + j) * 128 + (((my_chunk ^ odd) ^ (uint32_t)(j & 7)) << 4)) = word;
          }
        }
        if (sub == 1) tc::tc_fence_before();
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tc::mbar_arrive(&M.s_done[sub]);
          if (sub == 1) tc::mbar_arrive(&M.t_empty[buf]);
        }
        pc3 += tick() - e2;
      }
    }
    if (PROF && warp == 2 && lane == 0) {
      prof[blockIdx.x * 8 + 4] = tick() - t_begin;
      prof[blockIdx.x * 8 + 5] = pc1;
      prof[blockIdx.x * 8 + 6] = pc2;
      prof[blockIdx.x * 8 + 7] = pc3;
    }
    if (stats != nullptr && live) {   // partial[(which * Cout + channel) * STATS_STRIDE + slot], slot = NPART * CTA + warp of the quadrant
      const int slot = (int)blockIdx.x * NPART + part;
      stats[(int64_t)c * STATS_STRIDE + slot] = s_sum;
      stats[(int64_t)(P.cout + c) * STATS_STRIDE + slot] = s_sq;
    }
  } else {
    // ===== staging DMA (one thread): per region, TMA store once the epilogue warps are done with it, then -- as soon as
    // the store has read the shared memory -- hand the region back for the next tile, loaded with that tile's residual /
    // z rows if the launch has any =====
    const int nch = P.tn >> 5, nA = (nch + 1) >> 1;
    const int nh = P.cout / KC;                 // 64-channel halves of the staging tile in use
    const int px_of[2] = {0, nA * 32}, npx[2] = {nA * 32, (nch - nA) * 32};
    // halo bits of tile_ (whole warp: lane = pixel of a chunk) -> M.mask[parity]
    auto halo_mask = [&](int tile_, uint32_t parity) {
      for (int ch = 0; ch < nch; ++ch) {
        const int p = P.p_first + tile_ * P.tn + ch * 32 + lane;
        bool halo = true;
        if (p < P.P_total) {
          const int col = p % P.Wp, row = (p / P.Wp) % P.Hp;
          halo = col == 0 || col == P.Wp - 1 || row == 0 || row == P.Hp - 1;
        }
        const uint32_t bits = __ballot_sync(0xffffffffu, halo);
        if (lane == 0) M.mask[parity][ch] = bits;
      }
    };
    auto hand_over = [&](int sub, int tile_) {   // lane 0
      if (P.has_res || P.bstat) {
        tc::mbar_arrive_expect_tx(&M.r_full[sub], (uint32_t)(nh * npx[sub]) * 128u);
        for (int hf = 0; hf < nh; ++hf)
          tma::load_2d(sO + hf * HALF_BYTES + px_of[sub] * 128, sub ? &tm_rB : &tm_rA, &M.r_full[sub], hf * KC,
                       P.p_first + tile_ * P.tn + px_of[sub]);
      } else {
        tc::mbar_arrive(&M.r_full[sub]);
      }
    };
    if ((int)blockIdx.x < P.ntiles) {
      halo_mask(blockIdx.x, 0);
      if (lane == 0) {
        hand_over(0, blockIdx.x);
        hand_over(1, blockIdx.x);
      }
    }
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x, ++it) {
      const int p0 = P.p_first + tile * P.tn;
      const int next = tile + (int)gridDim.x;
      // mask[(it + 1) & 1] was last read for tile it - 1, whose epilogue is over: the warps have all arrived on
      // s_done[1] of that tile, which lane 0 waited for before coming here
      if (next < P.ntiles) halo_mask(next, (it + 1) & 1);
      if (lane == 0) {
        for (int sub = 0; sub < 2; ++sub) {
          tc::mbar_wait(&M.s_done[sub], it & 1);
          for (int hf = 0; hf < nh; ++hf)
            tma::store_2d(sub ? &tm_yB : &tm_yA, sO + hf * HALF_BYTES + px_of[sub] * 128, hf * KC, p0 + px_of[sub]);
          tma::store_commit();
          if (next < P.ntiles) {
            tma::store_wait_read<0>();
            hand_over(sub, next);
          }
        }
      }
      __syncwarp();
    }
    if (lane == 0) tma::store_wait_all<0>();
    __syncwarp();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace convt_<storage type>
namespace convt = RD_ACT_NS(convt);

// Pixels per tile.  One persistent CTA per SM walks the tiles, so the kernel's time is (tiles per CTA, rounded up) x (tile
// time).  Measured at 128 -> 128, B = 2 (scripts/conv_t_tiles.py): tile time = 2.3 us + 0.0234 us per pixel -- the constant is
// the 288 KB of weight tiles every CTA re-streams from L2 per tile, ~96 pixels' worth -- so narrower tiles pay only where
// they save a round: W = 664 is 346 tiles of 256 pixels = 2.3 per SM -> three rounds, at 224 pixels three shorter ones
// (32.2 -> 30.5 us); W = 332: 160 pixels (24.0 -> 19.3 us); W = 2656 / 1328 stay at 256.  RD_CONVT_TN=<n> or
// rd_set_conv_t(<n>) fix the width (A/B, tests).
static int pick_tn(int npx, int sms) {
  using convt::TN;
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("RD_CONVT_TN");
    forced = e ? atoi(e) : 0;
    if (forced % 32 || forced < 128 || forced > TN) forced = 0;
  }
  if (rd::conv_t_tile()) return rd::conv_t_tile();
  if (forced) return forced;
  int best = TN;
  int64_t best_cost = -1;
  for (int tn = TN; tn >= 160; tn -= 32) {
    const int tiles = (npx + tn - 1) / tn;
    const int rounds = (tiles + sms - 1) / sms;
    const int64_t cost = (int64_t)rounds * (tn + 96);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = tn; }
  }
  return best;
}

// Called by conv_tc.cu's dispatcher (same storage-type pass).  y = relu?(conv3x3(x) * scale + shift + residual), Cout = 128 or 64.
int RD_ACT_FN(rd_convt_run_, )(const void* x_pad, const void* w_packed, const float* scale, const float* shift,
                               const void* residual_pad, void* y_pad, int N, int H, int W, int Cin, int Cout, int relu,
                               cudaStream_t stream, float* stats, int* stats_slots, const void* bn_z, const float* bn_coef,
                               int bn_mask) {
  using namespace convt;
  RD_REQUIRE(Cout == 64 || Cout == 128, "rd_conv(T): Cout must be 64 or 128");
  RD_REQUIRE(!bn_z || (!residual_pad && stats && bn_coef && (bn_mask == 0 || bn_mask == 2)),
             "rd_conv(T): BatchNorm-backward sums need a statistics buffer, the coefficient block, mask_mode 0 or 2, and no residual");
  Params P;
  memset(&P, 0, sizeof(P));
  P.kh = Cin / KC;
  P.cout = Cout;
  P.Wp = W + 2;
  P.Hp = H + 2;
  const int64_t total = (int64_t)N * P.Hp * P.Wp;
  RD_REQUIRE(total < 0x7fffff00LL, "rd_conv(T): tensor too large for 32-bit flat pixel coordinates");
  P.P_total = (int)total;
  P.p_first = P.Wp + 1;
  const int p_last = P.P_total - P.Wp - 2;   // last interior pixel of the last image
  int dev = 0, sms = 0;
  RD_CUDA(cudaGetDevice(&dev));
  RD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  P.tn = pick_tn(p_last - P.p_first + 1, sms);
  P.ntiles = (p_last - P.p_first + 1 + P.tn - 1) / P.tn;
  P.relu = relu ? 1 : 0;
  P.has_res = residual_pad ? 1 : 0;
  P.bstat = bn_z ? (bn_mask == 2 ? 2 : 1) : 0;
  CUtensorMap tm_x, tm_x2, tm_w, tm_yA, tm_yB, tm_rA, tm_rB;
  {
    const uint64_t d[2] = {(uint64_t)Cin, (uint64_t)total};
    const uint64_t s[1] = {(uint64_t)Cin * 2};
    const uint32_t b[2] = {(uint32_t)KC, (uint32_t)P.tn}, b2[2] = {(uint32_t)KC, 8u};
    if (tma::make_map(&tm_x, RD_ACT_TMA_TYPE, x_pad, 2, d, s, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    if (tma::make_map(&tm_x2, RD_ACT_TMA_TYPE, x_pad, 2, d, s, b2, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    // box of 128 rows on a map with Cout rows: with Cout == 64 the upper half of every weight tile is TMA zero fill
    const uint64_t d[3] = {(uint64_t)Cin, (uint64_t)Cout, 9u};
    const uint64_t s[2] = {(uint64_t)Cin * 2, (uint64_t)Cin * Cout * 2};
    const uint32_t b[3] = {(uint32_t)KC, (uint32_t)CO, 1u};
    if (tma::make_map(&tm_w, RD_ACT_TMA_TYPE, w_packed, 3, d, s, b, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    const uint64_t d[2] = {(uint64_t)Cout, (uint64_t)total};
    const uint64_t s[1] = {(uint64_t)Cout * 2};
    const int nch = P.tn / 32, nA = (nch + 1) / 2;   // the two staging regions of a tile (see the epilogue)
    const uint32_t bA[2] = {(uint32_t)KC, (uint32_t)(nA * 32)}, bB[2] = {(uint32_t)KC, (uint32_t)((nch - nA) * 32)};
    const void* rsrc = residual_pad ? residual_pad : (bn_z ? bn_z : y_pad);
    if (tma::make_map(&tm_yA, RD_ACT_TMA_TYPE, y_pad, 2, d, s, bA, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    if (tma::make_map(&tm_yB, RD_ACT_TMA_TYPE, y_pad, 2, d, s, bB, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    if (tma::make_map(&tm_rA, RD_ACT_TMA_TYPE, rsrc, 2, d, s, bA, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    if (tma::make_map(&tm_rB, RD_ACT_TMA_TYPE, rsrc, 2, d, s, bB, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  }
  {
    // 227 KB: strips 33 KB each, weight tiles 16 KB each, staging 64 KB.  2 strips + 5 weight tiles; 2+4 and 3+3 measured
    // within 2 % (RD_CONVT_RINGS=ab overrides): the MMA warp's remaining waits (~4 k of 12 k cycles per tile) are the
    // L2 -> shared-memory stream itself (488 KB per tile and SM, 59 % of it the weight tiles every CTA re-reads).
    P.nsa = 2; P.nsb = 5;
    const char* e = getenv("RD_CONVT_RINGS");
    if (e && e[0] >= '2' && e[0] <= '3' && e[1] >= '2' && e[1] <= '5') { P.nsa = e[0] - '0'; P.nsb = e[1] - '0'; }
    P.off_w = P.nsa * STRIP_BYTES;
    P.off_o = P.off_w + P.nsb * WT_BYTES;
    P.off_misc = P.off_o + 2 * HALF_BYTES;
  }
  const size_t smem = (size_t)P.off_misc + sizeof(Misc) + 1024;
  RD_REQUIRE(smem <= 227 * 1024, "rd_conv(T): shared memory layout exceeds 227 KB (%zu)", smem);
  RD_CUDA(rd::smem_optin(convt_kernel<false>, smem));
  RD_CUDA(rd::smem_optin(convt_kernel<true>, smem));
  const int grid = P.ntiles < sms ? P.ntiles : sms;
  if (stats) {
    constexpr int NPART = EPI_WARPS / 4;
    RD_REQUIRE(NPART * grid <= STATS_STRIDE, "rd_conv(T) stats: %d partial slots exceed %d", NPART * grid, STATS_STRIDE);
    if (stats_slots) *stats_slots = NPART * grid;
  }
  static const bool want_prof = [] { const char* e = getenv("RD_CONVT_PROF"); return e && e[0] == '1'; }();
  if (want_prof) {   // diagnostic: per-role cycle counters, synchronous, printed to stderr
    static long long* d_prof = nullptr;
    if (!d_prof) RD_CUDA(cudaMalloc(&d_prof, 1024 * 8 * sizeof(long long)));
    RD_CUDA(cudaMemsetAsync(d_prof, 0, 1024 * 8 * sizeof(long long), stream));
    RD_CUDA(rd::launch(convt_kernel<true>, dim3(grid), dim3(NTHREADS), smem, stream, tm_x, tm_x2, tm_w, tm_yA, tm_yB, tm_rA, tm_rB, scale, shift, stats,
                       bn_coef, d_prof, P));
    RD_CUDA(cudaStreamSynchronize(stream));
    static long long h[1024 * 8];
    RD_CUDA(cudaMemcpy(h, d_prof, sizeof(long long) * 8 * grid, cudaMemcpyDeviceToHost));
    double m[8] = {0};
    for (int b = 0; b < grid; ++b)
      for (int k = 0; k < 8; ++k) m[k] += (double)h[b * 8 + k] / grid;
    const double tiles = (double)P.ntiles / grid;
    fprintf(stderr, "[rd_conv(T) prof] Cin=%d tiles/cta=%.1f | per tile: mma %.0f (wait a_full %.0f, b_full %.0f, t_empty %.0f) | epi %.0f "
            "(wait t_full %.0f, wait region %.0f, body %.0f)\n", Cin, tiles, m[0] / tiles, m[1] / tiles, m[2] / tiles, m[3] / tiles,
            m[4] / tiles, m[5] / tiles, m[6] / tiles, m[7] / tiles);
  } else {
    RD_CUDA(rd::launch(convt_kernel<false>, dim3(grid), dim3(NTHREADS), smem, stream, tm_x, tm_x2, tm_w, tm_yA, tm_yB, tm_rA, tm_rB, scale, shift, stats,
                       bn_coef, (long long*)nullptr, P));
  }
  rd::count_launch();
  return rd::check_launch("rd_conv(T)");
}
