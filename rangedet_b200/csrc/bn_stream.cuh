// TMA-staged streaming variants of the BatchNorm passes (included by bn_train.cu inside namespace bn).
//
// The register-staged kernels in bn_train.cu keep at most 4-6 16-byte loads per thread in flight
// (24-48 KB per SM): ncu shows them at 33 % (stats), 50 % (bwd_reduce) and 56 % (bwd_apply) of the DRAM peak with
// the warps parked on the long scoreboard.  Here one elected thread streams contiguous row segments
// (UP pixels x C channels, 22-48 KB per tensor) into a 3-stage shared-memory ring with cp.async.bulk + mbarrier
// complete_tx: up to ~190 KB per SM in flight at no register cost; 512 threads consume each stage with
// conflict-free 16-byte shared loads and write results with 16-byte global stores.  One persistent CTA per SM.
#pragma once

constexpr int SNT = 512;
constexpr int STAGES = 3;

struct SGeo {
  int N, H, W, C;
  int cgs, ppb;      // 8-channel groups per pixel; pixels per pass of the block (ppb * cgs <= SNT threads work)
  int up, chunks;    // pixels per unit (multiple of ppb), units per image row
  int nunits;
  uint32_t ubytes;   // bytes of a full unit of one tensor
};

static SGeo make_sgeo(int N, int H, int W, int C, int ntensors) {
  SGeo g;
  g.N = N; g.H = H; g.W = W; g.C = C;
  g.cgs = C / 8;
  g.ppb = SNT / g.cgs;
  // measured (scripts/bn_bench.py, 4x64x2656x128): 12 KB units 3.0 TB/s, 16 KB 4.6, 24 KB 5.5, 32 KB 5.9 TB/s for the
  // backward pair -- the ring must hold ~100-190 KB per SM to cover the HBM latency-bandwidth product
  int target = ntensors == 1 ? 49152 : (ntensors == 2 ? 32768 : 21840);
  {
    static int env_ub = -1;   // RD_BN_UB=<bytes>: tuning override of the per-tensor unit size (2 tensors; 3 use 2/3 of it)
    if (env_ub < 0) {
      const char* e = getenv("RD_BN_UB");
      env_ub = e ? atoi(e) : 0;
    }
    if (env_ub >= 2048 && env_ub <= 32768) target = ntensors == 1 ? env_ub * 3 / 2 : (ntensors == 2 ? env_ub : env_ub * 2 / 3);
  }
  int up = target / (2 * C) / g.ppb * g.ppb;
  if (up < g.ppb) up = g.ppb;
  g.up = up;
  g.chunks = (W + up - 1) / up;
  g.nunits = N * H * g.chunks;
  g.ubytes = (uint32_t)up * C * 2;
  return g;
}

static size_t sgeo_smem(const SGeo& g, int ntensors) { return (size_t)STAGES * ntensors * g.ubytes + 128; }

static int stream_grid(const SGeo& g) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return g.nunits < sms ? g.nunits : sms;
}

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void s_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void s_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = s_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}
// 1-D bulk copy global -> shared (TMA engine, no tensor map): 16-byte aligned addresses, size multiple of 16
__device__ __forceinline__ void s_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)),
               "l"(src), "r"(bytes), "r"(s_u32(bar))
               : "memory");
}

struct SUnit {
  int n, h, w0, npx;
  int64_t off;   // element offset of pixel w0 in a haloed (W halo 1) tensor
};
__device__ __forceinline__ SUnit s_unit(const SGeo& G, int u) {
  SUnit U;
  const int chunk = u % G.chunks, row = u / G.chunks;
  U.h = row % G.H;
  U.n = row / G.H;
  U.w0 = chunk * G.up;
  U.npx = min(G.up, G.W - U.w0);
  U.off = pix_off(U.n, U.h, U.w0, G.H, G.W, 1, G.C);
  return U;
}

// Ring of STAGES x nt unit buffers; thread 0 produces, everybody consumes.
struct SRing {
  unsigned char* buf;
  uint64_t* full;
  const act_t* src[3];
  int nt;
  int my_units;
  int rev;   // 1: walk the units from the far end of the tensor (see stream_rev())

  __device__ __forceinline__ void init(unsigned char* dsm, uint64_t* bars, const SGeo& G) {
    buf = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dsm) + 127) & ~(uintptr_t)127);
    full = bars;
    my_units = (int)blockIdx.x < G.nunits ? (G.nunits - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (threadIdx.x == 0) {
      rd::pdl_trigger();
      for (int s = 0; s < STAGES; ++s) s_mbar_init(&full[s], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    rd::pdl_wait();   // programmatic dependent launch: the predecessor's tensors are complete from here on
    if (threadIdx.x == 0)
      for (int k = 0; k < STAGES && k < my_units; ++k) issue(G, k);
  }
  __device__ __forceinline__ int unit_of(const SGeo& G, int k) const {
    const int u = (int)blockIdx.x + k * (int)gridDim.x;
    return rev ? G.nunits - 1 - u : u;
  }
  __device__ __forceinline__ void issue(const SGeo& G, int k) {
    const SUnit U = s_unit(G, unit_of(G, k));
    const int stage = k % STAGES;
    const uint32_t bytes = (uint32_t)U.npx * G.C * 2;
    s_mbar_expect_tx(&full[stage], bytes * nt);
    for (int t = 0; t < nt; ++t) s_bulk_g2s(buf + ((size_t)stage * nt + t) * G.ubytes, src[t] + U.off, bytes, &full[stage]);
  }
  __device__ __forceinline__ const uint4* wait(const SGeo& G, int k, int t) const {
    return reinterpret_cast<const uint4*>(buf + ((size_t)(k % STAGES) * nt + t) * G.ubytes);
  }
  __device__ __forceinline__ void acquire(int k) { s_mbar_wait(&full[k % STAGES], (uint32_t)(k / STAGES) & 1u); }
  // all threads have finished reading stage k: refill it with unit k + STAGES
  __device__ __forceinline__ void release(const SGeo& G, int k) {
    __syncthreads();
    if (threadIdx.x == 0 && k + STAGES < my_units) issue(G, k + STAGES);
  }
};

__device__ __forceinline__ void s_block_reduce_store(const float (&s)[8], const float (&q)[8], const SGeo& G, bool act,
                                                     float* __restrict__ partial) {
  __shared__ float red[SNT * 16];
  const int t = threadIdx.x;
  const int cg = t % G.cgs, pl = t / G.cgs;
  if (act) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      red[(pl * G.cgs + cg) * 16 + i] = s[i];
      red[(pl * G.cgs + cg) * 16 + 8 + i] = q[i];
    }
  }
  __syncthreads();
  const int nval = G.cgs * 16;  // = 2C
  for (int v = t; v < nval; v += SNT) {
    float a = 0.f;
    for (int p = 0; p < G.ppb; ++p) a += red[p * nval + v];
    const int cgi = v / 16, i = v % 16;
    partial[(int64_t)((i / 8) * G.C + cgi * 8 + (i % 8)) * MAX_BLOCKS + blockIdx.x] = a;
  }
}

// ---- forward statistics ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SNT, 1) s_stats_kernel(const act_t* __restrict__ z, SGeo G,
                                                         float* __restrict__ partial, int rev) {
  extern __shared__ unsigned char dsm[];
  __shared__ uint64_t bars[STAGES];
  SRing R;
  R.src[0] = z; R.src[1] = nullptr; R.src[2] = nullptr;
  R.nt = 1;
  R.rev = rev;
  R.init(dsm, bars, G);
  const int t = threadIdx.x, cg = t % G.cgs, pl = t / G.cgs;
  const bool act = pl < G.ppb;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  for (int k = 0; k < R.my_units; ++k) {
    const SUnit U = s_unit(G, R.unit_of(G, k));
    R.acquire(k);
    const uint4* sz = R.wait(G, k, 0);
    if (act) {
#pragma unroll 2
      for (int p = pl; p < U.npx; p += G.ppb) {
        float f[8];
        unpack8(sz[p * G.cgs + cg], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
      }
    }
    R.release(G, k);
  }
  s_block_reduce_store(s, q, G, act, partial);
}

// ---- forward apply: y = relu?(z*a + b + rb) + ra --------------------------------------------------------------
__global__ void __launch_bounds__(SNT, 1) s_fwd_apply_kernel(const act_t* __restrict__ z,
                                                             const float* __restrict__ coef,
                                                             const act_t* __restrict__ rb,
                                                             const act_t* __restrict__ ra,
                                                             act_t* __restrict__ y, SGeo G, int relu, int rev) {
  extern __shared__ unsigned char dsm[];
  __shared__ uint64_t bars[STAGES];
  SRing R;
  int nt = 0;
  R.src[0] = R.src[1] = R.src[2] = nullptr;
  R.src[nt++] = z;
  const int i_rb = rb ? nt : -1;
  if (rb) R.src[nt++] = rb;
  const int i_ra = ra ? nt : -1;
  if (ra) R.src[nt++] = ra;
  R.nt = nt;
  R.rev = rev;
  R.init(dsm, bars, G);
  const int t = threadIdx.x, cg = t % G.cgs, pl = t / G.cgs;
  const bool act = pl < G.ppb;
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = coef[cg * 8 + i]; b[i] = coef[G.C + cg * 8 + i]; }
  for (int k = 0; k < R.my_units; ++k) {
    const SUnit U = s_unit(G, R.unit_of(G, k));
    R.acquire(k);
    const uint4* sz = R.wait(G, k, 0);
    const uint4* srb = i_rb >= 0 ? R.wait(G, k, i_rb) : nullptr;
    const uint4* sra = i_ra >= 0 ? R.wait(G, k, i_ra) : nullptr;
    if (act) {
#pragma unroll 2
      for (int p = pl; p < U.npx; p += G.ppb) {
        const int e = p * G.cgs + cg;
        float f[8], r[8];
        unpack8(sz[e], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], a[i], b[i]);
        if (srb) {
          unpack8(srb[e], r);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] += r[i];
        }
        if (relu) {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
        }
        if (sra) {
          unpack8(sra[e], r);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] += r[i];
        }
        *reinterpret_cast<uint4*>(y + U.off + (int64_t)p * G.C + cg * 8) = pack8(f);
      }
    }
    R.release(G, k);
  }
}

// ---- backward reduce ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SNT, 1) s_bwd_reduce_kernel(const act_t* __restrict__ dy,
                                                              const act_t* __restrict__ ym,
                                                              const act_t* __restrict__ z,
                                                              const float* __restrict__ coef, SGeo G, int mask_mode,
                                                              float* __restrict__ partial, int rev) {
  extern __shared__ unsigned char dsm[];
  __shared__ uint64_t bars[STAGES];
  SRing R;
  R.src[0] = dy; R.src[1] = z; R.src[2] = mask_mode == 1 ? ym : nullptr;
  R.nt = mask_mode == 1 ? 3 : 2;
  R.rev = rev;
  R.init(dsm, bars, G);
  const int t = threadIdx.x, cg = t % G.cgs, pl = t / G.cgs;
  const bool act = pl < G.ppb;
  float a[8], b[8], mean[8], s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = coef[cg * 8 + i];
    b[i] = coef[G.C + cg * 8 + i];
    mean[i] = coef[2 * G.C + cg * 8 + i];
    s[i] = q[i] = 0.f;
  }
  for (int k = 0; k < R.my_units; ++k) {
    const SUnit U = s_unit(G, R.unit_of(G, k));
    R.acquire(k);
    const uint4* sd = R.wait(G, k, 0);
    const uint4* sz = R.wait(G, k, 1);
    const uint4* sm = mask_mode == 1 ? R.wait(G, k, 2) : nullptr;
    if (act) {
#pragma unroll 2
      for (int p = pl; p < U.npx; p += G.ppb) {
        const int e = p * G.cgs + cg;
        float g[8], zf[8];
        unpack8(sd[e], g);
        unpack8(sz[e], zf);
        masked_grad_v(g, zf, sm ? sm[e] : make_uint4(0, 0, 0, 0), mask_mode, a, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += g[i]; q[i] = fmaf(g[i], zf[i] - mean[i], q[i]); }
      }
    }
    R.release(G, k);
  }
  s_block_reduce_store(s, q, G, act, partial);
}

// ---- backward apply ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SNT, 1) s_bwd_apply_kernel(const act_t* __restrict__ dy,
                                                             const act_t* __restrict__ ym,
                                                             const act_t* __restrict__ z,
                                                             const float* __restrict__ coef,
                                                             const float* __restrict__ coef2, SGeo G, int mask_mode,
                                                             act_t* __restrict__ dz, int dz_halo,
                                                             act_t* __restrict__ g_out, int rev) {
  extern __shared__ unsigned char dsm[];
  __shared__ uint64_t bars[STAGES];
  SRing R;
  R.src[0] = dy; R.src[1] = z; R.src[2] = mask_mode == 1 ? ym : nullptr;
  R.nt = mask_mode == 1 ? 3 : 2;
  R.rev = rev;
  R.init(dsm, bars, G);
  const int t = threadIdx.x, cg = t % G.cgs, pl = t / G.cgs;
  const bool act = pl < G.ppb;
  float a[8], b[8], mean[8], c1[8], c2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = coef[cg * 8 + i];
    b[i] = coef[G.C + cg * 8 + i];
    mean[i] = coef[2 * G.C + cg * 8 + i];
    c1[i] = coef2[cg * 8 + i];
    c2[i] = coef2[G.C + cg * 8 + i];
  }
  for (int k = 0; k < R.my_units; ++k) {
    const SUnit U = s_unit(G, R.unit_of(G, k));
    R.acquire(k);
    const uint4* sd = R.wait(G, k, 0);
    const uint4* sz = R.wait(G, k, 1);
    const uint4* sm = mask_mode == 1 ? R.wait(G, k, 2) : nullptr;
    const int64_t obase = pix_off(U.n, U.h, U.w0, G.H, G.W, dz_halo, G.C) + cg * 8;
    if (act) {
#pragma unroll 2
      for (int p = pl; p < U.npx; p += G.ppb) {
        const int e = p * G.cgs + cg;
        float g[8], zf[8], d[8];
        unpack8(sd[e], g);
        unpack8(sz[e], zf);
        masked_grad_v(g, zf, sm ? sm[e] : make_uint4(0, 0, 0, 0), mask_mode, a, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = a[i] * (g[i] - c1[i] - (zf[i] - mean[i]) * c2[i]);
        *reinterpret_cast<uint4*>(dz + obase + (int64_t)p * G.C) = pack8(d);
        if (g_out) *reinterpret_cast<uint4*>(g_out + U.off + (int64_t)p * G.C + cg * 8) = pack8(g);
      }
    }
    R.release(G, k);
  }
}

static bool stream_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RD_BN_STREAM");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// Walk order of the units.  A tensor of the training step is 45-90 MB and the L2 holds 126 MB: a pass that starts where
// its producer (or the previous pass over the same tensors) stopped finds the most recently touched part still in L2,
// a pass that starts at the same end finds it evicted.  The convolutions write ascending; the forward apply and the
// backward reduce walk descending, the backward apply ascending again.  RD_BN_REV=<bitmask> overrides
// (1 stats, 2 fwd apply, 4 bwd reduce, 8 bwd apply).
static int stream_rev(int which) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RD_BN_REV");
    v = e ? atoi(e) : 6;
  }
  return (v >> which) & 1;
}

template <typename K>
static int stream_prepare(K kernel, size_t smem) {
  RD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return 0;
}
