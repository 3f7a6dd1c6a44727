// Diagnostic micro-benchmark (not part of the library): issue rate of tcgen05.mma kind::f16 on sm_100a
// as a function of tile shape, operand source and shared-memory layout.  One CTA per SM, one issuing
// thread, no TMA and no epilogue -- what it prints is the cost of the MMA itself (operand fetch
// included).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I rangedet_b200/csrc
//                         scripts/mma_bench.cu -o scripts/mma_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_common.cuh"

struct Cfg {
  int M, N;
  int swz;      // 1: SW128 K-major tiles (rows 128 B), 0: no-swizzle canonical (8x16B core matrices)
  int a_tmem;   // A operand from TMEM (.ts form)
  int nd;       // accumulators cycled through (1: one dependent chain)
  int nbuf;     // distinct operand buffers cycled through (power of two)
  int kslices;  // K=16 slices used per buffer (4 = the whole 64-wide tile)
  int iters;
  int a_off_rows;  // extra start offset of A in pixel rows (128 B) -- row-shifted views as in conv strip mode
  const char* name;
  int style;  // 0: `if (t == 0)` around the loop; 1: elect.sync branch around the loop; 2: converged warp, MMA
              // predicated on elect.sync inside the asm; 3: converged warp, `if (elect_one()) mma` per MMA
};

__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mma_ss_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) bench_kernel(Cfg c, long long* out) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int t = threadIdx.x;
  // operands: nbuf A tiles (M+8 rows x 128 B) then nbuf B tiles (N rows x 128 B), each 1024-aligned
  const uint32_t a_bytes = ((c.M + 8) * 128 + 1023) & ~1023u, b_bytes = (c.N * 128 + 1023) & ~1023u;
  unsigned char* sA = base;
  unsigned char* sB = base + c.nbuf * a_bytes;
  for (uint32_t i = t; i < (c.nbuf * (a_bytes + b_bytes)) / 4; i += 128)
    reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u + (i * 2654435761u & 0x00ff00ffu);  // small bf16 values
  if (t == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  if (t < 32) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = slot;
  const bool go = c.style == 0 ? (t == 0) : c.style == 1 ? (t < 32 && elect_one()) : (t < 32);
  if (go) {
    const uint32_t idesc = tc::make_idesc_bf16(c.M, c.N);
    const uint64_t hi = c.swz ? tc::make_smem_desc(0, 0, 1024, tc::LAYOUT_SW128)
                              : tc::make_smem_desc(0, 0, 0, tc::LAYOUT_NONE);
    // no-swizzle canonical: chunk-major (LBO = rows*16 between K chunks, SBO = 128 between 8-row groups)
    const uint64_t hiA = c.swz ? hi : tc::make_smem_desc(0, c.M * 16, 128, tc::LAYOUT_NONE);
    const uint64_t hiB = c.swz ? hi : tc::make_smem_desc(0, c.N * 16, 128, tc::LAYOUT_NONE);
    const uint32_t kstep = c.swz ? 2u : (uint32_t)(2 * c.M * 16) >> 4;   // address units (16 B) per K=16 slice
    const uint32_t kstepB = c.swz ? 2u : (uint32_t)(2 * c.N * 16) >> 4;
    const uint32_t a_lo = (tc::smem_u32(sA) >> 4) + c.a_off_rows * 8, b_lo = tc::smem_u32(sB) >> 4;
    const uint32_t a_tm = tmem + 256;  // A-in-TMEM: 8 columns per K=16 slice (contents irrelevant)
    const long long t0 = clock64();
    uint32_t i = 0;
    for (int it = 0; it < c.iters; ++it) {
      const uint32_t buf = it & (c.nbuf - 1);
      for (int ks = 0; ks < c.kslices; ++ks, ++i) {
        const uint32_t d = tmem + (i & (c.nd - 1)) * c.N;
        const uint64_t bd = hiB | (uint64_t)((b_lo + buf * (b_bytes >> 4) + ks * kstepB) & 0x3FFF);
        if (c.a_tmem) {
          mma_ts(d, a_tm + ks * 8, bd, idesc, 1u);
        } else {
          const uint64_t ad = hiA | (uint64_t)((a_lo + buf * (a_bytes >> 4) + ks * kstep) & 0x3FFF);
          if (c.style == 2) mma_ss_elect(d, ad, bd, idesc, 1u);
          else if (c.style == 3) { if (elect_one()) tc::mma_bf16_ss(d, ad, bd, idesc, 1u); __syncwarp(); }
          else tc::mma_bf16_ss(d, ad, bd, idesc, 1u);
        }
      }
    }
    if (c.style < 2 || elect_one()) tc::umma_commit(&bar);
    if (c.style >= 2) __syncwarp();
    tc::mbar_wait(&bar, 0);
    if (c.style < 2 || t == 0) out[blockIdx.x] = clock64() - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (t < 32) tc::tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
  int grid = argc > 1 ? atoi(argv[1]) : 148;
  setvbuf(stdout, nullptr, _IONBF, 0);
  const Cfg cfgs[] = {
      {128, 64, 1, 0, 1, 4, 4, 512, 0, "SS M128 N64  sw128"},
      {128, 128, 1, 0, 1, 4, 4, 512, 0, "SS M128 N128 sw128"},
      {128, 256, 1, 0, 1, 4, 4, 512, 0, "SS M128 N256 sw128"},
      {64, 256, 1, 0, 1, 4, 4, 512, 0, "SS M64  N256 sw128"},
      {128, 64, 1, 0, 2, 4, 4, 512, 0, "SS M128 N64  sw128 2 accumulators"},
      {128, 64, 1, 0, 4, 4, 4, 512, 0, "SS M128 N64  sw128 4 accumulators"},
      {128, 64, 1, 0, 1, 1, 4, 512, 0, "SS M128 N64  sw128 one buffer"},
      {128, 64, 1, 0, 1, 1, 1, 2048, 0, "SS M128 N64  sw128 same slice"},
      {128, 64, 1, 0, 1, 4, 4, 512, 1, "SS M128 N64  sw128 A shifted 1 row"},
      {128, 128, 1, 0, 1, 4, 4, 512, 2, "SS M128 N128 sw128 A shifted 2 rows"},
      {128, 64, 0, 0, 1, 4, 4, 512, 0, "SS M128 N64  no-swizzle"},
      {128, 128, 0, 0, 1, 4, 4, 512, 0, "SS M128 N128 no-swizzle"},
      {128, 256, 0, 0, 1, 2, 4, 512, 0, "SS M128 N256 no-swizzle"},
      {128, 64, 1, 1, 1, 4, 4, 512, 0, "TS M128 N64  sw128 (A in TMEM)"},
      {128, 128, 1, 1, 1, 4, 4, 512, 0, "TS M128 N128 sw128 (A in TMEM)"},
      {128, 256, 1, 1, 1, 4, 4, 512, 0, "TS M128 N256 sw128 (A in TMEM)"},
      {128, 64, 1, 0, 1, 4, 4, 512, 0, "SS M128 N64  style1 elect branch", 1},
      {128, 64, 1, 0, 1, 4, 4, 512, 0, "SS M128 N64  style2 converged @elect", 2},
      {128, 64, 1, 0, 1, 4, 4, 512, 0, "SS M128 N64  style3 converged if(elect)", 3},
      {128, 128, 1, 0, 1, 4, 4, 512, 0, "SS M128 N128 style2", 2},
      {128, 256, 1, 0, 1, 4, 4, 512, 0, "SS M128 N256 style2", 2},
      {128, 128, 1, 0, 1, 4, 4, 512, 0, "SS M128 N128 style3", 3},
      {128, 256, 1, 0, 1, 4, 4, 512, 0, "SS M128 N256 style3", 3},
      {128, 64, 0, 0, 1, 4, 4, 512, 0, "SS M128 N64  no-swizzle style2", 2},
      {128, 32, 1, 0, 1, 4, 4, 512, 0, "SS M128 N32  sw128"},
      {128, 16, 1, 0, 1, 4, 4, 512, 0, "SS M128 N16  sw128"},
  };
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 1024);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  printf("grid=%d\n%-42s %10s %10s %8s\n", grid, "config", "cyc/MMA", "floor", "ratio");
  for (const Cfg& c : cfgs) {
    const size_t smem = 1024 + (size_t)c.nbuf * ((((c.M + 8) * 128 + 1023) & ~1023) + ((c.N * 128 + 1023) & ~1023));
    if (smem > 220 * 1024) { printf("%-42s skipped (smem)\n", c.name); continue; }
    for (int rep = 0; rep < 2; ++rep) bench_kernel<<<grid, 128, smem>>>(c, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-42s ERROR %s\n", c.name, cudaGetErrorString(e)); return 1; }
    long long h[1024];
    cudaMemcpy(h, d_out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double mean = 0, mx = 0;
    for (int b = 0; b < grid; ++b) { mean += (double)h[b] / grid; if (h[b] > mx) mx = (double)h[b]; }
    const double n = (double)c.iters * c.kslices;
    const double floor = (c.M > 128 ? c.M : 128) * c.N / 256.0;
    printf("%-42s %10.1f %10.1f %8.2f   (max block %.1f)\n", c.name, mean / n, floor, mean / n / floor, mx / n);
  }
  return 0;
}
