// Diagnostic micro-benchmark #2 (not part of the library): how cheaply can ONE warp issue
// tcgen05.mma?  scripts/mma_bench.cu showed ~240 cycles per MMA independent of N, i.e. the issuing
// thread's own instruction latency is the bound.  Here the loop is specialised at compile time and
// the issue style varies.   Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I rangedet_b200/csrc
//                                        scripts/mma_bench2.cu -o scripts/mma_bench2
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_common.cuh"

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mma_ss_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void mma_ss_acc(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
  asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc),
               "r"(idesc)
               : "memory");
}

// STYLE 0: one thread (`t == 0`) runs the loop.  1: whole warp converged, elect.sync inside the asm.
// 2: whole warp converged, `if (elect_one())` around each group of 4 MMAs.  3: whole warp converged, leader
// elected once, `if (leader)` around each group.  NW = number of issuing warps (each its own accumulator).
template <int STYLE, int N, int NW>
__global__ void __launch_bounds__(128, 1) bench_kernel(int iters, long long* out) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  const int t = threadIdx.x, warp = t >> 5;
  constexpr uint32_t A_BYTES = 17408, B_BYTES = N * 128, NBUF = 4;
  for (uint32_t i = t; i < NBUF * (A_BYTES + B_BYTES) / 4; i += 128)
    reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u + (i * 2654435761u & 0x00ff00ffu);
  if (t == 0) { for (int i = 0; i < 4; ++i) tc::mbar_init(&bar[i], 1); tc::fence_mbar_init(); }
  if (t < 32) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = slot;
  if (warp < NW && (STYLE != 0 || (t & 31) == 0)) {
    constexpr uint32_t idesc = tc::make_idesc_bf16(128, N);
    const uint64_t hi = tc::make_smem_desc(0, 0, 1024, tc::LAYOUT_SW128);
    const uint32_t a_lo = tc::smem_u32(base) >> 4, b_lo = (tc::smem_u32(base) + NBUF * A_BYTES) >> 4;
    const uint32_t d = tmem + warp * N;
    bool leader = true;
    if (STYLE == 3) leader = elect_one();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
      const uint32_t buf = it & (NBUF - 1);
      const uint64_t ad = hi | (uint64_t)(a_lo + buf * (A_BYTES >> 4));
      const uint64_t bd = hi | (uint64_t)(b_lo + buf * (B_BYTES >> 4));
      if (STYLE == 1) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_ss_elect(d, ad + 2 * ks, bd + 2 * ks, idesc);
      } else if (STYLE == 2) {
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) mma_ss_acc(d, ad + 2 * ks, bd + 2 * ks, idesc);
        }
        __syncwarp();
      } else if (STYLE == 3) {
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) mma_ss_acc(d, ad + 2 * ks, bd + 2 * ks, idesc);
        }
        __syncwarp();
      } else {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_ss_acc(d, ad + 2 * ks, bd + 2 * ks, idesc);
      }
    }
    if (STYLE == 0 || elect_one()) tc::umma_commit(&bar[warp]);
    if (STYLE != 0) __syncwarp();
    tc::mbar_wait(&bar[warp], 0);
    if ((t & 31) == 0) out[blockIdx.x * 4 + warp] = clock64() - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (t < 32) tc::tmem_dealloc(tmem, 512);
}

template <int STYLE, int N, int NW>
static void run(const char* name, long long* d_out) {
  const int iters = 1024, grid = 148;
  const size_t smem = 1024 + 4 * (17408 + N * 128);
  cudaFuncSetAttribute(bench_kernel<STYLE, N, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) bench_kernel<STYLE, N, NW><<<grid, 128, smem>>>(iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-40s ERROR %s\n", name, cudaGetErrorString(e)); exit(1); }
  static long long h[148 * 4];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int b = 0; b < grid; ++b) mean += (double)h[b * 4] / grid;
  const double per = mean / (iters * 4.0) / NW;  // CTA-level cycles per MMA
  printf("%-40s N=%3d warps=%d  %7.1f cyc/MMA (floor %5.1f, x%.2f)\n", name, N, NW, per, N / 2.0, per / (N / 2.0));
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 148 * 4);
  run<0, 64, 1>("one thread", d_out);
  run<1, 64, 1>("converged, elect in asm", d_out);
  run<2, 64, 1>("converged, if(elect_one()) per group", d_out);
  run<3, 64, 1>("converged, if(leader) per group", d_out);
  run<0, 128, 1>("one thread", d_out);
  run<2, 128, 1>("converged, if(elect_one()) per group", d_out);
  run<3, 128, 1>("converged, if(leader) per group", d_out);
  run<0, 256, 1>("one thread", d_out);
  run<3, 256, 1>("converged, if(leader) per group", d_out);
  run<0, 64, 2>("one thread per warp, 2 warps", d_out);
  run<0, 64, 4>("one thread per warp, 4 warps", d_out);
  run<3, 64, 2>("if(leader), 2 warps", d_out);
  run<3, 64, 4>("if(leader), 4 warps", d_out);
  run<0, 128, 2>("one thread per warp, 2 warps", d_out);
  run<3, 128, 2>("if(leader), 2 warps", d_out);
  return 0;
}
