// rd_tc_probe_gemm: a one-CTA tcgen05 GEMM used by the tests to validate, on hardware, the
// descriptor encodings (tc_common.cuh) that the fused Meta-Kernel and conv kernels build on.
//   D(128 x n) = A(128 x k) . B(n x k)^T,  bf16 operands, fp32 accumulate in TMEM.
#include "../../include/rangedet_b200.h"
#include "rd_common.cuh"
#include "tc_common.cuh"

namespace {

__global__ void __launch_bounds__(128, 1)
tc_probe_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d, int n,
                int k, uint32_t tmem_cols, int mn_major) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base_slot;
  const int t = threadIdx.x, warp = t >> 5;
  const int kchunks = k / 8;  // 16-byte chunks along K
  // canonical no-swizzle K-major layout: chunk kc of row r at  kc * (rows*16) + r * 16
  unsigned char* sa = smem;                  // 128 rows
  unsigned char* sb = smem + kchunks * 128 * 16;  // n rows
  if (mn_major) {
    // MN-major canonical layout: 16-byte rows hold 8 consecutive M (or N) elements of one k;
    // element (mn, kk) at (mn/8)*SBO + (kk/8)*128 + (kk%8)*16 + (mn%8)*2, SBO = (k/8)*128
    const int sbo = (k / 8) * 128;
    for (int e = t; e < 16 * k; e += 128) {  // (mn block 0..15, kk)
      const int kk = e % k, mb = e / k;
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) w[q] = tc::pack_bf16x2(a[(mb * 8 + 2 * q) * k + kk], a[(mb * 8 + 2 * q + 1) * k + kk]);
      *reinterpret_cast<uint4*>(sa + mb * sbo + (kk / 8) * 128 + (kk % 8) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    for (int e = t; e < (n / 8) * k; e += 128) {
      const int kk = e % k, nb = e / k;
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) w[q] = tc::pack_bf16x2(b[(nb * 8 + 2 * q) * k + kk], b[(nb * 8 + 2 * q + 1) * k + kk]);
      *reinterpret_cast<uint4*>(sb + nb * sbo + (kk / 8) * 128 + (kk % 8) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  } else {
  for (int e = t; e < 128 * kchunks; e += 128) {
    const int r = e % 128, kc = e / 128;
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      w[q] = tc::pack_bf16x2(a[r * k + kc * 8 + 2 * q], a[r * k + kc * 8 + 2 * q + 1]);
    *reinterpret_cast<uint4*>(sa + kc * 128 * 16 + r * 16) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (int e = t; e < n * kchunks; e += 128) {
    const int r = e % n, kc = e / n;
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      w[q] = tc::pack_bf16x2(b[r * k + kc * 8 + 2 * q], b[r * k + kc * 8 + 2 * q + 1]);
    *reinterpret_cast<uint4*>(sb + kc * n * 16 + r * 16) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  }
  if (t == 0) {
    tc::mbar_init(&mbar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) {
    tc::tmem_alloc(&tmem_base_slot, tmem_cols);
    tc::tmem_relinquish();
  }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (t == 0) {
    const uint32_t idesc = tc::make_idesc_bf16(128, n, mn_major, mn_major);
    const uint32_t a_base = tc::smem_u32(sa), b_base = tc::smem_u32(sb);
    for (int ks = 0; ks < k / 16; ++ks) {
      uint64_t ad, bd;
      if (mn_major) {  // LBO = K-block (8 k) stride, SBO = MN-block (8 m/n) stride; a K step is two K blocks
        ad = tc::make_smem_desc(a_base + ks * 256, 128, (k / 8) * 128, tc::LAYOUT_NONE);
        bd = tc::make_smem_desc(b_base + ks * 256, 128, (k / 8) * 128, tc::LAYOUT_NONE);
      } else {
        ad = tc::make_smem_desc(a_base + ks * 2 * (128 * 16), 128 * 16, 128, tc::LAYOUT_NONE);
        bd = tc::make_smem_desc(b_base + ks * 2 * (n * 16), n * 16, 128, tc::LAYOUT_NONE);
      }
      tc::mma_bf16_ss(tmem_base, ad, bd, idesc, ks > 0 ? 1u : 0u);
    }
    tc::umma_commit(&mbar);
  }
  tc::mbar_wait(&mbar, 0);
  tc::tc_fence_after();
  const int row = warp * 32 + (t & 31);
  for (int c0 = 0; c0 < n; c0 += 16) {
    float v[16];
    tc::tmem_ld_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) d[row * n + c0 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace

extern "C" int rd_tc_probe_gemm(const float* a, const float* b, float* d, int n, int k, int mn_major,
                                rd_stream_t stream) {
  RD_REQUIRE(a && b && d, "rd_tc_probe_gemm: null pointer");
  RD_REQUIRE(k > 0 && k % 16 == 0 && k <= 128, "rd_tc_probe_gemm: k must be a multiple of 16, <= 128");
  RD_REQUIRE(n >= 16 && n % 16 == 0 && n <= 256, "rd_tc_probe_gemm: n must be a multiple of 16 in [16,256]");
  if (rd_check_device()) return 1;
  uint32_t cols = 32;
  while (cols < (uint32_t)n) cols <<= 1;
  const size_t smem = (size_t)(128 + n) * k * 2 + 1024;
  RD_CUDA(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_probe_kernel<<<1, 128, smem, rd::as_stream(stream)>>>(a, b, d, n, k, cols, mn_major ? 1 : 0);
  rd::count_launch();
  return rd::check_launch("rd_tc_probe_gemm");
}
