// rd_tma_probe: one-CTA TMA self-test.  Loads the box (box_w, 1, 64) of a (W, H, C) fp32 tensor at
// signed coordinates (c0, c1, c2) into shared memory (zero fill outside the tensor) and copies it to
// `dst` (64 x box_w floats) with plain stores; then stores the same tile back through a TMA store
// into `dst2` viewed as the same (W, H, C) tensor.  Validates tensor-map encoding, alignment rules,
// out-of-bound semantics and the mbarrier transaction count the fused kernels rely on.
#include <stdio.h>

#include "../../include/rangedet_b200.h"
#include "rd_common.cuh"
#include "tma_common.cuh"

namespace {

__global__ void __launch_bounds__(128, 1)
tma_probe_kernel(const __grid_constant__ CUtensorMap tm_src, const __grid_constant__ CUtensorMap tm_dst,
                 float* __restrict__ dst, int box_w, int c0, int c1, int c2, int verbose) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ uint64_t bar;
  const uint32_t a0 = tc::smem_u32(smem_raw);
  unsigned char* base = smem_raw + ((1024u - (a0 & 1023u)) & 1023u);
  float* tile = reinterpret_cast<float*>(base);
  const int t = threadIdx.x;
  if (t == 0) {
    if (verbose) printf("tma_probe: dyn smem base 0x%x, tile 0x%x, bar 0x%x\n", a0, tc::smem_u32(tile), tc::smem_u32(&bar));
    tc::mbar_init(&bar, 1);
    tc::fence_mbar_init();
  }
  __syncthreads();
  if (t == 0) {
    tc::mbar_arrive_expect_tx(&bar, (uint32_t)(64 * box_w * 4));
    tma::load_3d(tile, &tm_src, &bar, c0, c1, c2);
  }
  tc::mbar_wait(&bar, 0);
  for (int e = t; e < 64 * box_w; e += 128) dst[e] = tile[e];
  tc::fence_proxy_async_smem();
  __syncthreads();
  if (t == 0) {
    tma::store_3d(&tm_dst, tile, c0, c1, c2);
    tma::store_commit();
    tma::store_wait_all<0>();
  }
}

}  // namespace

extern "C" int rd_tma_probe(const float* src, float* dst, float* dst2, int W, int H, int C, int box_w, int c0,
                            int c1, int c2, rd_stream_t stream) {
  RD_REQUIRE(src && dst && dst2, "rd_tma_probe: null pointer");
  RD_REQUIRE(W % 4 == 0 && box_w % 4 == 0 && box_w >= 4 && box_w <= 256 && C >= 64, "rd_tma_probe: bad shape");
  if (rd_check_device()) return 1;
  CUtensorMap ms, md;
  const uint64_t dims[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)C};
  const uint64_t strides[2] = {(uint64_t)W * 4, (uint64_t)W * H * 4};
  const uint32_t box[3] = {(uint32_t)box_w, 1u, 64u};
  if (tma::make_map(&ms, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, src, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  if (tma::make_map(&md, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, dst2, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  const size_t smem = (size_t)64 * box_w * 4 + 1024;
  RD_CUDA(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const char* v = getenv("RD_TMA_PROBE_VERBOSE");
  tma_probe_kernel<<<1, 128, smem, rd::as_stream(stream)>>>(ms, md, dst, box_w, c0, c1, c2, v ? 1 : 0);
  rd::count_launch();
  return rd::check_launch("rd_tma_probe");
}
