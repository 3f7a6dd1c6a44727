// get_sorted_foreground on the GPU: masked top-K by score, descending, + gather.
//
// Replaces the Python CustomOp GetSortedFGOperator.forward, /root/reference
// operator_py/get_sorted_foreground.py:11-40 (call site rangedet/symbol/head/builder.py:512-521):
//   score = cls_score * mask                      (:20)
//   topk(score, k = num_fgs) then argsort desc    (:23,33-34)   -> order of the K best points
//   gather bbox_delta (8) / pc (3) / score        (:35-37)
// The reference does this with nd.topk + a Python loop over the batch + three fancy-index gathers.
// Here: ONE stable radix sort over 64-bit keys (row << 32 | ~ordered(score)) for the whole batch --
// rows stay contiguous, scores descend within a row, equal scores keep ascending point index (the
// order of a stable argsort; MXNet's tie order is unspecified) -- and one gather kernel.
// HBM-bound integer/byte work: 4+4 B/point in, 16 B/point of sort traffic per pass, 48 B per kept point.
#include <cub/cub.cuh>
#include <stdint.h>

#include "../../include/rangedet_b200.h"
#include "rd_common.cuh"

namespace sfg {

constexpr int NT = 256;

__device__ __forceinline__ uint32_t ordered_desc(float f) {
  f += 0.0f;  // -0.0 (negative logit x zero mask) -> +0.0: the reference compares them equal
  uint32_t u = __float_as_uint(f);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending unsigned order == ascending float order
  return ~u;                                        // ... and descending after the complement
}

__global__ void __launch_bounds__(NT)
make_keys_kernel(const float* __restrict__ cls_score, const float* __restrict__ mask, uint64_t* __restrict__ keys,
                 int32_t* __restrict__ idx, float* __restrict__ masked, int64_t total, int n) {
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
    const float s = __fmul_rn(cls_score[i], mask[i]);  // get_sorted_foreground.py:20
    const int64_t row = i / n;
    keys[i] = ((uint64_t)row << 32) | ordered_desc(s);
    idx[i] = (int32_t)(i - row * n);
    masked[i] = s;
  }
}

// one thread per (kept point, float): 1 score + 8 delta + 3 pc = 12 floats per point
__global__ void __launch_bounds__(NT)
gather_kernel(const int32_t* __restrict__ order, const float* __restrict__ masked, const float* __restrict__ bbox_delta,
              const float* __restrict__ pc, float* __restrict__ out_score, float* __restrict__ out_delta,
              float* __restrict__ out_pc, int B, int n, int k) {
  const int64_t total = (int64_t)B * k * 12;
  for (int64_t t = (int64_t)blockIdx.x * NT + threadIdx.x; t < total; t += (int64_t)gridDim.x * NT) {
    const int f = (int)(t % 12);
    const int64_t p = t / 12;
    const int b = (int)(p / k), j = (int)(p % k);
    const int64_t src = (int64_t)b * n + order[(int64_t)b * n + j];
    if (f == 0) out_score[p] = masked[src];
    else if (f < 9) out_delta[p * 8 + (f - 1)] = bbox_delta[src * 8 + (f - 1)];
    else out_pc[p * 3 + (f - 9)] = pc[src * 3 + (f - 9)];
  }
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static int row_bits(int B) {
  int b = 0;
  while ((1 << b) < B) ++b;
  return b;
}

}  // namespace sfg

extern "C" {

size_t rd_get_sorted_foreground_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  const size_t total = (size_t)B * N;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)total, 0, 32 + sfg::row_bits(B));
  return 2 * sfg::align256(total * 8) + 2 * sfg::align256(total * 4) + sfg::align256(total * 4) + sfg::align256(cub_bytes) + 256;
}

int rd_get_sorted_foreground(const float* cls_score, const float* bbox_delta, const float* pc, const float* mask, int B,
                             int N, int num_fgs, float* out_score, float* out_delta, float* out_pc, void* workspace,
                             size_t workspace_bytes, rd_stream_t stream) {
  RD_REQUIRE(cls_score && bbox_delta && pc && mask && out_score && out_delta && out_pc, "rd_get_sorted_foreground: null pointer");
  RD_REQUIRE(B > 0 && N > 0, "rd_get_sorted_foreground: bad shape");
  // infer_shape of the reference op: assert pc_shape[1] >= num_fgs (get_sorted_foreground.py:66)
  RD_REQUIRE(num_fgs > 0 && num_fgs <= N, "rd_get_sorted_foreground: num_fgs must be in [1, N] (got %d, N=%d)", num_fgs, N);
  RD_REQUIRE((int64_t)B * N <= 0x7fffffffLL, "rd_get_sorted_foreground: B*N too large");
  RD_REQUIRE(workspace && workspace_bytes >= rd_get_sorted_foreground_workspace_bytes(B, N),
             "rd_get_sorted_foreground: workspace too small");
  if (rd_check_device()) return 1;
  cudaStream_t st = rd::as_stream(stream);
  const size_t total = (size_t)B * N;
  char* w = static_cast<char*>(workspace);
  w += (256 - (reinterpret_cast<uintptr_t>(w) & 255)) & 255;
  uint64_t* keys_in = reinterpret_cast<uint64_t*>(w);   w += sfg::align256(total * 8);
  uint64_t* keys_out = reinterpret_cast<uint64_t*>(w);  w += sfg::align256(total * 8);
  int32_t* idx_in = reinterpret_cast<int32_t*>(w);      w += sfg::align256(total * 4);
  int32_t* idx_out = reinterpret_cast<int32_t*>(w);     w += sfg::align256(total * 4);
  float* masked = reinterpret_cast<float*>(w);          w += sfg::align256(total * 4);
  size_t cub_bytes = 0;
  const int end_bit = 32 + sfg::row_bits(B);
  RD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, keys_in, keys_out, idx_in, idx_out, (int)total, 0, end_bit, st));
  int dev = 0, sms = 0;
  RD_CUDA(cudaGetDevice(&dev));
  RD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid_cap = sms * 8;
  int grid = (int)((total + sfg::NT - 1) / sfg::NT);
  if (grid > grid_cap) grid = grid_cap;
  sfg::make_keys_kernel<<<grid, sfg::NT, 0, st>>>(cls_score, mask, keys_in, idx_in, masked, (int64_t)total, N);
  rd::count_launch();
  if (rd::check_launch("rd_get_sorted_foreground(keys)")) return 1;
  RD_CUDA(cub::DeviceRadixSort::SortPairs(w, cub_bytes, keys_in, keys_out, idx_in, idx_out, (int)total, 0, end_bit, st));
  rd::count_launch();
  const int64_t gt = (int64_t)B * num_fgs * 12;
  grid = (int)((gt + sfg::NT - 1) / sfg::NT);
  if (grid > grid_cap) grid = grid_cap;
  sfg::gather_kernel<<<grid, sfg::NT, 0, st>>>(idx_out, masked, bbox_delta, pc, out_score, out_delta, out_pc, B, N, num_fgs);
  rd::count_launch();
  return rd::check_launch("rd_get_sorted_foreground(gather)");
}

}  // extern "C"
