// Optimizer update (CFFM.py:517-529) [TF-1.14]:
//  - IndexedSlices gradients of the three gathered tables are de-duplicated by a deterministic
//    sort-by-row + segmented sum (summation in order of appearance, like unique +
//    unsorted_segment_sum) and applied with SparseApplyAdagrad to the touched rows only;
//  - every other trainable variable gets the dense ApplyAdagrad.  acc0 = 1e-8, no epsilon (Q11).
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <string>

#include "common.cuh"
#include "kernels.h"

namespace cffm {

__global__ void k_iota(int32_t* v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (int32_t)i;
}

__global__ void k_head_flags(const int32_t* __restrict__ sorted, int64_t n, uint8_t* __restrict__ flags) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = (i == 0 || sorted[i] != sorted[i - 1]) ? 1 : 0;
}

static size_t cub_bytes(int64_t cap) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)cap, 0, 32);
  thrust::counting_iterator<int32_t> it(0);
  cub::DeviceSelect::Flagged(nullptr, b, it, (const uint8_t*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr, (int)cap);
  return (a > b ? a : b) + 256;
}

int sparse_work_alloc(SparseWork* w, int64_t cap, std::string* err) {
  w->cap = cap;
  w->cub_tmp_bytes = cub_bytes(cap);
  cudaError_t e;
#define SW_ALLOC(ptr, bytes)                                                        \
  e = cudaMalloc((void**)&(ptr), (bytes));                                           \
  if (e != cudaSuccess) { if (err) *err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return CFFM_ERR_NOMEM; }
  SW_ALLOC(w->keys_out, sizeof(int32_t) * cap);
  SW_ALLOC(w->vals, sizeof(int32_t) * cap);
  SW_ALLOC(w->vals_out, sizeof(int32_t) * cap);
  SW_ALLOC(w->seg_start, sizeof(int32_t) * (cap + 1));
  SW_ALLOC(w->n_uniq, sizeof(int32_t) * 4);
  SW_ALLOC(w->flags, cap);
  SW_ALLOC(w->cub_tmp, w->cub_tmp_bytes);
#undef SW_ALLOC
  // positions 0..cap-1 never change
  k_iota<<<(int)((cap + 255) / 256), 256>>>(w->vals, cap);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { if (err) *err = std::string("k_iota: ") + cudaGetErrorString(e); return CFFM_ERR_CUDA; }
  return CFFM_OK;
}

void sparse_work_free(SparseWork* w) {
  void* p[] = {w->keys_out, w->vals, w->vals_out, w->seg_start, w->n_uniq, w->flags, w->cub_tmp};
  for (void* q : p) if (q) cudaFree(q);
  *w = SparseWork();
}

int sparse_sort_segments(SparseWork* w, const int32_t* ids, int64_t n, int features_M, cudaStream_t s, int64_t* launches) {
  if (n > w->cap) return CFFM_ERR_INVALID;
  int bits = 1;
  while (bits < 31 && (1ll << bits) < (long long)features_M) ++bits;
  size_t tmp = w->cub_tmp_bytes;
  // LSD radix sort is stable: equal ids keep their order of appearance.
  cub::DeviceRadixSort::SortPairs(w->cub_tmp, tmp, ids, w->keys_out, w->vals, w->vals_out, (int)n, 0, bits, s);
  k_head_flags<<<(int)((n + 255) / 256), 256, 0, s>>>(w->keys_out, n, w->flags);
  thrust::counting_iterator<int32_t> it(0);
  tmp = w->cub_tmp_bytes;
  cub::DeviceSelect::Flagged(w->cub_tmp, tmp, it, w->flags, w->seg_start, w->n_uniq, (int)n, s);
  if (launches) *launches += 5;
  return cudaGetLastError() == cudaSuccess ? CFFM_OK : CFFM_ERR_CUDA;
}

// One warp per unique row.  Lane k owns column k (and k+32 for K = 64); gradient rows of the
// segment are added one after the other in order of appearance, four loads in flight.
__device__ __forceinline__ void seg_update(float* __restrict__ tab, float* __restrict__ acc,
                                           const float* __restrict__ grads, int K, int64_t row,
                                           const int32_t* __restrict__ pos, int start, int end, float lr, int lane) {
  for (int k = lane; k < K; k += 32) {
    float g = 0.f;
    int t = start;
    for (; t + 4 <= end; t += 4) {
      const float g0 = __ldg(grads + (int64_t)__ldg(pos + t) * K + k);
      const float g1 = __ldg(grads + (int64_t)__ldg(pos + t + 1) * K + k);
      const float g2 = __ldg(grads + (int64_t)__ldg(pos + t + 2) * K + k);
      const float g3 = __ldg(grads + (int64_t)__ldg(pos + t + 3) * K + k);
      g += g0; g += g1; g += g2; g += g3;
    }
    for (; t < end; ++t) g += __ldg(grads + (int64_t)__ldg(pos + t) * K + k);
    const int64_t o = row * K + k;
    const float a = acc[o] + g * g;
    acc[o] = a;
    tab[o] -= lr * g * __frsqrt_rn(a);
  }
}

__global__ void k_sparse_adagrad(const int32_t* __restrict__ sorted_ids, const int32_t* __restrict__ pos,
                                 const int32_t* __restrict__ seg_start, const int32_t* __restrict__ n_uniq, int n,
                                 SparseTables t, float lr) {
  const int lane = threadIdx.x & 31;
  const int U = *n_uniq;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int sgm = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; sgm < U; sgm += nwarps) {
    const int start = seg_start[sgm];
    const int end = (sgm + 1 < U) ? seg_start[sgm + 1] : n;
    const int64_t row = sorted_ids[start];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if (t.tab[j]) seg_update(t.tab[j], t.acc[j], t.grads[j], t.K[j], row, pos, start, end, lr, lane);
  }
}

void launch_sparse_adagrad(const SparseWork* w, const SparseTables& t, int64_t n, float lr, cudaStream_t s, int64_t* launches) {
  if (n <= 0) return;
  int64_t warps = n;  // upper bound on unique rows
  int blocks = (int)((warps * 32 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_sparse_adagrad<<<blocks, 256, 0, s>>>(w->keys_out, w->vals_out, w->seg_start, w->n_uniq, (int)n, t, lr);
  if (launches) *launches += 1;
}

__global__ void k_dense_adagrad(float* __restrict__ w, float* __restrict__ acc, const float* __restrict__ g, int64_t n, float lr) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const float gi = g[i];
    const float a = acc[i] + gi * gi;
    acc[i] = a;
    w[i] -= lr * gi * __frsqrt_rn(a);
  }
}

void launch_dense_adagrad(float* w, float* acc, const float* g, int64_t n, float lr, cudaStream_t s) {
  if (n <= 0) return;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_dense_adagrad<<<blocks, 256, 0, s>>>(w, acc, g, n, lr);
}

}  // namespace cffm
