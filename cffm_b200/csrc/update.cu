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

// ---------------------------------------------------------------------------------------------
// Optimizers of CFFM.py:517-529 [TF-1.14].  s1 / s2 are the slots: Adagrad accumulator (init 1e-8,
// no epsilon, Q11); Momentum accumulator (0.95); Adam m and v (0.9 / 0.999 / 1e-8, step size lr_t).
__device__ __forceinline__ void opt_apply(int opt, float& w, float& s1, float& s2, float g, float lr) {
  switch (opt) {
    case CFFM_OPT_ADAGRAD: s1 += g * g; w -= lr * g * __frsqrt_rn(s1); break;
    case CFFM_OPT_SGD: w -= lr * g; break;
    case CFFM_OPT_MOMENTUM: s1 = 0.95f * s1 + g; w -= lr * s1; break;
    default:
      s1 = 0.9f * s1 + 0.1f * g;
      s2 = 0.999f * s2 + 0.001f * g * g;
      w -= lr * s1 / (sqrtf(s2) + 1e-8f);
      break;
  }
}

// sum of the gradient rows of one segment, column k, in order of appearance (four loads in flight)
__device__ __forceinline__ float seg_sum(const float* __restrict__ grads, int K, int k, const int32_t* __restrict__ pos,
                                         int start, int end) {
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
  return g;
}

// One warp per unique row: IndexedSlices de-duplication + SparseApply* on the touched rows only.
__global__ void k_sparse_update(const int32_t* __restrict__ sorted_ids, const int32_t* __restrict__ pos,
                                const int32_t* __restrict__ seg_start, const int32_t* __restrict__ n_uniq, int n,
                                SparseTables t, int opt, float lr, const float* __restrict__ lr_dev) {
  const int lane = threadIdx.x & 31;
  const int U = *n_uniq;
  if (lr_dev) lr = *lr_dev;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int sgm = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; sgm < U; sgm += nwarps) {
    const int start = seg_start[sgm];
    const int end = (sgm + 1 < U) ? seg_start[sgm + 1] : n;
    const int64_t row = sorted_ids[start];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (!t.tab[j] || t.dense[j]) continue;
      const int K = t.K[j];
      for (int k = lane; k < K; k += 32) {
        const float g = seg_sum(t.grads[j], K, k, pos, start, end);
        const int64_t o = row * K + k;
        float w = t.tab[j][o], s1 = t.acc[j] ? t.acc[j][o] : 0.f, s2 = t.acc2[j] ? t.acc2[j][o] : 0.f;
        opt_apply(opt, w, s1, s2, g, lr);
        t.tab[j][o] = w;
        if (t.acc[j]) t.acc[j][o] = s1;
        if (t.acc2[j]) t.acc2[j][o] = s2;
      }
    }
  }
}

// rowmap[row] = segment index for the touched rows (reset to -1 afterwards)
__global__ void k_rowmap_set(const int32_t* __restrict__ sorted_ids, const int32_t* __restrict__ seg_start,
                             const int32_t* __restrict__ n_uniq, int32_t* __restrict__ rowmap, int clear) {
  const int U = *n_uniq;
  for (int sgm = blockIdx.x * blockDim.x + threadIdx.x; sgm < U; sgm += gridDim.x * blockDim.x)
    rowmap[sorted_ids[seg_start[sgm]]] = clear ? -1 : sgm;
}

// Dense pass over a whole table: g = reg * w (+ the segment sum of the row if it was touched).  Used when the
// table gradient is dense: the l2 regulariser (lamda > 0, CFFM.py:489-491, Q9) or Adam's sparse apply, which
// decays m, v and moves every row [TF-1.14].  One warp per row.
__global__ void k_table_dense_update(float* __restrict__ tab, float* __restrict__ s1p, float* __restrict__ s2p,
                                     const float* __restrict__ grads, int K, int64_t M, const int32_t* __restrict__ rowmap,
                                     const int32_t* __restrict__ pos, const int32_t* __restrict__ seg_start,
                                     const int32_t* __restrict__ n_uniq, int n, float reg, int opt, float lr,
                                     const float* __restrict__ lr_dev) {
  const int lane = threadIdx.x & 31;
  const int U = *n_uniq;
  if (lr_dev) lr = *lr_dev;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < M; row += nwarps) {
    const int sgm = rowmap[row];
    int start = 0, end = 0;
    if (sgm >= 0) { start = seg_start[sgm]; end = (sgm + 1 < U) ? seg_start[sgm + 1] : n; }
    for (int k = lane; k < K; k += 32) {
      const int64_t o = row * K + k;
      float w = tab[o], s1 = s1p ? s1p[o] : 0.f, s2 = s2p ? s2p[o] : 0.f;
      float g = reg * w;
      if (sgm >= 0) g += seg_sum(grads, K, k, pos, start, end);
      opt_apply(opt, w, s1, s2, g, lr);
      tab[o] = w;
      if (s1p) s1p[o] = s1;
      if (s2p) s2p[o] = s2;
    }
  }
}

void launch_sparse_update(const SparseWork* w, const SparseTables& t, int64_t n, int opt, float lr, const float* lr_dev,
                          cudaStream_t s, int64_t* launches) {
  if (n <= 0) return;
  bool any_sparse = false, any_dense = false;
  for (int j = 0; j < 3; ++j) if (t.tab[j]) { if (t.dense[j]) any_dense = true; else any_sparse = true; }
  if (any_sparse) {
    int blocks = (int)((n * 32 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_sparse_update<<<blocks, 256, 0, s>>>(w->keys_out, w->vals_out, w->seg_start, w->n_uniq, (int)n, t, opt, lr, lr_dev);
    if (launches) *launches += 1;
  }
  if (any_dense) {
    int ub = (int)((n + 255) / 256); if (ub > 148 * 4) ub = 148 * 4;
    k_rowmap_set<<<ub, 256, 0, s>>>(w->keys_out, w->seg_start, w->n_uniq, t.rowmap, 0);
    for (int j = 0; j < 3; ++j) {
      if (!t.tab[j] || !t.dense[j]) continue;
      k_table_dense_update<<<148 * 8, 256, 0, s>>>(t.tab[j], t.acc[j], t.acc2[j], t.grads[j], t.K[j], t.M, t.rowmap, w->vals_out,
                                                    w->seg_start, w->n_uniq, (int)n, t.reg[j], opt, lr, lr_dev);
      if (launches) *launches += 1;
    }
    k_rowmap_set<<<ub, 256, 0, s>>>(w->keys_out, w->seg_start, w->n_uniq, t.rowmap, 1);
    if (launches) *launches += 2;
  }
}

__global__ void k_dense_update(float* __restrict__ w, float* __restrict__ s1p, float* __restrict__ s2p,
                               const float* __restrict__ g, int64_t n, int opt, float lr, const float* __restrict__ lr_dev) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (lr_dev) lr = *lr_dev;
  for (; i < n; i += stride) {
    float wi = w[i], s1 = s1p ? s1p[i] : 0.f, s2 = s2p ? s2p[i] : 0.f;
    opt_apply(opt, wi, s1, s2, g[i], lr);
    w[i] = wi;
    if (s1p) s1p[i] = s1;
    if (s2p) s2p[i] = s2;
  }
}

void launch_dense_update(float* w, float* s1, float* s2, const float* g, int64_t n, int opt, float lr, const float* lr_dev,
                         cudaStream_t s) {
  if (n <= 0) return;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_dense_update<<<blocks, 256, 0, s>>>(w, s1, s2, g, n, opt, lr, lr_dev);
}

// Adam: t += 1; lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) -> scalars[4]   (scalars[5] holds t)
__global__ void k_adam_tick(float* __restrict__ scalars, float lr) {
  const float t = scalars[5] + 1.f;
  scalars[5] = t;
  scalars[4] = lr * sqrtf(1.f - powf(0.999f, t)) / (1.f - powf(0.9f, t));
}
void launch_adam_tick(float* scalars, float lr, cudaStream_t s) { k_adam_tick<<<1, 1, 0, s>>>(scalars, lr); }

// sum of squares of a table (l2 regulariser term of the loss), deterministic two-stage reduction
__global__ void k_sumsq_partial(const float* __restrict__ x, int64_t n, float* __restrict__ partial) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s = fmaf(x[i], x[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}
__global__ void k_sumsq_final(const float* __restrict__ partial, int n, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) *out = v;
  }
}
void launch_sumsq(const float* x, int64_t n, float* partial512, float* out, cudaStream_t s) {
  k_sumsq_partial<<<512, 256, 0, s>>>(x, n, partial512);
  k_sumsq_final<<<1, 512, 0, s>>>(partial512, 512, out);
}

}  // namespace cffm
