// Optimizer update (CFFM.py:517-529) [TF-1.14]:
//  - IndexedSlices gradients of the three gathered tables are de-duplicated by a deterministic
//    (hand-written, stable radix) sort-by-row + segmented sum (summation in order of appearance, like unique +
//    unsorted_segment_sum) and applied with SparseApplyAdagrad to the touched rows only;
//  - every other trainable variable gets the dense ApplyAdagrad.  acc0 = 1e-8, no epsilon (Q11).
#include <string>

#include "common.cuh"
#include "kernels.h"

namespace cffm {

__global__ void k_iota(int32_t* v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (int32_t)i;
}

// ---------------------------------------------------------------------------------------------
// Stable LSD radix sort of (row id, source position) pairs, 8 bits per pass, hand-written:
//   k_rs_hist    per-tile digit histograms
//   k_rs_scan    exclusive scan in (digit, tile) order -> first output slot of every (tile, digit)
//   k_rs_scatter stable scatter: a tile is walked in rounds of 256 elements; inside a round the rank of
//                an element among equal digits is popc(match_any & lanes below) + counts of lower warps
// Equal ids keep their order of appearance, which is what makes the segmented sum below reproduce
// unique + unsorted_segment_sum's summation order.
constexpr int RS_THREADS = 256, RS_ITEMS = 8, RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int SEG_LT = 32;           // entries per chunk of the segmented reduction (one warp each: short chunks = more warps in flight)
constexpr int SEG_SHORT_LIST = 32768; // up to this many entries: one warp per segment (k_seg_rows)
constexpr int SEG_MAXC = 2;          // columns per lane and table (K <= 64)
constexpr int SEG_LONG = 64;         // short lists: segments longer than this are summed by a whole block (k_seg_long)

__global__ void k_rs_hist(const int32_t* __restrict__ keys, int n, int shift, int32_t* __restrict__ hist) {
  __shared__ int32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int t0 = blockIdx.x * RS_TILE;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int i = t0 + r * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255], 1);
  }
  __syncthreads();
  hist[blockIdx.x * 256 + threadIdx.x] = h[threadIdx.x];
}

__global__ void k_rs_scan(int32_t* __restrict__ hist, int ntiles) {  // in place: counts -> first output slots
  __shared__ int32_t tot[256];
  const int d = threadIdx.x;
  int s = 0;
  for (int b = 0; b < ntiles; ++b) s += hist[b * 256 + d];
  tot[d] = s;
  __syncthreads();
  if (d == 0) { int run = 0; for (int q = 0; q < 256; ++q) { const int c = tot[q]; tot[q] = run; run += c; } }
  __syncthreads();
  int run = tot[d];
  for (int b = 0; b < ntiles; ++b) { const int c = hist[b * 256 + d]; hist[b * 256 + d] = run; run += c; }
}

__global__ void k_rs_scatter(const int32_t* __restrict__ keys, const int32_t* __restrict__ vals, int n, int shift,
                             const int32_t* __restrict__ offs, int32_t* __restrict__ keys_out, int32_t* __restrict__ vals_out) {
  __shared__ int32_t base[256];
  __shared__ int32_t wcnt[RS_THREADS / 32][256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  base[tid] = offs[blockIdx.x * 256 + tid];
  const int t0 = blockIdx.x * RS_TILE;
  for (int r = 0; r < RS_ITEMS; ++r) {
#pragma unroll
    for (int w8 = 0; w8 < RS_THREADS / 32; ++w8) wcnt[w8][tid] = 0;
    __syncthreads();
    const int i = t0 + r * RS_THREADS + tid;
    const bool ok = i < n;
    const int32_t key = ok ? keys[i] : 0;
    const int d = ok ? ((key >> shift) & 255) : 256;            // 256: tail lanes, never written
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    if (ok && rank == 0) wcnt[warp][d] = __popc(peers);
    __syncthreads();
    if (ok) {
      int prior = 0;
      for (int w8 = 0; w8 < warp; ++w8) prior += wcnt[w8][d];
      const int dst = base[d] + prior + rank;
      keys_out[dst] = key;
      vals_out[dst] = vals[i];
    }
    __syncthreads();
    {
      int add = 0;
#pragma unroll
      for (int w8 = 0; w8 < RS_THREADS / 32; ++w8) add += wcnt[w8][tid];
      base[tid] += add;
    }
    __syncthreads();
  }
}

// ---- unique rows: head flags -> exclusive scan -> first sorted position of every segment ----
__global__ void k_seg_count(const int32_t* __restrict__ sorted, int n, int32_t* __restrict__ tile_heads) {
  __shared__ int32_t cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const int t0 = blockIdx.x * RS_TILE;
  int c = 0;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int i = t0 + r * RS_THREADS + threadIdx.x;
    if (i < n && (i == 0 || sorted[i] != sorted[i - 1])) ++c;
  }
  c = (int)warp_sum((float)c);   // <= 2048 per tile: exact in fp32
  if ((threadIdx.x & 31) == 0) atomicAdd(&cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) tile_heads[blockIdx.x] = cnt;
}
__global__ void k_seg_scan(int32_t* __restrict__ tile_heads, int ntiles, int32_t* __restrict__ n_uniq) {  // one thread: ntiles is small
  int run = 0;
  for (int b = 0; b < ntiles; ++b) { const int c = tile_heads[b]; tile_heads[b] = run; run += c; }
  *n_uniq = run;
}
__global__ void k_seg_compact(const int32_t* __restrict__ sorted, int n, const int32_t* __restrict__ tile_off,
                              int32_t* __restrict__ seg_start) {
  __shared__ int32_t wtot[RS_THREADS / 32];
  __shared__ int32_t run;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) run = tile_off[blockIdx.x];
  __syncthreads();
  const int t0 = blockIdx.x * RS_TILE;
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int i = t0 + r * RS_THREADS + tid;
    const bool head = i < n && (i == 0 || sorted[i] != sorted[i - 1]);
    const unsigned m = __ballot_sync(0xffffffffu, head);
    if (lane == 0) wtot[warp] = __popc(m);
    __syncthreads();
    int prior = run;
    for (int w8 = 0; w8 < warp; ++w8) prior += wtot[w8];
    if (head) seg_start[prior + __popc(m & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w8 = 0; w8 < RS_THREADS / 32; ++w8) t += wtot[w8]; run += t; }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Short lists (the reference's own datasets: Frappe 2 560 ids per step, ml-tag 3 072, Book-Crossing 3 072): the whole
// sort + unique in ONE block instead of 3 launches per radix pass + 3 -- at these sizes the twelve launches cost more
// than the work.  Same algorithm (stable LSD radix, 8 bits per pass, ranks by match_any), keys and positions ping-pong
// in shared memory; then head flags -> scan -> seg_start / n_uniq.  Bit-identical output to the multi-kernel path.
constexpr int SS_MAX = 4096, SS_THREADS = 1024;
constexpr int SS_SMEM = (4 * SS_MAX + 256 + (SS_THREADS / 32) * 256 + 64) * 4;
__global__ void __launch_bounds__(SS_THREADS, 1) k_small_sort_segments(const int32_t* __restrict__ ids, int n, int passes,
                                                                       int32_t* __restrict__ keys_out, int32_t* __restrict__ vals_out,
                                                                       int32_t* __restrict__ seg_start, int32_t* __restrict__ n_uniq) {
  extern __shared__ int32_t ssm[];
  int32_t* kA = ssm; int32_t* vA = kA + SS_MAX; int32_t* kB = vA + SS_MAX; int32_t* vB = kB + SS_MAX;
  int32_t* base = vB + SS_MAX;                       // [256] next output slot of every digit
  int32_t* wcnt = base + 256;                        // [32 warps][256]
  int32_t* wtot = wcnt + (SS_THREADS / 32) * 256;    // [32] + running total
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < n; i += SS_THREADS) { kA[i] = ids[i]; vA[i] = i; }
  __syncthreads();
  for (int p = 0; p < passes; ++p) {
    const int shift = 8 * p;
    if (tid < 256) base[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += SS_THREADS) atomicAdd(&base[(kA[i] >> shift) & 255], 1);
    __syncthreads();
    if (warp == 0) {                                  // exclusive scan of the 256 digit counts: 8 per lane
      int c[8], s = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = base[lane * 8 + j]; s += c[j]; }
      int incl = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      int run = incl - s;
#pragma unroll
      for (int j = 0; j < 8; ++j) { base[lane * 8 + j] = run; run += c[j]; }
    }
    __syncthreads();
    for (int r0 = 0; r0 < n; r0 += SS_THREADS) {      // stable scatter, one round of 1024 elements at a time
      for (int e = tid; e < (SS_THREADS / 32) * 256; e += SS_THREADS) wcnt[e] = 0;
      __syncthreads();
      const int i = r0 + tid;
      const bool ok = i < n;
      const int32_t key = ok ? kA[i] : 0;
      const int d = ok ? ((key >> shift) & 255) : 256;
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      const int rank = __popc(peers & ((1u << lane) - 1u));
      if (ok && rank == 0) wcnt[warp * 256 + d] = __popc(peers);
      __syncthreads();
      if (ok) {
        int prior = 0;
        for (int w8 = 0; w8 < warp; ++w8) prior += wcnt[w8 * 256 + d];
        const int dst = base[d] + prior + rank;
        kB[dst] = key; vB[dst] = vA[i];
      }
      __syncthreads();
      if (tid < 256) {
        int add = 0;
#pragma unroll 8
        for (int w8 = 0; w8 < SS_THREADS / 32; ++w8) add += wcnt[w8 * 256 + tid];
        base[tid] += add;
      }
      __syncthreads();
    }
    int32_t* t0 = kA; kA = kB; kB = t0; t0 = vA; vA = vB; vB = t0;
  }
  for (int i = tid; i < n; i += SS_THREADS) { keys_out[i] = kA[i]; vals_out[i] = vA[i]; }
  // unique rows: head flags, block scan in rounds of 1024
  if (tid == 0) wtot[32] = 0;
  __syncthreads();
  for (int r0 = 0; r0 < n; r0 += SS_THREADS) {
    const int i = r0 + tid;
    const bool head = i < n && (i == 0 || kA[i] != kA[i - 1]);
    const unsigned mk = __ballot_sync(0xffffffffu, head);
    if (lane == 0) wtot[warp] = __popc(mk);
    __syncthreads();
    int prior = wtot[32];
    for (int w8 = 0; w8 < warp; ++w8) prior += wtot[w8];
    if (head) seg_start[prior + __popc(mk & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w8 = 0; w8 < SS_THREADS / 32; ++w8) t += wtot[w8]; wtot[32] += t; }
    __syncthreads();
  }
  if (tid == 0) *n_uniq = wtot[32];
}

int sparse_work_alloc(SparseWork* w, int64_t cap, int gcols, std::string* err) {
  w->cap = cap;
  const int64_t ntiles = (cap + RS_TILE - 1) / RS_TILE;
  w->cub_tmp_bytes = sizeof(int32_t) * (size_t)(ntiles * 256 + ntiles + 16);   // digit offsets + per-tile head counts
  cudaError_t e;
#define SW_ALLOC(ptr, bytes)                                                        \
  e = dev_malloc((void**)&(ptr), (bytes));                                           \
  if (e != cudaSuccess) { if (err) *err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return CFFM_ERR_NOMEM; }
  SW_ALLOC(w->keys_out, sizeof(int32_t) * cap);
  SW_ALLOC(w->vals, sizeof(int32_t) * cap);
  SW_ALLOC(w->vals_out, sizeof(int32_t) * cap);
  SW_ALLOC(w->seg_start, sizeof(int32_t) * (cap + 1));
  SW_ALLOC(w->n_uniq, sizeof(int32_t) * 4);
  SW_ALLOC(w->flags, sizeof(int32_t) * 2 * cap);   // ping-pong buffers of the sort: keys | values
  SW_ALLOC(w->cub_tmp, w->cub_tmp_bytes);
  {
    const int64_t chunks = (cap + SEG_LT - 1) / SEG_LT + 1;
    SW_ALLOC(w->pieces, sizeof(float) * chunks * 2 * 3 * SEG_MAXC * 32);
    if (gcols > 0) SW_ALLOC(w->gsum, sizeof(float) * cap * gcols);   // per-unique-row sums of the three tables
    SW_ALLOC(w->chunk_flags, sizeof(int32_t) * chunks);
  }
#undef SW_ALLOC
  // positions 0..cap-1 never change
  k_iota<<<(int)((cap + 255) / 256), 256>>>(w->vals, cap);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { if (err) *err = std::string("k_iota: ") + cudaGetErrorString(e); return CFFM_ERR_CUDA; }
  return CFFM_OK;
}

void sparse_work_free(SparseWork* w) {
  void* p[] = {w->keys_out, w->vals, w->vals_out, w->seg_start, w->n_uniq, w->flags, w->cub_tmp, w->pieces, w->chunk_flags, w->gsum};
  for (void* q : p) if (q) dev_free(q);
  *w = SparseWork();
}

int sparse_sort_segments(SparseWork* w, const int32_t* ids, int64_t n64, int features_M, cudaStream_t s, int64_t* launches) {
  if (n64 > w->cap || n64 < 1) return CFFM_ERR_INVALID;
  const int n = (int)n64;
  int bits = 1;
  while (bits < 31 && (1ll << bits) < (long long)features_M) ++bits;
  const int passes = (bits + 7) / 8;
  if (n <= SS_MAX) {
    static PerDeviceOnce attr_once;
    bool& attr_done = attr_once();
    if (!attr_done) {
      if (cudaFuncSetAttribute(k_small_sort_segments, cudaFuncAttributeMaxDynamicSharedMemorySize, SS_SMEM) != cudaSuccess) return CFFM_ERR_CUDA;
      attr_done = true;
    }
    k_small_sort_segments<<<1, SS_THREADS, SS_SMEM, s>>>(ids, n, passes, w->keys_out, w->vals_out, w->seg_start, w->n_uniq);
    if (launches) *launches += 1;
    return cudaGetLastError() == cudaSuccess ? CFFM_OK : CFFM_ERR_CUDA;
  }
  const int ntiles = (n + RS_TILE - 1) / RS_TILE;
  int32_t* offs = reinterpret_cast<int32_t*>(w->cub_tmp);
  int32_t* tile_heads = offs + (size_t)ntiles * 256;
  int32_t* tmp_k = reinterpret_cast<int32_t*>(w->flags);
  int32_t* tmp_v = tmp_k + w->cap;
  // ping-pong so that the last pass lands in keys_out / vals_out
  const int32_t* src_k = ids; const int32_t* src_v = w->vals;
  for (int p = 0; p < passes; ++p) {
    const bool to_out = ((passes - 1 - p) % 2) == 0;
    int32_t* dst_k = to_out ? w->keys_out : tmp_k;
    int32_t* dst_v = to_out ? w->vals_out : tmp_v;
    k_rs_hist<<<ntiles, RS_THREADS, 0, s>>>(src_k, n, 8 * p, offs);
    k_rs_scan<<<1, 256, 0, s>>>(offs, ntiles);
    k_rs_scatter<<<ntiles, RS_THREADS, 0, s>>>(src_k, src_v, n, 8 * p, offs, dst_k, dst_v);
    src_k = dst_k; src_v = dst_v;
  }
  k_seg_count<<<ntiles, RS_THREADS, 0, s>>>(w->keys_out, n, tile_heads);
  k_seg_scan<<<1, 1, 0, s>>>(tile_heads, ntiles, w->n_uniq);
  k_seg_compact<<<ntiles, RS_THREADS, 0, s>>>(w->keys_out, n, tile_heads, w->seg_start);
  if (launches) *launches += 3 * passes + 3;
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

// IndexedSlices de-duplication + SparseApply* on the touched rows only, in two phases.
//
// Phase 1 (segment sums): the gradient rows of every unique id are added up -- in a summation order that is a fixed
//   function of the sorted list, so the update is deterministic -- into a compact buffer G[segment][column] per table:
//   * lists up to SEG_SHORT_LIST entries: one warp per segment, rows in order of appearance; a segment longer than
//     SEG_LONG entries (fields with two or three values put one id into most samples of a batch) is cut into eight
//     consecutive pieces summed by the eight warps of a block, then the pieces are added in order;
//   * longer lists: the sorted list is cut into chunks of SEG_LT entries, one warp per chunk adds the rows of every
//     segment piece inside its chunk in order of appearance; pieces of segments that cross chunk borders go to scratch
//     and k_seg_sums_fixup adds them up, chunk after chunk, from the chunk the segment starts in.
// Phase 2 (apply): one warp per unique row reads its sum and the row's variable / slots, applies the optimizer, writes
//   back.  No warp ever waits for a table row while it still has gradient rows to add.
// What the split buys: the positions and keys of a chunk come from one coalesced load (shuffled to the lanes that need
// them), eight gradient rows are in flight per lane, and a chunk is 1 + 4 dependent memory round trips instead of 2 per
// four entries plus 2 per segment end (ncu, Criteo shape: the single-phase kernel spent 70 % of its samples on
// long_scoreboard with 8.6 % of DRAM bandwidth in use).
template <int MAXC> struct SegAccT { float v[3][MAXC]; };
struct GSums { float* g[3]; };   // per table: [segment][K]

template <int MAXC> __device__ __forceinline__ void sa_zero(SegAccT<MAXC>& a) {
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int c = 0; c < MAXC; ++c) a.v[j][c] = 0.f;
}
template <int MAXC> __device__ __forceinline__ void sa_row(SegAccT<MAXC>& a, const SparseTables& t, int64_t p, int lane) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (!t.tab[j] || t.dense[j]) continue;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int k = lane + 32 * c;
      if (k < t.K[j]) a.v[j][c] = __ldg(t.grads[j] + p * t.K[j] + k);
    }
  }
}
template <int MAXC> __device__ __forceinline__ void sa_add(SegAccT<MAXC>& a, const SegAccT<MAXC>& b) {
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int c = 0; c < MAXC; ++c) a.v[j][c] += b.v[j][c];
}
template <int MAXC> __device__ __forceinline__ void sa_store_g(const SegAccT<MAXC>& a, const SparseTables& t, const GSums& G, int64_t sgm, int lane) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (!t.tab[j] || t.dense[j]) continue;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const int k = lane + 32 * c;
      if (k < t.K[j]) G.g[j][sgm * t.K[j] + k] = a.v[j][c];
    }
  }
}
// piece scratch: [chunk][slot 0/1][table][MAXC][32 lanes]
template <int MAXC> __device__ __forceinline__ float* piece_ptr(float* pieces, int chunk, int slot) { return pieces + ((int64_t)chunk * 2 + slot) * (3 * MAXC * 32); }
template <int MAXC> __device__ __forceinline__ void sa_store_p(const SegAccT<MAXC>& a, float* p, int lane) {
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int c = 0; c < MAXC; ++c) p[(j * MAXC + c) * 32 + lane] = a.v[j][c];
}
template <int MAXC> __device__ __forceinline__ void sa_add_p(SegAccT<MAXC>& a, const float* p, int lane) {
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int c = 0; c < MAXC; ++c) a.v[j][c] += p[(j * MAXC + c) * 32 + lane];
}
// index of the segment that contains sorted position t (warp-uniform)
__device__ __forceinline__ int seg_of(const int32_t* __restrict__ seg_start, int U, int t) {
  int lo = 0, hi = U - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (seg_start[mid] <= t) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// flags[chunk]: bit0 = slot 1 holds the open tail piece of a segment that starts in this chunk;
//               bit1 = slot 0 holds a middle piece (the segment covers the whole chunk and goes on)
template <int MAXC>
__global__ void k_seg_sums_chunks(const int32_t* __restrict__ sorted, const int32_t* __restrict__ pos, const int32_t* __restrict__ seg_start,
                                  const int32_t* __restrict__ n_uniq, int n, SparseTables t, GSums G, float* __restrict__ pieces,
                                  int32_t* __restrict__ flags) {
  static_assert(SEG_LT == 32, "one entry per lane");
  const int lane = threadIdx.x & 31;
  const int chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int t0 = chunk * SEG_LT;
  if (t0 >= n) return;
  const int t1 = min(n, t0 + SEG_LT);
  // one coalesced load brings the chunk's positions and keys; the lanes get them by shuffle
  const int my_pos = t0 + lane < t1 ? pos[t0 + lane] : 0;
  const int my_key = t0 + lane < t1 ? sorted[t0 + lane] : -1;
  const int key_after = t0 + SEG_LT < n ? sorted[t0 + SEG_LT] : -1;
  const bool head_open = t0 > 0 && sorted[t0 - 1] == sorted[t0];
  int sgm = seg_of(seg_start, *n_uniq, t0);
  int fl = 0;
  SegAccT<MAXC> acc; sa_zero(acc);
  bool first_piece = true;
  for (int tb = 0; tb < SEG_LT; tb += 8) {
    if (t0 + tb >= t1) break;
    SegAccT<MAXC> rows8[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int p = __shfl_sync(0xffffffffu, my_pos, tb + u);
      sa_zero(rows8[u]);
      if (t0 + tb + u < t1) sa_row(rows8[u], t, (int64_t)p, lane);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int key = __shfl_sync(0xffffffffu, my_key, tb + u);
      const int nxt_in = __shfl_sync(0xffffffffu, my_key, (tb + u + 1) & 31);
      const int tt = t0 + tb + u;
      if (tt < t1) {                                           // warp-uniform
        const int nxt = tb + u + 1 < SEG_LT ? nxt_in : key_after;
        sa_add(acc, rows8[u]);
        const bool last_in_chunk = tt + 1 == t1;
        if (last_in_chunk || nxt != key) {
          const bool closes = nxt != key;                        // the segment ends with this entry
          const bool started_inside = !(first_piece && head_open);
          if (started_inside && closes) sa_store_g(acc, t, G, (int64_t)sgm, lane);
          else if (!started_inside && closes) sa_store_p(acc, piece_ptr<MAXC>(pieces, chunk, 0), lane);
          else if (started_inside && !closes) { sa_store_p(acc, piece_ptr<MAXC>(pieces, chunk, 1), lane); fl |= 1; }
          else { sa_store_p(acc, piece_ptr<MAXC>(pieces, chunk, 0), lane); fl |= 2; }
          sa_zero(acc);
          first_piece = false;
          if (closes) ++sgm;
        }
      }
    }
  }
  if (lane == 0) flags[chunk] = fl;
}

template <int MAXC>
__global__ void k_seg_sums_fixup(const int32_t* __restrict__ seg_start, const int32_t* __restrict__ n_uniq, int n, SparseTables t, GSums G,
                                 const float* __restrict__ pieces, const int32_t* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const int chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int t0 = chunk * SEG_LT;
  if (t0 >= n) return;
  if (!(flags[chunk] & 1)) return;                              // no segment starts here and runs on
  const int t1 = min(n, t0 + SEG_LT);
  SegAccT<MAXC> acc; sa_zero(acc);
  sa_add_p(acc, piece_ptr<MAXC>(const_cast<float*>(pieces), chunk, 1), lane);
  for (int cc = chunk + 1;; ++cc) {                            // the pieces of the following chunks, in order
    sa_add_p(acc, piece_ptr<MAXC>(const_cast<float*>(pieces), cc, 0), lane);
    if (!(flags[cc] & 2)) break;                               // a first piece: the segment ends inside chunk cc
  }
  sa_store_g(acc, t, G, (int64_t)seg_of(seg_start, *n_uniq, t1 - 1), lane);
}

// Short lists (the small workloads: a few thousand entries): one warp per segment, all segments in parallel; rows are
// added in order of appearance.  Positions come 32 at a time from one coalesced load.
template <int MAXC>
__global__ void k_seg_sums_rows(const int32_t* __restrict__ pos, const int32_t* __restrict__ seg_start, const int32_t* __restrict__ n_uniq,
                                int n, SparseTables t, GSums G) {
  const int lane = threadIdx.x & 31;
  const int U = *n_uniq;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int sgm = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; sgm < U; sgm += nwarps) {
    const int start = seg_start[sgm], end = sgm + 1 < U ? seg_start[sgm + 1] : n;
    if (end - start > SEG_LONG) continue;         // k_seg_sums_long: one id that fills a good part of a batch
    SegAccT<MAXC> acc; sa_zero(acc);
    for (int b0 = start; b0 < end; b0 += 32) {
      const int my_pos = b0 + lane < end ? pos[b0 + lane] : 0;
      for (int tb = 0; tb < 32 && b0 + tb < end; tb += 8) {
        SegAccT<MAXC> rows8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int p = __shfl_sync(0xffffffffu, my_pos, tb + u);
          sa_zero(rows8[u]);
          if (b0 + tb + u < end) sa_row(rows8[u], t, (int64_t)p, lane);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) sa_add(acc, rows8[u]);       // (rows beyond the segment are zero)
      }
    }
    sa_store_g(acc, t, G, (int64_t)sgm, lane);
  }
}

// Long segments of a short list: a block per segment, its eight warps sum eight consecutive pieces (order of
// appearance inside a piece), warp 0 adds the pieces in order.
template <int MAXC>
__global__ void k_seg_sums_long(const int32_t* __restrict__ pos, const int32_t* __restrict__ seg_start, const int32_t* __restrict__ n_uniq,
                                int n, SparseTables t, GSums G) {
  __shared__ float part[8][3 * MAXC * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int U = *n_uniq;
  for (int sgm = blockIdx.x; sgm < U; sgm += gridDim.x) {
    const int start = seg_start[sgm], end = sgm + 1 < U ? seg_start[sgm + 1] : n;
    if (end - start <= SEG_LONG) continue;        // block-uniform
    const int piece = (((end - start + 7) >> 3) + 7) & ~7;
    const int s0 = start + warp * piece, s1 = min(end, s0 + piece);
    SegAccT<MAXC> acc; sa_zero(acc);
    for (int b0 = s0; b0 < s1; b0 += 32) {
      const int my_pos = b0 + lane < s1 ? pos[b0 + lane] : 0;
      for (int tb = 0; tb < 32 && b0 + tb < s1; tb += 8) {
        SegAccT<MAXC> rows8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int p = __shfl_sync(0xffffffffu, my_pos, tb + u);
          sa_zero(rows8[u]);
          if (b0 + tb + u < s1) sa_row(rows8[u], t, (int64_t)p, lane);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) sa_add(acc, rows8[u]);
      }
    }
    sa_store_p(acc, part[warp], lane);
    __syncthreads();
    if (warp == 0) {
      SegAccT<MAXC> tot; sa_zero(tot);
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) sa_add_p(tot, part[w8], lane);
      sa_store_g(tot, t, G, (int64_t)sgm, lane);
    }
    __syncthreads();
  }
}

// Phase 2: SparseApply* on the unique rows, one warp per row.
template <int MAXC>
__global__ void k_apply_rows(const int32_t* __restrict__ sorted, const int32_t* __restrict__ seg_start, const int32_t* __restrict__ n_uniq,
                             SparseTables t, GSums G, int opt, float lr, const float* __restrict__ lr_dev) {
  const int lane = threadIdx.x & 31;
  const int U = *n_uniq;
  if (lr_dev) lr = *lr_dev;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int sgm = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; sgm < U; sgm += nwarps) {
    const int64_t row = sorted[seg_start[sgm]];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (!t.tab[j] || t.dense[j]) continue;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        const int k = lane + 32 * c;
        if (k >= t.K[j]) continue;
        const int64_t o = row * t.K[j] + k;
        const float g = G.g[j][(int64_t)sgm * t.K[j] + k];
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

static GSums gsums_of(const SparseWork* w, const SparseTables& t) {
  GSums G; int64_t off = 0;
  for (int j = 0; j < 3; ++j) { G.g[j] = w->gsum ? w->gsum + off : nullptr; off += w->cap * (int64_t)(t.tab[j] ? t.K[j] : 0); }
  return G;
}

// Phase 1 alone: per-unique-row sums of the gradient rows into w->gsum (table j at offset cap * sum_{i<j} K_i, row stride
// K_j, one row per segment of the sorted list).  Also used by the row-sharded exchange, which ships these sums.
template <int MAXC>
static void segment_sums_t(const SparseWork* w, const SparseTables& t, int64_t n, cudaStream_t s, int64_t* launches) {
  const GSums G = gsums_of(w, t);
  if (n <= SEG_SHORT_LIST) {
    int blocks = (int)((n * 32 + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
    k_seg_sums_rows<MAXC><<<blocks, 256, 0, s>>>(w->vals_out, w->seg_start, w->n_uniq, (int)n, t, G);
    int lb = (int)((n + SEG_LONG - 1) / SEG_LONG); if (lb > 148 * 4) lb = 148 * 4;   // at most n / SEG_LONG long segments exist
    if (n > SEG_LONG) { k_seg_sums_long<MAXC><<<lb, 256, 0, s>>>(w->vals_out, w->seg_start, w->n_uniq, (int)n, t, G); if (launches) *launches += 1; }
    if (launches) *launches += 1;
  } else {
    const int chunks = (int)((n + SEG_LT - 1) / SEG_LT);
    const int blocks = (chunks * 32 + 255) / 256;
    k_seg_sums_chunks<MAXC><<<blocks, 256, 0, s>>>(w->keys_out, w->vals_out, w->seg_start, w->n_uniq, (int)n, t, G, w->pieces, w->chunk_flags);
    k_seg_sums_fixup<MAXC><<<blocks, 256, 0, s>>>(w->seg_start, w->n_uniq, (int)n, t, G, w->pieces, w->chunk_flags);
    if (launches) *launches += 2;
  }
}
static int maxc_of(const SparseTables& t) {
  int k = 1;
  for (int j = 0; j < 3; ++j) if (t.tab[j] && !t.dense[j] && t.K[j] > k) k = t.K[j];
  return k > 32 ? 2 : 1;
}
void launch_segment_sums(const SparseWork* w, const SparseTables& t, int64_t n, cudaStream_t s, int64_t* launches) {
  if (n <= 0) return;
  if (maxc_of(t) == 2) segment_sums_t<2>(w, t, n, s, launches); else segment_sums_t<1>(w, t, n, s, launches);
}

void launch_sparse_update(const SparseWork* w, const SparseTables& t, int64_t n, int opt, float lr, const float* lr_dev,
                          cudaStream_t s, int64_t* launches) {
  if (n <= 0) return;
  bool any_sparse = false, any_dense = false;
  for (int j = 0; j < 3; ++j) if (t.tab[j]) { if (t.dense[j]) any_dense = true; else any_sparse = true; }
  if (any_sparse) {
    launch_segment_sums(w, t, n, s, launches);
    const GSums G = gsums_of(w, t);
    int blocks = (int)((n * 32 + 255) / 256); if (blocks > 148 * 64) blocks = 148 * 64;   // a row per warp: many short chains
    if (maxc_of(t) == 2) k_apply_rows<2><<<blocks, 256, 0, s>>>(w->keys_out, w->seg_start, w->n_uniq, t, G, opt, lr, lr_dev);
    else k_apply_rows<1><<<blocks, 256, 0, s>>>(w->keys_out, w->seg_start, w->n_uniq, t, G, opt, lr, lr_dev);
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
