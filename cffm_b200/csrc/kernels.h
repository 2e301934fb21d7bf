// Cross-file launcher declarations.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace cffm {
struct Model;

void launch_gather_rows(const float* table, const int32_t* ids, int64_t n, int K, float* out, cudaStream_t s);
int forward_setup_attrs(Model* m);
int backward_setup_attrs(Model* m);

// loss reduction pieces (forward.cu)
void launch_loss_sum(Model* m, int B, cudaStream_t s);
void launch_loss_finish(Model* m, int B, cudaStream_t s);

// sparse update (update.cu)
struct SparseWork {
  int64_t cap = 0;
  int32_t *keys_out = nullptr, *vals = nullptr, *vals_out = nullptr, *seg_start = nullptr, *n_uniq = nullptr;
  uint8_t* flags = nullptr;       // ping-pong buffers of the radix sort
  void* cub_tmp = nullptr;        // digit offsets + per-tile segment-head counts
  float* pieces = nullptr;        // partial sums of segment pieces that cross chunk borders
  float* gsum = nullptr;          // [cap * gcols] per-unique-row gradient sums (phase 1 of the update), table after table
  int32_t* chunk_flags = nullptr;
  size_t cub_tmp_bytes = 0;
};
// gcols = total columns of the tables whose gradient rows will be summed (Ki + Ko + 1)
int sparse_work_alloc(SparseWork* w, int64_t cap, int gcols, std::string* err);
void sparse_work_free(SparseWork* w);
// sort ids, find segments; afterwards w->keys_out (sorted ids), w->vals_out (source positions),
// w->seg_start (first sorted position of each unique row) and w->n_uniq are valid on the stream.
int sparse_sort_segments(SparseWork* w, const int32_t* ids, int64_t n, int features_M, cudaStream_t s, int64_t* launches);
// segment-sum the gradient rows in order of appearance and apply SparseApplyAdagrad to up to three tables
struct SparseTables {
  float *tab[3] = {nullptr, nullptr, nullptr};
  float *acc[3] = {nullptr, nullptr, nullptr};    // slot 1 (Adagrad accumulator / momentum / Adam m); may be null
  float *acc2[3] = {nullptr, nullptr, nullptr};   // slot 2 (Adam v); may be null
  const float* grads[3] = {nullptr, nullptr, nullptr};
  int K[3] = {0, 0, 0};
  bool dense[3] = {false, false, false};           // whole-table pass (l2 regulariser / Adam)
  float reg[3] = {0.f, 0.f, 0.f};                  // regulariser strength of the dense pass
  int32_t* rowmap = nullptr;                       // [M], -1 = untouched (dense passes only)
  int64_t M = 0;
};
// phase 1 only: sums per unique row into w->gsum (table j at cap * sum_{i<j} K_i, row stride K_j)
void launch_segment_sums(const SparseWork* w, const SparseTables& t, int64_t n, cudaStream_t s, int64_t* launches);
void launch_sparse_update(const SparseWork* w, const SparseTables& t, int64_t n, int opt, float lr, const float* lr_dev,
                          cudaStream_t s, int64_t* launches);
void launch_dense_update(float* w, float* s1, float* s2, const float* g, int64_t n, int opt, float lr, const float* lr_dev,
                         cudaStream_t s);
void launch_adam_tick(float* scalars, float lr, cudaStream_t s);
void launch_sumsq(const float* x, int64_t n, float* partial512, float* out, cudaStream_t s);

}  // namespace cffm
