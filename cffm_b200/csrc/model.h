// Internal state behind a cffm_handle: parameter registry, device buffers, workspaces.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/cffm.h"
#include "kernels.h"

namespace cffm {

enum ParamKind { PK_DENSE = 0, PK_TABLE_INNER = 1, PK_TABLE_OUTER = 2, PK_TABLE_BIAS = 3 };

struct ParamInfo {
  std::string name;
  int64_t shape[4];
  int ndim;
  int64_t numel;
  int kind;       // ParamKind
  int64_t offset; // into the flat dense block (PK_DENSE only)
  bool trainable; // receives a gradient (false: dead variables)
};

constexpr int kMaxConv = 8;

// Offsets (in floats) into the flat dense parameter / gradient / accumulator blocks.
struct DenseLayout {
  int64_t iconv_w = -1, iconv_b = -1, din_k = -1, din_b = -1;  // inner path
  int64_t conv_w[kMaxConv], conv_b[kMaxConv];                  // outer conv stack
  int64_t outer_W = -1, outer_b = -1;                          // dead
  int64_t d1_k = -1, d1_b = -1, d2_k = -1, d2_b = -1;          // outer head
  int64_t att_W = -1, att_b = -1, d3_k = -1, d3_b = -1;        // linear attention
  int64_t bias = -1;
  int64_t total = 0;
};

// tables and ids a step computes on: the handle's own tables, or (row-sharded tables) the rows received from their
// owners with the ids renumbered to positions in that list
struct TableView { const float *inner = nullptr, *outer = nullptr, *fbias = nullptr; const int32_t* ids = nullptr; };

struct Comm;        // NCCL state (comm.cu)
struct ShardState;  // row-sharded tables: exchange buffers and per-step counts (shard.cu)

struct Model {
  cffm_config cfg;
  int F, P, Ki, Ko, M;
  // row-sharded tables (cfg.shard_world > 1): this handle stores rows r with r % shard_world == shard_rank as local
  // row r / shard_world; Mloc rows here, at most Mloc_max on any rank.  Replicated: Mloc == Mloc_max == M.
  int shard_world = 1, shard_rank = 0;
  int64_t Mloc = 0, Mloc_max = 0;
  ShardState* shard = nullptr;
  TableView view;        // of the last forward pass (the backward pass reads the same rows)
  int conv_depth;  // int(log2 Ko), CFFM.py:373
  int n_live;      // conv layers that reach the output = conv_depth - 1
  int t1_dim;      // sum_{l<conv_depth} Ko >> l   (= 2K-2)
  int max_batch;
  int device;
  cudaStream_t stream = nullptr;
  // Side streams, forked from and joined to the step's stream -- inside a stream capture parallel branches of the step
  // graph.  [0]: the inner path + linear term (and, on one GPU, the sort of the step's ids), which do not depend on the
  // outer path until the head / the update; [1]: on small steps the weight gradients of the conv stack, which nothing
  // waits for until the dense update, beside the chain of data gradients.  At the reference's dataset shapes (a few dozen CTAs per
  // kernel) the branches really overlap.  CFFM_SIDE_STREAM=0 keeps everything on one stream.
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork[2] = {nullptr, nullptr}, ev_join[2] = {nullptr, nullptr};
  std::string err;
  int64_t launches = 0;

  std::vector<ParamInfo> params;
  DenseLayout lay;

  // parameters + Adagrad accumulators
  float *inner_tab = nullptr, *outer_tab = nullptr, *fbias_tab = nullptr;
  float *inner_acc = nullptr, *outer_acc = nullptr, *fbias_acc = nullptr;
  float *dense_w = nullptr, *dense_acc = nullptr, *dense_g = nullptr;
  // second optimizer slot (Adam v) and the row -> segment map of dense table passes (Adam / lamda > 0)
  float *inner_acc2 = nullptr, *outer_acc2 = nullptr, *fbias_acc2 = nullptr, *dense_acc2 = nullptr;
  int32_t* rowmap = nullptr;
  float* sumsq_partial = nullptr;   // [512]

  // pair tables
  int *pair_i = nullptr, *pair_j = nullptr;  // [P]

  // ---- forward workspaces (sized by max_batch) ----
  int32_t* ids_buf = nullptr;   // [B,F] staging for host entry points (device)
  float* labels_buf = nullptr;  // [B]
  float* outer_rows = nullptr;  // [B,F,Ko] gathered outer rows
  float* Y[kMaxConv] = {};      // pre-activation conv outputs, NHWC [B,H_l+1,H_l+1,P]
  float* t1 = nullptr;          // [B,t1_dim]
  float* hid = nullptr;         // [B,32]
  float* comp_inner = nullptr;  // [B] final2
  float* comp_outer = nullptr;  // [B] final (after beta)
  float* comp_lin = nullptr;    // [B]
  float* out = nullptr;         // [B] raw sum (pre-sigmoid)
  float* pred = nullptr;        // [B] what sess.run(self.out) returns
  float* loss_terms = nullptr;  // [B]
  float* scalars = nullptr;     // [16] device scalars: 0 loss sum, 1 loss, 2 gscale, 3 global B ...
  float* loss_out = nullptr;    // [1]

  // ---- backward workspaces (allocated on first train step) ----
  bool train_ready = false;
  float* gout = nullptr;          // [B] dLoss/dout_raw
  float* dY[kMaxConv] = {};       // gradients w.r.t. Y_l
  float* g_inner_rows = nullptr;  // [B,F,Ki]
  float* g_outer_rows = nullptr;  // [B,F,Ko]
  float* g_bias_rows = nullptr;   // [B,F]
  float* v_head = nullptr;        // [t1_dim] W1.W2
  float* rowbuf = nullptr;        // [B, n_small] per-sample rows to be column-summed
  int n_small = 0;
  float* partials = nullptr;      // scratch for split reductions
  int64_t partials_cap = 0;
  int64_t aux_off = 0;            // start of the aux sums region inside dense_g: q[t1_dim], G, rowsums[n_small]
  void* reduce_descs = nullptr;   // device array of ReduceDesc
  int n_reduce_descs = 0;
  float* fb_buf = nullptr;        // [B,F] gathered feature_bias

  // ---- sparse update workspaces ----
  int64_t upd_cap = 0;            // capacity in ids (world * max_batch * F)
  int32_t *all_ids = nullptr;     // [upd_cap] ids of the global batch (== ids when world==1)
  float *all_g_inner = nullptr, *all_g_outer = nullptr, *all_g_bias = nullptr;  // gathered grads (DP)
  SparseWork sw;                  // sort / segment scratch

  // ---- host staging (pinned) ----
  int32_t* h_ids[2] = {nullptr, nullptr};
  float* h_labels[2] = {nullptr, nullptr};
  float* h_loss[2] = {nullptr, nullptr};
  float* h_out = nullptr;
  cudaEvent_t slot_done[2] = {nullptr, nullptr};
  int pending = 0;   // submitted-but-unreported steps (0..1)
  int slot = 0;

  // ---- CUDA graph cache for the train step (keyed by B) ----
  cudaGraphExec_t step_graph = nullptr;
  int64_t step_graph_B = -1;
  int64_t step_graph_launches = 0;

  Comm* comm = nullptr;
  int world = 1, rank = 0;
  void* tcs = nullptr;      // TCState: bf16 activations / operand copies of the tensor-core path (conv_tc.cu)
  // per-kernel CUDA-event timing (cffm_profile_enable / cffm_profile_report)
  bool prof_on = false;
  struct ProfEv { std::string tag; cudaEvent_t a, b; };
  std::vector<ProfEv> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  std::map<std::string, std::pair<int64_t, double>> prof_acc;  // tag -> (launches, total ms)
  int64_t last_B = 0;       // batch of the last forward / train step (for cffm_debug_fetch)
  bool use_graph = true;    // CFFM_GRAPH=0 disables CUDA-graph replay of the train step

  // evaluate scratch
  double* eval_acc = nullptr;  // [8] device doubles

  // resident training set (cffm_dataset_*)
  int32_t *ds_ids = nullptr, *ds_ids_tmp = nullptr;
  float *ds_labels = nullptr, *ds_labels_tmp = nullptr;
  int64_t* ds_perm = nullptr;
  int64_t ds_N = 0;
};

// the conv stack runs on the tcgen05 tensor cores (bf16 operands, or split bf16 = hi + lo for fp32-class accuracy)
inline bool tc_path(const Model* m) { return m->cfg.precision != CFFM_PREC_FP32; }

// ---- implemented across the .cu files ----
int model_build_layout(Model* m);
int model_alloc(Model* m);
int model_alloc_train(Model* m);
void model_free(Model* m);
int model_init_params(Model* m, uint64_t seed);
const ParamInfo* model_find(const Model* m, const char* name);

// labels == nullptr: scoring only; otherwise loss terms and dLoss/dout are produced as well
int run_forward(Model* m, const int32_t* ids_dev, const float* labels_dev, int64_t B, cudaStream_t s);
int run_backward_update(Model* m, const int32_t* ids_dev, const float* labels_dev, int64_t B, cudaStream_t s);

// tensor-core conv stack (CFFM_PREC_BF16), conv_tc.cu
int tc_supported(Model* m);
int tc_train_supported(Model* m);
int tc_alloc(Model* m, bool train);
void tc_free(Model* m);
int tc_conv_forward(Model* m, int B, cudaStream_t s);
int tc_conv_backward(Model* m, int B, cudaStream_t s);
int tc_debug_fetch(Model* m, bool grad, int l, float* dev_out, int64_t rows);

inline bool sharded(const Model* m) { return m->shard_world > 1; }
int shard_forward_exchange(Model* m, const int32_t* ids_dev, int64_t B, cudaStream_t s, TableView* view);
int shard_backward_update(Model* m, int64_t B, cudaStream_t s);
void shard_free(Model* m);

int comm_allreduce_f32(Model* m, float* buf, int64_t n, cudaStream_t s);
int comm_send(Model* m, const void* buf, int64_t bytes, int peer, cudaStream_t s);
int comm_recv(Model* m, void* buf, int64_t bytes, int peer, cudaStream_t s);
int comm_allgather(Model* m, const void* send, void* recv, int64_t bytes_per_rank, cudaStream_t s);
int comm_group_begin(Model* m);
int comm_group_end(Model* m);
void comm_destroy(Model* m);

// side_fork: the side stream, ordered after everything enqueued on `s` so far (or `s` itself when there is none);
// side_join: `s` continues after everything enqueued on the side stream
// Small steps (layer 0's gradient tensor within 48 MB: the reference's dataset shapes) are a chain of kernels of a few
// dozen CTAs: there the tensor-core kernels of one step are put on parallel branches too.  Big steps keep them on one
// stream: the persistent kernels count on all their CTAs running together (L2 sharing between neighbouring CTAs), and two
// of them competing for the SMs measured 59.8 -> 70.7 ms at the Criteo shape.
inline bool small_step(const Model* m, int64_t B) {
  const int64_t H = m->Ko / 2, Pp = (m->P + 63) & ~63;
  return B * H * H * Pp * 2 <= ((int64_t)48 << 20);
}
cudaStream_t side_fork(Model* m, cudaStream_t s, int which = 0);
int side_join(Model* m, cudaStream_t side, cudaStream_t s, int which = 0);

// Brackets one launch with CUDA events on its stream when profiling is enabled.
struct ProfScope {
  Model* m; cudaStream_t s; int idx;
  ProfScope(Model* m_, const char* tag, cudaStream_t s_);
  ~ProfScope();
};
#define CFFM_PROF_CAT2(a, b) a##b
#define CFFM_PROF_CAT(a, b) CFFM_PROF_CAT2(a, b)
#define CFFM_PROF(m, tag, s) ProfScope CFFM_PROF_CAT(_prof_scope_, __LINE__)((m), (tag), (s))

#define CFFM_CUDA_OK(m, call)                                                                 \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      (m)->err = std::string(#call) + ": " + cudaGetErrorString(_e);                          \
      return CFFM_ERR_CUDA;                                                                   \
    }                                                                                         \
  } while (0)

}  // namespace cffm

struct cffm_handle { cffm::Model m; };
