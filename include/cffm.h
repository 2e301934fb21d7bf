/*
 * cffm.h -- C ABI of the B200-native CFFM hot path (libcffm_b200.so).
 *
 * The reference (Anony-CFFM/CFFM, TensorFlow 1.14 script) has no plugin / FFI interface; its
 * "operator boundary" for the hot path is the set of tf.Session.run fetches in CFFM.py plus the
 * LoadData object.  Each entry point below names the reference interface it replaces
 * (file:line relative to the reference tree).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success and a negative cffm_status on failure; the message is
 *     available from cffm_last_error(h) (h may be NULL for failures of cffm_create / loaders);
 *   - no C++ exception crosses this boundary;
 *   - the caller owns every buffer it passes; the library owns parameters, optimizer state and
 *     workspaces (device memory, allocated with cudaMalloc on cfg.device);
 *   - *_dev entry points take DEVICE pointers and are asynchronous on the given CUDA stream
 *     (results valid after the stream is synchronised); *_host entry points take HOST pointers,
 *     copy in/out on the handle's own stream and return after the result has landed;
 *   - a handle is bound to one device and is not thread-safe; distinct handles are independent
 *     (one per rank for data-parallel training);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     CFFM_ERR_CUDA.
 */
#ifndef CFFM_H_
#define CFFM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFFM_ABI_VERSION 2

typedef enum cffm_status {
  CFFM_OK = 0,
  CFFM_ERR_INVALID = -1,  /* bad argument / unsupported configuration */
  CFFM_ERR_CUDA = -2,     /* CUDA runtime error or no device          */
  CFFM_ERR_NOMEM = -3,
  CFFM_ERR_IO = -4,       /* libfm loader: file errors                */
  CFFM_ERR_COMM = -5,     /* NCCL errors                              */
  CFFM_ERR_UNSUPPORTED = -6
} cffm_status;

/* --activation (CFFM.py:75-76, :132-141) */
typedef enum cffm_activation {
  CFFM_ACT_RELU = 0, CFFM_ACT_ELU = 1, CFFM_ACT_SELU = 2, CFFM_ACT_PRELU = 3, CFFM_ACT_GELU = 4
} cffm_activation;

/* --loss_type (CFFM.py:46-47, :486-514) */
typedef enum cffm_loss {
  CFFM_LOSS_SQUARE = 0, /* lamda==0: sqrt(mean((y-out)^2)+1e-10); lamda>0: l2_loss + regularisers */
  CFFM_LOSS_LOG = 1,    /* out=sigmoid(out); tf.contrib.losses.log_loss, eps 1e-7                 */
  CFFM_LOSS_MSE = 2,
  CFFM_LOSS_MAE = 3,
  CFFM_LOSS_HYBRID = 4
} cffm_loss;

/* --optimizer (CFFM.py:48-49, :517-529) */
typedef enum cffm_optimizer {
  CFFM_OPT_ADAGRAD = 0, /* lr, initial_accumulator_value 1e-8, no epsilon */
  CFFM_OPT_SGD = 1,
  CFFM_OPT_MOMENTUM = 2, /* momentum 0.95 */
  CFFM_OPT_ADAM = 3      /* 0.9 / 0.999 / 1e-8 */
} cffm_optimizer;

/* Arithmetic of the conv contraction. */
typedef enum cffm_precision {
  CFFM_PREC_FP32 = 0, /* fp32 SIMT contraction (reference arithmetic)                         */
  CFFM_PREC_BF16 = 1, /* bf16 operands on tcgen05 tensor cores, fp32 accumulation in TMEM,   */
                      /* fp32 master weights / optimizer state (logits within 1e-2)           */
  CFFM_PREC_BF16X3 = 2 /* split bf16 on the same tensor cores: every operand as hi + lo bf16,  */
                      /* a*b = hi*hi + lo*hi + hi*lo in one fp32 TMEM accumulator: fp32-class  */
                      /* accuracy (logits within 1e-4 of the fp32 graph) at tensor-core speed  */
} cffm_precision;

/* Mirrors the constructor arguments of class CFFM (CFFM.py:98-101). */
typedef struct cffm_config {
  int32_t abi_version;   /* CFFM_ABI_VERSION */
  int32_t features_M;    /* rows of the three tables (LoadData.features_M)   */
  int32_t num_field;     /* --num_field                                        */
  int32_t inner_dims;    /* --inner_dims (power of two, 4..64)                 */
  int32_t outer_dims;    /* --outer_dims (power of two, 4..64)                 */
  int32_t inner_conv;    /* --inner_conv 0/1                                   */
  int32_t outer_conv;    /* --outer_conv 0/1                                   */
  int32_t linear_att;    /* --linear_att 0/1                                   */
  int32_t activation;    /* cffm_activation                                    */
  int32_t loss_type;     /* cffm_loss                                          */
  int32_t optimizer;     /* cffm_optimizer                                     */
  int32_t precision;     /* cffm_precision                                     */
  float lr;              /* --lr                                               */
  float lamda;           /* --lamda (lamda_bilinear)                           */
  float lamda_att;       /* --lamda_att (softmax temperature; also Q9)         */
  float beta_outer;      /* --beta_outer                                       */
  int32_t max_batch;     /* largest B a single launch will see (workspace size) */
  int32_t device;        /* CUDA device ordinal                                */
  uint64_t seed;         /* parameter initialisation seed (reference: unseeded) */
  /* Row-sharded tables (new; the reference is single-device): with shard_world = G > 1 this handle owns the rows
   * r of inner_embeddings / outer_embeddings / feature_bias (and of their optimizer slots) with r % G == shard_rank,
   * stored densely as local row r / G; every gather and every update of a step then goes through an all-to-all
   * with the owners (cffm_comm_init must be called with the same rank / world before the first forward).
   * 0 or 1: tables replicated on every rank. */
  int32_t shard_world;
  int32_t shard_rank;
} cffm_config;

typedef struct cffm_handle cffm_handle;

/* ---- lifetime: tf.Session + build_graph + init_op (CFFM.py:158-161, :531-541) ------------- */
int cffm_create(const cffm_config* cfg, cffm_handle** out);
int cffm_destroy(cffm_handle* h);
const char* cffm_last_error(const cffm_handle* h);
/* CUDA runtime/device probe: 0 if a usable device exists, CFFM_ERR_CUDA otherwise. */
int cffm_device_available(void);

/* ---- variables: self.weights + the four tf.layers.dense layers (CFFM.py:257-284, :323,
 *      :376, Q6 names "dense/kernel" ...); accumulators = Adagrad slots (CFFM.py:523-524) ---- */
int cffm_param_count(const cffm_handle* h);
/* name_cap bytes at name; shape has room for 4 entries */
int cffm_param_info(const cffm_handle* h, int index, char* name, int name_cap, int64_t* shape,
                    int32_t* ndim, int64_t* numel, int32_t* trainable);
int cffm_get_param(cffm_handle* h, const char* name, float* host_dst, int64_t numel);
int cffm_set_param(cffm_handle* h, const char* name, const float* host_src, int64_t numel);
int cffm_get_accum(cffm_handle* h, const char* name, float* host_dst, int64_t numel);
int cffm_set_accum(cffm_handle* h, const char* name, const float* host_src, int64_t numel);
/* re-run the initialisers of CFFM.py:257-284 / :459-467 with a seed */
int cffm_init_params(cffm_handle* h, uint64_t seed);
/* Optimizer step counter: the `t` of AdamOptimizer's bias correction (beta1_power / beta2_power
 * non-slot variables [TF-1.14], CFFM.py:519-520).  The other optimizers have no step-dependent
 * state: they report 0 and ignore the setter.
 * A checkpoint has to carry it together with the slots ("<name>" = slot 1, "<name>:2" = Adam v
 * through cffm_get_accum / cffm_set_accum) for a resumed run to equal an uninterrupted one
 * (the reference's saver, CFFM.py:159 / :226-228, saves all global variables). */
int cffm_get_opt_step(cffm_handle* h, int64_t* step);
int cffm_set_opt_step(cffm_handle* h, int64_t step);

/* ---- sess.run(self.out) (CFFM.py:596): ids int32 [B, num_field] row-major -> out float [B] -- */
int cffm_forward_dev(cffm_handle* h, const int32_t* ids_dev, int64_t B, float* out_dev, void* stream);
int cffm_forward_host(cffm_handle* h, const int32_t* ids_host, int64_t B, float* out_host);

/* ---- sess.run((self.loss, self.optimizer)) (CFFM.py:200): one fwd + bwd + update ----------- */
int cffm_train_step_dev(cffm_handle* h, const int32_t* ids_dev, const float* labels_dev, int64_t B,
                        float* loss_dev, void* stream);
int cffm_train_step_host(cffm_handle* h, const int32_t* ids_host, const float* labels_host, int64_t B,
                         float* loss_host);
/* Pipelined form of the host entry point for the training loop (CFFM.py:186-200): the batch is
 * copied into one of two pinned staging slots and the step is enqueued; the loss of step t is
 * returned by the call that submits step t+1 (or by cffm_train_flush).  Returns the number of
 * losses written to loss_host (0 or 1) through *n_losses. */
int cffm_train_submit_host(cffm_handle* h, const int32_t* ids_host, const float* labels_host, int64_t B,
                           float* loss_host, int32_t* n_losses);
int cffm_train_flush(cffm_handle* h, float* loss_host, int32_t* n_losses);

/* ---- evaluate() (CFFM.py:583-615): ordered blocks of `batch`, predictions clipped to
 *      [min y, max y], RMSE and R2 reduced on the device; only the scalars come back -------- */
int cffm_evaluate_host(cffm_handle* h, const int32_t* ids_host, const float* labels_host, int64_t N,
                       int64_t batch, double* rmse, double* r2);

/* ---- resident training set: the loop of CFFM.py:181-200 without per-step host->device copies.
 *      The split is uploaded once; an epoch's shuffle (sklearn shuffle, CFFM.py:183) is a device gather
 *      with the host-made permutation (new[i] = old[perm[i]]); a step trains on the contiguous block
 *      [start, start+B) of the current order (get_random_block_from_data, CFFM.py:560-581), enqueued
 *      asynchronously; cffm_dataset_evaluate is evaluate() (CFFM.py:583-615) over the resident split. */
int cffm_dataset_upload(cffm_handle* h, const int32_t* ids_host, const float* labels_host, int64_t N);
int cffm_dataset_permute(cffm_handle* h, const int64_t* perm_host);
int cffm_train_block(cffm_handle* h, int64_t start, int64_t B);
int cffm_last_loss(cffm_handle* h, float* loss_host);
int cffm_dataset_evaluate(cffm_handle* h, int64_t batch, double* rmse, double* r2);

int cffm_synchronize(cffm_handle* h);
/* 1 while training steps replay from a CUDA graph (default), 0 after CFFM_GRAPH=0 or a failed capture */
int cffm_uses_graph(const cffm_handle* h);
/* number of kernels launched by the library on this handle since creation */
int64_t cffm_launch_count(const cffm_handle* h);

/* ---- per-kernel device timing: while enabled every launch of the library is bracketed by CUDA
 *      events on its stream (steps run eagerly, not from the CUDA graph).  cffm_profile_report
 *      synchronises and writes one line per kernel tag: "<tag> <launches> <total_ms>\n";
 *      returns the buffer size needed. */
int cffm_profile_enable(cffm_handle* h, int32_t on);
int64_t cffm_profile_report(cffm_handle* h, char* buf, int64_t cap, int32_t reset);

/* ---- operator-level entry points (the gather / scatter halves of the path) ----------------- */
/* tf.nn.embedding_lookup (CFFM.py:303, :354, :422): out[n, K] = table[ids[n], :] */
int cffm_op_gather_dev(const float* table_dev, const int32_t* ids_dev, int64_t n, int32_t K,
                       float* out_dev, void* stream);
/* IndexedSlices de-duplication + SparseApplyAdagrad [TF-1.14] (CFFM.py:523-524):
 * rows of `table` named by ids are updated with the per-row sum of grads (summed in order of
 * appearance).  Optionally returns the sorted unique rows (uniq_dev, capacity n) and their count. */
int cffm_op_sparse_adagrad_dev(float* table_dev, float* accum_dev, int32_t features_M, int32_t K,
                               const int32_t* ids_dev, const float* grads_dev, int64_t n, float lr,
                               int32_t* uniq_dev, int32_t* n_uniq_dev, void* stream);

/* bf16 tensor-core GEMM used by the conv stack in CFFM_PREC_BF16 mode, exposed for validation:
 * C[M,N] (fp32, row-major) = A[M,K] . B[N,K]^T, A and B bf16 row-major (K contiguous, K % 8 == 0).
 * tcgen05.mma with TMA-staged operands and fp32 accumulation in TMEM. */
int cffm_op_gemm_bf16_dev(const void* a_dev, const void* b_dev, float* c_dev, int32_t M, int32_t N, int32_t K,
                          void* stream);
/* Same pipeline with both operands transposed (MN-major UMMA descriptors), the form the conv
 * weight gradient takes: C[M,N] = sum_r A[r,M] * B[r,N]; A [R,M], B [R,N] bf16 row-major. */
int cffm_op_gemm_bf16_tn_dev(const void* a_dev, const void* b_dev, float* c_dev, int32_t M, int32_t N, int32_t R,
                             void* stream);
/* message of the last failed tensor-core launch on this thread */
const char* cffm_tc_last_error(void);

/* ---- own bounds check (compute-sanitizer is not available on the GPU pool this was built on): with CFFM_GUARD=1 in
 *      the environment every device allocation of the library carries a 4 KB pattern band on either side; this call
 *      synchronises and returns the number of allocations whose bands were written to (0 = clean, also when the
 *      guard is off; negative: error).  msg (may be NULL) receives a description of the first damaged one. */
int cffm_debug_check_guards(char* msg, int32_t cap);
/* checks the checker: 0 = a write one element before / after a guarded buffer is reported; -1 = CFFM_GUARD is off */
int cffm_debug_guard_selftest(void);

/* ---- intermediate tensors of the last forward / train step, for parity tests ---------------
 * what: "out", "final2", "final", "linear", "t1", "outer_rows", "conv_<l>" (pre-activation Y_l),
 *       "grad_out", "grad_inner_rows", "grad_outer_rows", "grad_bias_rows", "dense_grads",
 *       "sorted_ids", "uniq_rows".  Copies min(cap, n) floats and reports n. */
int cffm_debug_fetch(cffm_handle* h, const char* what, float* host_dst, int64_t cap, int64_t* n);
/* gradient of the flat dense parameter block for one named variable (after a train step) */
int cffm_debug_dense_grad(cffm_handle* h, const char* name, float* host_dst, int64_t numel);

/* ---- data-parallel training over NCCL (new; the reference is single-device, CFFM.py:19) ----
 * rank 0 calls cffm_comm_unique_id and ships the 128 bytes to the other ranks (the Python host
 * does that with torch.distributed); every rank then calls cffm_comm_init.  After that
 * cffm_train_step_* treats its batch as one shard of a global batch of world*B samples:
 * dense gradients and the loss sum are all-reduced; replicated tables: touched rows are all-gathered and every
 * rank applies the identical update; row-sharded tables (cfg.shard_world > 1): per step
 *   forward   unique ids of the local batch, bucketed by owner -> all-to-all -> owners gather their rows ->
 *             all-to-all back -> the step runs on the received rows
 *   backward  gradient rows summed per unique id on the requesting rank FIRST -> all-to-all to the owners ->
 *             owner-side sort + segment sum (ranks in order) + sparse optimizer update of its shard.
 * With sharded tables cffm_forward_* / cffm_evaluate_* are collective too: every rank must make the same
 * sequence of calls (batch sizes may differ). */
int cffm_comm_unique_id(char id_out[128]);
int cffm_comm_init(cffm_handle* h, const char id[128], int32_t rank, int32_t world);

/* ---- LoadData (LoadData.py:25-112): libfm text -> first-appearance ids -> CSR in pinned host
 *      memory.  Files are scanned for the vocabulary in the order train, test, validation. ---- */
typedef struct cffm_libfm cffm_libfm;
int cffm_libfm_load(const char* train_path, const char* test_path, const char* validation_path,
                    cffm_libfm** out);
int64_t cffm_libfm_features_M(const cffm_libfm* d);
/* split: 0 train, 1 validation, 2 test.  Rows are ordered by (stable) ascending row length
 * (LoadData.py:105-112).  labels_raw = float(items[0]); labels_log = 1 if >0 else 0. */
int cffm_libfm_split(const cffm_libfm* d, int split, int64_t* n_rows, const int64_t** row_ptr,
                     const int32_t** ids, const float** labels_raw, const float** labels_log);
/* vocabulary token of a feature id (for tests); returns length or negative */
int cffm_libfm_token(const cffm_libfm* d, int64_t feature_id, char* buf, int cap);
int cffm_libfm_free(cffm_libfm* d);
/* message of the last failed cffm_libfm_load on this thread */
const char* cffm_libfm_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* CFFM_H_ */
