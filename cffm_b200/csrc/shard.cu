// Row-sharded embedding tables (cfg.shard_world = G > 1; SURVEY §8(e), BASELINE.json configs[3]): rank r owns the
// rows with row % G == r of inner_embeddings / outer_embeddings / feature_bias and of their optimizer slots
// (CFFM.py:257-266, :276-277 create the full tables; the reference is single-device).  Every step moves rows,
// not tables:
//
//   forward   ids of the local batch -> owner-major keys -> stable radix sort -> UNIQUE rows, grouped by owner
//             -> counts to every rank (all-gather of G integers, the one host synchronisation of a step)
//             -> all-to-all of row numbers -> owners gather the rows of the three tables -> all-to-all back.
//             The step then computes on the received rows ("mini tables", one row per unique id of the local
//             batch) with the ids renumbered to positions in that list: every kernel of forward.cu / backward.cu
//             runs unchanged.
//   backward  per-sample gradient rows -> summed per unique id on the requesting rank FIRST (phase 1 of update.cu:
//             the summation order of a single device) -> all-to-all of the sums to the owners -> owner: sort by local row + segmented sum over the
//             requesting ranks in rank order + sparse optimizer update of its shard (update.cu, unchanged).
//
// Traffic per rank and step: U * (4 + 4*(Ki+Ko+1)) bytes each way for U unique ids (vs. world * B * F rows for the
// replicated all-gather).  The all-to-alls are grouped ncclSend / ncclRecv pairs with exact counts; the rank's own
// part is a device-to-device copy.
#include <stdio.h>

#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "model.h"

namespace cffm {

struct ShardState {
  int64_t cap = 0;       // ids of one local batch: max_batch * F
  int64_t own_cap = 0;   // rows this rank can be asked for in one step: world * cap
  int32_t* keys = nullptr;        // [cap] owner-major keys (id % G) * Mloc_max + id / G
  SparseWork sw_req;              // requester-side sort of the keys
  int32_t* uniq_rows = nullptr;   // [cap] row number at its owner of every unique key, in key order
  int32_t* counts = nullptr;      // [G] unique rows per owner
  int32_t* all_counts = nullptr;  // [G][G] counts of every rank
  int32_t* h_counts = nullptr;    // pinned copy
  int32_t* ids_remap = nullptr;   // [cap] position of each id's row in the received list
  float *mini_inner = nullptr, *mini_outer = nullptr, *mini_bias = nullptr;   // [cap, K] rows received from the owners
  // (the gradient sums per unique row of the backward exchange live in sw_req.gsum: phase 1 of update.cu)
  int32_t* req_rows = nullptr;    // [own_cap] local rows the ranks asked this rank for, rank-major
  float *x_inner = nullptr, *x_outer = nullptr, *x_bias = nullptr;            // [own_cap, K] rows out (forward) / gradient rows in (backward)
  std::vector<int64_t> send_cnt, send_off, recv_cnt, recv_off;               // of the last forward exchange
  int64_t U = 0, R = 0;
};

// ---------------------------------------------------------------------------------------------
__global__ void k_shard_keys(const int32_t* __restrict__ ids, int n, int G, int Mloc_max, int32_t* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int id = ids[i];
  const int owner = id % G;
  keys[i] = owner * Mloc_max + id / G;
}

// unique key u (segment u of the sorted list) -> row number at its owner
__global__ void k_shard_uniq(const int32_t* __restrict__ sorted, const int32_t* __restrict__ seg_start, const int32_t* __restrict__ n_uniq,
                             int Mloc_max, int32_t* __restrict__ uniq_rows) {
  const int U = *n_uniq;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < U; u += gridDim.x * blockDim.x) uniq_rows[u] = sorted[seg_start[u]] % Mloc_max;
}

// counts[o] = number of unique keys of owner o: lower bounds of o * Mloc_max in the unique key list
__global__ void k_shard_bounds(const int32_t* __restrict__ sorted, const int32_t* __restrict__ seg_start, const int32_t* __restrict__ n_uniq,
                               int Mloc_max, int G, int32_t* __restrict__ counts) {
  __shared__ int32_t lb[129];
  const int U = *n_uniq;
  const int o = threadIdx.x;
  if (o <= G) {
    const int64_t want = (int64_t)o * Mloc_max;
    int lo = 0, hi = U;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((int64_t)sorted[seg_start[mid]] < want) lo = mid + 1; else hi = mid;
    }
    lb[o] = lo;
  }
  __syncthreads();
  if (o < G) counts[o] = lb[o + 1] - lb[o];
}

// ids_remap[source position] = index of the id's unique key
__global__ void k_shard_remap(const int32_t* __restrict__ pos, const int32_t* __restrict__ seg_start, const int32_t* __restrict__ n_uniq,
                              int n, int32_t* __restrict__ remap) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int U = *n_uniq;
  int lo = 0, hi = U - 1;                       // largest u with seg_start[u] <= t
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (seg_start[mid] <= t) lo = mid; else hi = mid - 1;
  }
  remap[pos[t]] = lo;
}

__global__ void k_gather_scalar(const float* __restrict__ tab, const int32_t* __restrict__ rows, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __ldg(tab + rows[i]);
}

// ---------------------------------------------------------------------------------------------
template <class T>
static int smalloc(Model* m, T** p, int64_t n) {
  cudaError_t e = dev_malloc((void**)p, sizeof(T) * (size_t)(n > 0 ? n : 1));
  if (e != cudaSuccess) { m->err = std::string("cudaMalloc (sharded tables): ") + cudaGetErrorString(e); *p = nullptr; return CFFM_ERR_NOMEM; }
  return CFFM_OK;
}
#define STRY(x) do { int _r = (x); if (_r != CFFM_OK) return _r; } while (0)

static int shard_alloc(Model* m) {
  if (m->shard) return CFFM_OK;
  if (!m->comm || m->world != m->shard_world || m->rank != m->shard_rank) {
    m->err = "row-sharded tables: call cffm_comm_init with rank == cfg.shard_rank and world == cfg.shard_world first";
    return CFFM_ERR_COMM;
  }
  ShardState* ss = new ShardState();
  m->shard = ss;
  const int G = m->shard_world;
  if (G > 128) { m->err = "shard_world > 128"; return CFFM_ERR_UNSUPPORTED; }
  ss->cap = (int64_t)m->max_batch * m->F;
  ss->own_cap = ss->cap * G;
  const int Ki = m->cfg.inner_conv ? m->Ki : 0, Ko = m->cfg.outer_conv ? m->Ko : 0;
  STRY(smalloc(m, &ss->keys, ss->cap));
  if (sparse_work_alloc(&ss->sw_req, ss->cap, Ki + Ko + 1, &m->err) != CFFM_OK) return CFFM_ERR_NOMEM;
  STRY(smalloc(m, &ss->uniq_rows, ss->cap));
  STRY(smalloc(m, &ss->counts, G));
  STRY(smalloc(m, &ss->all_counts, (int64_t)G * G));
  CFFM_CUDA_OK(m, cudaHostAlloc((void**)&ss->h_counts, sizeof(int32_t) * G * G, cudaHostAllocDefault));
  STRY(smalloc(m, &ss->ids_remap, ss->cap));
  if (Ki) { STRY(smalloc(m, &ss->mini_inner, ss->cap * Ki)); STRY(smalloc(m, &ss->x_inner, ss->own_cap * Ki)); }
  if (Ko) { STRY(smalloc(m, &ss->mini_outer, ss->cap * Ko)); STRY(smalloc(m, &ss->x_outer, ss->own_cap * Ko)); }
  STRY(smalloc(m, &ss->mini_bias, ss->cap)); STRY(smalloc(m, &ss->x_bias, ss->own_cap));
  STRY(smalloc(m, &ss->req_rows, ss->own_cap));
  ss->send_cnt.assign(G, 0); ss->send_off.assign(G + 1, 0); ss->recv_cnt.assign(G, 0); ss->recv_off.assign(G + 1, 0);
  return CFFM_OK;
}

void shard_free(Model* m) {
  ShardState* ss = m->shard;
  if (!ss) return;
  void* p[] = {ss->keys, ss->uniq_rows, ss->counts, ss->all_counts, ss->ids_remap, ss->mini_inner, ss->mini_outer, ss->mini_bias,
               ss->req_rows, ss->x_inner, ss->x_outer, ss->x_bias};
  for (void* q : p) if (q) dev_free(q);
  if (ss->h_counts) cudaFreeHost(ss->h_counts);
  sparse_work_free(&ss->sw_req);
  delete ss;
  m->shard = nullptr;
}

// One all-to-all: this rank sends `mine + send_off[p] * width` (send_cnt[p] rows) to every p and receives recv_cnt[p]
// rows from p at `theirs + recv_off[p] * width`; `width` in bytes per row.  swap = the reverse direction (rows back to
// the requesters / gradient sums to the owners use the transposed counts).
static int all_to_all_rows(Model* m, const void* mine, void* theirs, int64_t width, bool reverse, cudaStream_t s) {
  ShardState* ss = m->shard;
  const int G = m->shard_world, me = m->shard_rank;
  const std::vector<int64_t>& scnt = reverse ? ss->recv_cnt : ss->send_cnt;
  const std::vector<int64_t>& soff = reverse ? ss->recv_off : ss->send_off;
  const std::vector<int64_t>& rcnt = reverse ? ss->send_cnt : ss->recv_cnt;
  const std::vector<int64_t>& roff = reverse ? ss->send_off : ss->recv_off;
  if (scnt[me] > 0)
    CFFM_CUDA_OK(m, cudaMemcpyAsync((char*)theirs + roff[me] * width, (const char*)mine + soff[me] * width, (size_t)(scnt[me] * width),
                                    cudaMemcpyDeviceToDevice, s));
  struct Guard { Model* m; bool open; ~Guard() { if (open) comm_group_end(m); } } guard{m, false};
  int r = comm_group_begin(m); if (r != CFFM_OK) return r;
  guard.open = true;
  for (int p = 0; p < G; ++p) {
    if (p == me) continue;
    if (scnt[p] > 0) { r = comm_send(m, (const char*)mine + soff[p] * width, scnt[p] * width, p, s); if (r != CFFM_OK) return r; }
    if (rcnt[p] > 0) { r = comm_recv(m, (char*)theirs + roff[p] * width, rcnt[p] * width, p, s); if (r != CFFM_OK) return r; }
  }
  guard.open = false;
  return comm_group_end(m);
}

int shard_forward_exchange(Model* m, const int32_t* ids, int64_t B, cudaStream_t s, TableView* view) {
  int r = shard_alloc(m); if (r != CFFM_OK) return r;
  ShardState* ss = m->shard;
  const int G = m->shard_world, me = m->shard_rank;
  const int n = (int)(B * m->F);
  const int Mmax = (int)m->Mloc_max;
  {
    CFFM_PROF(m, "shard_sort_unique", s);
    k_shard_keys<<<ceil_div(n, 256), 256, 0, s>>>(ids, n, G, Mmax, ss->keys);
    m->launches++;
    r = sparse_sort_segments(&ss->sw_req, ss->keys, n, (int)std::min<int64_t>((int64_t)Mmax * G, 0x7fffffff), s, &m->launches);
    if (r != CFFM_OK) { m->err = "sharded tables: sort of the row keys failed"; return r; }
    int ub = ceil_div(n, 256); if (ub > 148 * 4) ub = 148 * 4;
    k_shard_uniq<<<ub, 256, 0, s>>>(ss->sw_req.keys_out, ss->sw_req.seg_start, ss->sw_req.n_uniq, Mmax, ss->uniq_rows);
    k_shard_bounds<<<1, 160, 0, s>>>(ss->sw_req.keys_out, ss->sw_req.seg_start, ss->sw_req.n_uniq, Mmax, G, ss->counts);
    k_shard_remap<<<ceil_div(n, 256), 256, 0, s>>>(ss->sw_req.vals_out, ss->sw_req.seg_start, ss->sw_req.n_uniq, n, ss->ids_remap);
    m->launches += 3;
  }
  {
    // counts of every rank: the sizes of the all-to-alls are host arguments, so this is the step's one host sync
    CFFM_PROF(m, "shard_counts", s);
    r = comm_allgather(m, ss->counts, ss->all_counts, sizeof(int32_t) * G, s); if (r != CFFM_OK) return r;
    CFFM_CUDA_OK(m, cudaMemcpyAsync(ss->h_counts, ss->all_counts, sizeof(int32_t) * G * G, cudaMemcpyDeviceToHost, s));
  }
  CFFM_CUDA_OK(m, cudaStreamSynchronize(s));
  ss->U = 0; ss->R = 0;
  for (int p = 0; p < G; ++p) {
    ss->send_cnt[p] = ss->h_counts[me * G + p]; ss->send_off[p] = ss->U; ss->U += ss->send_cnt[p];
    ss->recv_cnt[p] = ss->h_counts[p * G + me]; ss->recv_off[p] = ss->R; ss->R += ss->recv_cnt[p];
  }
  ss->send_off[G] = ss->U; ss->recv_off[G] = ss->R;
  if (ss->U < 1 || ss->U > ss->cap || ss->R > ss->own_cap) { m->err = "sharded tables: inconsistent row counts"; return CFFM_ERR_COMM; }
  { CFFM_PROF(m, "shard_ids_a2a", s);
    r = all_to_all_rows(m, ss->uniq_rows, ss->req_rows, sizeof(int32_t), false, s); if (r != CFFM_OK) return r; }
  if (ss->R > 0) {
    CFFM_PROF(m, "shard_gather_rows", s);
    if (m->cfg.inner_conv) { launch_gather_rows(m->inner_tab, ss->req_rows, ss->R, m->Ki, ss->x_inner, s); m->launches++; }
    if (m->cfg.outer_conv) { launch_gather_rows(m->outer_tab, ss->req_rows, ss->R, m->Ko, ss->x_outer, s); m->launches++; }
    k_gather_scalar<<<ceil_div(ss->R, 256), 256, 0, s>>>(m->fbias_tab, ss->req_rows, ss->R, ss->x_bias);
    m->launches++;
  }
  {
    CFFM_PROF(m, "shard_rows_a2a", s);
    if (m->cfg.inner_conv) { r = all_to_all_rows(m, ss->x_inner, ss->mini_inner, sizeof(float) * m->Ki, true, s); if (r != CFFM_OK) return r; }
    if (m->cfg.outer_conv) { r = all_to_all_rows(m, ss->x_outer, ss->mini_outer, sizeof(float) * m->Ko, true, s); if (r != CFFM_OK) return r; }
    r = all_to_all_rows(m, ss->x_bias, ss->mini_bias, sizeof(float), true, s); if (r != CFFM_OK) return r;
  }
  view->inner = ss->mini_inner; view->outer = ss->mini_outer; view->fbias = ss->mini_bias; view->ids = ss->ids_remap;
  CFFM_CUDA_OK(m, cudaGetLastError());
  return CFFM_OK;
}

// After the backward kernels: per-sample gradient rows -> unique-row sums -> owners -> update of the local shard.
int shard_backward_update(Model* m, int64_t B, cudaStream_t s) {
  ShardState* ss = m->shard;
  if (!ss) { m->err = "sharded tables: no forward exchange before the update"; return CFFM_ERR_INVALID; }
  int r = CFFM_OK;
  const int n = (int)(B * m->F);
  // per-unique-id sums on the requesting rank (phase 1 of the update, the summation order of a single device) ...
  SparseTables tq;
  int64_t goff[3] = {0, 0, 0};
  {
    int j = 0; int64_t off = 0;
    if (m->cfg.inner_conv) { tq.tab[j] = ss->mini_inner; tq.grads[j] = m->g_inner_rows; tq.K[j] = m->Ki; goff[j] = off; off += ss->cap * m->Ki; ++j; }
    if (m->cfg.outer_conv) { tq.tab[j] = ss->mini_outer; tq.grads[j] = m->g_outer_rows; tq.K[j] = m->Ko; goff[j] = off; off += ss->cap * m->Ko; ++j; }
    tq.tab[j] = ss->mini_bias; tq.grads[j] = m->g_bias_rows; tq.K[j] = 1; goff[j] = off; ++j;
  }
  { CFFM_PROF(m, "shard_grad_reduce", s);
    launch_segment_sums(&ss->sw_req, tq, n, s, &m->launches); }
  {  // ... and on to the owners
    CFFM_PROF(m, "shard_grad_a2a", s);
    int j = 0;
    if (m->cfg.inner_conv) { r = all_to_all_rows(m, ss->sw_req.gsum + goff[j], ss->x_inner, sizeof(float) * m->Ki, false, s); if (r != CFFM_OK) return r; ++j; }
    if (m->cfg.outer_conv) { r = all_to_all_rows(m, ss->sw_req.gsum + goff[j], ss->x_outer, sizeof(float) * m->Ko, false, s); if (r != CFFM_OK) return r; ++j; }
    r = all_to_all_rows(m, ss->sw_req.gsum + goff[j], ss->x_bias, sizeof(float), false, s); if (r != CFFM_OK) return r;
  }
  const int opt = m->cfg.optimizer;
  const bool adam = opt == CFFM_OPT_ADAM;
  const bool l2 = m->cfg.lamda > 0.f && m->cfg.loss_type == CFFM_LOSS_SQUARE;
  if (adam) { launch_adam_tick(m->scalars, m->cfg.lr, s); m->launches++; }
  const float* lr_dev = adam ? m->scalars + 4 : nullptr;
  if (ss->R > 0 || adam || l2) {
    // the rows of this shard that any rank touched: sort by local row, sum the ranks' contributions in rank order
    if (ss->R > 0) {
      CFFM_PROF(m, "sort_segments", s);
      r = sparse_sort_segments(&m->sw, ss->req_rows, ss->R, (int)m->Mloc_max, s, &m->launches);
      if (r != CFFM_OK) { m->err = "sparse_sort_segments failed"; return r; }
    }
    SparseTables t;
    t.rowmap = m->rowmap; t.M = m->Mloc;
    int j = 0;
    if (m->cfg.inner_conv) { t.tab[j] = m->inner_tab; t.acc[j] = m->inner_acc; t.acc2[j] = m->inner_acc2; t.grads[j] = ss->x_inner; t.K[j] = m->Ki;
                             t.dense[j] = adam || l2; t.reg[j] = l2 ? m->cfg.lamda : 0.f; ++j; }
    if (m->cfg.outer_conv) { t.tab[j] = m->outer_tab; t.acc[j] = m->outer_acc; t.acc2[j] = m->outer_acc2; t.grads[j] = ss->x_outer; t.K[j] = m->Ko;
                             t.dense[j] = adam || l2; t.reg[j] = l2 ? m->cfg.lamda_att : 0.f; ++j; }
    t.tab[j] = m->fbias_tab; t.acc[j] = m->fbias_acc; t.acc2[j] = m->fbias_acc2; t.grads[j] = ss->x_bias; t.K[j] = 1; t.dense[j] = adam; ++j;
    if (opt == CFFM_OPT_SGD) for (int q = 0; q < 3; ++q) t.acc[q] = nullptr;
    if (ss->R > 0) {
      CFFM_PROF(m, "sparse_adagrad", s);
      launch_sparse_update(&m->sw, t, ss->R, opt, m->cfg.lr, lr_dev, s, &m->launches);
    } else {
      m->err = "sharded tables: a dense table pass (Adam / lamda > 0) needs at least one requested row per rank and step";
      return CFFM_ERR_UNSUPPORTED;
    }
  }
  CFFM_CUDA_OK(m, cudaGetLastError());
  return CFFM_OK;
}

}  // namespace cffm
