// Parameter registry, device allocation and initialisers.
// Restates CFFM.py:239-293 (initialize_variables), :323 / :376-377 (conv weights), the four
// tf.layers.dense layers (:339, :409, :410, :441) and the initial values of Q7.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "model.h"

#include <stdlib.h>

#include <mutex>
#include <unordered_map>

namespace cffm {

// ---- guarded device allocations (common.cuh) -------------------------------------------------
namespace {
constexpr size_t kGuard = 4096;
constexpr uint32_t kPattern = 0xA5C3F00Du;
struct GuardRec { void* base; size_t bytes; };
std::mutex g_guard_mu;
std::unordered_map<void*, GuardRec> g_guard;   // user pointer -> allocation
bool guard_on() { static const bool on = [] { const char* e = getenv("CFFM_GUARD"); return e && e[0] == '1'; }(); return on; }
__global__ void k_guard_fill(uint32_t* p, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = kPattern;
}
__global__ void k_guard_check(const uint32_t* p, size_t n, int* bad) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && p[i] != kPattern) atomicAdd(bad, 1);
}
}  // namespace

cudaError_t dev_malloc(void** p, size_t bytes) {
  if (!guard_on()) return cudaMalloc(p, bytes);
  const size_t body = (bytes + 255) & ~size_t(255);
  void* base = nullptr;
  cudaError_t e = cudaMalloc(&base, body + 2 * kGuard);
  if (e != cudaSuccess) { *p = nullptr; return e; }
  const size_t nw = kGuard / 4;
  k_guard_fill<<<(unsigned)((nw + 255) / 256), 256>>>(reinterpret_cast<uint32_t*>(base), nw);
  // the tail band starts right after the requested bytes (rounded up to 4): an overrun by one element is seen
  const size_t tail_off = kGuard + ((bytes + 3) & ~size_t(3));
  k_guard_fill<<<(unsigned)((nw + 255) / 256), 256>>>(reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(base) + tail_off), nw);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { cudaFree(base); *p = nullptr; return e; }
  *p = reinterpret_cast<char*>(base) + kGuard;
  std::lock_guard<std::mutex> lk(g_guard_mu);
  g_guard[*p] = GuardRec{base, bytes};
  return cudaSuccess;
}

cudaError_t dev_free(void* p) {
  if (!p) return cudaSuccess;
  if (guard_on()) {
    std::lock_guard<std::mutex> lk(g_guard_mu);
    auto it = g_guard.find(p);
    if (it != g_guard.end()) { void* base = it->second.base; g_guard.erase(it); return cudaFree(base); }
  }
  return cudaFree(p);
}

int dev_check_guards(char* msg, int cap) {
  if (msg && cap > 0) msg[0] = 0;
  if (!guard_on()) return 0;
  if (cudaDeviceSynchronize() != cudaSuccess) { if (msg) snprintf(msg, cap, "device error before the guard check"); return -1; }
  int* bad_d = nullptr;
  if (cudaMalloc((void**)&bad_d, sizeof(int)) != cudaSuccess) return -1;
  int damaged = 0;
  std::lock_guard<std::mutex> lk(g_guard_mu);
  for (auto& kv : g_guard) {
    const GuardRec& r = kv.second;
    const size_t nw = kGuard / 4;
    const size_t tail_off = kGuard + ((r.bytes + 3) & ~size_t(3));
    int bad[2] = {0, 0};
    for (int side = 0; side < 2; ++side) {
      cudaMemset(bad_d, 0, sizeof(int));
      const char* at = reinterpret_cast<const char*>(r.base) + (side ? tail_off : 0);
      k_guard_check<<<(unsigned)((nw + 255) / 256), 256>>>(reinterpret_cast<const uint32_t*>(at), nw, bad_d);
      cudaMemcpy(&bad[side], bad_d, sizeof(int), cudaMemcpyDeviceToHost);
    }
    if (bad[0] || bad[1]) {
      if (!damaged && msg) snprintf(msg, cap, "allocation of %zu bytes: %d words damaged below, %d above", r.bytes, bad[0], bad[1]);
      ++damaged;
    }
  }
  cudaFree(bad_d);
  return damaged;
}

// self-test of the checker: a guarded buffer of 100 floats, element 100 (one past the end) and element -1 written
__global__ void k_guard_poke(float* p, int i) { p[i] = 1.f; }
int dev_guard_selftest() {
  if (!guard_on()) return -1;
  float* p = nullptr;
  if (dev_malloc((void**)&p, 100 * sizeof(float)) != cudaSuccess) return -2;
  char msg[128];
  const int before = dev_check_guards(msg, sizeof(msg));
  k_guard_poke<<<1, 1>>>(p, 100);
  const int after_hi = dev_check_guards(msg, sizeof(msg));
  k_guard_poke<<<1, 1>>>(p, -1);
  const int after_lo = dev_check_guards(msg, sizeof(msg));
  dev_free(p);
  const int end = dev_check_guards(msg, sizeof(msg));
  return (after_hi == before + 1 && after_lo == before + 1 && end == before) ? 0 : 1;
}

static int64_t align4(int64_t x) { return (x + 3) & ~int64_t(3); }

static void add_param(Model* m, const std::string& name, std::initializer_list<int64_t> shape, int kind,
                      int64_t* off_slot, bool trainable) {
  ParamInfo pi;
  pi.name = name;
  pi.ndim = (int)shape.size();
  pi.numel = 1;
  int i = 0;
  for (auto s : shape) { pi.shape[i++] = s; pi.numel *= s; }
  for (; i < 4; ++i) pi.shape[i] = 1;
  pi.kind = kind;
  pi.trainable = trainable;
  pi.offset = -1;
  if (kind == PK_DENSE) {
    pi.offset = m->lay.total;
    m->lay.total = align4(m->lay.total + pi.numel);
    if (off_slot) *off_slot = pi.offset;
  }
  m->params.push_back(pi);
}

int model_build_layout(Model* m) {
  const cffm_config& c = m->cfg;
  m->F = c.num_field; m->Ki = c.inner_dims; m->Ko = c.outer_dims; m->M = c.features_M;
  m->P = m->F * (m->F - 1) / 2;
  m->max_batch = c.max_batch; m->device = c.device;
  auto pow2 = [](int x) { return x >= 4 && x <= 64 && (x & (x - 1)) == 0; };
  if (c.abi_version != CFFM_ABI_VERSION) { m->err = "abi_version mismatch"; return CFFM_ERR_INVALID; }
  if (m->F < 2 || m->F > 64) { m->err = "num_field must be in [2,64]"; return CFFM_ERR_INVALID; }
  if (m->M < 1) { m->err = "features_M must be positive"; return CFFM_ERR_INVALID; }
  if (c.inner_conv && !pow2(m->Ki)) { m->err = "inner_dims must be a power of two in [4,64]"; return CFFM_ERR_INVALID; }
  if (c.outer_conv && !pow2(m->Ko)) { m->err = "outer_dims must be a power of two in [4,64]"; return CFFM_ERR_INVALID; }
  if (c.max_batch < 1) { m->err = "max_batch must be positive"; return CFFM_ERR_INVALID; }
  m->shard_world = c.shard_world > 1 ? c.shard_world : 1;
  m->shard_rank = c.shard_world > 1 ? c.shard_rank : 0;
  if (m->shard_rank < 0 || m->shard_rank >= m->shard_world) { m->err = "shard_rank outside [0, shard_world)"; return CFFM_ERR_INVALID; }
  m->Mloc_max = ((int64_t)m->M + m->shard_world - 1) / m->shard_world;
  m->Mloc = ((int64_t)m->M - m->shard_rank + m->shard_world - 1) / m->shard_world;   // rows r < M with r % world == rank
  if (m->Mloc < 1) { m->err = "features_M is smaller than shard_world"; return CFFM_ERR_INVALID; }
  if (m->Mloc_max * m->shard_world >= (1ll << 31)) { m->err = "features_M too large for 32-bit row keys"; return CFFM_ERR_INVALID; }
  if (c.activation < 0 || c.activation > CFFM_ACT_GELU) { m->err = "unknown activation"; return CFFM_ERR_INVALID; }
  if (c.loss_type < 0 || c.loss_type > CFFM_LOSS_HYBRID) { m->err = "unknown loss_type"; return CFFM_ERR_INVALID; }
  if (c.optimizer < 0 || c.optimizer > CFFM_OPT_ADAM) { m->err = "unknown optimizer"; return CFFM_ERR_INVALID; }
  if (c.lamda > 0.f && c.loss_type != CFFM_LOSS_SQUARE) {
    // the reference's log_loss + lamda branch indexes variables that do not exist (KeyError);
    // the other losses ignore lamda
    if (c.loss_type == CFFM_LOSS_LOG) { m->err = "log_loss with lamda > 0 raises KeyError in the reference"; return CFFM_ERR_UNSUPPORTED; }
  }
  if (c.lamda < 0.f) { m->err = "lamda must be >= 0"; return CFFM_ERR_INVALID; }
  if (!(c.lamda_att != 0.f)) { m->err = "lamda_att must be non-zero"; return CFFM_ERR_INVALID; }
  m->conv_depth = 0; m->n_live = 0; m->t1_dim = 0;
  if (c.outer_conv) {
    int d = 0; while ((1 << (d + 1)) <= m->Ko) ++d;  // int(log2 K)
    m->conv_depth = d; m->n_live = d - 1;
    for (int l = 0; l < d; ++l) m->t1_dim += m->Ko >> l;
  }
  if (c.precision != CFFM_PREC_FP32 && c.precision != CFFM_PREC_BF16 && c.precision != CFFM_PREC_BF16X3) { m->err = "unknown precision"; return CFFM_ERR_INVALID; }
  if (c.precision != CFFM_PREC_FP32) { int r = tc_supported(m); if (r != CFFM_OK) return r; }
  DenseLayout& L = m->lay;
  L = DenseLayout();
  for (int i = 0; i < kMaxConv; ++i) L.conv_w[i] = L.conv_b[i] = -1;
  const int64_t F = m->F, P = m->P, M = m->Mloc;   // table shapes are those of the local shard
  int dense_idx = 0;
  auto dense_name = [&]() { std::string n = dense_idx == 0 ? "dense" : "dense_" + std::to_string(dense_idx); ++dense_idx; return n; };
  if (c.inner_conv) add_param(m, "inner_embeddings", {M, m->Ki}, PK_TABLE_INNER, nullptr, true);
  if (c.outer_conv) {
    add_param(m, "outer_embeddings", {M, m->Ko}, PK_TABLE_OUTER, nullptr, true);
    add_param(m, "outer_W", {P, 1}, PK_DENSE, &L.outer_W, false);
    add_param(m, "outer_b", {1}, PK_DENSE, &L.outer_b, false);
  }
  add_param(m, "feature_bias", {M, 1}, PK_TABLE_BIAS, nullptr, true);
  if (c.linear_att) {
    add_param(m, "bias_W", {F, F}, PK_DENSE, &L.att_W, true);
    add_param(m, "bias_b", {F}, PK_DENSE, &L.att_b, true);
  }
  add_param(m, "bias", {}, PK_DENSE, &L.bias, true);
  m->params.back().ndim = 0;
  if (c.inner_conv) {
    add_param(m, "inner_layer_conv_weight_0", {1, 2, 1, 2}, PK_DENSE, &L.iconv_w, true);
    add_param(m, "inner_layer_conv_bias_0", {2}, PK_DENSE, &L.iconv_b, true);
    std::string n = dense_name();
    add_param(m, n + "/kernel", {P * m->Ki, 1}, PK_DENSE, &L.din_k, true);
    add_param(m, n + "/bias", {1}, PK_DENSE, &L.din_b, true);
  }
  if (c.outer_conv) {
    for (int l = 0; l < m->conv_depth; ++l) {
      bool live = l < m->n_live;
      add_param(m, "outer_layer_conv_weight_" + std::to_string(l), {2, 2, P, P}, PK_DENSE, &L.conv_w[l], live);
      add_param(m, "outer_layer_conv_bias_" + std::to_string(l), {P}, PK_DENSE, &L.conv_b[l], live);
    }
    std::string n1 = dense_name(), n2 = dense_name();
    add_param(m, n1 + "/kernel", {m->t1_dim, 32}, PK_DENSE, &L.d1_k, true);
    add_param(m, n1 + "/bias", {32}, PK_DENSE, &L.d1_b, true);
    add_param(m, n2 + "/kernel", {32, 1}, PK_DENSE, &L.d2_k, true);
    add_param(m, n2 + "/bias", {1}, PK_DENSE, &L.d2_b, true);
  }
  if (c.linear_att) {
    std::string n = dense_name();
    add_param(m, n + "/kernel", {F, 1}, PK_DENSE, &L.d3_k, true);
    add_param(m, n + "/bias", {1}, PK_DENSE, &L.d3_b, true);
  }
  return CFFM_OK;
}

const ParamInfo* model_find(const Model* m, const char* name) {
  for (auto& p : m->params)
    if (p.name == name) return &p;
  return nullptr;
}

// ---------------------------------------------------------------------------------------------
// Counter-based RNG: element index + stream key -> splitmix64 -> uniforms -> Box-Muller.
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ float u01(uint64_t bits) {  // (0,1)
  return ((float)(bits >> 40) + 0.5f) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float normal_at(uint64_t key, uint64_t idx, uint32_t attempt) {
  uint64_t h = splitmix64(key ^ splitmix64(idx * 0x100000001B3ull + attempt));
  uint64_t h2 = splitmix64(h);
  float u1 = u01(h), u2 = u01(h2);
  return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

enum InitKind { INIT_CONST = 0, INIT_NORMAL = 1, INIT_TRUNC_NORMAL = 2, INIT_UNIFORM = 3 };

// Tables of a sharded handle (world > 1): local element (lr, c) of a [Mloc, K] shard is element
// ((lr * world + rank) * K + c) of the full table, so every rank draws exactly the values the replicated layout has.
__global__ void k_init(float* __restrict__ dst, int64_t n, int kind, float a, uint64_t key, int K, int world, int rank) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; e < n; e += stride) {
    int64_t i = e;
    if (world > 1) { const int64_t lr = e / K; i = (lr * world + rank) * K + (e - lr * K); }
    float v;
    if (kind == INIT_CONST) v = a;
    else if (kind == INIT_NORMAL) v = a * normal_at(key, (uint64_t)i, 0);
    else if (kind == INIT_TRUNC_NORMAL) {  // tf.truncated_normal: redraw until |z| <= 2 (CFFM.py:460)
      uint32_t t = 0; float z;
      do { z = normal_at(key, (uint64_t)i, t++); } while (fabsf(z) > 2.f && t < 64);
      v = a * z;
    } else {  // uniform(-a, a): glorot_uniform of tf.layers.dense [TF-1.14]
      v = a * (2.f * u01(splitmix64(key ^ splitmix64((uint64_t)i))) - 1.f);
    }
    dst[e] = v;
  }
}

static void launch_init(Model* m, float* dst, int64_t n, int kind, float a, uint64_t key, int table_K = 0) {
  if (n <= 0) return;
  int blocks = (int)((n + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
  k_init<<<blocks, 256, 0, m->stream>>>(dst, n, kind, a, key, table_K > 0 ? table_K : 1, table_K > 0 ? m->shard_world : 1, m->shard_rank);
  m->launches++;
}

int model_init_params(Model* m, uint64_t seed) {
  const DenseLayout& L = m->lay;
  const int64_t F = m->F, P = m->P, M = m->Mloc;
  uint64_t k = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
  auto key = [&](int i) { return k + 0x632BE59BD9B4E019ull * (uint64_t)(i + 1); };
  float* w = m->dense_w;
  CFFM_CUDA_OK(m, cudaMemsetAsync(m->dense_w, 0, sizeof(float) * m->lay.total, m->stream));
  if (m->cfg.inner_conv) launch_init(m, m->inner_tab, M * m->Ki, INIT_NORMAL, 0.1f, key(0), m->Ki);     // :257-259
  if (m->cfg.outer_conv) {
    launch_init(m, m->outer_tab, M * m->Ko, INIT_NORMAL, 0.01f, key(1), m->Ko);                         // :264-266
    launch_init(m, w + L.outer_W, P, INIT_TRUNC_NORMAL, 1.f, key(2));                            // :271
    launch_init(m, w + L.outer_b, 1, INIT_TRUNC_NORMAL, 1.f, key(3));                            // :272
  }
  launch_init(m, m->fbias_tab, M, INIT_CONST, 0.f, 0);                                           // :276-277
  if (m->cfg.linear_att) {
    launch_init(m, w + L.att_W, F * F, INIT_TRUNC_NORMAL, 1.f, key(4));                          // :281
    launch_init(m, w + L.att_b, F, INIT_TRUNC_NORMAL, 1.f, key(5));                              // :282
    launch_init(m, w + L.d3_k, F, INIT_UNIFORM, sqrtf(6.f / (float)(F + 1)), key(6));            // :441
  }
  if (m->cfg.inner_conv) {
    launch_init(m, w + L.iconv_w, 4, INIT_TRUNC_NORMAL, 1.f, key(7));                            // :323
    launch_init(m, w + L.iconv_b, 2, INIT_CONST, 0.01f, 0);                                      // :466
    launch_init(m, w + L.din_k, P * m->Ki, INIT_UNIFORM, sqrtf(6.f / (float)(P * m->Ki + 1)), key(8));  // :339
  }
  if (m->cfg.outer_conv) {
    for (int l = 0; l < m->conv_depth; ++l) {                                                    // :375-377
      launch_init(m, w + L.conv_w[l], 4 * P * P, INIT_TRUNC_NORMAL, 1.f, key(10 + l));
      launch_init(m, w + L.conv_b[l], P, INIT_CONST, 0.01f, 0);
    }
    launch_init(m, w + L.d1_k, (int64_t)m->t1_dim * 32, INIT_UNIFORM, sqrtf(6.f / (float)(m->t1_dim + 32)), key(30));  // :409
    launch_init(m, w + L.d2_k, 32, INIT_UNIFORM, sqrtf(6.f / 33.f), key(31));                    // :410
  }
  // optimizer slots: Adagrad initial_accumulator_value = 1e-8 (CFFM.py:523-524); momentum / Adam slots start at 0
  const float s0 = m->cfg.optimizer == CFFM_OPT_ADAGRAD ? 1e-8f : 0.f;
  launch_init(m, m->dense_acc, m->lay.total, INIT_CONST, s0, 0);
  if (m->inner_acc) launch_init(m, m->inner_acc, M * m->Ki, INIT_CONST, s0, 0);
  if (m->outer_acc) launch_init(m, m->outer_acc, M * m->Ko, INIT_CONST, s0, 0);
  launch_init(m, m->fbias_acc, M, INIT_CONST, s0, 0);
  if (m->dense_acc2) launch_init(m, m->dense_acc2, m->lay.total, INIT_CONST, 0.f, 0);
  if (m->inner_acc2) launch_init(m, m->inner_acc2, M * m->Ki, INIT_CONST, 0.f, 0);
  if (m->outer_acc2) launch_init(m, m->outer_acc2, M * m->Ko, INIT_CONST, 0.f, 0);
  if (m->fbias_acc2) launch_init(m, m->fbias_acc2, M, INIT_CONST, 0.f, 0);
  CFFM_CUDA_OK(m, cudaMemsetAsync(m->scalars, 0, 16 * sizeof(float), m->stream));  // Adam step counter etc.
  CFFM_CUDA_OK(m, cudaGetLastError());
  CFFM_CUDA_OK(m, cudaStreamSynchronize(m->stream));
  return CFFM_OK;
}

// ---------------------------------------------------------------------------------------------
template <class T>
static int dmalloc(Model* m, T** p, int64_t n) {
  if (n <= 0) n = 1;
  cudaError_t e = dev_malloc((void**)p, sizeof(T) * (size_t)n);
  if (e != cudaSuccess) {
    m->err = std::string("cudaMalloc(") + std::to_string(sizeof(T) * (size_t)n) + " B): " + cudaGetErrorString(e);
    *p = nullptr;
    return e == cudaErrorMemoryAllocation ? CFFM_ERR_NOMEM : CFFM_ERR_CUDA;
  }
  return CFFM_OK;
}
#define TRY(x) do { int _r = (x); if (_r != CFFM_OK) return _r; } while (0)

int model_alloc(Model* m) {
  const int64_t B = m->max_batch, F = m->F, P = m->P, M = m->Mloc;
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  CFFM_CUDA_OK(m, cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  {
    const char* e = getenv("CFFM_SIDE_STREAM");
    if (!(e && !strcmp(e, "0"))) {
      for (int i = 0; i < 2; ++i) {
        CFFM_CUDA_OK(m, cudaStreamCreateWithFlags(&m->side[i], cudaStreamNonBlocking));
        CFFM_CUDA_OK(m, cudaEventCreateWithFlags(&m->ev_fork[i], cudaEventDisableTiming));
        CFFM_CUDA_OK(m, cudaEventCreateWithFlags(&m->ev_join[i], cudaEventDisableTiming));
      }
    }
  }
  if (m->cfg.inner_conv) { TRY(dmalloc(m, &m->inner_tab, M * m->Ki)); TRY(dmalloc(m, &m->inner_acc, M * m->Ki)); }
  if (m->cfg.outer_conv) { TRY(dmalloc(m, &m->outer_tab, M * m->Ko)); TRY(dmalloc(m, &m->outer_acc, M * m->Ko)); }
  TRY(dmalloc(m, &m->fbias_tab, M)); TRY(dmalloc(m, &m->fbias_acc, M));
  if (m->cfg.optimizer == CFFM_OPT_ADAM) {
    if (m->cfg.inner_conv) TRY(dmalloc(m, &m->inner_acc2, M * m->Ki));
    if (m->cfg.outer_conv) TRY(dmalloc(m, &m->outer_acc2, M * m->Ko));
    TRY(dmalloc(m, &m->fbias_acc2, M));
  }
  if (m->cfg.optimizer == CFFM_OPT_ADAM || m->cfg.lamda > 0.f) {
    TRY(dmalloc(m, &m->rowmap, M));
    CFFM_CUDA_OK(m, cudaMemset(m->rowmap, 0xFF, sizeof(int32_t) * M));  // -1: untouched
  }
  TRY(dmalloc(m, &m->sumsq_partial, 512));
  TRY(dmalloc(m, &m->dense_w, m->lay.total)); TRY(dmalloc(m, &m->dense_acc, m->lay.total));
  if (m->cfg.optimizer == CFFM_OPT_ADAM) TRY(dmalloc(m, &m->dense_acc2, m->lay.total));
  // dense_g = [gradients of the dense block | aux batch sums: q[t1_dim], G, pad[3], rowsums[n_small]]
  m->n_small = 2 * m->F + 6;
  m->aux_off = m->lay.total;
  const int64_t g_total = m->lay.total + align4(m->t1_dim + 4 + m->n_small);
  TRY(dmalloc(m, &m->dense_g, g_total));
  CFFM_CUDA_OK(m, cudaMemset(m->dense_g, 0, sizeof(float) * g_total));
  // pair tables, order of the loops at CFFM.py:304-305
  std::vector<int> pi(P), pj(P);
  int p = 0;
  for (int i = 0; i < F; ++i) for (int j = i + 1; j < F; ++j) { pi[p] = i; pj[p] = j; ++p; }
  TRY(dmalloc(m, &m->pair_i, P)); TRY(dmalloc(m, &m->pair_j, P));
  if (P > 0) {
    CFFM_CUDA_OK(m, cudaMemcpy(m->pair_i, pi.data(), sizeof(int) * P, cudaMemcpyHostToDevice));
    CFFM_CUDA_OK(m, cudaMemcpy(m->pair_j, pj.data(), sizeof(int) * P, cudaMemcpyHostToDevice));
  }
  TRY(dmalloc(m, &m->ids_buf, B * F)); TRY(dmalloc(m, &m->labels_buf, B));
  if (m->cfg.outer_conv) {
    TRY(dmalloc(m, &m->outer_rows, B * F * m->Ko));
    if (tc_path(m)) {
      TRY(tc_alloc(m, false));
    } else {
      for (int l = 0; l < m->n_live; ++l) {
        int64_t H = m->Ko >> (l + 1);
        TRY(dmalloc(m, &m->Y[l], B * H * H * P));
      }
    }
    TRY(dmalloc(m, &m->t1, B * m->t1_dim)); TRY(dmalloc(m, &m->hid, B * 32));
  }
  TRY(dmalloc(m, &m->comp_inner, B)); TRY(dmalloc(m, &m->comp_outer, B)); TRY(dmalloc(m, &m->comp_lin, B));
  TRY(dmalloc(m, &m->out, B)); TRY(dmalloc(m, &m->pred, B)); TRY(dmalloc(m, &m->loss_terms, B));
  TRY(dmalloc(m, &m->scalars, 16)); TRY(dmalloc(m, &m->loss_out, 1));
  TRY(dmalloc(m, &m->eval_acc, 8));
  TRY(dmalloc(m, &m->fb_buf, B * F));
  CFFM_CUDA_OK(m, cudaMemset(m->scalars, 0, 16 * sizeof(float)));
  for (int s = 0; s < 2; ++s) {
    CFFM_CUDA_OK(m, cudaHostAlloc((void**)&m->h_ids[s], sizeof(int32_t) * B * F, cudaHostAllocDefault));
    CFFM_CUDA_OK(m, cudaHostAlloc((void**)&m->h_labels[s], sizeof(float) * B, cudaHostAllocDefault));
    CFFM_CUDA_OK(m, cudaHostAlloc((void**)&m->h_loss[s], sizeof(float) * 4, cudaHostAllocDefault));
    CFFM_CUDA_OK(m, cudaEventCreateWithFlags(&m->slot_done[s], cudaEventDisableTiming));
  }
  CFFM_CUDA_OK(m, cudaHostAlloc((void**)&m->h_out, sizeof(float) * B, cudaHostAllocDefault));
  return CFFM_OK;
}

void model_free(Model* m) {
  if (m->device >= 0) cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  for (int i = 0; i < 2; ++i) if (m->side[i]) cudaStreamSynchronize(m->side[i]);
  if (m->step_graph) cudaGraphExecDestroy(m->step_graph);
  void* dev[] = {m->inner_tab, m->outer_tab, m->fbias_tab, m->inner_acc, m->outer_acc, m->fbias_acc, m->dense_w,
                 m->dense_acc, m->dense_g, m->pair_i, m->pair_j, m->ids_buf, m->labels_buf, m->outer_rows, m->t1,
                 m->hid, m->comp_inner, m->comp_outer, m->comp_lin, m->out, m->pred, m->loss_terms, m->scalars,
                 m->loss_out, m->gout, m->g_inner_rows, m->g_outer_rows, m->g_bias_rows, m->v_head, m->rowbuf,
                 m->partials, m->fb_buf, m->all_ids, m->all_g_inner, m->all_g_outer, m->all_g_bias, m->reduce_descs,
                 m->eval_acc, m->inner_acc2, m->outer_acc2, m->fbias_acc2, m->dense_acc2, m->rowmap, m->sumsq_partial};
  sparse_work_free(&m->sw);
  shard_free(m);
  tc_free(m);
  void* ds[] = {m->ds_ids, m->ds_ids_tmp, m->ds_labels, m->ds_labels_tmp, m->ds_perm};
  for (void* p : ds) if (p) dev_free(p);
  for (void* p : dev) if (p) dev_free(p);
  for (int l = 0; l < kMaxConv; ++l) { if (m->Y[l]) dev_free(m->Y[l]); if (m->dY[l]) dev_free(m->dY[l]); }
  for (int s = 0; s < 2; ++s) {
    if (m->h_ids[s]) cudaFreeHost(m->h_ids[s]);
    if (m->h_labels[s]) cudaFreeHost(m->h_labels[s]);
    if (m->h_loss[s]) cudaFreeHost(m->h_loss[s]);
    if (m->slot_done[s]) cudaEventDestroy(m->slot_done[s]);
  }
  if (m->h_out) cudaFreeHost(m->h_out);
  for (int i = 0; i < 2; ++i) {
    if (m->ev_fork[i]) cudaEventDestroy(m->ev_fork[i]);
    if (m->ev_join[i]) cudaEventDestroy(m->ev_join[i]);
    if (m->side[i]) cudaStreamDestroy(m->side[i]);
  }
  if (m->stream) cudaStreamDestroy(m->stream);
}

cudaStream_t side_fork(Model* m, cudaStream_t s, int which) {
  if (!m->side[which]) return s;
  if (cudaEventRecord(m->ev_fork[which], s) != cudaSuccess || cudaStreamWaitEvent(m->side[which], m->ev_fork[which], 0) != cudaSuccess) {
    cudaGetLastError();
    return s;   // nothing has been enqueued on the side stream: the work stays on the step's stream
  }
  return m->side[which];
}

int side_join(Model* m, cudaStream_t side, cudaStream_t s, int which) {
  if (side == s) return CFFM_OK;
  CFFM_CUDA_OK(m, cudaEventRecord(m->ev_join[which], side));
  CFFM_CUDA_OK(m, cudaStreamWaitEvent(s, m->ev_join[which], 0));
  return CFFM_OK;
}

}  // namespace cffm
