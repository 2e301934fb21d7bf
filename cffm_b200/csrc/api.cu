// C ABI (include/cffm.h) over the kernels.  No C++ exception leaves this file.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>

#include "common.cuh"
#include "kernels.h"
#include "model.h"

using namespace cffm;


static thread_local std::string g_err;

#define API_BEGIN try {
#define API_END(h)                                                        \
  } catch (const std::bad_alloc&) {                                       \
    if (h) (h)->m.err = "out of host memory"; else g_err = "out of host memory"; \
    return CFFM_ERR_NOMEM;                                                \
  } catch (const std::exception& e) {                                     \
    if (h) (h)->m.err = e.what(); else g_err = e.what();                  \
    return CFFM_ERR_INVALID;                                              \
  } catch (...) {                                                         \
    if (h) (h)->m.err = "unknown C++ exception"; else g_err = "unknown C++ exception"; \
    return CFFM_ERR_INVALID;                                              \
  }

extern "C" const char* cffm_last_error(const cffm_handle* h) { return h ? h->m.err.c_str() : g_err.c_str(); }

extern "C" int cffm_device_available(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n < 1) {
    g_err = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    cudaGetLastError();
    return CFFM_ERR_CUDA;
  }
  return CFFM_OK;
}

extern "C" int cffm_create(const cffm_config* cfg, cffm_handle** out) {
  cffm_handle* h = nullptr;
  API_BEGIN
  if (!cfg || !out) { g_err = "null argument"; return CFFM_ERR_INVALID; }
  *out = nullptr;
  if (cffm_device_available() != CFFM_OK) return CFFM_ERR_CUDA;  // no CPU fallback
  h = new cffm_handle();
  Model* m = &h->m;
  m->cfg = *cfg;
  int r = model_build_layout(m);
  if (r == CFFM_OK) r = model_alloc(m);
  if (r == CFFM_OK) r = forward_setup_attrs(m);
  if (r == CFFM_OK) r = backward_setup_attrs(m);
  if (r == CFFM_OK) r = model_init_params(m, cfg->seed);
  if (r != CFFM_OK) { g_err = m->err; model_free(m); delete h; return r; }
  const char* eg = getenv("CFFM_GRAPH");
  m->use_graph = !(eg && eg[0] == '0');
  if (sharded(m)) m->use_graph = false;   // the row exchange reads its counts on the host: the step is not one graph
  *out = h;
  return CFFM_OK;
  API_END((cffm_handle*)nullptr)
}

extern "C" int cffm_destroy(cffm_handle* h) {
  if (!h) return CFFM_OK;
  Model* m = &h->m;
  if (m->device >= 0) cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  cudaDeviceSynchronize();
  // a captured step holds NCCL kernels: the graph has to go before the communicator
  if (m->step_graph) { cudaGraphExecDestroy(m->step_graph); m->step_graph = nullptr; }
  comm_destroy(m);
  model_free(m);
  delete h;
  return CFFM_OK;
}

extern "C" int cffm_synchronize(cffm_handle* h) {
  if (!h) return CFFM_ERR_INVALID;
  CFFM_CUDA_OK(&h->m, cudaStreamSynchronize(h->m.stream));
  return CFFM_OK;
}

extern "C" int cffm_debug_check_guards(char* msg, int32_t cap) { return dev_check_guards(msg, cap); }
extern "C" int cffm_debug_guard_selftest(void) { return dev_guard_selftest(); }

extern "C" int cffm_uses_graph(const cffm_handle* h) { return h && h->m.use_graph ? 1 : 0; }

extern "C" int64_t cffm_launch_count(const cffm_handle* h) { return h ? h->m.launches : 0; }

// ---- per-kernel timing ------------------------------------------------------------------------
namespace cffm {
ProfScope::ProfScope(Model* m_, const char* tag, cudaStream_t s_) : m(m_), s(s_), idx(-1) {
  if (!m->prof_on) return;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return;
  Model::ProfEv ev;
  ev.tag = tag;
  for (cudaEvent_t* e : {&ev.a, &ev.b}) {
    if (!m->prof_pool.empty()) { *e = m->prof_pool.back(); m->prof_pool.pop_back(); }
    else if (cudaEventCreate(e) != cudaSuccess) return;
  }
  cudaEventRecord(ev.a, s);
  m->prof_pending.push_back(ev);
  idx = (int)m->prof_pending.size() - 1;
}
ProfScope::~ProfScope() { if (idx >= 0) cudaEventRecord(m->prof_pending[idx].b, s); }
}  // namespace cffm

extern "C" int cffm_profile_enable(cffm_handle* h, int32_t on) {
  if (!h) return CFFM_ERR_INVALID;
  h->m.prof_on = on != 0;
  return CFFM_OK;
}

extern "C" int64_t cffm_profile_report(cffm_handle* h, char* buf, int64_t cap, int32_t reset) {
  if (!h) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  cudaSetDevice(m->device);
  cudaDeviceSynchronize();
  for (auto& ev : m->prof_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev.a, ev.b) == cudaSuccess) { auto& acc = m->prof_acc[ev.tag]; acc.first += 1; acc.second += ms; }
    m->prof_pool.push_back(ev.a); m->prof_pool.push_back(ev.b);
  }
  m->prof_pending.clear();
  cudaGetLastError();
  std::string out;
  for (auto& kv : m->prof_acc) {
    char line[256];
    snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), (long long)kv.second.first, kv.second.second);
    out += line;
  }
  if (buf && cap > 0) { strncpy(buf, out.c_str(), (size_t)cap - 1); buf[cap - 1] = 0; }
  if (reset) m->prof_acc.clear();
  return (int64_t)out.size() + 1;
}

// ---------------------------------------------------------------------------------------------
extern "C" int cffm_param_count(const cffm_handle* h) { return h ? (int)h->m.params.size() : CFFM_ERR_INVALID; }

extern "C" int cffm_param_info(const cffm_handle* h, int index, char* name, int name_cap, int64_t* shape, int32_t* ndim,
                               int64_t* numel, int32_t* trainable) {
  if (!h || index < 0 || index >= (int)h->m.params.size()) return CFFM_ERR_INVALID;
  const ParamInfo& p = h->m.params[index];
  if (name && name_cap > 0) { strncpy(name, p.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (shape) for (int i = 0; i < 4; ++i) shape[i] = p.shape[i];
  if (ndim) *ndim = p.ndim;
  if (numel) *numel = p.numel;
  if (trainable) *trainable = p.trainable ? 1 : 0;
  return CFFM_OK;
}

// slot 0: the variable; 1: first optimizer slot; 2: second slot (Adam v)
static float* param_ptr(Model* m, const ParamInfo* p, int slot) {
  switch (p->kind) {
    case PK_TABLE_INNER: return slot == 0 ? m->inner_tab : (slot == 1 ? m->inner_acc : m->inner_acc2);
    case PK_TABLE_OUTER: return slot == 0 ? m->outer_tab : (slot == 1 ? m->outer_acc : m->outer_acc2);
    case PK_TABLE_BIAS: return slot == 0 ? m->fbias_tab : (slot == 1 ? m->fbias_acc : m->fbias_acc2);
    default: {
      float* base = slot == 0 ? m->dense_w : (slot == 1 ? m->dense_acc : m->dense_acc2);
      return base ? base + p->offset : nullptr;
    }
  }
}

static int param_copy(cffm_handle* h, const char* name, float* host, int64_t numel, bool accum, bool to_host) {
  if (!h || !name || !host) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  std::string nm(name);
  int slot = accum ? 1 : 0;
  if (accum && nm.size() > 2 && nm.compare(nm.size() - 2, 2, ":2") == 0) { slot = 2; nm.resize(nm.size() - 2); }  // "<name>:2" = Adam v
  const ParamInfo* p = model_find(m, nm.c_str());
  if (!p) { m->err = std::string("unknown variable: ") + name; return CFFM_ERR_INVALID; }
  if (numel != p->numel) { m->err = std::string("size mismatch for ") + name; return CFFM_ERR_INVALID; }
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  CFFM_CUDA_OK(m, cudaStreamSynchronize(m->stream));
  float* d = param_ptr(m, p, slot);
  if (!d) { m->err = std::string("this optimizer has no such slot: ") + name; return CFFM_ERR_INVALID; }
  if (to_host) CFFM_CUDA_OK(m, cudaMemcpy(host, d, sizeof(float) * numel, cudaMemcpyDeviceToHost));
  else CFFM_CUDA_OK(m, cudaMemcpy(d, host, sizeof(float) * numel, cudaMemcpyHostToDevice));
  return CFFM_OK;
}
extern "C" int cffm_get_param(cffm_handle* h, const char* n, float* d, int64_t k) { return param_copy(h, n, d, k, false, true); }
extern "C" int cffm_set_param(cffm_handle* h, const char* n, const float* s, int64_t k) { return param_copy(h, n, const_cast<float*>(s), k, false, false); }
extern "C" int cffm_get_accum(cffm_handle* h, const char* n, float* d, int64_t k) { return param_copy(h, n, d, k, true, true); }
extern "C" int cffm_set_accum(cffm_handle* h, const char* n, const float* s, int64_t k) { return param_copy(h, n, const_cast<float*>(s), k, true, false); }

// Adam's step counter lives in scalars[5] (a float: exact up to 2^24 steps), advanced by k_adam_tick
extern "C" int cffm_get_opt_step(cffm_handle* h, int64_t* step) {
  if (!h || !step) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  *step = 0;
  if (m->cfg.optimizer != CFFM_OPT_ADAM) return CFFM_OK;
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  CFFM_CUDA_OK(m, cudaStreamSynchronize(m->stream));
  float t = 0.f;
  CFFM_CUDA_OK(m, cudaMemcpy(&t, m->scalars + 5, sizeof(float), cudaMemcpyDeviceToHost));
  *step = (int64_t)t;
  return CFFM_OK;
}
extern "C" int cffm_set_opt_step(cffm_handle* h, int64_t step) {
  if (!h || step < 0 || step >= (1 << 24)) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  if (m->cfg.optimizer != CFFM_OPT_ADAM) return CFFM_OK;
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  CFFM_CUDA_OK(m, cudaStreamSynchronize(m->stream));
  const float t = (float)step;
  CFFM_CUDA_OK(m, cudaMemcpy(m->scalars + 5, &t, sizeof(float), cudaMemcpyHostToDevice));
  return CFFM_OK;
}

extern "C" int cffm_init_params(cffm_handle* h, uint64_t seed) {
  if (!h) return CFFM_ERR_INVALID;
  CFFM_CUDA_OK(&h->m, cudaSetDevice(h->m.device));
  return model_init_params(&h->m, seed);
}

// ---------------------------------------------------------------------------------------------
static int check_batch(Model* m, int64_t B) {
  if (B < 1 || B > m->max_batch) { m->err = "batch size outside [1, max_batch]"; return CFFM_ERR_INVALID; }
  return CFFM_OK;
}

__global__ void k_copy_f32(const float* __restrict__ src, float* __restrict__ dst, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

extern "C" int cffm_forward_dev(cffm_handle* h, const int32_t* ids_dev, int64_t B, float* out_dev, void* stream) {
  API_BEGIN
  if (!h || !ids_dev || !out_dev) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  int r = check_batch(m, B); if (r != CFFM_OK) return r;
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  cudaStream_t s = (cudaStream_t)stream;
  r = run_forward(m, ids_dev, nullptr, B, s); if (r != CFFM_OK) return r;
  k_copy_f32<<<ceil_div(B, 256), 256, 0, s>>>(m->pred, out_dev, B);
  m->launches++;
  m->last_B = B;
  CFFM_CUDA_OK(m, cudaGetLastError());
  return CFFM_OK;
  API_END(h)
}

extern "C" int cffm_forward_host(cffm_handle* h, const int32_t* ids_host, int64_t N, float* out_host) {
  API_BEGIN
  if (!h || !ids_host || !out_host || N < 1) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  // the pinned staging slot may still feed the H2D copy of a pipelined training step
  if (m->pending) { m->err = "cffm_train_flush the pipelined steps first"; return CFFM_ERR_INVALID; }
  const int F = m->F;
  for (int64_t o = 0; o < N; o += m->max_batch) {  // ordered blocks, last one partial (CFFM.py:617-629)
    const int64_t B = std::min<int64_t>(m->max_batch, N - o);
    CFFM_CUDA_OK(m, cudaEventSynchronize(m->slot_done[0]));
    memcpy(m->h_ids[0], ids_host + o * F, sizeof(int32_t) * B * F);
    CFFM_CUDA_OK(m, cudaMemcpyAsync(m->ids_buf, m->h_ids[0], sizeof(int32_t) * B * F, cudaMemcpyHostToDevice, m->stream));
    int r = run_forward(m, m->ids_buf, nullptr, B, m->stream); if (r != CFFM_OK) return r;
    CFFM_CUDA_OK(m, cudaMemcpyAsync(m->h_out, m->pred, sizeof(float) * B, cudaMemcpyDeviceToHost, m->stream));
    CFFM_CUDA_OK(m, cudaStreamSynchronize(m->stream));
    memcpy(out_host + o, m->h_out, sizeof(float) * B);
    m->last_B = B;
  }
  return CFFM_OK;
  API_END(h)
}

// fwd + bwd + update on stream s with device-resident inputs
static int train_step_on(Model* m, const int32_t* ids, const float* labels, int64_t B, cudaStream_t s) {
  int r = run_forward(m, ids, labels, B, s);
  if (r != CFFM_OK) return r;
  return run_backward_update(m, ids, labels, B, s);
}

static int enqueue_staged_step(Model* m, int64_t B, cudaStream_t run_stream);

extern "C" int cffm_train_step_dev(cffm_handle* h, const int32_t* ids_dev, const float* labels_dev, int64_t B,
                                   float* loss_dev, void* stream) {
  API_BEGIN
  if (!h || !ids_dev || !labels_dev) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  int r = check_batch(m, B); if (r != CFFM_OK) return r;
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  r = model_alloc_train(m); if (r != CFFM_OK) return r;
  cudaStream_t s = (cudaStream_t)stream;
  if (m->use_graph && !m->prof_on) {
    // stage into the library's buffers (device-to-device) so that the step can replay from its CUDA graph
    CFFM_CUDA_OK(m, cudaMemcpyAsync(m->ids_buf, ids_dev, sizeof(int32_t) * B * m->F, cudaMemcpyDeviceToDevice, s));
    CFFM_CUDA_OK(m, cudaMemcpyAsync(m->labels_buf, labels_dev, sizeof(float) * B, cudaMemcpyDeviceToDevice, s));
    r = enqueue_staged_step(m, B, s);
  } else {
    r = train_step_on(m, ids_dev, labels_dev, B, s);
  }
  if (r != CFFM_OK) return r;
  if (loss_dev) { k_copy_f32<<<1, 32, 0, s>>>(m->loss_out, loss_dev, 1); m->launches++; }
  m->last_B = B;
  CFFM_CUDA_OK(m, cudaGetLastError());
  return CFFM_OK;
  API_END(h)
}

// Enqueue one step that reads the device staging buffers; replayed from a CUDA graph when the
// batch size repeats (the reference loop uses one fixed batch size, CFFM.py:186-200).
static int enqueue_staged_step(Model* m, int64_t B, cudaStream_t run_stream) {
  if (m->prof_on) return train_step_on(m, m->ids_buf, m->labels_buf, B, run_stream);  // eager, event-bracketed
  if (m->use_graph && m->step_graph && m->step_graph_B == B) {
    CFFM_CUDA_OK(m, cudaGraphLaunch(m->step_graph, run_stream));
    m->launches += m->step_graph_launches;
    return CFFM_OK;
  }
  if (m->use_graph) {
    if (m->step_graph) { cudaGraphExecDestroy(m->step_graph); m->step_graph = nullptr; m->step_graph_B = -1; }
    const int64_t before = m->launches;
    cudaGraph_t graph = nullptr;
    CFFM_CUDA_OK(m, cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal));
    int r = train_step_on(m, m->ids_buf, m->labels_buf, B, m->stream);
    cudaError_t e = cudaStreamEndCapture(m->stream, &graph);
    if (r == CFFM_OK && e == cudaSuccess && graph) {
      e = cudaGraphInstantiate(&m->step_graph, graph, 0);
      cudaGraphDestroy(graph);
      if (e == cudaSuccess) {
        m->step_graph_B = B;
        m->step_graph_launches = m->launches - before;
        m->launches = before;
        CFFM_CUDA_OK(m, cudaGraphLaunch(m->step_graph, run_stream));
        m->launches += m->step_graph_launches;
        return CFFM_OK;
      }
    }
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    m->launches = before;
    m->use_graph = false;  // capture is not possible here: run eagerly from now on
    if (r != CFFM_OK) return r;
  }
  return train_step_on(m, m->ids_buf, m->labels_buf, B, run_stream);
}

static int submit(Model* m, const int32_t* ids_host, const float* labels_host, int64_t B) {
  const int sl = m->slot;
  const int F = m->F;
  // the slot's previous use (two submissions ago) must have been consumed by the device
  CFFM_CUDA_OK(m, cudaEventSynchronize(m->slot_done[sl]));
  memcpy(m->h_ids[sl], ids_host, sizeof(int32_t) * B * F);
  memcpy(m->h_labels[sl], labels_host, sizeof(float) * B);
  CFFM_CUDA_OK(m, cudaMemcpyAsync(m->ids_buf, m->h_ids[sl], sizeof(int32_t) * B * F, cudaMemcpyHostToDevice, m->stream));
  CFFM_CUDA_OK(m, cudaMemcpyAsync(m->labels_buf, m->h_labels[sl], sizeof(float) * B, cudaMemcpyHostToDevice, m->stream));
  int r = enqueue_staged_step(m, B, m->stream); if (r != CFFM_OK) return r;
  CFFM_CUDA_OK(m, cudaMemcpyAsync(m->h_loss[sl], m->loss_out, sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  CFFM_CUDA_OK(m, cudaEventRecord(m->slot_done[sl], m->stream));
  m->last_B = B;
  return CFFM_OK;
}

extern "C" int cffm_train_step_host(cffm_handle* h, const int32_t* ids_host, const float* labels_host, int64_t B,
                                    float* loss_host) {
  API_BEGIN
  if (!h || !ids_host || !labels_host) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  int r = check_batch(m, B); if (r != CFFM_OK) return r;
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  r = model_alloc_train(m); if (r != CFFM_OK) return r;
  if (m->pending) { m->err = "cffm_train_flush the pipelined steps first"; return CFFM_ERR_INVALID; }
  r = submit(m, ids_host, labels_host, B); if (r != CFFM_OK) return r;
  CFFM_CUDA_OK(m, cudaEventSynchronize(m->slot_done[m->slot]));
  if (loss_host) *loss_host = m->h_loss[m->slot][0];
  m->slot ^= 1;
  return CFFM_OK;
  API_END(h)
}

extern "C" int cffm_train_submit_host(cffm_handle* h, const int32_t* ids_host, const float* labels_host, int64_t B,
                                      float* loss_host, int32_t* n_losses) {
  API_BEGIN
  if (!h || !ids_host || !labels_host) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  int r = check_batch(m, B); if (r != CFFM_OK) return r;
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  r = model_alloc_train(m); if (r != CFFM_OK) return r;
  if (n_losses) *n_losses = 0;
  r = submit(m, ids_host, labels_host, B); if (r != CFFM_OK) return r;
  const int prev = m->slot ^ 1;
  if (m->pending) {  // report the step submitted by the previous call
    CFFM_CUDA_OK(m, cudaEventSynchronize(m->slot_done[prev]));
    if (loss_host) *loss_host = m->h_loss[prev][0];
    if (n_losses) *n_losses = 1;
  }
  m->pending = 1;
  m->slot ^= 1;
  return CFFM_OK;
  API_END(h)
}

extern "C" int cffm_train_flush(cffm_handle* h, float* loss_host, int32_t* n_losses) {
  if (!h) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  if (n_losses) *n_losses = 0;
  if (m->pending) {
    const int prev = m->slot ^ 1;
    CFFM_CUDA_OK(m, cudaEventSynchronize(m->slot_done[prev]));
    if (loss_host) *loss_host = m->h_loss[prev][0];
    if (n_losses) *n_losses = 1;
    m->pending = 0;
  }
  CFFM_CUDA_OK(m, cudaStreamSynchronize(m->stream));
  return CFFM_OK;
}

// ---------------------------------------------------------------------------------------------
// evaluate() (CFFM.py:583-615): clip to [min y, max y], RMSE and R2, reduced on the device.
// acc: 0 min, 1 max, 2 sum y, 3 sum y^2, 4 sum (y - clip(pred))^2
__global__ void k_label_stats(const float* __restrict__ y, int64_t n, double* __restrict__ acc) {
  __shared__ double s_min[32], s_max[32], s_sum[32], s_sq[32];
  double mn = 1e300, mx = -1e300, su = 0.0, sq = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = y[i];
    mn = v < mn ? v : mn; mx = v > mx ? v : mx; su += v; sq += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
    mn = a < mn ? a : mn; mx = b > mx ? b : mx;
    su += __shfl_xor_sync(0xffffffffu, su, o); sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s_min[w] = mn; s_max[w] = mx; s_sum[w] = su; s_sq[w] = sq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
      mn = s_min[i] < mn ? s_min[i] : mn; mx = s_max[i] > mx ? s_max[i] : mx; su += s_sum[i]; sq += s_sq[i];
    }
    acc[0] = mn; acc[1] = mx; acc[2] = su; acc[3] = sq; acc[4] = 0.0;
  }
}
__global__ void k_sse_clipped(const float* __restrict__ pred, const float* __restrict__ y, int64_t n, double* __restrict__ acc) {
  __shared__ double red[32];
  const double lo = acc[0], hi = acc[1];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    double p = pred[i];
    p = p < lo ? lo : (p > hi ? hi : p);
    const double d = (double)y[i] - p;
    s += d * d;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    acc[4] += s;  // single block per launch, launches are stream-ordered: deterministic
  }
}

extern "C" int cffm_evaluate_host(cffm_handle* h, const int32_t* ids_host, const float* labels_host, int64_t N,
                                  int64_t batch, double* rmse, double* r2) {
  API_BEGIN
  if (!h || !ids_host || !labels_host || N < 1) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  if (batch < 1 || batch > m->max_batch) batch = m->max_batch;
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  if (m->pending) { m->err = "cffm_train_flush the pipelined steps first"; return CFFM_ERR_INVALID; }
  const int F = m->F;
  float* y_dev = nullptr;
  CFFM_CUDA_OK(m, dev_malloc((void**)&y_dev, sizeof(float) * N));
  cudaError_t e = cudaMemcpyAsync(y_dev, labels_host, sizeof(float) * N, cudaMemcpyHostToDevice, m->stream);
  if (e != cudaSuccess) { dev_free(y_dev); m->err = cudaGetErrorString(e); return CFFM_ERR_CUDA; }
  k_label_stats<<<1, 1024, 0, m->stream>>>(y_dev, N, m->eval_acc);
  m->launches++;
  int rc = CFFM_OK;
  int sl = 0;
  for (int64_t o = 0; o < N && rc == CFFM_OK; o += batch) {
    const int64_t B = std::min<int64_t>(batch, N - o);
    cudaEventSynchronize(m->slot_done[sl]);
    memcpy(m->h_ids[sl], ids_host + o * F, sizeof(int32_t) * B * F);
    cudaMemcpyAsync(m->ids_buf, m->h_ids[sl], sizeof(int32_t) * B * F, cudaMemcpyHostToDevice, m->stream);
    cudaEventRecord(m->slot_done[sl], m->stream);
    rc = run_forward(m, m->ids_buf, nullptr, B, m->stream);
    k_sse_clipped<<<1, 1024, 0, m->stream>>>(m->pred, y_dev + o, B, m->eval_acc);
    m->launches++;
    sl ^= 1;
    m->last_B = B;
  }
  double acc[5] = {0, 0, 0, 0, 0};
  e = cudaMemcpyAsync(acc, m->eval_acc, sizeof(acc), cudaMemcpyDeviceToHost, m->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
  dev_free(y_dev);
  if (rc != CFFM_OK) return rc;
  if (e != cudaSuccess) { m->err = cudaGetErrorString(e); return CFFM_ERR_CUDA; }
  const double sse = acc[4], n = (double)N;
  const double sst = acc[3] - acc[2] * acc[2] / n;
  if (rmse) *rmse = sqrt(sse / n);
  if (r2) *r2 = sst > 0 ? 1.0 - sse / sst : 0.0;
  return CFFM_OK;
  API_END(h)
}

// ---------------------------------------------------------------------------------------------
// Resident training set.
__global__ void k_permute_rows(const int32_t* __restrict__ ids, const float* __restrict__ y, const int64_t* __restrict__ perm,
                               int64_t N, int F, int32_t* __restrict__ ids_out, float* __restrict__ y_out) {
  const int64_t total = N * F;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / F; const int f = (int)(e - r * F);
    const int64_t src = perm[r];
    ids_out[e] = ids[src * F + f];
    if (f == 0) y_out[r] = y[src];
  }
}

extern "C" int cffm_dataset_upload(cffm_handle* h, const int32_t* ids_host, const float* labels_host, int64_t N) {
  API_BEGIN
  if (!h || !ids_host || !labels_host || N < 1) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  CFFM_CUDA_OK(m, cudaStreamSynchronize(m->stream));
  void* old[] = {m->ds_ids, m->ds_ids_tmp, m->ds_labels, m->ds_labels_tmp, m->ds_perm};
  for (void* p : old) if (p) dev_free(p);
  m->ds_ids = m->ds_ids_tmp = nullptr; m->ds_labels = m->ds_labels_tmp = nullptr; m->ds_perm = nullptr; m->ds_N = 0;
  const int F = m->F;
  CFFM_CUDA_OK(m, dev_malloc((void**)&m->ds_ids, sizeof(int32_t) * N * F));
  CFFM_CUDA_OK(m, dev_malloc((void**)&m->ds_ids_tmp, sizeof(int32_t) * N * F));
  CFFM_CUDA_OK(m, dev_malloc((void**)&m->ds_labels, sizeof(float) * N));
  CFFM_CUDA_OK(m, dev_malloc((void**)&m->ds_labels_tmp, sizeof(float) * N));
  CFFM_CUDA_OK(m, dev_malloc((void**)&m->ds_perm, sizeof(int64_t) * N));
  CFFM_CUDA_OK(m, cudaMemcpy(m->ds_ids, ids_host, sizeof(int32_t) * N * F, cudaMemcpyHostToDevice));
  CFFM_CUDA_OK(m, cudaMemcpy(m->ds_labels, labels_host, sizeof(float) * N, cudaMemcpyHostToDevice));
  m->ds_N = N;
  return CFFM_OK;
  API_END(h)
}

extern "C" int cffm_dataset_permute(cffm_handle* h, const int64_t* perm_host) {
  API_BEGIN
  if (!h || !perm_host) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  if (!m->ds_N) { m->err = "no resident dataset"; return CFFM_ERR_INVALID; }
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  CFFM_CUDA_OK(m, cudaMemcpyAsync(m->ds_perm, perm_host, sizeof(int64_t) * m->ds_N, cudaMemcpyHostToDevice, m->stream));
  const int64_t total = m->ds_N * m->F;
  int blocks = (int)((total + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
  k_permute_rows<<<blocks, 256, 0, m->stream>>>(m->ds_ids, m->ds_labels, m->ds_perm, m->ds_N, m->F, m->ds_ids_tmp, m->ds_labels_tmp);
  m->launches++;
  std::swap(m->ds_ids, m->ds_ids_tmp); std::swap(m->ds_labels, m->ds_labels_tmp);
  CFFM_CUDA_OK(m, cudaStreamSynchronize(m->stream));  // perm_host may be freed by the caller
  return CFFM_OK;
  API_END(h)
}

extern "C" int cffm_train_block(cffm_handle* h, int64_t start, int64_t B) {
  API_BEGIN
  if (!h) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  int r = check_batch(m, B); if (r != CFFM_OK) return r;
  if (start < 0 || start + B > m->ds_N) { m->err = "block outside the resident dataset"; return CFFM_ERR_INVALID; }
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  r = model_alloc_train(m); if (r != CFFM_OK) return r;
  CFFM_CUDA_OK(m, cudaMemcpyAsync(m->ids_buf, m->ds_ids + start * m->F, sizeof(int32_t) * B * m->F, cudaMemcpyDeviceToDevice, m->stream));
  CFFM_CUDA_OK(m, cudaMemcpyAsync(m->labels_buf, m->ds_labels + start, sizeof(float) * B, cudaMemcpyDeviceToDevice, m->stream));
  r = enqueue_staged_step(m, B, m->stream);
  m->last_B = B;
  return r;
  API_END(h)
}

extern "C" int cffm_last_loss(cffm_handle* h, float* loss_host) {
  if (!h || !loss_host) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  CFFM_CUDA_OK(m, cudaMemcpyAsync(m->h_loss[0] + 2, m->loss_out, sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  CFFM_CUDA_OK(m, cudaStreamSynchronize(m->stream));
  *loss_host = m->h_loss[0][2];
  return CFFM_OK;
}

static int evaluate_device(Model* m, const int32_t* ids_dev, const float* y_dev, int64_t N, int64_t batch, double* rmse, double* r2) {
  k_label_stats<<<1, 1024, 0, m->stream>>>(y_dev, N, m->eval_acc);
  m->launches++;
  int rc = CFFM_OK;
  for (int64_t o = 0; o < N && rc == CFFM_OK; o += batch) {
    const int64_t B = std::min<int64_t>(batch, N - o);
    rc = run_forward(m, ids_dev + o * m->F, nullptr, B, m->stream);
    k_sse_clipped<<<1, 1024, 0, m->stream>>>(m->pred, y_dev + o, B, m->eval_acc);
    m->launches++;
    m->last_B = B;
  }
  if (rc != CFFM_OK) return rc;
  double acc[5] = {0, 0, 0, 0, 0};
  CFFM_CUDA_OK(m, cudaMemcpyAsync(acc, m->eval_acc, sizeof(acc), cudaMemcpyDeviceToHost, m->stream));
  CFFM_CUDA_OK(m, cudaStreamSynchronize(m->stream));
  const double sse = acc[4], n = (double)N;
  const double sst = acc[3] - acc[2] * acc[2] / n;
  if (rmse) *rmse = sqrt(sse / n);
  if (r2) *r2 = sst > 0 ? 1.0 - sse / sst : 0.0;
  return CFFM_OK;
}

extern "C" int cffm_dataset_evaluate(cffm_handle* h, int64_t batch, double* rmse, double* r2) {
  API_BEGIN
  if (!h) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  if (!m->ds_N) { m->err = "no resident dataset"; return CFFM_ERR_INVALID; }
  if (batch < 1 || batch > m->max_batch) batch = m->max_batch;
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  return evaluate_device(m, m->ds_ids, m->ds_labels, m->ds_N, batch, rmse, r2);
  API_END(h)
}

// ---------------------------------------------------------------------------------------------
extern "C" int cffm_op_gather_dev(const float* table_dev, const int32_t* ids_dev, int64_t n, int32_t K, float* out_dev,
                                  void* stream) {
  if (!table_dev || !ids_dev || !out_dev || n < 0 || K < 4 || (K & 3)) { g_err = "bad argument (K must be a multiple of 4)"; return CFFM_ERR_INVALID; }
  launch_gather_rows(table_dev, ids_dev, n, K, out_dev, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return CFFM_ERR_CUDA; }
  return CFFM_OK;
}

__global__ void k_copy_uniq(const int32_t* __restrict__ sorted, const int32_t* __restrict__ seg_start,
                            const int32_t* __restrict__ n_uniq, int32_t* __restrict__ uniq_out, int32_t* __restrict__ n_out) {
  const int U = *n_uniq;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < U; i += gridDim.x * blockDim.x) uniq_out[i] = sorted[seg_start[i]];
  if (blockIdx.x == 0 && threadIdx.x == 0 && n_out) *n_out = U;
}

extern "C" int cffm_op_sparse_adagrad_dev(float* table_dev, float* accum_dev, int32_t features_M, int32_t K,
                                          const int32_t* ids_dev, const float* grads_dev, int64_t n, float lr,
                                          int32_t* uniq_dev, int32_t* n_uniq_dev, void* stream) {
  if (!table_dev || !accum_dev || !ids_dev || !grads_dev || n < 1 || K < 1) { g_err = "bad argument"; return CFFM_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  SparseWork w;
  std::string err;
  cudaStreamSynchronize(s);
  int r = sparse_work_alloc(&w, n, K, &err);
  if (r != CFFM_OK) { g_err = err; sparse_work_free(&w); return r; }
  r = sparse_sort_segments(&w, ids_dev, n, features_M, s, nullptr);
  if (r == CFFM_OK) {
    SparseTables t;
    t.tab[0] = table_dev; t.acc[0] = accum_dev; t.grads[0] = grads_dev; t.K[0] = K;
    launch_sparse_update(&w, t, n, CFFM_OPT_ADAGRAD, lr, nullptr, s, nullptr);
    if (uniq_dev) k_copy_uniq<<<64, 256, 0, s>>>(w.keys_out, w.seg_start, w.n_uniq, uniq_dev, n_uniq_dev);
  }
  cudaError_t e = cudaStreamSynchronize(s);
  sparse_work_free(&w);
  if (r != CFFM_OK) { g_err = "sort failed"; return r; }
  if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return CFFM_ERR_CUDA; }
  return CFFM_OK;
}

// ---------------------------------------------------------------------------------------------
__global__ void k_i32_to_f32(const int32_t* __restrict__ src, float* __restrict__ dst, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (float)src[i];
}

extern "C" int cffm_debug_fetch(cffm_handle* h, const char* what, float* host_dst, int64_t cap, int64_t* n_out) {
  API_BEGIN
  if (!h || !what) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  CFFM_CUDA_OK(m, cudaStreamSynchronize(m->stream));
  CFFM_CUDA_OK(m, cudaDeviceSynchronize());
  const int64_t B = m->last_B, F = m->F, P = m->P;
  const float* src = nullptr; int64_t n = 0;
  const int32_t* isrc = nullptr;
  std::string w(what);
  if (w == "out") { src = m->out; n = B; }
  else if (w == "pred") { src = m->pred; n = B; }
  else if (w == "final2") { src = m->comp_inner; n = B; }
  else if (w == "final") { src = m->comp_outer; n = B; }
  else if (w == "linear") { src = m->comp_lin; n = B; }
  else if (w == "t1") { src = m->t1; n = B * m->t1_dim; }
  else if (w == "outer_rows") { src = m->outer_rows; n = B * F * m->Ko; }
  else if (w == "grad_out") { src = m->gout; n = B; }
  else if (w == "grad_inner_rows") { src = m->g_inner_rows; n = B * F * m->Ki; }
  else if (w == "grad_outer_rows") { src = m->g_outer_rows; n = B * F * m->Ko; }
  else if (w == "grad_bias_rows") { src = m->g_bias_rows; n = B * F; }
  else if (w == "dense_grads") { src = m->dense_g; n = m->lay.total; }
  else if (w == "loss") { src = m->loss_out; n = 1; }
  else if (w == "sorted_ids") { isrc = m->sw.keys_out; n = B * F * m->world; }
  else if (w == "seg_start") { isrc = m->sw.seg_start; n = B * F * m->world; }
  else if (w == "n_uniq") { isrc = m->sw.n_uniq; n = 1; }
  else if (w.rfind("conv_", 0) == 0 || w.rfind("dconv_", 0) == 0) {
    const bool grad = w[0] == 'd';
    const int l = atoi(w.c_str() + (grad ? 6 : 5));
    if (l < 0 || l >= m->n_live) { m->err = "no such conv layer"; return CFFM_ERR_INVALID; }
    const int64_t H = m->Ko >> (l + 1);
    n = B * H * H * P;
    if (tc_path(m)) {  // bf16 mode stores X_{l+1} = phi(Y_l) (and dY_l) padded, in bf16
      if (n_out) *n_out = n;
      if (!host_dst || cap <= 0) return CFFM_OK;
      float* tmp = nullptr;
      CFFM_CUDA_OK(m, dev_malloc((void**)&tmp, sizeof(float) * n));
      int r = tc_debug_fetch(m, grad, l, tmp, B * H * H);
      if (r == CFFM_OK) cudaMemcpy(host_dst, tmp, sizeof(float) * std::min(cap, n), cudaMemcpyDeviceToHost);
      dev_free(tmp);
      return r;
    }
    src = grad ? m->dY[l] : m->Y[l];
  } else { m->err = std::string("unknown tensor: ") + what; return CFFM_ERR_INVALID; }
  if (n_out) *n_out = n;
  if ((!src && !isrc) || n == 0) { m->err = std::string("tensor not available: ") + what; return CFFM_ERR_INVALID; }
  const int64_t k = std::min(cap, n);
  if (host_dst && k > 0) {
    if (src) CFFM_CUDA_OK(m, cudaMemcpy(host_dst, src, sizeof(float) * k, cudaMemcpyDeviceToHost));
    else {
      float* tmp = nullptr;
      CFFM_CUDA_OK(m, dev_malloc((void**)&tmp, sizeof(float) * k));
      k_i32_to_f32<<<ceil_div(k, 256), 256>>>(isrc, tmp, k);
      cudaError_t e = cudaMemcpy(host_dst, tmp, sizeof(float) * k, cudaMemcpyDeviceToHost);
      dev_free(tmp);
      if (e != cudaSuccess) { m->err = cudaGetErrorString(e); return CFFM_ERR_CUDA; }
    }
  }
  return CFFM_OK;
  API_END(h)
}

extern "C" int cffm_debug_dense_grad(cffm_handle* h, const char* name, float* host_dst, int64_t numel) {
  if (!h || !name || !host_dst) return CFFM_ERR_INVALID;
  Model* m = &h->m;
  const ParamInfo* p = model_find(m, name);
  if (!p || p->kind != PK_DENSE) { m->err = std::string("not a dense variable: ") + name; return CFFM_ERR_INVALID; }
  if (numel != p->numel) { m->err = "size mismatch"; return CFFM_ERR_INVALID; }
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  CFFM_CUDA_OK(m, cudaStreamSynchronize(m->stream));
  CFFM_CUDA_OK(m, cudaDeviceSynchronize());
  CFFM_CUDA_OK(m, cudaMemcpy(host_dst, m->dense_g + p->offset, sizeof(float) * numel, cudaMemcpyDeviceToHost));
  return CFFM_OK;
}
