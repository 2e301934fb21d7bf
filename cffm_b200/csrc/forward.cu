// Forward pass of the CFFM graph (CFFM.py:296-453) + loss terms (CFFM.py:486-514).
#include <stdio.h>

#include <string>

#include "common.cuh"
#include "gemm_simt.cuh"
#include "kernels.h"
#include "model.h"

namespace cffm {

// ---------------------------------------------------------------------------------------------
// tf.nn.embedding_lookup (CFFM.py:303, :354, :422): out[n, :] = table[ids[n], :], 128-bit lanes.
__global__ void k_gather_rows(const float* __restrict__ table, const int32_t* __restrict__ ids, int64_t n,
                              int K4, float4* __restrict__ out) {
  const int64_t total = n * K4;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / K4;
    const int c = (int)(t - r * K4);
    const int64_t row = __ldg(ids + r);
    out[t] = __ldg(reinterpret_cast<const float4*>(table) + row * K4 + c);
  }
}

void launch_gather_rows(const float* table, const int32_t* ids, int64_t n, int K, float* out, cudaStream_t s) {
  if (n <= 0) return;
  const int K4 = K / 4;
  int64_t total = n * K4;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_gather_rows<<<blocks, 256, 0, s>>>(table, ids, n, K4, reinterpret_cast<float4*>(out));
}

// ---------------------------------------------------------------------------------------------
// Inner path (CFFM.py:301-343) + linear term (CFFM.py:422-446), one warp per sample.  The rows
// of inner_embeddings / feature_bias are gathered straight into shared memory; the P x K inner
// product map, its activation, the 1x2 conv, the max-pool of the activated input and the dense
// projection all stay in registers.
struct InnerLinArgs {
  const int32_t* ids; int B, F, P, K, lgK;
  int inner_conv, linear_att;
  const float *tab, *fbias;                 // inner_embeddings [M,K], feature_bias [M]
  const float *cw, *cb, *Wd, *bd;           // inner conv filter [1,2,1,2], bias [2]; dense kernel [P*K], bias
  const float *attW, *attb, *w3, *b3;       // bias_W [F,F], bias_b [F]; dense_3 kernel [F], bias
  const int *pair_i, *pair_j;
  float tau;                                // lamda_att
  float *comp_inner, *comp_lin;             // [B]
  float* fb_out;                            // [B,F] gathered feature_bias (kept for the backward pass)
};

template <int ACT>
__global__ void k_inner_linear_fwd(const InnerLinArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int F = a.F, K = a.K, P = a.P;
  int* s_pi = reinterpret_cast<int*>(sm);
  int* s_pj = s_pi + P;
  // keep every warp's row block 16-byte aligned for the float4 stores
  float* wbase = sm + ((2 * P + 3) & ~3) + (size_t)warp * (F * K + ((2 * F + 3) & ~3));
  float* e = wbase; float* fb = e + F * K; float* z = fb + F;
  for (int t = threadIdx.x; t < P; t += blockDim.x) { s_pi[t] = a.pair_i[t]; s_pj[t] = a.pair_j[t]; }
  __syncthreads();
  const int b = blockIdx.x * wpb + warp;
  if (b >= a.B) return;
  const int32_t* id = a.ids + (int64_t)b * F;
  if (a.inner_conv) {
    const int K4 = K >> 2;
    for (int t = lane; t < F * K4; t += 32) {
      const int f = t / K4, c = t - f * K4;
      const float4 v = __ldg(reinterpret_cast<const float4*>(a.tab + (int64_t)__ldg(id + f) * K) + c);
      *reinterpret_cast<float4*>(e + f * K + 4 * c) = v;
    }
    __syncwarp();
    // a lane takes two neighbouring k (the two taps of one conv position, K is even): both output channels of the
    // position come out of one lane, no exchange between lanes, 64 elements per pass
    const float w00 = __ldg(a.cw + 0), w01 = __ldg(a.cw + 1), w10 = __ldg(a.cw + 2), w11 = __ldg(a.cw + 3);  // W[0,t,0,o] -> t*2+o
    const float c0 = __ldg(a.cb + 0), c1 = __ldg(a.cb + 1);
    const int PK = P * K;
    float acc = 0.f;
#pragma unroll 4   // the dense-weight loads of four passes go out together: a pass no longer waits for its own
    for (int base = 0; base < PK; base += 64) {
      const int idx = base + 2 * lane;
      if (idx < PK) {
        const int p = idx >> a.lgK, k = idx & (K - 1);
        const float2 ei = *reinterpret_cast<const float2*>(e + s_pi[p] * K + k);
        const float2 ej = *reinterpret_cast<const float2*>(e + s_pj[p] * K + k);
        const float a0 = act_f<ACT>(ei.x * ej.x), a1 = act_f<ACT>(ei.y * ej.y);   // :310, :319  taps 2w, 2w+1
        const float y0 = fmaf(a1, w10, a0 * w00) + c0;                            // :327 conv + bias, o = 0
        const float y1 = fmaf(a1, w11, a0 * w01) + c1;                            //                   o = 1
        const float mx = fmaxf(a0, a1);                                           // :331-332
        const float2 wd = __ldg(reinterpret_cast<const float2*>(a.Wd + idx));     // :333 flatten (p,w,o), :339
        acc = fmaf(phi_f<ACT>(y0) + mx, wd.x, acc);                               // :478, :330
        acc = fmaf(phi_f<ACT>(y1) + mx, wd.y, acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) a.comp_inner[b] = acc + __ldg(a.bd);
  }
  for (int f = lane; f < F; f += 32) {
    const float v = __ldg(a.fbias + __ldg(id + f));  // :422
    fb[f] = v;
    if (a.fb_out) a.fb_out[(int64_t)b * F + f] = v;
  }
  __syncwarp();
  float lin = 0.f;
  if (a.linear_att) {
    float zmax = -INFINITY;
    for (int q = lane; q < F; q += 32) {
      float s = 0.f;
      for (int f = 0; f < F; ++f) s = fmaf(fb[f], __ldg(a.attW + f * F + q), s);  // :432
      s = (s + __ldg(a.attb + q)) / a.tau;                                          // :434
      z[q] = s; zmax = fmaxf(zmax, s);
    }
    zmax = warp_max(zmax);
    float se = 0.f;
    for (int q = lane; q < F; q += 32) { const float ex = expf(z[q] - zmax); z[q] = ex; se += ex; }
    se = warp_sum(se);                                                               // :436
    for (int q = lane; q < F; q += 32) lin = fmaf(fb[q] * (z[q] / se), __ldg(a.w3 + q), lin);  // :438, :441
    lin = warp_sum(lin) + __ldg(a.b3);
  } else {
    for (int q = lane; q < F; q += 32) lin += fb[q];  // :444
    lin = warp_sum(lin);
  }
  if (lane == 0) a.comp_lin[b] = lin;
}

// ---------------------------------------------------------------------------------------------
// sum_pooling[0] (CFFM.py:381) without the cube: sum_{c,p} o_i[h] o_j[c] = sum_i o_i[h] T_i,
// T_i = sum_{j>i} sum_c o_j[c].  One warp per sample.
__global__ void k_sumpool0(const float* __restrict__ rows, int B, int F, int K, float* __restrict__ t1, int t1_dim) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int b = blockIdx.x * wpb + warp;
  if (b >= B) return;
  float* S = sm + warp * 2 * F; float* T = S + F;
  const float* r = rows + (int64_t)b * F * K;
  for (int f = 0; f < F; ++f) {
    float s = 0.f;
    for (int c = lane; c < K; c += 32) s += r[f * K + c];
    s = warp_sum(s);
    if (lane == 0) S[f] = s;
  }
  __syncwarp();
  for (int i = lane; i < F; i += 32) { float t = 0.f; for (int j = i + 1; j < F; ++j) t += S[j]; T[i] = t; }
  __syncwarp();
  for (int h = lane; h < K; h += 32) {
    float v = 0.f;
    for (int i = 0; i < F - 1; ++i) v = fmaf(r[i * K + h], T[i], v);
    t1[(int64_t)b * t1_dim + h] = v;
  }
}

// sum_pooling[l+1] = reduce_sum(phi(Y_l), axis=[2,3]) (CFFM.py:390): one warp per (b,h).
template <int ACT>
__global__ void k_sumpool(const float* __restrict__ Y, int64_t BH, int H, int P, float* __restrict__ t1,
                          int t1_dim, int off) {
  const int lane = threadIdx.x & 31;
  const int64_t wg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wg >= BH) return;
  const float* run = Y + wg * H * P;
  const int n = H * P;
  float s = 0.f;
  for (int t = lane; t < n; t += 32) s += phi_f<ACT>(run[t]);
  s = warp_sum(s);
  if (lane == 0) {
    const int64_t b = wg / H; const int h = (int)(wg - b * H);
    t1[b * t1_dim + off + h] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// Head: dense(32), dense(1), beta_outer (CFFM.py:409-414), add_n (:453), prediction and the
// per-sample loss term / unscaled dLoss/dout (:486-514).  One warp per sample, lane = hidden unit.
struct HeadArgs {
  int B, t1_dim, inner_conv, outer_conv, loss_type, l2mode;
  const float *t1, *W1, *b1, *W2, *b2, *bias;
  const float *comp_inner, *comp_lin;
  float beta, invB;
  float *comp_outer, *out, *pred;
  const float* labels;   // null for scoring
  float *loss_terms, *gout;
};

__global__ void k_head_fwd(const HeadArgs a) {
  const int lane = threadIdx.x & 31;
  const int b = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (b >= a.B) return;
  float out = 0.f;
  if (a.inner_conv) out = a.comp_inner[b];
  if (a.outer_conv) {
    const float* t = a.t1 + (int64_t)b * a.t1_dim;
    float h = 0.f;
    for (int i = 0; i < a.t1_dim; ++i) h = fmaf(__ldg(t + i), __ldg(a.W1 + i * 32 + lane), h);
    h += __ldg(a.b1 + lane);
    float fin = warp_sum(h * __ldg(a.W2 + lane)) + __ldg(a.b2);
    fin *= a.beta;
    if (lane == 0) a.comp_outer[b] = fin;
    out += fin;
  }
  out += a.comp_lin[b];
  out += __ldg(a.bias);
  if (lane != 0) return;
  a.out[b] = out;
  const bool logl = a.loss_type == CFFM_LOSS_LOG;
  const float p = logl ? 1.f / (1.f + expf(-out)) : out;
  a.pred[b] = p;
  if (!a.labels) return;
  const float y = a.labels[b];
  const float d = out - y;
  float term, g;
  const float eps = 1e-7f;
  switch (a.loss_type) {
    case CFFM_LOSS_SQUARE:                                                   // :493 (lamda == 0) / :489 l2_loss (lamda > 0)
      term = a.l2mode ? 0.5f * d * d : d * d; g = d; break;
    case CFFM_LOSS_MSE: term = d * d; g = 2.f * d; break;                    // :506
    case CFFM_LOSS_MAE: term = fabsf(d); g = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); break;  // :508
    case CFFM_LOSS_LOG: {                                                    // :496-504
      term = -y * logf(p + eps) - (1.f - y) * logf(1.f - p + eps);
      g = (-y / (p + eps) + (1.f - y) / (1.f - p + eps)) * p * (1.f - p);
    } break;
    default: {                                                               // :511-513 hybrid
      const float ll = -y * logf(out + eps) - (1.f - y) * logf(1.f - out + eps);
      term = 0.25f * d * d + 0.5f * a.invB * ll;
      g = 0.5f * d + 0.5f * a.invB * (-y / (out + eps) + (1.f - y) / (1.f - out + eps));
    } break;
  }
  a.loss_terms[b] = term;
  a.gout[b] = g;
}

// Deterministic single-block sum of the loss terms -> scalars[0].
__global__ void k_loss_sum(const float* __restrict__ terms, int B, float* __restrict__ scalars) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) s += terms[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) scalars[0] = v;
  }
}

// loss value and the scale of dLoss/dout from the (global) loss sum; then gout *= scale.
__global__ void k_loss_finish(float* __restrict__ scalars, int loss_type, float invB, float* __restrict__ gout,
                              int B, float* __restrict__ loss_out, float reg_inner, float reg_outer) {
  const float sum = scalars[0];
  float loss, scale;
  if (loss_type == CFFM_LOSS_SQUARE && (reg_inner > 0.f || reg_outer > 0.f)) {  // :489-491: l2_loss + regularisers
    loss = sum + 0.5f * reg_inner * scalars[6] + 0.5f * reg_outer * scalars[7]; scale = 1.f;
  } else if (loss_type == CFFM_LOSS_SQUARE) { loss = sqrtf(sum * invB + 1e-10f); scale = invB / loss; }
  else if (loss_type == CFFM_LOSS_HYBRID) { loss = sum; scale = 1.f; }
  else { loss = sum * invB; scale = invB; }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) gout[i] *= scale;
  if (i == 0) { scalars[1] = loss; scalars[2] = scale; loss_out[0] = loss; }
}

void launch_loss_sum(Model* m, int B, cudaStream_t s) {
  CFFM_PROF(m, "loss_sum", s);
  k_loss_sum<<<1, 1024, 0, s>>>(m->loss_terms, B, m->scalars);
  m->launches++;
}
void launch_loss_finish(Model* m, int B, cudaStream_t s) {
  const float invB = 1.f / (float)((int64_t)B * m->world);
  CFFM_PROF(m, "loss_finish", s);
  const bool l2 = m->cfg.lamda > 0.f && m->cfg.loss_type == CFFM_LOSS_SQUARE;
  k_loss_finish<<<ceil_div(B, 256), 256, 0, s>>>(m->scalars, m->cfg.loss_type, invB, m->gout, B, m->loss_out,
                                                 l2 && m->cfg.inner_conv ? m->cfg.lamda : 0.f,
                                                 l2 && m->cfg.outer_conv ? m->cfg.lamda_att : 0.f);
  m->launches++;
}

// ---------------------------------------------------------------------------------------------
static size_t inner_smem(int F, int P, int K, int wpb) {
  return sizeof(float) * ((size_t)((2 * P + 3) & ~3) + (size_t)wpb * ((size_t)F * K + ((2 * F + 3) & ~3)));
}

int forward_setup_attrs(Model* m) {
  // opt in to large dynamic shared memory for the per-sample kernels
  const int maxsm = 200 * 1024;
  CFFM_DISPATCH_ACT(m->cfg.activation,
    CFFM_CUDA_OK(m, cudaFuncSetAttribute(k_inner_linear_fwd<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm)));
  return CFFM_OK;
}

template <class Prob>
static void launch_gemm(Model* m, const Prob& prob, int nsplit, cudaStream_t s) {
  const int tiles = ceil_div(prob.M, GBM) * ceil_div(prob.N, GBN);
  dim3 grid(tiles, nsplit);
  k_gemm_simt<Prob><<<grid, GTHREADS, 0, s>>>(prob);
  m->launches++;
}

static int ilog2(int x) { int l = 0; while ((1 << (l + 1)) <= x) ++l; return l; }

int run_forward(Model* m, const int32_t* ids_in, const float* labels, int64_t B64, cudaStream_t s) {
  const bool for_train = labels != nullptr;
  const int B = (int)B64, F = m->F, P = m->P;
  const DenseLayout& L = m->lay;
  const float* w = m->dense_w;
  // row-sharded tables: fetch the rows of this batch from their owners; the step then runs on that list
  TableView tv; tv.inner = m->inner_tab; tv.outer = m->outer_tab; tv.fbias = m->fbias_tab; tv.ids = ids_in;
  if (sharded(m)) { int r = shard_forward_exchange(m, ids_in, B64, s, &tv); if (r != CFFM_OK) return r; }
  m->view = tv;
  const int32_t* ids = tv.ids;
  // ---- inner path + linear term ----
  cudaStream_t side = s;
  {
    InnerLinArgs a;
    a.ids = ids; a.B = B; a.F = F; a.P = P; a.K = m->cfg.inner_conv ? m->Ki : 4; a.lgK = ilog2(a.K);
    a.inner_conv = m->cfg.inner_conv; a.linear_att = m->cfg.linear_att;
    a.tab = tv.inner; a.fbias = tv.fbias;
    a.cw = w + L.iconv_w; a.cb = w + L.iconv_b; a.Wd = w + L.din_k; a.bd = w + L.din_b;
    a.attW = w + L.att_W; a.attb = w + L.att_b; a.w3 = w + L.d3_k; a.b3 = w + L.d3_b;
    a.pair_i = m->pair_i; a.pair_j = m->pair_j; a.tau = m->cfg.lamda_att;
    a.comp_inner = m->comp_inner; a.comp_lin = m->comp_lin; a.fb_out = for_train ? m->fb_buf : nullptr;
    int wpb = 8;
    while (wpb > 1 && inner_smem(F, P, a.K, wpb) > 96 * 1024) wpb >>= 1;
    const size_t smem = inner_smem(F, P, a.K, wpb);
    side = side_fork(m, s);   // joined before the head
    CFFM_PROF(m, "inner_linear_fwd", side);
    CFFM_DISPATCH_ACT(m->cfg.activation,
      k_inner_linear_fwd<ACT><<<ceil_div(B, wpb), wpb * 32, smem, side>>>(a));
    m->launches++;
  }
  // ---- outer path ----
  if (m->cfg.outer_conv) {
    const int K = m->Ko;
    { CFFM_PROF(m, "gather_outer", s); launch_gather_rows(tv.outer, ids, (int64_t)B * F, K, m->outer_rows, s); }
    m->launches++;
    { CFFM_PROF(m, "sumpool0", s);
    k_sumpool0<<<ceil_div(B, 8), 256, 8 * 2 * F * sizeof(float), s>>>(m->outer_rows, B, F, K, m->t1, m->t1_dim); }
    m->launches++;
    if (tc_path(m)) {
      int r = tc_conv_forward(m, B, s);
      if (r != CFFM_OK) return r;
    }
    int off = K;
    for (int l = 0; l < (tc_path(m) ? 0 : m->n_live); ++l) {
      const int Hin = K >> l, Ho = Hin >> 1;
      const std::string tag_f = "conv_fwd_l" + std::to_string(l), tag_s = "sumpool_l" + std::to_string(l + 1);
      if (l == 0) {
        CFFM_PROF(m, tag_f.c_str(), s);
        Conv0FwdProb pr;
        pr.M = B * Ho * Ho; pr.N = P; pr.Kd = 4 * P;
        pr.g.P = P; pr.g.F = F; pr.g.K = K; pr.g.lgHo = ilog2(Ho);
        pr.g.rows = m->outer_rows; pr.g.pair_i = m->pair_i; pr.g.pair_j = m->pair_j;
        pr.W = w + L.conv_w[0]; pr.bias = w + L.conv_b[0]; pr.Yout = m->Y[0];
        launch_gemm(m, pr, 1, s);
      } else {
        CFFM_PROF(m, tag_f.c_str(), s);
        CFFM_DISPATCH_ACT(m->cfg.activation, {
          ConvFwdProb<ACT> pr;
          pr.M = B * Ho * Ho; pr.N = P; pr.Kd = 4 * P;
          pr.g.P = P; pr.g.Hin = Hin; pr.g.lgHo = ilog2(Ho);
          pr.Yprev = m->Y[l - 1]; pr.W = w + L.conv_w[l]; pr.bias = w + L.conv_b[l]; pr.Yout = m->Y[l];
          launch_gemm(m, pr, 1, s);
        });
      }
      const int64_t BH = (int64_t)B * Ho;
      CFFM_PROF(m, tag_s.c_str(), s);
      CFFM_DISPATCH_ACT(m->cfg.activation,
        k_sumpool<ACT><<<ceil_div(BH * 32, 256), 256, 0, s>>>(m->Y[l], BH, Ho, P, m->t1, m->t1_dim, off));
      m->launches++;
      off += Ho;
    }
  }
  // ---- head + loss terms ----
  { int r = side_join(m, side, s); if (r != CFFM_OK) return r; }
  {
    HeadArgs a;
    a.B = B; a.t1_dim = m->t1_dim; a.inner_conv = m->cfg.inner_conv; a.outer_conv = m->cfg.outer_conv;
    a.loss_type = m->cfg.loss_type; a.l2mode = (m->cfg.lamda > 0.f && m->cfg.loss_type == CFFM_LOSS_SQUARE) ? 1 : 0;
    a.t1 = m->t1; a.W1 = w + L.d1_k; a.b1 = w + L.d1_b; a.W2 = w + L.d2_k; a.b2 = w + L.d2_b; a.bias = w + L.bias;
    a.comp_inner = m->comp_inner; a.comp_lin = m->comp_lin; a.beta = m->cfg.beta_outer;
    a.invB = 1.f / (float)((int64_t)B * m->world);
    a.comp_outer = m->comp_outer; a.out = m->out; a.pred = m->pred;
    a.labels = labels; a.loss_terms = m->loss_terms; a.gout = m->gout;
    CFFM_PROF(m, "head_fwd", s);
    k_head_fwd<<<ceil_div((int64_t)B * 32, 256), 256, 0, s>>>(a);
    m->launches++;
  }
  CFFM_CUDA_OK(m, cudaGetLastError());
  return CFFM_OK;
}

}  // namespace cffm
