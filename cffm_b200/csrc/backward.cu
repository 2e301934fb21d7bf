// Backward pass + update of one training step (what Optimizer.minimize adds to the graph,
// CFFM.py:517-529) [TF-1.14].  The maths is SURVEY.md Appendix A; every batch reduction is a
// fixed-order two-stage sum (chunk partials, then chunks in order) so a step is reproducible.
#include <stdio.h>

#include <string>
#include <vector>

#include "common.cuh"
#include "gemm_simt.cuh"
#include "kernels.h"
#include "model.h"

namespace cffm {

struct ReduceDesc { int64_t src_off; int64_t dst_off; int64_t n; int32_t C; int32_t pad; };

// ---------------------------------------------------------------------------------------------
// v_head[t] = beta * sum_n dense_1/kernel[t,n] * dense_2/kernel[n]: there is no activation between
// the two dense layers (CFFM.py:409-410), so d out / d t1[b,t] = g_b * v_head[t].
__global__ void k_head_prep(const float* __restrict__ W1, const float* __restrict__ W2, float beta, int t1_dim,
                            float* __restrict__ v) {
  for (int t = threadIdx.x; t < t1_dim; t += blockDim.x) {
    float s = 0.f;
    for (int n = 0; n < 32; ++n) s = fmaf(W1[t * 32 + n], W2[n], s);
    v[t] = beta * s;
  }
}

// partial[c][col] = sum over the rows of chunk c of (wgt[r] *) X[r, col]
__global__ void k_colsum(const float* __restrict__ X, int64_t rows, int ld, int n, const float* __restrict__ wgt,
                         float* __restrict__ partial, int C) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const int c = blockIdx.y;
  const int64_t rpc = (rows + C - 1) / C;
  const int64_t r0 = (int64_t)c * rpc;
  const int64_t r1 = r0 + rpc < rows ? r0 + rpc : rows;
  float s = 0.f;
  if (col < n)
    for (int64_t r = r0 + warp; r < r1; r += 8) {
      const float x = X[r * ld + col];
      s += wgt ? wgt[r] * x : x;
    }
  red[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && col < n) {
    float t = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) t += red[w8][lane];
    partial[(int64_t)c * n + col] = t;
  }
}

__global__ void k_reduce_partials(const ReduceDesc* __restrict__ descs, const float* __restrict__ partials,
                                  float* __restrict__ g) {
  const ReduceDesc d = descs[blockIdx.y];
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < d.n; j += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < d.C; ++c) s += partials[d.src_off + (int64_t)c * d.n + j];
    g[d.dst_off + j] = s;
  }
}

// Gradients that follow from the aux batch sums q[t] = sum_b g_b t1[b,t], G = sum_b g_b and the
// column sums of the per-sample rows written by k_inner_linear_bwd.
struct HeadGradArgs {
  int F, t1_dim, inner_conv, outer_conv, linear_att;
  float beta;
  const float *W1, *b1, *W2;
  const float* aux;  // q[t1_dim], G, pad3, rowsums[2F+6]
  float* g;          // dense gradient block
  int64_t d1_k, d1_b, d2_k, d2_b, bias, din_b, d3_k, d3_b, att_b, iconv_w, iconv_b;
};
__global__ void k_head_grads(const HeadGradArgs a) {
  const float* q = a.aux;
  const float G = a.aux[a.t1_dim];
  const float* rs = a.aux + a.t1_dim + 4;
  const int tid = threadIdx.x;
  if (a.outer_conv) {
    for (int e = tid; e < a.t1_dim * 32; e += blockDim.x) a.g[a.d1_k + e] = a.beta * q[e >> 5] * a.W2[e & 31];
    if (tid < 32) {
      a.g[a.d1_b + tid] = a.beta * G * a.W2[tid];
      float s = G * a.b1[tid];
      for (int t = 0; t < a.t1_dim; ++t) s = fmaf(q[t], a.W1[t * 32 + tid], s);
      a.g[a.d2_k + tid] = a.beta * s;
    }
    if (tid == 0) a.g[a.d2_b] = a.beta * G;
  }
  if (tid == 0) a.g[a.bias] = G;
  if (a.inner_conv) {
    if (tid == 0) a.g[a.din_b] = G;
    if (tid < 4) a.g[a.iconv_w + tid] = rs[2 * a.F + tid];
    if (tid < 2) a.g[a.iconv_b + tid] = rs[2 * a.F + 4 + tid];
  }
  if (a.linear_att) {
    if (tid == 0) a.g[a.d3_b] = G;
    for (int f = tid; f < a.F; f += blockDim.x) { a.g[a.att_b + f] = rs[f]; a.g[a.d3_k + f] = rs[a.F + f]; }
  }
}

// dY of the last live conv layer: X_{d-1} = phi(Y_{d-2}) only feeds sum_pooling[d-1].
template <int ACT>
__global__ void k_dy_top(const float* __restrict__ Y, const float* __restrict__ gout, const float* __restrict__ v_lvl,
                         int H, int P, int64_t total, float* __restrict__ dY) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int64_t pix = e / P;              // (b*H + h)*H + w
  const int64_t bh = pix / H;
  const int h = (int)(bh % H);
  const int64_t b = bh / H;
  dY[e] = gout[b] * v_lvl[h] * phi_df<ACT>(Y[e]);
}

// ---------------------------------------------------------------------------------------------
// Layer-0 data gradient contracted straight into the outer-embedding rows:
//   dX0[2h+dh,2w+dw,p] = sum_q dY0[h,w,q] W0[dh,dw,p,q] + dsp0[2h+dh]
//   d o_i[a] += sum_c dX0[a,c,p] o_j[c] ;  d o_j[c] += sum_a dX0[a,c,p] o_i[a]
// One CTA per sample, one thread per output position (h,w) holding the 2x2 taps of PG pairs; the
// cube gradient never exists in memory.  The reductions over w (shuffles) and h (shared memory,
// fixed order) are deterministic.
constexpr int D0_PG = 4, D0_QC = 16;
struct Dgrad0Args {
  const float *dY0, *W0, *rows, *gout, *v_head;
  const int *pair_i, *pair_j;
  float* g_rows;
  int F, P, K, lgHo;
};
// (K = 64: one thread per output position is 1024 threads, so the kernel has to fit 64 registers)
__global__ void __launch_bounds__(1024, 1) k_dgrad0(const Dgrad0Args a) {
  extern __shared__ __align__(16) float sm[];
  const int T = blockDim.x, t = threadIdx.x, K = a.K, F = a.F, P = a.P;
  const int Ho = 1 << a.lgHo;
  float* o = sm;
  float* dOi = o + F * K;
  float* dOj = dOi + F * K;
  float* dYs = dOj + F * K;                  // [QC][T+1]
  float* Ws = dYs + D0_QC * (T + 1);         // [QC][PG*4]
  float* red = Ws + D0_QC * D0_PG * 4;       // [PG][Ho][K]
  const int b = blockIdx.x;
  const int h = t >> a.lgHo, w = t & (Ho - 1);
  const unsigned mask = T >= 32 ? 0xffffffffu : ((1u << T) - 1u);
  for (int e = t; e < F * K; e += T) { o[e] = a.rows[(int64_t)b * F * K + e]; dOi[e] = 0.f; dOj[e] = 0.f; }
  const float gb = a.gout[b];
  const float dsp0 = gb * a.v_head[2 * h], dsp1 = gb * a.v_head[2 * h + 1];
  const float* dYb = a.dY0 + (int64_t)b * T * P;
  for (int p0 = 0; p0 < P; p0 += D0_PG) {
    float acc[D0_PG * 4];
#pragma unroll
    for (int c = 0; c < D0_PG * 4; ++c) acc[c] = 0.f;
    for (int q0 = 0; q0 < P; q0 += D0_QC) {
      __syncthreads();
#pragma unroll
      for (int r = 0; r < D0_QC; ++r) {
        const int e = t + r * T;
        const int qq = e & (D0_QC - 1), tt = e >> 4;
        const int q = q0 + qq;
        dYs[qq * (T + 1) + tt] = q < P ? __ldg(dYb + (int64_t)tt * P + q) : 0.f;
      }
      for (int e = t; e < D0_QC * D0_PG * 4; e += T) {
        const int qq = e & (D0_QC - 1), c = e >> 4;
        const int p = p0 + (c >> 2), tap = c & 3, q = q0 + qq;
        Ws[qq * (D0_PG * 4) + c] = (p < P && q < P) ? __ldg(a.W0 + ((int64_t)tap * P + p) * P + q) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int qq = 0; qq < D0_QC; ++qq) {
        const float d = dYs[qq * (T + 1) + t];
#pragma unroll
        for (int j4 = 0; j4 < D0_PG; ++j4) {
          const float4 wv = *reinterpret_cast<const float4*>(Ws + qq * (D0_PG * 4) + 4 * j4);
          acc[4 * j4 + 0] = fmaf(d, wv.x, acc[4 * j4 + 0]);
          acc[4 * j4 + 1] = fmaf(d, wv.y, acc[4 * j4 + 1]);
          acc[4 * j4 + 2] = fmaf(d, wv.z, acc[4 * j4 + 2]);
          acc[4 * j4 + 3] = fmaf(d, wv.w, acc[4 * j4 + 3]);
        }
      }
    }
#pragma unroll
    for (int pp = 0; pp < D0_PG; ++pp) {
      const int p = p0 + pp;
      float ci0 = 0.f, ci1 = 0.f, cj0 = 0.f, cj1 = 0.f;
      int i = 0;
      if (p < P) {
        i = a.pair_i[p];
        const int j = a.pair_j[p];
        const float D00 = acc[pp * 4 + 0] + dsp0, D01 = acc[pp * 4 + 1] + dsp0;   // tap = dh*2+dw
        const float D10 = acc[pp * 4 + 2] + dsp1, D11 = acc[pp * 4 + 3] + dsp1;
        const float oj0 = o[j * K + 2 * w], oj1 = o[j * K + 2 * w + 1];
        const float oi0 = o[i * K + 2 * h], oi1 = o[i * K + 2 * h + 1];
        ci0 = fmaf(D01, oj1, D00 * oj0); ci1 = fmaf(D11, oj1, D10 * oj0);
        cj0 = fmaf(D10, oi1, D00 * oi0); cj1 = fmaf(D11, oi1, D01 * oi0);
      }
      for (int off = Ho >> 1; off > 0; off >>= 1) {
        ci0 += __shfl_xor_sync(mask, ci0, off);
        ci1 += __shfl_xor_sync(mask, ci1, off);
      }
      if (w == 0 && p < P) { dOi[i * K + 2 * h] += ci0; dOi[i * K + 2 * h + 1] += ci1; }
      red[(pp * Ho + h) * K + 2 * w] = cj0;
      red[(pp * Ho + h) * K + 2 * w + 1] = cj1;
    }
    __syncthreads();
    if (t < K) {
      for (int pp = 0; pp < D0_PG; ++pp) {
        const int p = p0 + pp;
        if (p < P) {
          float s = 0.f;
          for (int hh = 0; hh < Ho; ++hh) s += red[(pp * Ho + hh) * K + t];
          dOj[a.pair_j[p] * K + t] += s;
        }
      }
    }
  }
  __syncthreads();
  for (int e = t; e < F * K; e += T) a.g_rows[(int64_t)b * F * K + e] = dOi[e] + dOj[e];
}

static size_t dgrad0_smem(int F, int K) {
  const int Ho = K / 2, T = Ho * Ho;
  return sizeof(float) * ((size_t)3 * F * K + (size_t)D0_QC * (T + 1) + D0_QC * D0_PG * 4 + (size_t)D0_PG * Ho * K);
}

// ---------------------------------------------------------------------------------------------
// Inner path + linear term backward, one warp per sample, forward recomputed.
// Writes the gradient rows of inner_embeddings / feature_bias for this sample and a row of
// per-sample terms whose column sums are dense gradients:
//   rowbuf[b] = [ dz/tau (F) | g*u (F) | d conv filter (4: t*2+o) | d conv bias (2) ]
struct InnerLinBwdArgs {
  const int32_t* ids; int B, F, P, K, lgK, n_small;
  int inner_conv, linear_att;
  const float *tab, *fbias, *cw, *cb, *Wd, *attW, *attb, *w3;
  const int *pair_i, *pair_j;
  float tau;
  const float* gout;
  float *g_inner_rows, *g_bias_rows, *rowbuf;
};

template <int ACT>
__global__ void k_inner_linear_bwd(const InnerLinBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int F = a.F, K = a.K, P = a.P;
  int* s_pi = reinterpret_cast<int*>(sm);
  int* s_pj = s_pi + P;
  const int pairs_pad = (2 * P + 3) & ~3;
  float* wbase = sm + pairs_pad + (size_t)warp * (2 * F * K + 3 * F + ((4 - (3 * F) % 4) % 4));
  float* e = wbase; float* de = e + F * K; float* fb = de + F * K; float* z = fb + F; float* dzt = z + F;
  for (int t = threadIdx.x; t < P; t += blockDim.x) { s_pi[t] = a.pair_i[t]; s_pj[t] = a.pair_j[t]; }
  __syncthreads();
  const int b = blockIdx.x * wpb + warp;
  if (b >= a.B) return;
  const int32_t* id = a.ids + (int64_t)b * F;
  const float g = a.gout[b];
  float* rb = a.rowbuf + (int64_t)b * a.n_small;
  if (a.inner_conv) {
    const int K4 = K >> 2;
    for (int t = lane; t < F * K4; t += 32) {
      const int f = t / K4, c = t - f * K4;
      const float4 v = __ldg(reinterpret_cast<const float4*>(a.tab + (int64_t)__ldg(id + f) * K) + c);
      *reinterpret_cast<float4*>(e + f * K + 4 * c) = v;
      *reinterpret_cast<float4*>(de + f * K + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    const int o = lane & 1;
    const float wt0 = __ldg(a.cw + o), wt1 = __ldg(a.cw + 2 + o), cbo = __ldg(a.cb + o);
    const float w_self = __ldg(a.cw + o * 2 + o), w_oth = __ldg(a.cw + o * 2 + (1 - o));  // w_t[t], w_t[1-t], t = o
    const int PK = P * K;
    const int nph = K >= 32 ? 1 : 32 / K;
    float gw0 = 0.f, gw1 = 0.f, gc = 0.f;
    auto pass = [&](int base, float wdv) {
      const int idx = base + lane;
      const bool ok = idx < PK;
      float I = 0.f, ei = 0.f, ej = 0.f;
      int oi = 0, oj = 0;
      if (ok) {
        const int p = idx >> a.lgK, k = idx & (K - 1);
        oi = s_pi[p] * K + k; oj = s_pj[p] * K + k;
        ei = e[oi]; ej = e[oj]; I = ei * ej;
      }
      const float A = ok ? act_f<ACT>(I) : 0.f;
      const float Ao = __shfl_xor_sync(0xffffffffu, A, 1);
      const float a0 = o ? Ao : A, a1 = o ? A : Ao;
      const float y = fmaf(a1, wt1, a0 * wt0) + cbo;
      const float dR = ok ? g * wdv : 0.f;
      const float dYv = dR * phi_df<ACT>(y);
      gw0 = fmaf(dYv, a0, gw0); gw1 = fmaf(dYv, a1, gw1); gc += dYv;
      const float dYo = __shfl_xor_sync(0xffffffffu, dYv, 1);
      const float dRo = __shfl_xor_sync(0xffffffffu, dR, 1);
      const int amax = (a0 >= a1) ? 0 : 1;  // max-pool gradient goes to the first maximum (Q15)
      const float dA = fmaf(dYv, w_self, dYo * w_oth) + ((o == amax) ? dR + dRo : 0.f);
      const float dI = dA * act_df<ACT>(I);
      if (nph == 1) {
        if (ok) { de[oi] = fmaf(dI, ej, de[oi]); de[oj] = fmaf(dI, ei, de[oj]); }
      } else {  // several pairs per warp pass: take turns so that a column has one writer at a time
        for (int ph = 0; ph < nph; ++ph) {
          if (ok && (lane >> a.lgK) == ph) { de[oi] = fmaf(dI, ej, de[oi]); de[oj] = fmaf(dI, ei, de[oj]); }
          __syncwarp();
        }
      }
    };
    // The dense-layer weight of an element is the only global load of a pass and everything depends on it:
    // fetch it four passes (128 elements) ahead so that its latency is off the loop-carried path.
    float wn[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i2 = u * 32 + lane; wn[u] = i2 < PK ? __ldg(a.Wd + i2) : 0.f; }
    for (int base = 0; base < PK; base += 128) {
      float wc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        wc[u] = wn[u];
        const int i2 = base + 128 + u * 32 + lane;
        wn[u] = i2 < PK ? __ldg(a.Wd + i2) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (base + u * 32 < PK) pass(base + u * 32, wc[u]);
    }
#pragma unroll
    for (int off = 16; off > 1; off >>= 1) {
      gw0 += __shfl_xor_sync(0xffffffffu, gw0, off);
      gw1 += __shfl_xor_sync(0xffffffffu, gw1, off);
      gc += __shfl_xor_sync(0xffffffffu, gc, off);
    }
    if (lane < 2) { rb[2 * F + lane] = gw0; rb[2 * F + 2 + lane] = gw1; rb[2 * F + 4 + lane] = gc; }
    __syncwarp();
    float* dst = a.g_inner_rows + (int64_t)b * F * K;
    for (int t = lane; t < F * K; t += 32) dst[t] = de[t];
  } else if (lane < 6) {
    rb[2 * F + lane] = 0.f;
  }
  // ---- linear term ----
  for (int f = lane; f < F; f += 32) fb[f] = __ldg(a.fbias + __ldg(id + f));
  __syncwarp();
  if (a.linear_att) {
    float zmax = -INFINITY;
    for (int q = lane; q < F; q += 32) {
      float s = 0.f;
      for (int f = 0; f < F; ++f) s = fmaf(fb[f], __ldg(a.attW + f * F + q), s);
      s = (s + __ldg(a.attb + q)) / a.tau;
      z[q] = s; zmax = fmaxf(zmax, s);
    }
    zmax = warp_max(zmax);
    float se = 0.f;
    for (int q = lane; q < F; q += 32) { const float ex = expf(z[q] - zmax); z[q] = ex; se += ex; }
    se = warp_sum(se);
    float dot = 0.f;
    for (int q = lane; q < F; q += 32) {
      const float sq = z[q] / se; z[q] = sq;
      dot = fmaf(g * __ldg(a.w3 + q) * fb[q], sq, dot);
    }
    dot = warp_sum(dot);
    for (int q = lane; q < F; q += 32) {
      const float sq = z[q];
      const float ds = g * __ldg(a.w3 + q) * fb[q];
      const float d = sq * (ds - dot) / a.tau;
      dzt[q] = d; rb[q] = d; rb[F + q] = g * fb[q] * sq;
    }
    __syncwarp();
    for (int f = lane; f < F; f += 32) {
      float s = g * __ldg(a.w3 + f) * z[f];
      for (int q = 0; q < F; ++q) s = fmaf(dzt[q], __ldg(a.attW + f * F + q), s);
      a.g_bias_rows[(int64_t)b * F + f] = s;
    }
  } else {
    for (int f = lane; f < F; f += 32) { a.g_bias_rows[(int64_t)b * F + f] = g; rb[f] = 0.f; rb[F + f] = 0.f; }
  }
}

static size_t inner_bwd_smem(int F, int P, int K, int wpb) {
  const size_t pairs_pad = (2 * (size_t)P + 3) & ~(size_t)3;
  const size_t per_warp = 2 * (size_t)F * K + 3 * F + ((4 - (3 * F) % 4) % 4);
  return sizeof(float) * (pairs_pad + wpb * per_warp);
}

// d dense/kernel[idx] = sum_b g_b R[b, idx]: one thread per flattened (p,k) element,
// the batch split into gridDim.y chunks; R is recomputed from the table rows.
struct InnerDenseGradArgs {
  const int32_t* ids; int B, F, P, K, lgK;
  const float *tab, *cw, *cb, *gout;
  const int *pair_i, *pair_j;
  float* partial; int C;
};
template <int ACT>
__global__ void k_inner_dense_grad(const InnerDenseGradArgs a) {
  // a thread takes two neighbouring flattened elements (p, w, o = 0 / 1) = the two taps of one conv position
  const int idx = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int PK = a.P * a.K, K = a.K, F = a.F;
  if (idx >= PK) return;
  const int p = idx >> a.lgK, k = idx & (K - 1);
  const int fi = a.pair_i[p], fj = a.pair_j[p];
  const float w00 = __ldg(a.cw + 0), w01 = __ldg(a.cw + 1), w10 = __ldg(a.cw + 2), w11 = __ldg(a.cw + 3);
  const float c0 = __ldg(a.cb + 0), c1 = __ldg(a.cb + 1);
  const int c = blockIdx.y;
  const int rpc = (a.B + a.C - 1) / a.C;
  const int b0 = c * rpc, b1 = min(a.B, b0 + rpc);
  float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
  for (int b = b0; b < b1; ++b) {   // two dependent loads per sample (id, then row): keep four samples in flight
    const float2 ei = __ldg(reinterpret_cast<const float2*>(a.tab + (int64_t)__ldg(a.ids + (int64_t)b * F + fi) * K + k));
    const float2 ej = __ldg(reinterpret_cast<const float2*>(a.tab + (int64_t)__ldg(a.ids + (int64_t)b * F + fj) * K + k));
    const float a0 = act_f<ACT>(ei.x * ej.x), a1 = act_f<ACT>(ei.y * ej.y);
    const float y0 = fmaf(a1, w10, a0 * w00) + c0, y1 = fmaf(a1, w11, a0 * w01) + c1;
    const float mx = fmaxf(a0, a1), g = __ldg(a.gout + b);
    s0 = fmaf(g, phi_f<ACT>(y0) + mx, s0);
    s1 = fmaf(g, phi_f<ACT>(y1) + mx, s1);
  }
  *reinterpret_cast<float2*>(a.partial + (int64_t)c * PK + idx) = make_float2(s0, s1);
}

// d bias_W[f,q] = sum_b fb[b,f] * dz[b,q]/tau
__global__ void k_att_outer(const float* __restrict__ fb, const float* __restrict__ rowbuf, int B, int F, int n_small,
                            float* __restrict__ partial, int C) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= F * F) return;
  const int f = e / F, q = e - f * F;
  const int c = blockIdx.y;
  const int rpc = (B + C - 1) / C;
  const int b0 = c * rpc, b1 = min(B, b0 + rpc);
  float s = 0.f;
  for (int b = b0; b < b1; ++b) s = fmaf(fb[(int64_t)b * F + f], rowbuf[(int64_t)b * n_small + q], s);
  partial[(int64_t)c * F * F + e] = s;
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

static int ilog2(int x) { int l = 0; while ((1 << (l + 1)) <= x) ++l; return l; }

// chunk counts / split factors are fixed at allocation time (from max_batch)
struct TrainPlan {
  int Cb;                      // chunks over the batch for per-sample reductions
  int nsplit[kMaxConv];        // wgrad split-K factors
  int Cl[kMaxConv];            // chunks for the conv bias column sums
  int64_t off_q, off_G, off_rows, off_Wd, off_attW, off_wg[kMaxConv], off_bg[kMaxConv], total;
};
static TrainPlan make_plan(const Model* m) {
  TrainPlan pl;
  const int64_t B = m->max_batch, P = m->P, F = m->F;
  pl.Cb = (int)std::min<int64_t>(64, std::max<int64_t>(1, (B + 31) / 32));
  int64_t off = 0;
  auto take = [&](int64_t n, int C) { int64_t o = off; off += ((n * C + 3) & ~int64_t(3)); return o; };
  pl.off_q = take(m->t1_dim, pl.Cb);
  pl.off_G = take(1, pl.Cb);
  pl.off_rows = take(m->n_small, pl.Cb);
  pl.off_Wd = m->cfg.inner_conv ? take(P * m->Ki, pl.Cb) : 0;
  pl.off_attW = m->cfg.linear_att ? take(F * F, pl.Cb) : 0;
  for (int l = 0; l < (tc_path(m) ? 0 : m->n_live); ++l) {
    const int64_t Ho = m->Ko >> (l + 1);
    const int64_t rows = B * Ho * Ho;
    const int tiles = ceil_div(4 * P, GBM) * ceil_div(P, GBN);
    const int64_t ksteps = (rows + GBK - 1) / GBK;
    int ns = (int)std::max<int64_t>(1, std::min<int64_t>(ksteps, (2 * 148 + tiles - 1) / tiles));
    pl.nsplit[l] = ns;
    pl.Cl[l] = (int)std::min<int64_t>(64, std::max<int64_t>(1, (rows + 63) / 64));
    pl.off_wg[l] = take(4 * P * P, ns);
    pl.off_bg[l] = take(P, pl.Cl[l]);
  }
  pl.total = off;
  return pl;
}

int model_alloc_train(Model* m) {
  if (m->train_ready) return CFFM_OK;
  const int64_t B = m->max_batch, F = m->F;
  CFFM_CUDA_OK(m, cudaSetDevice(m->device));
  TRY(dmalloc(m, &m->gout, B));
  if (m->cfg.outer_conv) {
    if (tc_path(m)) {
      TRY(tc_alloc(m, true));
    } else {
      for (int l = 0; l < m->n_live; ++l) {
        const int64_t H = m->Ko >> (l + 1);
        TRY(dmalloc(m, &m->dY[l], B * H * H * m->P));
      }
    }
    TRY(dmalloc(m, &m->g_outer_rows, B * F * m->Ko));
    TRY(dmalloc(m, &m->v_head, m->t1_dim));
  }
  if (m->cfg.inner_conv) TRY(dmalloc(m, &m->g_inner_rows, B * F * m->Ki));
  TRY(dmalloc(m, &m->g_bias_rows, B * F));
  TRY(dmalloc(m, &m->rowbuf, B * m->n_small));
  const TrainPlan pl = make_plan(m);
  m->partials_cap = pl.total;
  TRY(dmalloc(m, &m->partials, pl.total));
  // reduction descriptors
  std::vector<ReduceDesc> d;
  const DenseLayout& L = m->lay;
  auto add = [&](int64_t src, int64_t dst, int64_t n, int C) { ReduceDesc r; r.src_off = src; r.dst_off = dst; r.n = n; r.C = C; r.pad = 0; d.push_back(r); };
  if (m->cfg.outer_conv) add(pl.off_q, m->aux_off, m->t1_dim, pl.Cb);
  add(pl.off_G, m->aux_off + m->t1_dim, 1, pl.Cb);
  add(pl.off_rows, m->aux_off + m->t1_dim + 4, m->n_small, pl.Cb);
  if (m->cfg.inner_conv) add(pl.off_Wd, L.din_k, (int64_t)m->P * m->Ki, pl.Cb);
  if (m->cfg.linear_att) add(pl.off_attW, L.att_W, F * F, pl.Cb);
  for (int l = 0; l < (tc_path(m) ? 0 : m->n_live); ++l) {  // bf16: conv_tc.cu writes these
    add(pl.off_wg[l], L.conv_w[l], 4ll * m->P * m->P, pl.nsplit[l]);
    add(pl.off_bg[l], L.conv_b[l], m->P, pl.Cl[l]);
  }
  m->n_reduce_descs = (int)d.size();
  TRY(dmalloc(m, (ReduceDesc**)&m->reduce_descs, (int64_t)d.size()));
  CFFM_CUDA_OK(m, cudaMemcpy(m->reduce_descs, d.data(), sizeof(ReduceDesc) * d.size(), cudaMemcpyHostToDevice));
  // sparse update scratch; under data parallelism the global batch is updated on every rank
  m->upd_cap = (int64_t)m->world * B * F;
  if (sparse_work_alloc(&m->sw, m->upd_cap, (m->cfg.inner_conv ? m->Ki : 0) + (m->cfg.outer_conv ? m->Ko : 0) + 1, &m->err) != CFFM_OK) return CFFM_ERR_NOMEM;
  if (m->world > 1 && !sharded(m)) {
    TRY(dmalloc(m, &m->all_ids, m->upd_cap));
    if (m->cfg.inner_conv) TRY(dmalloc(m, &m->all_g_inner, m->upd_cap * m->Ki));
    if (m->cfg.outer_conv) TRY(dmalloc(m, &m->all_g_outer, m->upd_cap * m->Ko));
    TRY(dmalloc(m, &m->all_g_bias, m->upd_cap));
  }
  m->train_ready = true;
  return CFFM_OK;
}

int backward_setup_attrs(Model* m) {
  const int maxsm = 200 * 1024;
  CFFM_DISPATCH_ACT(m->cfg.activation,
    CFFM_CUDA_OK(m, cudaFuncSetAttribute(k_inner_linear_bwd<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm)));
  CFFM_CUDA_OK(m, cudaFuncSetAttribute(k_dgrad0, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
  return CFFM_OK;
}

template <class Prob>
static void launch_gemm(Model* m, const Prob& prob, int nsplit, cudaStream_t s) {
  const int tiles = ceil_div(prob.M, GBM) * ceil_div(prob.N, GBN);
  dim3 grid(tiles, nsplit);
  k_gemm_simt<Prob><<<grid, GTHREADS, 0, s>>>(prob);
  m->launches++;
}

static void launch_colsum(Model* m, const float* X, int64_t rows, int ld, int n, const float* wgt, float* partial, int C,
                          cudaStream_t s) {
  CFFM_PROF(m, "colsum", s);
  dim3 grid(ceil_div(n, 32), C);
  k_colsum<<<grid, 256, 0, s>>>(X, rows, ld, n, wgt, partial, C);
  m->launches++;
}

// Everything after the forward pass of a training step.  The loss sum is already in scalars[0].
int run_backward_update(Model* m, const int32_t* ids_in, const float* labels, int64_t B64, cudaStream_t s) {
  (void)labels;
  // the rows the forward pass computed on (row-sharded tables: the received list and the renumbered ids)
  const TableView tv = m->view;
  const int32_t* ids = sharded(m) ? tv.ids : ids_in;
  const int B = (int)B64, F = m->F, P = m->P;
  const DenseLayout& L = m->lay;
  const float* w = m->dense_w;
  float* g = m->dense_g;
  const TrainPlan pl = make_plan(m);
  float* part = m->partials;
  const int act = m->cfg.activation;

  // ---- loss: (global) sum -> loss value, scale of dLoss/dout ----
  if (m->cfg.lamda > 0.f && m->cfg.loss_type == CFFM_LOSS_SQUARE) {  // regulariser terms of CFFM.py:489-491
    CFFM_PROF(m, "l2_reg_sums", s);
    if (m->cfg.inner_conv) launch_sumsq(m->inner_tab, m->Mloc * m->Ki, m->sumsq_partial, m->scalars + 6, s);
    if (m->cfg.outer_conv) launch_sumsq(m->outer_tab, m->Mloc * m->Ko, m->sumsq_partial, m->scalars + 7, s);
    m->launches += 4;
    if (sharded(m)) { int r = comm_allreduce_f32(m, m->scalars + 6, 2, s); if (r != CFFM_OK) return r; }   // shards -> whole tables
  }
  launch_loss_sum(m, B, s);
  if (m->world > 1) { int r = comm_allreduce_f32(m, m->scalars, 1, s); if (r != CFFM_OK) return r; }
  launch_loss_finish(m, B, s);

  // ---- inner path + linear term (side stream: independent of the conv stack until the partial sums are folded) ----
  const cudaStream_t side = side_fork(m, s);
  {
    InnerLinBwdArgs a;
    a.ids = ids; a.B = B; a.F = F; a.P = P; a.K = m->cfg.inner_conv ? m->Ki : 4; a.lgK = ilog2(a.K);
    a.n_small = m->n_small; a.inner_conv = m->cfg.inner_conv; a.linear_att = m->cfg.linear_att;
    a.tab = tv.inner; a.fbias = tv.fbias; a.cw = w + L.iconv_w; a.cb = w + L.iconv_b; a.Wd = w + L.din_k;
    a.attW = w + L.att_W; a.attb = w + L.att_b; a.w3 = w + L.d3_k;
    a.pair_i = m->pair_i; a.pair_j = m->pair_j; a.tau = m->cfg.lamda_att; a.gout = m->gout;
    a.g_inner_rows = m->g_inner_rows; a.g_bias_rows = m->g_bias_rows; a.rowbuf = m->rowbuf;
    int wpb = 8;
    while (wpb > 1 && inner_bwd_smem(F, P, a.K, wpb) > 96 * 1024) wpb >>= 1;
    { CFFM_PROF(m, "inner_linear_bwd", side);
    CFFM_DISPATCH_ACT(act, k_inner_linear_bwd<ACT><<<ceil_div(B, wpb), wpb * 32, inner_bwd_smem(F, P, a.K, wpb), side>>>(a)); }
    m->launches++;
    launch_colsum(m, m->rowbuf, B, m->n_small, m->n_small, nullptr, part + pl.off_rows, pl.Cb, side);
    if (m->cfg.inner_conv) {
      InnerDenseGradArgs d;
      d.ids = ids; d.B = B; d.F = F; d.P = P; d.K = m->Ki; d.lgK = ilog2(m->Ki);
      d.tab = tv.inner; d.cw = w + L.iconv_w; d.cb = w + L.iconv_b; d.gout = m->gout;
      d.pair_i = m->pair_i; d.pair_j = m->pair_j; d.partial = part + pl.off_Wd; d.C = pl.Cb;
      dim3 grid(ceil_div((int64_t)P * m->Ki / 2, 256), pl.Cb);
      CFFM_PROF(m, "inner_dense_grad", side);
      CFFM_DISPATCH_ACT(act, k_inner_dense_grad<ACT><<<grid, 256, 0, side>>>(d));
      m->launches++;
    }
    if (m->cfg.linear_att) {
      dim3 grid(ceil_div(F * F, 256), pl.Cb);
      CFFM_PROF(m, "att_outer", side);
      k_att_outer<<<grid, 256, 0, side>>>(m->fb_buf, m->rowbuf, B, F, m->n_small, part + pl.off_attW, pl.Cb);
      m->launches++;
    }
  }
  // one GPU: the sort of the step's ids needs nothing but the ids
  const bool sort_on_side = m->world == 1 && !sharded(m);
  if (sort_on_side) {
    int r;
    { CFFM_PROF(m, "sort_segments", side); r = sparse_sort_segments(&m->sw, ids, (int64_t)B * F, m->M, side, &m->launches); }
    if (r != CFFM_OK) { m->err = "sparse_sort_segments failed"; return r; }
  }
  // ---- head ----
  launch_colsum(m, m->gout, B, 1, 1, nullptr, part + pl.off_G, pl.Cb, s);
  if (m->cfg.outer_conv) {
    const int K = m->Ko;
    { CFFM_PROF(m, "head_prep", s);
    k_head_prep<<<1, 128, 0, s>>>(w + L.d1_k, w + L.d2_k, m->cfg.beta_outer, m->t1_dim, m->v_head); }
    m->launches++;
    launch_colsum(m, m->t1, B, m->t1_dim, m->t1_dim, m->gout, part + pl.off_q, pl.Cb, s);
    // offsets of the pooling levels inside t1
    int lvl_off[kMaxConv + 1]; lvl_off[0] = 0;
    for (int l = 0; l < m->conv_depth; ++l) lvl_off[l + 1] = lvl_off[l] + (K >> l);
    if (tc_path(m)) {
      int r = tc_conv_backward(m, B, s);
      if (r != CFFM_OK) return r;
    }
    // ---- top of the conv stack ----
    if (!tc_path(m)) {
      const int l = m->n_live - 1;
      const int H = K >> (l + 1);
      const int64_t total = (int64_t)B * H * H * P;
      CFFM_PROF(m, "dy_top", s);
      CFFM_DISPATCH_ACT(act, k_dy_top<ACT><<<ceil_div(total, 256), 256, 0, s>>>(
          m->Y[l], m->gout, m->v_head + lvl_off[l + 1], H, P, total, m->dY[l]));
      m->launches++;
    }
    for (int l = (tc_path(m) ? -1 : m->n_live - 1); l >= 0; --l) {
      const int Hin = K >> l, Ho = Hin >> 1;
      const int rows = B * Ho * Ho;
      launch_colsum(m, m->dY[l], rows, P, P, nullptr, part + pl.off_bg[l], pl.Cl[l], s);
      const std::string tag_w = "conv_wgrad_l" + std::to_string(l), tag_d = "conv_dgrad_l" + std::to_string(l);
      if (l > 0) {
        CFFM_DISPATCH_ACT(act, {
          ConvWgradProb<ACT> wp;
          wp.M = 4 * P; wp.N = P; wp.Kd = rows;
          wp.g.P = P; wp.g.Hin = Hin; wp.g.lgHo = ilog2(Ho);
          wp.Yprev = m->Y[l - 1]; wp.dY = m->dY[l]; wp.partial = part + pl.off_wg[l];
          { CFFM_PROF(m, tag_w.c_str(), s); launch_gemm(m, wp, pl.nsplit[l], s); }
          ConvDgradProb<ACT> dp;
          dp.M = rows; dp.N = 4 * P; dp.Kd = P;
          dp.g.P = P; dp.g.Hin = Hin; dp.g.lgHo = ilog2(Ho);
          dp.dY = m->dY[l]; dp.W = w + L.conv_w[l]; dp.Yprev = m->Y[l - 1]; dp.dYprev = m->dY[l - 1];
          dp.gout = m->gout; dp.v_head = m->v_head; dp.sp_off = lvl_off[l];
          { CFFM_PROF(m, tag_d.c_str(), s); launch_gemm(m, dp, 1, s); }
        });
      } else {
        Conv0WgradProb wp;
        wp.M = 4 * P; wp.N = P; wp.Kd = rows;
        wp.g.P = P; wp.g.F = F; wp.g.K = K; wp.g.lgHo = ilog2(Ho);
        wp.g.rows = m->outer_rows; wp.g.pair_i = m->pair_i; wp.g.pair_j = m->pair_j;
        wp.dY = m->dY[0]; wp.partial = part + pl.off_wg[0];
        { CFFM_PROF(m, tag_w.c_str(), s); launch_gemm(m, wp, pl.nsplit[0], s); }
        Dgrad0Args da;
        da.dY0 = m->dY[0]; da.W0 = w + L.conv_w[0]; da.rows = m->outer_rows; da.gout = m->gout; da.v_head = m->v_head;
        da.pair_i = m->pair_i; da.pair_j = m->pair_j; da.g_rows = m->g_outer_rows;
        da.F = F; da.P = P; da.K = K; da.lgHo = ilog2(Ho);
        CFFM_PROF(m, tag_d.c_str(), s);
        k_dgrad0<<<B, Ho * Ho, dgrad0_smem(F, K), s>>>(da);
        m->launches++;
      }
    }
  }
  // ---- fold the partial sums into the dense gradient block, then the derived head gradients ----
  { int r = side_join(m, side, s); if (r != CFFM_OK) return r; }
  {
    dim3 grid(2 * 148, m->n_reduce_descs);
    { CFFM_PROF(m, "reduce_partials", s);
    k_reduce_partials<<<grid, 256, 0, s>>>((const ReduceDesc*)m->reduce_descs, part, g); }
    m->launches++;
    HeadGradArgs h;
    h.F = F; h.t1_dim = m->t1_dim; h.inner_conv = m->cfg.inner_conv; h.outer_conv = m->cfg.outer_conv;
    h.linear_att = m->cfg.linear_att; h.beta = m->cfg.beta_outer;
    h.W1 = w + L.d1_k; h.b1 = w + L.d1_b; h.W2 = w + L.d2_k; h.aux = g + m->aux_off; h.g = g;
    h.d1_k = L.d1_k; h.d1_b = L.d1_b; h.d2_k = L.d2_k; h.d2_b = L.d2_b; h.bias = L.bias; h.din_b = L.din_b;
    h.d3_k = L.d3_k; h.d3_b = L.d3_b; h.att_b = L.att_b; h.iconv_w = L.iconv_w; h.iconv_b = L.iconv_b;
    CFFM_PROF(m, "head_grads", s);
    k_head_grads<<<1, 256, 0, s>>>(h);
    m->launches++;
  }
  // ---- data parallel: sum dense gradients, gather the touched rows of every rank ----
  const int32_t* upd_ids = ids;
  const float *gi = m->g_inner_rows, *go = m->g_outer_rows, *gbr = m->g_bias_rows;
  int64_t n_upd = (int64_t)B * F;
  if (sharded(m)) {
    // row-sharded tables: dense gradients are summed over the ranks; gradient rows go to the owners of the rows
    { CFFM_PROF(m, "dp_allreduce", s);
      int r = comm_allreduce_f32(m, g, L.total, s); if (r != CFFM_OK) return r; }
    int r = shard_backward_update(m, B64, s); if (r != CFFM_OK) return r;
    const int opt = m->cfg.optimizer;
    const float* lr_dev = opt == CFFM_OPT_ADAM ? m->scalars + 4 : nullptr;
    CFFM_PROF(m, "dense_adagrad", s);
    launch_dense_update(m->dense_w, opt == CFFM_OPT_SGD ? nullptr : m->dense_acc, m->dense_acc2, g, L.total, opt, m->cfg.lr, lr_dev, s);
    m->launches++;
    CFFM_CUDA_OK(m, cudaGetLastError());
    return CFFM_OK;
  }
  if (m->world > 1) {
    CFFM_PROF(m, "dp_allreduce_allgather", s);
    // the group is closed on every exit path (an open NCCL group inside a stream capture poisons the communicator)
    struct GroupGuard {
      Model* m; bool open;
      ~GroupGuard() { if (open) comm_group_end(m); }
    } guard{m, false};
    int r = comm_group_begin(m); if (r != CFFM_OK) return r;
    guard.open = true;
    r = comm_allreduce_f32(m, g, L.total, s); if (r != CFFM_OK) return r;
    r = comm_allgather(m, ids, m->all_ids, sizeof(int32_t) * n_upd, s); if (r != CFFM_OK) return r;
    if (m->cfg.inner_conv) { r = comm_allgather(m, gi, m->all_g_inner, sizeof(float) * n_upd * m->Ki, s); if (r != CFFM_OK) return r; }
    if (m->cfg.outer_conv) { r = comm_allgather(m, go, m->all_g_outer, sizeof(float) * n_upd * m->Ko, s); if (r != CFFM_OK) return r; }
    r = comm_allgather(m, gbr, m->all_g_bias, sizeof(float) * n_upd, s); if (r != CFFM_OK) return r;
    guard.open = false;
    r = comm_group_end(m); if (r != CFFM_OK) return r;
    upd_ids = m->all_ids; gi = m->all_g_inner; go = m->all_g_outer; gbr = m->all_g_bias;
    n_upd *= m->world;
  }
  // ---- sparse update of the three tables (one sort shared by all of them) ----
  {
    int r = CFFM_OK;
    if (!sort_on_side) { CFFM_PROF(m, "sort_segments", s); r = sparse_sort_segments(&m->sw, upd_ids, n_upd, m->M, s, &m->launches); }
    if (r != CFFM_OK) { m->err = "sparse_sort_segments failed"; return r; }
    const int opt = m->cfg.optimizer;
    const bool adam = opt == CFFM_OPT_ADAM;
    const bool l2 = m->cfg.lamda > 0.f && m->cfg.loss_type == CFFM_LOSS_SQUARE;
    if (adam) { launch_adam_tick(m->scalars, m->cfg.lr, s); m->launches++; }
    const float* lr_dev = adam ? m->scalars + 4 : nullptr;
    SparseTables t;
    t.rowmap = m->rowmap; t.M = m->Mloc;
    int j = 0;
    // Adam's sparse apply moves every row; the l2 regulariser makes the two embedding gradients dense (Q9:
    // the outer table is regularised by lamda_att)
    if (m->cfg.inner_conv) { t.tab[j] = m->inner_tab; t.acc[j] = m->inner_acc; t.acc2[j] = m->inner_acc2; t.grads[j] = gi; t.K[j] = m->Ki;
                             t.dense[j] = adam || l2; t.reg[j] = l2 ? m->cfg.lamda : 0.f; ++j; }
    if (m->cfg.outer_conv) { t.tab[j] = m->outer_tab; t.acc[j] = m->outer_acc; t.acc2[j] = m->outer_acc2; t.grads[j] = go; t.K[j] = m->Ko;
                             t.dense[j] = adam || l2; t.reg[j] = l2 ? m->cfg.lamda_att : 0.f; ++j; }
    t.tab[j] = m->fbias_tab; t.acc[j] = m->fbias_acc; t.acc2[j] = m->fbias_acc2; t.grads[j] = gbr; t.K[j] = 1; t.dense[j] = adam; ++j;
    if (opt == CFFM_OPT_SGD) for (int q = 0; q < 3; ++q) t.acc[q] = nullptr;
    { CFFM_PROF(m, "sparse_adagrad", s);
      launch_sparse_update(&m->sw, t, n_upd, opt, m->cfg.lr, lr_dev, s, &m->launches); }
    // ---- dense update ----
    CFFM_PROF(m, "dense_adagrad", s);
    launch_dense_update(m->dense_w, opt == CFFM_OPT_SGD ? nullptr : m->dense_acc, m->dense_acc2, g, L.total, opt, m->cfg.lr, lr_dev, s);
    m->launches++;
  }
  CFFM_CUDA_OK(m, cudaGetLastError());
  return CFFM_OK;
}

}  // namespace cffm
