// fp32 SIMT contraction used by the conv stack in CFFM_PREC_FP32 mode (and for pair counts too
// small for tensor-core tiles).  One CTA computes a 128x64 tile of C = A.B with 8x4 register
// tiles; the operands are produced by problem functors so that the interaction cube
// (CFFM.py:355-367) and the im2col view of a 2x2/stride-2 conv (CFFM.py:385-386) are never
// materialised: they are re-indexed on the fly while the tile is staged into shared memory.
#pragma once
#include "common.cuh"

namespace cffm {

constexpr int GBM = 128, GBN = 64, GBK = 16, GTHREADS = 256;

template <class Prob>
__global__ void __launch_bounds__(GTHREADS) k_gemm_simt(const Prob prob) {
  __shared__ __align__(16) float As[GBK][GBM + 4];
  __shared__ __align__(16) float Bs[GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int tiles_n = (prob.N + GBN - 1) / GBN;
  const int tile_m = blockIdx.x / tiles_n, tile_n = blockIdx.x - tile_m * tiles_n;
  const int m0 = tile_m * GBM, n0 = tile_n * GBN;
  const int split = blockIdx.y;
  const int ksteps = (prob.Kd + GBK - 1) / GBK;
  const int per = (ksteps + gridDim.y - 1) / gridDim.y;
  const int ks_begin = split * per;
  const int ks_end = min(ksteps, ks_begin + per);

  // ---- staging maps ----
  typename Prob::RowCtx rc[Prob::A_KFAST ? 8 : 1];
  int a_kk, a_mm;  // fixed coordinate of this thread inside the A tile
  if constexpr (Prob::A_KFAST) {
    a_kk = tid & 15; a_mm = tid >> 4;
#pragma unroll
    for (int r = 0; r < 8; ++r) rc[r] = prob.row_ctx(m0 + a_mm + 16 * r);
  } else {
    a_mm = tid & 127; a_kk = tid >> 7;
    rc[0] = prob.row_ctx(m0 + a_mm);
  }
  int b_kk, b_nn;
  if constexpr (Prob::B_KFAST) { b_kk = tid & 15; b_nn = tid >> 4; }
  else { b_nn = tid & 63; b_kk = tid >> 6; }

  float a_reg[8], b_reg[4];
  auto fetch = [&](int ks) {
    const int k0 = ks * GBK;
    if constexpr (Prob::A_KFAST) {
      typename Prob::RedCtx kc = prob.red_ctx(k0 + a_kk);
#pragma unroll
      for (int r = 0; r < 8; ++r) a_reg[r] = prob.loadA(rc[r], kc);
    } else {
#pragma unroll
      for (int r = 0; r < 8; ++r) a_reg[r] = prob.loadA(rc[0], prob.red_ctx(k0 + a_kk + 2 * r));
    }
    if constexpr (Prob::B_KFAST) {
#pragma unroll
      for (int r = 0; r < 4; ++r) b_reg[r] = prob.loadB(k0 + b_kk, n0 + b_nn + 16 * r);
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) b_reg[r] = prob.loadB(k0 + b_kk + 4 * r, n0 + b_nn);
    }
  };
  auto stash = [&]() {
    if constexpr (Prob::A_KFAST) {
#pragma unroll
      for (int r = 0; r < 8; ++r) As[a_kk][a_mm + 16 * r] = a_reg[r];
    } else {
#pragma unroll
      for (int r = 0; r < 8; ++r) As[a_kk + 2 * r][a_mm] = a_reg[r];
    }
    if constexpr (Prob::B_KFAST) {
#pragma unroll
      for (int r = 0; r < 4; ++r) Bs[b_kk][b_nn + 16 * r] = b_reg[r];
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) Bs[b_kk + 4 * r][b_nn] = b_reg[r];
    }
  };

  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  if (ks_begin < ks_end) fetch(ks_begin);
  for (int ks = ks_begin; ks < ks_end; ++ks) {
    stash();
    __syncthreads();
    if (ks + 1 < ks_end) fetch(ks + 1);
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + tx * 4 + j;
    if (n < prob.N) {
      typename Prob::ColCtx cc = prob.col_ctx(n);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m < prob.M) prob.store(m, cc, acc[i][j], split);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Index helpers shared by the conv problems.  Spatial sizes are powers of two.
struct ConvGeom {
  int P;      // channels (= pairs)
  int Hin;    // input spatial size of this layer
  int lgHo;   // log2(Hin/2)
  __device__ __forceinline__ void pos(int m, int& b, int& h, int& w) const {
    const int Ho = 1 << lgHo;
    w = m & (Ho - 1); h = (m >> lgHo) & (Ho - 1); b = m >> (2 * lgHo);
  }
  // offset of X[b, 2h, 2w, 0]
  __device__ __forceinline__ int64_t base(int m) const {
    int b, h, w; pos(m, b, h, w);
    return (((int64_t)b * Hin + 2 * h) * Hin + 2 * w) * P;
  }
  // offset of tap (dh,dw), channel p relative to base; k = (dh*2+dw)*P + p
  __device__ __forceinline__ int tapoff(int k) const {
    const int tap = k / P, p = k - tap * P;
    return ((tap >> 1) * Hin + (tap & 1)) * P + p;
  }
};

struct Ctx64 { int64_t off; bool ok; };
struct Ctx32 { int off; bool ok; };
struct Ctx2 { int a, b; bool ok; };
struct Ctx2L { int64_t a, b; bool ok; };

// ---- forward, layer l >= 1: Y_l[m, q] = sum_k phi(Y_{l-1})[im2col(m, k)] W_l[k, q] + b_l[q] --
template <int ACT>
struct ConvFwdProb {
  static constexpr bool A_KFAST = true, B_KFAST = false;
  typedef Ctx64 RowCtx; typedef Ctx32 RedCtx; typedef int ColCtx;
  int M, N, Kd;
  ConvGeom g;
  const float* __restrict__ Yprev; const float* __restrict__ W; const float* __restrict__ bias;
  float* __restrict__ Yout;
  __device__ RowCtx row_ctx(int m) const { return {m < M ? g.base(m) : 0, m < M}; }
  __device__ RedCtx red_ctx(int k) const { return {k < Kd ? g.tapoff(k) : 0, k < Kd}; }
  __device__ float loadA(const RowCtx& r, const RedCtx& k) const {
    return (r.ok && k.ok) ? phi_f<ACT>(__ldg(Yprev + r.off + k.off)) : 0.f;
  }
  __device__ float loadB(int k, int n) const { return (k < Kd && n < N) ? __ldg(W + (int64_t)k * N + n) : 0.f; }
  __device__ ColCtx col_ctx(int n) const { return n; }
  __device__ void store(int m, int n, float v, int) const { Yout[(int64_t)m * N + n] = v + __ldg(bias + n); }
};

// ---- forward, layer 0: the A operand is the interaction cube, synthesised from the gathered
//      outer rows: X0[b, 2h+dh, 2w+dw, p] = o_i[2h+dh] * o_j[2w+dw] (CFFM.py:355-367) ----------
struct Conv0Geom {
  int P, F, K, lgHo;  // K = outer_dims, Ho = K/2
  const float* __restrict__ rows;  // [B, F, K]
  const int* __restrict__ pair_i; const int* __restrict__ pair_j;
  __device__ __forceinline__ Ctx2L rowc(int m, bool ok) const {
    const int Ho = 1 << lgHo;
    const int w = m & (Ho - 1), h = (m >> lgHo) & (Ho - 1), b = m >> (2 * lgHo);
    const int64_t sb = (int64_t)b * F * K;
    return {sb + 2 * h, sb + 2 * w, ok};
  }
  __device__ __forceinline__ Ctx2 redc(int k, bool ok) const {
    if (!ok) return {0, 0, false};
    const int tap = k / P, p = k - tap * P;
    return {__ldg(pair_i + p) * K + (tap >> 1), __ldg(pair_j + p) * K + (tap & 1), true};
  }
  __device__ __forceinline__ float cube(const Ctx2L& r, const Ctx2& k) const {
    return (r.ok && k.ok) ? __ldg(rows + r.a + k.a) * __ldg(rows + r.b + k.b) : 0.f;
  }
};

struct Conv0FwdProb {
  static constexpr bool A_KFAST = true, B_KFAST = false;
  typedef Ctx2L RowCtx; typedef Ctx2 RedCtx; typedef int ColCtx;
  int M, N, Kd;
  Conv0Geom g;
  const float* __restrict__ W; const float* __restrict__ bias; float* __restrict__ Yout;
  __device__ RowCtx row_ctx(int m) const { return g.rowc(m, m < M); }
  __device__ RedCtx red_ctx(int k) const { return g.redc(k, k < Kd); }
  __device__ float loadA(const RowCtx& r, const RedCtx& k) const { return g.cube(r, k); }
  __device__ float loadB(int k, int n) const { return (k < Kd && n < N) ? __ldg(W + (int64_t)k * N + n) : 0.f; }
  __device__ ColCtx col_ctx(int n) const { return n; }
  __device__ void store(int m, int n, float v, int) const { Yout[(int64_t)m * N + n] = v + __ldg(bias + n); }
};

// ---- weight gradient, layer l >= 1: dW_l[k, q] = sum_m phi(Y_{l-1})[im2col(m, k)] dY_l[m, q];
//      the reduction over m is split across gridDim.y, partial sums land in `partial` ---------
template <int ACT>
struct ConvWgradProb {
  static constexpr bool A_KFAST = false, B_KFAST = false;
  typedef Ctx32 RowCtx; typedef Ctx64 RedCtx; typedef int ColCtx;
  int M, N, Kd;  // M = 4P (rows k), N = P, Kd = B*Ho*Ho (reduction over positions)
  ConvGeom g;
  const float* __restrict__ Yprev; const float* __restrict__ dY; float* __restrict__ partial;
  __device__ RowCtx row_ctx(int k) const { return {k < M ? g.tapoff(k) : 0, k < M}; }
  __device__ RedCtx red_ctx(int m) const { return {m < Kd ? g.base(m) : 0, m < Kd}; }
  __device__ float loadA(const RowCtx& r, const RedCtx& k) const {
    return (r.ok && k.ok) ? phi_f<ACT>(__ldg(Yprev + k.off + r.off)) : 0.f;
  }
  __device__ float loadB(int m, int n) const { return (m < Kd && n < N) ? __ldg(dY + (int64_t)m * N + n) : 0.f; }
  __device__ ColCtx col_ctx(int n) const { return n; }
  __device__ void store(int k, int n, float v, int split) const {
    partial[((int64_t)split * M + k) * N + n] = v;
  }
};

struct Conv0WgradProb {
  static constexpr bool A_KFAST = false, B_KFAST = false;
  typedef Ctx2 RowCtx; typedef Ctx2L RedCtx; typedef int ColCtx;
  int M, N, Kd;
  Conv0Geom g;
  const float* __restrict__ dY; float* __restrict__ partial;
  __device__ RowCtx row_ctx(int k) const { return g.redc(k, k < M); }
  __device__ RedCtx red_ctx(int m) const { return g.rowc(m, m < Kd); }
  __device__ float loadA(const RowCtx& r, const RedCtx& k) const { return g.cube(k, r); }
  __device__ float loadB(int m, int n) const { return (m < Kd && n < N) ? __ldg(dY + (int64_t)m * N + n) : 0.f; }
  __device__ ColCtx col_ctx(int n) const { return n; }
  __device__ void store(int k, int n, float v, int split) const {
    partial[((int64_t)split * M + k) * N + n] = v;
  }
};

// ---- data gradient, layer l >= 1:
//      dX_l[b,2h+dh,2w+dw,p] = sum_q dY_l[m,q] W_l[(dh,dw,p),q] + dsp_l[b,2h+dh]   
//      dY_{l-1} = dX_l * phi'(Y_{l-1}); windows do not overlap, so every element is written once.
template <int ACT>
struct ConvDgradProb {
  static constexpr bool A_KFAST = true, B_KFAST = true;
  typedef Ctx64 RowCtx; typedef Ctx32 RedCtx;
  struct ColCtx { int off; int dh; };
  int M, N, Kd;  // M = B*Ho*Ho, N = 4P, Kd = P
  ConvGeom g;
  const float* __restrict__ dY; const float* __restrict__ W; const float* __restrict__ Yprev;
  float* __restrict__ dYprev;
  const float* __restrict__ gout;    // [B] dLoss/dout
  const float* __restrict__ v_head;  // [t1_dim] (dense_1 . dense_2), already scaled by beta_outer
  int sp_off;                        // offset of level l inside t1
  __device__ RowCtx row_ctx(int m) const { return {(int64_t)m * Kd, m < M}; }
  __device__ RedCtx red_ctx(int q) const { return {q, q < Kd}; }
  __device__ float loadA(const RowCtx& r, const RedCtx& k) const { return (r.ok && k.ok) ? __ldg(dY + r.off + k.off) : 0.f; }
  __device__ float loadB(int q, int n) const { return (q < Kd && n < N) ? __ldg(W + (int64_t)n * Kd + q) : 0.f; }
  __device__ ColCtx col_ctx(int n) const {
    const int tap = n / g.P, p = n - tap * g.P;
    return {((tap >> 1) * g.Hin + (tap & 1)) * g.P + p, tap >> 1};
  }
  __device__ void store(int m, const ColCtx& c, float v, int) const {
    int b, h, w; g.pos(m, b, h, w);
    const int64_t idx = (((int64_t)b * g.Hin + 2 * h) * g.Hin + 2 * w) * g.P + c.off;
    const float dsp = __ldg(gout + b) * __ldg(v_head + sp_off + 2 * h + c.dh);
    dYprev[idx] = (v + dsp) * phi_df<ACT>(__ldg(Yprev + idx));
  }
};

}  // namespace cffm
