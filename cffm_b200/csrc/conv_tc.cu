// Tensor-core (tcgen05) conv stack for CFFM_PREC_BF16: the 2x2/stride-2 conv layers of the outer
// path (CFFM.py:373-391) as implicit GEMMs with bf16 operands and fp32 accumulation in TMEM.
//
//   layer 0   factorised form (conv0_fact.cuh, conv0_dfact.cuh, conv0_wfact.cuh: the cube is rank one per pair,
//             so the layer is a bilinear form in the rows themselves) from 16 fields / 512 samples on; otherwise
//             the direct form below: A operand = interaction cube (CFFM.py:355-367), synthesised by producer
//             warps straight into tensor memory / the UMMA shared-memory layout.  Either way the cube never
//             exists in HBM.
//   layer >=1 A operand = im2col view of the stored activations: non-overlapping 2x2 windows make
//             im2col a pure re-index, expressed as a 5-D TMA box over the NHWC tensor.
//   epilogues read the accumulator from TMEM and fuse bias + relu + activation + bf16 store +
//             the sum-pooling row sums (forward), the pooling-gradient broadcast and activation
//             mask (data gradient), or the contraction of the cube gradient with the embedding
//             rows (layer-0 data gradient).
//   weight gradients use MN-major descriptors (both operands "transposed"), split over the
//             position axis, reduced in a fixed order.
// Activations X_l = phi(Y_{l-1}) and gradients dY_l are bf16, NHWC with the channel (pair) axis
// padded to a multiple of 64; master weights, optimizer state, pooling sums and all reductions
// stay fp32.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "model.h"
#include "tc_kernel.cuh"

namespace cffm {
namespace tc {

typedef __nv_bfloat16 bf16;

struct Geom {
  int B, P, Pp, F, K;      // batch, pairs, padded pairs, fields, outer dims
  int Ho, lgHo, Hin;       // output / input spatial size of this layer
  int M;                   // B*Ho*Ho output positions
  int BN, tiles_n;         // UMMA N and number of N tiles
  __device__ __forceinline__ void pos(int m, int& b, int& h, int& w) const {
    w = m & (Ho - 1); h = (m >> lgHo) & (Ho - 1); b = m >> (2 * lgHo);
  }
};

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <int ACT>
__device__ __forceinline__ float phi_scale() { return ACT == CFFM_ACT_SELU ? kSeluScale : 1.f; }

// ---- split-bf16 ("bf16x3", CFFM_PREC_BF16X3): fp32-class accuracy on the bf16 tensor cores -------------------
// Every operand is kept as hi + lo with hi = bf16(x), lo = bf16(x - hi) (|x - hi - lo| <= 2^-18 |x|) and a product
// a*b is taken as a_hi*b_hi + a_lo*b_hi + a_hi*b_lo (the dropped lo*lo term is 2^-18 relative as well), all three
// accumulated in the same fp32 TMEM accumulator.  In the pipeline this is a GEMM with a three times longer
// reduction axis, [A_hi | A_lo | A_hi] . [B_hi ; B_hi ; B_lo]: reduction chunk kc of a unit is chunk kc / 3 of
// the plain operands in part kc % 3, so only the loaders (which tensor map / which half to synthesise) and the
// epilogues (which write hi and lo) know about it; k_tc itself is unchanged.
__device__ __forceinline__ uint32_t pack2_lo(float lo, float hi) {   // the lo halves of two values
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  const float2 hf = __bfloat1622float2(h);
  __nv_bfloat162 v = __floats2bfloat162_rn(lo - hf.x, hi - hf.y);
  return *reinterpret_cast<uint32_t*>(&v);
}
struct SplitK {
  int nparts = 1;                 // 1: plain bf16, 3: split
  __device__ __forceinline__ int part(int kc) const { return nparts == 1 ? 0 : kc % 3; }
  __device__ __forceinline__ int base(int kc) const { return nparts == 1 ? kc : kc / 3; }
  __device__ __forceinline__ bool a_lo(int kc) const { return part(kc) == 1; }
  __device__ __forceinline__ bool b_lo(int kc) const { return part(kc) == 2; }
};

// =================================================================================================
// Forward: Y_l = conv(X_l) + b_l ; X_{l+1} = phi(Y_l) (bf16) ; sum_pooling[l+1][b,h] = sum_{w,q} X_{l+1}
// =================================================================================================
template <int ACT, bool L0, bool SPLIT = false>
struct ConvFwdTC : KMajorA, KMajorB {
  static constexpr bool kSynthA = L0;
  static constexpr int kStages = 4, kExtraBytes = L0 ? 28 * 1024 : 0, kATiles = 1, kAccBufs = 2, kEpiWarps = L0 ? 8 : 4;
  // layer 0: the synthesised cube slab goes to tensor memory (tcgen05.st) and the MMA reads A from there:
  // the kernel was bound by shared-memory bandwidth (STS of the slab + UMMA reads of A and B + LDS of the rows)
  // ... and every slab feeds two N tiles (kBPair): the producers, not the tensor pipe, bounded the kernel
  static constexpr bool kATmem = L0, kSynthAlternate = L0, kBPair = L0, kEpiPrefetch = false;
  CUtensorMap mapA, mapB, mapA2, mapB2;   // *2: the lo halves (split mode)
  Geom g;
  SplitK sk; bf16* Xout_lo;
  bf16* Dphi;               // gelu, training: phi'(Y) next to X = phi(Y) (its derivative does not follow from X's sign)
  const float* bias; bf16* Xout; float* t1; int t1_dim, sp_off;
  float* pool_part;         // l >= 1, tiles_n > 1: pooled sums per N tile [tiles_n][B*Ho], added up by k_pool_parts
  const float* rows; const int* pair_i; const int* pair_j;
  __device__ uint32_t idesc() const { return umma_idesc_bf16(BM, g.BN); }
  __device__ int bn() const { return g.BN; }
  __device__ int m_tiles() const { return (g.M + BM - 1) / BM; }
  // Layer 0 (A synthesised on chip): a CTA walks all units of its M tile.  Layers >= 1 read A (the im2col view
  // of X_l) from memory once per N tile: neighbouring CTAs take the N tiles of ONE M tile at the same time, so
  // the tile comes from DRAM once and from L2 for the others (a CTA walking its own M tile alone re-read it
  // from DRAM: 148 tiles of 4*Pp*256 B do not fit L2).
  __device__ int n_iters(int cta, int ncta) const {
    const int mt = m_tiles();
    if (L0) return (cta < mt ? (mt - cta + ncta - 1) / ncta : 0) * units_n();
    const int n = mt * g.tiles_n;
    return cta < n ? (n - cta + ncta - 1) / ncta : 0;
  }
  __device__ int units_n() const { return L0 ? g.tiles_n >> 1 : g.tiles_n; }   // layer 0: a unit = two N tiles
  __device__ Unit unit(int cta, int ncta, int it) const {
    if (L0) return {cta + (it / units_n()) * ncta, it % units_n(), 0};
    const int u = cta + it * ncta;
    return {u / g.tiles_n, u % g.tiles_n, 0};
  }
  __device__ int k_chunks(Unit) const { return (4 * g.Pp / BK) * sk.nparts; }
  __device__ uint32_t tx_bytes() const { return (uint32_t)(L0 ? 2 * g.BN * BK * 2 : A_STAGE_BYTES + g.BN * BK * 2); }
  __device__ void prefetch() const {
    if (!L0) prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    if (sk.nparts > 1) { if (!L0) prefetch_tmap(&mapA2); prefetch_tmap(&mapB2); }
  }
  __device__ void load_a(uint8_t* s, uint64_t* bar, Unit un, int kc) const {
    const int m0 = un.m_tile * BM;
    const int b0 = m0 >> (2 * g.lgHo), h0 = (m0 >> g.lgHo) & (g.Ho - 1);
    const int k0 = sk.base(kc) * BK, dh = k0 / (2 * g.Pp), c0 = k0 - dh * 2 * g.Pp;
    tma_load_5d(s, sk.a_lo(kc) ? &mapA2 : &mapA, bar, c0, 0, dh, h0, b0);
  }
  __device__ void load_b(uint8_t* s, uint64_t* bar, Unit un, int kc) const {
    tma_load_2d(s, sk.b_lo(kc) ? &mapB2 : &mapB, bar, sk.base(kc) * BK, un.n_tile * g.BN);
  }
  // ---- layer 0: stage the sample's outer rows, then build 128 x 64 slabs of the cube ----
  // Shared scratch of the producers: rows o[F+1][K] (row F = zeros, the target of padded pairs) and a
  // table with one entry per pair of every stage: byte offsets of rows i and j (lo / hi 16 bits).
  // k = p*4 + dh*2 + dw: 16 consecutive pairs per 64-wide slab; no branches in the inner loop.
  struct SynthState {};
  // A 128-row tile covers 128 / Ho^2 samples: one (K = 32: 256 positions per sample, K = 64: 1024) or two (K = 16: 64).
  __device__ int samples_per_tile() const { const int s = BM >> (2 * g.lgHo); return s < 1 ? 1 : s; }
  __device__ int sample_bytes() const { return (g.F + 1) * g.K * 4; }
  __device__ void synth_begin(Unit un, uint8_t* ex, int t, SynthState&) const {
    const int S = samples_per_tile();
    uint32_t* tab = reinterpret_cast<uint32_t*>(ex + S * sample_bytes());
    const int b0 = (un.m_tile * BM) >> (2 * g.lgHo);
    for (int sl = 0; sl < S; ++sl) {
      float* o = reinterpret_cast<float*>(ex + sl * sample_bytes());
      const int b = b0 + sl;
      const float4* src = reinterpret_cast<const float4*>(rows + (int64_t)(b < g.B ? b : 0) * g.F * g.K);
      for (int e = t; e < g.F * g.K / 4; e += 256) reinterpret_cast<float4*>(o)[e] = b < g.B ? __ldg(src + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int e = t; e < g.K / 4; e += 256) reinterpret_cast<float4*>(o + g.F * g.K)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const uint32_t zero_row = (uint32_t)(g.F * g.K * 4);
    for (int e = t; e < g.Pp; e += 256)
      tab[e] = e < g.P ? ((uint32_t)(pair_i[e] * g.K * 4) | ((uint32_t)(pair_j[e] * g.K * 4) << 16)) : (zero_row | (zero_row << 16));
  }
  // 8 pairs x 4 taps = 32 bf16 = 16 packed columns of this thread's row (columns half*16 .. +15 of the stage)
  __device__ void synth_regs(uint32_t (&pk)[16], Unit un, int kc_in, int t256, const uint8_t* ex) const {
    const int kc = sk.base(kc_in);
    const bool lo_part = sk.a_lo(kc_in);
    const int t = t256 & 127, half = t256 >> 7;
    const uint8_t* o = ex + (t >> (2 * g.lgHo)) * sample_bytes();      // this row's sample inside the tile
    const uint32_t* tab = reinterpret_cast<const uint32_t*>(ex + samples_per_tile() * sample_bytes());
    const int m = un.m_tile * BM + t;
    const int h = (m >> g.lgHo) & (g.Ho - 1), w = m & (g.Ho - 1);
    const uint8_t* oh = o + 8 * h;
    const uint8_t* ow = o + 8 * w;
    const uint4* te = reinterpret_cast<const uint4*>(tab + kc * 16 + half * 8);
    const uint4 e0 = te[0], e1 = te[1];
    const uint32_t ent[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    uint32_t cur_i = 0xFFFFFFFFu;
    float2 oi = make_float2(0.f, 0.f);
#pragma unroll
    for (int pr = 0; pr < 8; ++pr) {
      const uint32_t en = ent[pr];
      if ((en & 0xFFFFu) != cur_i) { cur_i = en & 0xFFFFu; oi = *reinterpret_cast<const float2*>(oh + cur_i); }  // uniform: pairs run i-major
      const float2 oj = *reinterpret_cast<const float2*>(ow + (en >> 16));
      if (SPLIT && lo_part) {   // warp-uniform
        pk[pr * 2 + 0] = pack2_lo(oi.x * oj.x, oi.x * oj.y);
        pk[pr * 2 + 1] = pack2_lo(oi.y * oj.x, oi.y * oj.y);
      } else {
        pk[pr * 2 + 0] = pack2(oi.x * oj.x, oi.x * oj.y);
        pk[pr * 2 + 1] = pack2(oi.y * oj.x, oi.y * oj.y);
      }
    }
  }
  __device__ void synth_a(uint8_t* sA, Unit un, int kc, int t256, const uint8_t* ex, SynthState&) const {
    uint32_t pk[16];
    synth_regs(pk, un, kc, t256, ex);
    const int t = t256 & 127, half = t256 >> 7;
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4)
      *reinterpret_cast<uint4*>(sA + sw128_offset(t, half * 4 + c4)) = make_uint4(pk[4 * c4], pk[4 * c4 + 1], pk[4 * c4 + 2], pk[4 * c4 + 3]);
  }
  __device__ void synth_a_tmem(uint32_t taddr, Unit un, int kc, int t256, const uint8_t* ex, SynthState&) const {
    uint32_t pk[16];
    synth_regs(pk, un, kc, t256, ex);
    tmem_st16(taddr + (uint32_t)((t256 >> 7) * 16), pk);
  }
  struct Epilogue {
    // layer 0: even N tiles belong to epilogue warps 0..3, odd ones to warps 4..7; the two partial
    // row sums meet in shared memory (double-buffered on the M tile's parity)
    // The accumulators of a unit are not double-buffered there, so the drain is on the critical path:
    // the bias sits in shared memory (zero beyond P: no per-element guards, broadcast LDS.128).
    const ConvFwdTC& p; int row, ew; float rowsum; float* xch; const float* sbias; int flip;
    __device__ Epilogue(const ConvFwdTC& p_, uint8_t* ex, int row_, int ew_)
        : p(p_), row(row_), ew(ew_), rowsum(0.f), xch(reinterpret_cast<float*>(ex)), sbias(xch + 2 * BM), flip(0) {
      if (L0) {
        float* sb = xch + 2 * BM;
        const int ncol = p.g.tiles_n * p.g.BN;
        for (int e = ew * 32 + (threadIdx.x & 31); e < ncol; e += 256) sb[e] = e < p.g.P ? __ldg(p.bias + e) : 0.f;
        asm volatile("bar.sync 2, 256;" ::: "memory");
      }
    }
    __device__ void begin(Unit un) { if (!L0 || un.n_tile < 2) rowsum = 0.f; }
    __device__ void chunk(Unit un, int c0, const float (&v)[32]) {
      const int m = un.m_tile * BM + row;
      const int n0 = un.n_tile * p.g.BN + c0;
      if (L0 && n0 >= p.g.Pp) return;     // N tiles may overhang the padded channel count (zero weights)
      uint32_t pk[16];
      uint32_t pl[SPLIT ? 16 : 1];
      uint32_t pd[ACT == CFFM_ACT_GELU ? 16 : 1];
      float bv[32];
      if (L0) {
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 b4 = *reinterpret_cast<const float4*>(sbias + n0 + 4 * q4);
          bv[4 * q4] = b4.x; bv[4 * q4 + 1] = b4.y; bv[4 * q4 + 2] = b4.z; bv[4 * q4 + 3] = b4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) bv[j] = n0 + j < p.g.P ? __ldg(p.bias + n0 + j) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float y0 = v[j] + bv[j];
        const float y1 = v[j + 1] + bv[j + 1];
        const float x0 = phi_f<ACT>(y0), x1 = phi_f<ACT>(y1);
        rowsum += x0 + x1;
        pk[j >> 1] = pack2(x0, x1);
        if constexpr (SPLIT) pl[j >> 1] = pack2_lo(x0, x1);
        if constexpr (ACT == CFFM_ACT_GELU) pd[j >> 1] = pack2(phi_df<ACT>(y0), phi_df<ACT>(y1));
      }
      if (m < p.g.M) {
        uint4* dst = reinterpret_cast<uint4*>(p.Xout + (int64_t)m * p.g.Pp + n0);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) dst[q4] = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
        if constexpr (SPLIT) {
          uint4* dl = reinterpret_cast<uint4*>(p.Xout_lo + (int64_t)m * p.g.Pp + n0);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) dl[q4] = make_uint4(pl[4 * q4], pl[4 * q4 + 1], pl[4 * q4 + 2], pl[4 * q4 + 3]);
        }
        if constexpr (ACT == CFFM_ACT_GELU) {
          if (p.Dphi) {
            uint4* dd = reinterpret_cast<uint4*>(p.Dphi + (int64_t)m * p.g.Pp + n0);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) dd[q4] = make_uint4(pd[4 * q4], pd[4 * q4 + 1], pd[4 * q4 + 2], pd[4 * q4 + 3]);
          }
        }
      }
    }
    __device__ void end(Unit un) {
      if (L0 && un.n_tile < p.g.tiles_n - 2) return;
      const int m = un.m_tile * BM + row;
      float s = m < p.g.M ? rowsum : 0.f;
      for (int off = p.g.Ho >> 1; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (L0) {
        float* slot = xch + flip * BM + row;
        flip ^= 1;
        if (ew >= 4) *slot = s;
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (ew >= 4) return;
        s += *slot;
      }
      int b, h, w; p.g.pos(m, b, h, w);
      if (w == 0 && m < p.g.M) {
        if (L0 || p.g.tiles_n == 1) p.t1[(int64_t)b * p.t1_dim + p.sp_off + h] = s;
        else p.pool_part[(int64_t)un.n_tile * p.g.B * p.g.Ho + (int64_t)b * p.g.Ho + h] = s;
      }
    }
    __device__ void finish() {}
  };
};

// t1[b, sp_off + h] = sum over N tiles (fixed order) of the pooled partial sums of forward layers >= 1
__global__ void k_pool_parts(const float* __restrict__ part, int tiles_n, int B, int Ho, float* __restrict__ t1, int t1_dim, int sp_off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Ho) return;
  float s = 0.f;
  for (int n = 0; n < tiles_n; ++n) s += part[(int64_t)n * B * Ho + i];
  t1[(int64_t)(i / Ho) * t1_dim + sp_off + i % Ho] = s;
}

// =================================================================================================
// Data gradient, layer l >= 1:
//   dX_l[b,2h+dh,2w+dw,p] = sum_q dY_l[m,q] W_l[(dh,dw,p),q] + g_b v[sp_off + 2h+dh]
//   dY_{l-1} = dX_l * phi'(Y_{l-1}); phi' follows from the sign of X_l = phi(Y_{l-1})
// =================================================================================================
template <int ACT, bool SPLIT = false>
struct ConvDgradTC : KMajorA, KMajorB {
  static constexpr bool kSynthA = false;
  // eight epilogue warps (two per TMEM lane quarter, alternate 32-column chunks); scratch = their staging tiles
  static constexpr int kStages = 4, kExtraBytes = 8 * 32 * 80, kATiles = 1, kAccBufs = 2, kEpiWarps = 8;
  static constexpr bool kATmem = false, kSynthAlternate = false, kBPair = false, kEpiPrefetch = true;
  CUtensorMap mapA, mapB, mapA2, mapB2;   // A: dY_l dims (Pp, M) box (64,128); B: Wd dims (Pp, 4Pp) box (64, BN); *2: lo halves
  Geom g;                   // BN divides Pp; tiles_n = 4*Pp/BN
  SplitK sk; bf16* dYprev_lo;
  const bf16* X; bf16* dYprev; const float* gout; const float* v_head; int sp_off;
  __device__ uint32_t idesc() const { return umma_idesc_bf16(BM, g.BN); }
  __device__ int bn() const { return g.BN; }
  __device__ int n_units() const { return ((g.M + BM - 1) / BM) * g.tiles_n; }
  __device__ int n_iters(int cta, int ncta) const { const int n = n_units(); return cta < n ? (n - cta + ncta - 1) / ncta : 0; }
  __device__ Unit unit(int cta, int ncta, int it) const { const int u = cta + it * ncta; return {u / g.tiles_n, u % g.tiles_n, 0}; }
  __device__ int k_chunks(Unit) const { return (g.Pp / BK) * sk.nparts; }
  __device__ uint32_t tx_bytes() const { return (uint32_t)(A_STAGE_BYTES + g.BN * BK * 2); }
  __device__ void prefetch() const {
    prefetch_tmap(&mapA); prefetch_tmap(&mapB);
    if (sk.nparts > 1) { prefetch_tmap(&mapA2); prefetch_tmap(&mapB2); }
  }
  __device__ void load_a(uint8_t* s, uint64_t* bar, Unit un, int kc) const {
    tma_load_2d(s, sk.a_lo(kc) ? &mapA2 : &mapA, bar, sk.base(kc) * BK, un.m_tile * BM);
  }
  __device__ void load_b(uint8_t* s, uint64_t* bar, Unit un, int kc) const {
    tma_load_2d(s, sk.b_lo(kc) ? &mapB2 : &mapB, bar, sk.base(kc) * BK, un.n_tile * g.BN);
  }
  struct SynthState {};
  __device__ void synth_begin(Unit, uint8_t*, int, SynthState&) const {}
  __device__ void synth_a(uint8_t*, Unit, int, int, const uint8_t*, SynthState&) const {}
  struct Epilogue {
    // K is short here (Pp/64 stages per unit), so the drain decides the speed, and a lane that loads and
    // stores 64 B of its own row costs one L1 wavefront per lane.  Instead: the mask of the whole unit
    // (this warp's four chunks) is requested -- before the accumulator is complete -- in a coalesced
    // layout (4 lanes per row piece), the packed result goes through a per-warp staging tile in shared
    // memory into the same layout, is masked there with bit operations and stored 8 rows per instruction.
    static constexpr int STG = 80;   // staging row stride (64 B + 16: conflict-free 16-byte stores)
    const ConvDgradTC& p; int row, sub, lane; float dsp;
    uint8_t* stg;
    int64_t cbase[4]; bool cok[4];   // rows (lane>>2) + 8i of this warp's 32: element offset of this lane's 8 channels
    // split mode: three times the MMA time per unit covers the drain, and the lo halves double its register need:
    // the mask is then fetched chunk by chunk instead of for the whole unit ahead of the accumulator
    uint4 mk[SPLIT ? 1 : MAX_BN / 64][4];
    __device__ Epilogue(const ConvDgradTC& p_, uint8_t* ex, int row_, int ew)
        : p(p_), row(row_), sub(ew >> 2), lane(threadIdx.x & 31), dsp(0.f), stg(ex + ew * 32 * STG) {}
    __device__ void begin(Unit) {}
    __device__ void prefetch(Unit un) {
      const int n0 = un.n_tile * p.g.BN;
      const int tap = n0 / p.g.Pp, pb = n0 - tap * p.g.Pp;
      const int dh = tap >> 1, dw = tap & 1;
      {
        const int m = un.m_tile * BM + row;
        int b, h, w; p.g.pos(m < p.g.M ? m : 0, b, h, w);
        dsp = m < p.g.M ? __ldg(p.gout + b) * __ldg(p.v_head + p.sp_off + 2 * h + dh) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = un.m_tile * BM + (row & ~31) + (lane >> 2) + 8 * i;
        cok[i] = m < p.g.M;
        int b, h, w; p.g.pos(cok[i] ? m : 0, b, h, w);
        cbase[i] = (((int64_t)b * p.g.Hin + 2 * h + dh) * p.g.Hin + 2 * w + dw) * p.g.Pp + pb + (lane & 3) * 8;
      }
      if constexpr (!SPLIT) {
#pragma unroll
        for (int ci = 0; ci < MAX_BN / 64; ++ci) {
          const int c0 = (2 * ci + sub) * 32;
          if (c0 < p.g.BN) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              mk[ci][i] = cok[i] ? __ldg(reinterpret_cast<const uint4*>(p.X + cbase[i] + c0)) : make_uint4(0u, 0u, 0u, 0u);
          }
        }
      }
    }
    static __device__ __forceinline__ uint32_t mul_bf16x2(uint32_t a, uint32_t b) {
      const __nv_bfloat162 r = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
      return *reinterpret_cast<const uint32_t*>(&r);
    }
    static __device__ __forceinline__ uint32_t keep_bits(uint32_t x) {   // X > 0  <=>  Y > 0, per bf16 half
      return ((x & 0x7FFFu) ? 0xFFFFu : 0u) | ((x & 0x7FFF0000u) ? 0xFFFF0000u : 0u);
    }
    __device__ __forceinline__ void chunk_i(Unit, int ci, int c0, const float (&v)[32]) {
      uint32_t o[16];
      if constexpr (SPLIT) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          mk[0][i] = cok[i] ? __ldg(reinterpret_cast<const uint4*>(p.X + cbase[i] + c0)) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int j = 0; j < 32; j += 2) o[j >> 1] = pack2((v[j] + dsp) * phi_scale<ACT>(), (v[j + 1] + dsp) * phi_scale<ACT>());
      stage_store(ci, c0, o, p.dYprev);
      if constexpr (SPLIT) {                          // split mode: the lo halves take the same route
#pragma unroll
        for (int j = 0; j < 32; j += 2) o[j >> 1] = pack2_lo((v[j] + dsp) * phi_scale<ACT>(), (v[j + 1] + dsp) * phi_scale<ACT>());
        stage_store(ci, c0, o, p.dYprev_lo);
      }
    }
    __device__ __forceinline__ void stage_store(int ci, int c0, const uint32_t (&o)[16], bf16* dst) {
      __syncwarp();                                   // the previous chunk has been read out of the staging tile
      uint4* mine = reinterpret_cast<uint4*>(stg + lane * STG);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) mine[q4] = make_uint4(o[4 * q4], o[4 * q4 + 1], o[4 * q4 + 2], o[4 * q4 + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 d = *reinterpret_cast<const uint4*>(stg + ((lane >> 2) + 8 * i) * STG + (lane & 3) * 16);
        const uint4 x = mk[SPLIT ? 0 : ci][i];
        if constexpr (ACT == CFFM_ACT_GELU) {   // p.X points at phi'(Y): multiply (bf16 x bf16 -> bf16)
          d.x = mul_bf16x2(d.x, x.x); d.y = mul_bf16x2(d.y, x.y); d.z = mul_bf16x2(d.z, x.z); d.w = mul_bf16x2(d.w, x.w);
        } else {
          d.x &= keep_bits(x.x); d.y &= keep_bits(x.y); d.z &= keep_bits(x.z); d.w &= keep_bits(x.w);
        }
        if (cok[i]) *reinterpret_cast<uint4*>(dst + cbase[i] + c0) = d;
      }
    }
    __device__ void end(Unit) {}
    __device__ void finish() {}
  };
};

// =================================================================================================
// Data gradient, layer 0, contracted with the embedding rows.  A CTA owns whole
// samples (two 128-row tiles x all N tiles) and accumulates d o[f][a] in shared memory:
//   column n = p*4 + dh*2 + dw;  D = acc + g_b v[2h+dh]
//   d o_i[2h+dh] += sum_{w,dw} D o_j[2w+dw]     (16 lanes of one h: halving butterfly)
//   d o_j[2w+dw] += sum_{h,dh} D o_i[2h+dh]     (two h per warp: one exchange; per-warp slices)
// Every accumulator has a single writer and a fixed order: the result is deterministic.
// =================================================================================================
struct Conv0DgradTC : KMajorA, KMajorB {
  static constexpr bool kSynthA = false;
  static constexpr int kStages = 3, kExtraBytes = 72 * 1024, kATiles = 1, kAccBufs = 2, kEpiWarps = 8;
  static constexpr bool kATmem = false, kSynthAlternate = false, kBPair = false, kEpiPrefetch = false;
  CUtensorMap mapA, mapB, mapA2, mapB2;   // A: dY_0 dims (Pp, M) box (64,128); B: Wd0 dims (Pp, 4Pp) box (64, BN); *2: lo halves
  Geom g;                   // Ho = 16: 256 rows per sample
  SplitK sk;
  const float* rows; const float* gout; const float* v_head; const int* pair_i; const int* pair_j; float* g_rows;
  __device__ uint32_t idesc() const { return umma_idesc_bf16(BM, g.BN); }
  __device__ int bn() const { return g.BN; }
  __device__ int per_sample() const { return 2 * g.tiles_n; }
  __device__ int n_iters(int cta, int ncta) const { return (cta < g.B ? (g.B - cta + ncta - 1) / ncta : 0) * per_sample(); }
  __device__ Unit unit(int cta, int ncta, int it) const {
    const int ps = per_sample();
    const int b = cta + (it / ps) * ncta, r = it % ps;
    return {b * 2 + r / g.tiles_n, r % g.tiles_n, r};  // z = position inside the sample's unit sequence
  }
  __device__ int k_chunks(Unit) const { return (g.Pp / BK) * sk.nparts; }
  __device__ uint32_t tx_bytes() const { return (uint32_t)(A_STAGE_BYTES + g.BN * BK * 2); }
  __device__ void prefetch() const {
    prefetch_tmap(&mapA); prefetch_tmap(&mapB);
    if (sk.nparts > 1) { prefetch_tmap(&mapA2); prefetch_tmap(&mapB2); }
  }
  __device__ void load_a(uint8_t* s, uint64_t* bar, Unit un, int kc) const {
    tma_load_2d(s, sk.a_lo(kc) ? &mapA2 : &mapA, bar, sk.base(kc) * BK, un.m_tile * BM);
  }
  __device__ void load_b(uint8_t* s, uint64_t* bar, Unit un, int kc) const {
    tma_load_2d(s, sk.b_lo(kc) ? &mapB2 : &mapB, bar, sk.base(kc) * BK, un.n_tile * g.BN);
  }
  struct SynthState {};
  __device__ void synth_begin(Unit, uint8_t*, int, SynthState&) const {}
  __device__ void synth_a(uint8_t*, Unit, int, int, const uint8_t*, SynthState&) const {}
  struct Epilogue {
    const Conv0DgradTC& p; int row, ew, lane;
    float *o, *dOi, *dOj;   // smem: rows [F*K], d o_i [2][F*K] (one per warp of a lane quarter), per-warp d o_j slices [8][F*K]
    uint8_t *pi, *pj;
    float dsp0, dsp1; int h, w;
    __device__ Epilogue(const Conv0DgradTC& p_, uint8_t* ex, int row_, int ew_) : p(p_), row(row_), ew(ew_) {
      const int FK = p.g.F * p.g.K;
      lane = threadIdx.x & 31;
      o = reinterpret_cast<float*>(ex); dOi = o + FK; dOj = dOi + 2 * FK;
      pi = reinterpret_cast<uint8_t*>(dOj + 8 * FK); pj = pi + ((p.g.P + 15) & ~15);
      const int t = ew * 32 + lane;
      for (int e = t; e < p.g.P; e += 256) { pi[e] = (uint8_t)p.pair_i[e]; pj[e] = (uint8_t)p.pair_j[e]; }
      dsp0 = dsp1 = 0.f; h = w = 0;
    }
    __device__ void epi_sync() { asm volatile("bar.sync 2, 256;" ::: "memory"); }
    __device__ void begin(Unit un) {
      const int FK = p.g.F * p.g.K, t = ew * 32 + lane;
      const int b = un.m_tile >> 1;
      if (un.z == 0) {  // first unit of a sample: stage its rows, clear the accumulators
        epi_sync();
        const float* src = p.rows + (int64_t)b * FK;
        for (int e = t; e < FK; e += 256) o[e] = __ldg(src + e);
        for (int e = t; e < 10 * FK; e += 256) dOi[e] = 0.f;   // both d o_i copies and the eight d o_j slices
        epi_sync();
      }
      const int r = (un.m_tile & 1) * BM + row;   // position inside the sample
      h = r >> 4; w = r & 15;
      const float gb = __ldg(p.gout + b);
      dsp0 = gb * __ldg(p.v_head + 2 * h); dsp1 = gb * __ldg(p.v_head + 2 * h + 1);
    }
    // 8 pairs per 32-column chunk.  Pairs are enumerated i-major (CFFM.py:304-305), so a chunk almost
    // always covers at most two runs of equal first field: the d o_i terms of a run are summed in
    // registers before one 16-lane reduction, and every shared-memory update below has exactly one
    // lane per address per instruction (adds to one address come from one lane, in program order).
    __device__ void chunk(Unit un, int c0, const float (&v)[32]) {
      const int K = p.g.K, F = p.g.F;
      const int pbase = (un.n_tile * p.g.BN + c0) >> 2;
      if (pbase >= p.g.P) return;                          // padding channels only (warp-uniform)
      const int i0 = pi[pbase], j0 = pj[pbase];
      const int n0 = min(8, F - j0);                       // pairs of the first run inside this chunk
      const int i1 = i0 + 1;
      const int n1 = n0 < 8 ? min(8 - n0, F - 1 - i1) : 0;
      if (pbase + 8 > p.g.P || n0 + n1 < 8) { chunk_generic(pbase, v); return; }
      const float2 oiA = *reinterpret_cast<const float2*>(o + i0 * K + 2 * h);
      const float2 oiB = *reinterpret_cast<const float2*>(o + (n1 ? i1 : i0) * K + 2 * h);
      float s4[4] = {0.f, 0.f, 0.f, 0.f};                  // run A dh0, run A dh1, run B dh0, run B dh1
      float cj[16];
      const bool up = (lane & 16) != 0;
      float* slice = dOj + ew * F * K + 2 * w + (up ? 1 : 0);
#pragma unroll
      for (int pp = 0; pp < 8; ++pp) {
        const bool inA = pp < n0;
        const int j = inA ? j0 + pp : i1 + 1 + (pp - n0);
        const float2 oi = inA ? oiA : oiB;
        const float2 oj = *reinterpret_cast<const float2*>(o + j * K + 2 * w);
        const float D00 = v[4 * pp] + dsp0, D01 = v[4 * pp + 1] + dsp0, D10 = v[4 * pp + 2] + dsp1, D11 = v[4 * pp + 3] + dsp1;
        const float c0v = fmaf(D01, oj.y, D00 * oj.x), c1v = fmaf(D11, oj.y, D10 * oj.x);
        if (inA) { s4[0] += c0v; s4[1] += c1v; } else { s4[2] += c0v; s4[3] += c1v; }
        cj[2 * pp] = fmaf(D10, oi.y, D00 * oi.x); cj[2 * pp + 1] = fmaf(D11, oi.y, D01 * oi.x);
      }
      // d o_i: 4 values over the 16 lanes of this h: two halving steps (lane bits 3, 2 pick the value), two plain ones
      {
        const bool b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
        float x0 = (b3 ? s4[2] : s4[0]) + __shfl_xor_sync(0xffffffffu, b3 ? s4[0] : s4[2], 8);
        float x1 = (b3 ? s4[3] : s4[1]) + __shfl_xor_sync(0xffffffffu, b3 ? s4[1] : s4[3], 8);
        float x = (b2 ? x1 : x0) + __shfl_xor_sync(0xffffffffu, b2 ? x0 : x1, 4);
        x += __shfl_xor_sync(0xffffffffu, x, 2);
        x += __shfl_xor_sync(0xffffffffu, x, 1);
        if ((lane & 3) == 0 && (!b3 || n1)) atomicAdd(dOi + (ew >> 2) * F * K + (b3 ? i1 : i0) * K + 2 * h + (b2 ? 1 : 0), x);
      }
      // d o_j: the warp's two h rows are lanes l and l^16; the lower half keeps dw = 0, the upper dw = 1
#pragma unroll
      for (int pp = 0; pp < 8; ++pp) {
        const float send = up ? cj[2 * pp] : cj[2 * pp + 1];
        const float keep = up ? cj[2 * pp + 1] : cj[2 * pp];
        const float tot = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        const int j = pp < n0 ? j0 + pp : i1 + 1 + (pp - n0);
        atomicAdd(slice + j * K, tot);
      }
    }
    // Generic path: any mix of pairs (run boundaries, padding); serialised shared-memory updates.
    __device__ void chunk_generic(int pbase, const float (&v)[32]) {
      const int K = p.g.K;
      float ci[16], cj[16];
#pragma unroll
      for (int pp = 0; pp < 8; ++pp) {
        const int pr = pbase + pp;
        float2 oi = make_float2(0.f, 0.f), oj = make_float2(0.f, 0.f);
        if (pr < p.g.P) {
          oi = *reinterpret_cast<const float2*>(o + pi[pr] * K + 2 * h);
          oj = *reinterpret_cast<const float2*>(o + pj[pr] * K + 2 * w);
        }
        const float D00 = v[4 * pp] + dsp0, D01 = v[4 * pp + 1] + dsp0, D10 = v[4 * pp + 2] + dsp1, D11 = v[4 * pp + 3] + dsp1;
        ci[2 * pp] = fmaf(D01, oj.y, D00 * oj.x); ci[2 * pp + 1] = fmaf(D11, oj.y, D10 * oj.x);
        cj[2 * pp] = fmaf(D10, oi.y, D00 * oi.x); cj[2 * pp + 1] = fmaf(D11, oi.y, D01 * oi.x);
      }
      // d o_i: reduce the 16 values over the 16 lanes (w) of this h; lane r ends with value r
#pragma unroll
      for (int s = 8; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int k = 0; k < s; ++k) {
          const float send = up ? ci[k] : ci[k + s];
          const float keep = up ? ci[k + s] : ci[k];
          ci[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
      }
      {  // consecutive pairs usually share their first field: one pair at a time, so an address has one writer
        const int r = lane & 15, pr = pbase + (r >> 1);
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
          if ((r >> 1) == pp && pr < p.g.P) dOi[(ew >> 2) * p.g.F * K + pi[pr] * K + 2 * h + (r & 1)] += ci[0];
          __syncwarp();
        }
      }
      // d o_j: add the warp's two h rows (lanes l and l^16); lanes 0-15 keep pairs 0-3, 16-31 pairs 4-7
      {
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float send = up ? cj[k] : cj[k + 8];
          const float keep = up ? cj[k + 8] : cj[k];
          cj[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
        float* slice = dOj + ew * p.g.F * K;
#pragma unroll
        for (int half = 0; half < 2; ++half) {  // pairs 4 apart may share their second field: halves take turns
          if ((int)up == half) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int pr = pbase + (up ? 4 : 0) + (k >> 1);
              if (pr < p.g.P) slice[pj[pr] * K + 2 * w + (k & 1)] += cj[k];
            }
          }
          __syncwarp();
        }
      }
    }
    __device__ void end(Unit un) {
      if (un.z != p.per_sample() - 1) return;
      const int FK = p.g.F * p.g.K, t = ew * 32 + lane;
      const int b = un.m_tile >> 1;
      epi_sync();
      float* dst = p.g_rows + (int64_t)b * FK;
      for (int e = t; e < FK; e += 256) {
        float sj = 0.f;
#pragma unroll
        for (int sl = 0; sl < 8; ++sl) sj += dOj[sl * FK + e];
        dst[e] = (dOi[e] + dOi[FK + e]) + sj;
      }
    }
    __device__ void finish() {}
  };
};

// =================================================================================================
// Weight gradient: dW_l[k, q] = sum_m A[m, k] dY_l[m, q], the position axis m split over gridDim
// (unit.z); partial tiles land in fp32 scratch [split][4Pp][Pp] and are reduced in order afterwards.
// B (= dY_l) is MN-major via TMA; A is MN-major via the im2col TMA box (l >= 1) or synthesised
// K-major by the producer warps (layer 0: the cube).
// =================================================================================================
template <int ACT, bool L0, bool SPLIT = false>
struct ConvWgradTC : KMajorA, MNMajorB {
  static constexpr bool kSynthA = L0;
  // layer 0: two 128-row A tiles (256 cube channels) share every dY stage -> half the L2 traffic of B;
  // their accumulators sit side by side in TMEM (2 x 256 columns, one buffer: the units are long)
  static constexpr int kATiles = L0 ? 2 : 1, kAccBufs = L0 ? 1 : 2, kEpiWarps = 4;
  static constexpr bool kATmem = false, kSynthAlternate = false, kBPair = false, kEpiPrefetch = false;
  static constexpr int kStages = L0 ? 3 : 4, kExtraBytes = L0 ? 24 * 1024 : 0;
  static constexpr int kRows = kATiles * BM;   // cube channels (rows of dW) per unit
  CUtensorMap mapA, mapB, mapA2, mapB2;   // A (l>=1): 5-D im2col map, box = 64 channels x 64 positions; B: dY dims (Pp, M) box (64,64); *2: lo halves
  Geom g;                   // BN divides Pp, multiple of 64
  SplitK sk;
  int chunks_total, chunks_per_split, n_split;
  float* partial;
  const float* rows; const int* pair_i; const int* pair_j;
  __device__ uint64_t a_desc(uint32_t addr, int k) const {
    if (L0) return umma_desc_k_sw128(addr) + (uint64_t)(k * 2);
    return umma_desc_mn_sw128(addr + k * 2048, 8192, 1024);
  }
  __device__ uint32_t idesc() const { return umma_idesc_bf16(BM, g.BN, !L0, true); }
  __device__ int bn() const { return g.BN; }
  __device__ int m_tiles() const { return 4 * g.Pp / kRows; }
  __device__ int n_units() const { return m_tiles() * g.tiles_n * n_split; }
  __device__ int n_iters(int cta, int ncta) const { const int n = n_units(); return cta < n ? (n - cta + ncta - 1) / ncta : 0; }
  __device__ Unit unit(int cta, int ncta, int it) const {
    const int u = cta + it * ncta;
    const int tiles = m_tiles() * g.tiles_n;
    const int z = u / tiles, r = u - z * tiles;
    return {r / g.tiles_n, r % g.tiles_n, z};
  }
  __device__ int k_chunks(Unit un) const {
    const int c0 = un.z * chunks_per_split;
    const int c1 = min(chunks_total, c0 + chunks_per_split);
    return (c1 - c0) * sk.nparts;
  }
  __device__ uint32_t tx_bytes() const { return (uint32_t)((L0 ? 0 : A_STAGE_BYTES) + g.BN * BK * 2); }
  __device__ void prefetch() const {
    if (!L0) prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    if (sk.nparts > 1) { if (!L0) prefetch_tmap(&mapA2); prefetch_tmap(&mapB2); }
  }
  __device__ void load_a(uint8_t* s, uint64_t* bar, Unit un, int kc) const {
    const int m0 = (un.z * chunks_per_split + sk.base(kc)) * BK;   // first of the 64 positions of this stage
    const int b0 = m0 >> (2 * g.lgHo), h0 = (m0 >> g.lgHo) & (g.Ho - 1);
    const CUtensorMap* map = sk.a_lo(kc) ? &mapA2 : &mapA;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int kk0 = (un.m_tile * (kRows / 64) + i) * 64;
      const int dh = kk0 / (2 * g.Pp), c0 = kk0 - dh * 2 * g.Pp;
      tma_load_5d(s + i * 8192, map, bar, c0, 0, dh, h0, b0);
    }
  }
  __device__ void load_b(uint8_t* s, uint64_t* bar, Unit un, int kc) const {
    const int m0 = (un.z * chunks_per_split + sk.base(kc)) * BK;
    const CUtensorMap* map = sk.b_lo(kc) ? &mapB2 : &mapB;
    for (int i = 0; i < g.BN / 64; ++i) tma_load_2d(s + i * 8192, map, bar, un.n_tile * g.BN + i * 64, m0);
  }
  // ---- layer 0: rows of the A stage are cube channels k = p*4 + dh*2 + dw, columns 64 positions ----
  // A thread owns one channel row (pair, dh, dw) for the whole unit.  Its 16 values o_j[2w+dw] only
  // change with the sample (every 4 stages), so they live in registers; per stage it reads two
  // o_i[2h+dh] scalars and writes 4 x 16 bytes.
  struct SynthState { float oj[16]; int sample; int oi_off; int oj_off; };
  __device__ void synth_begin(Unit un, uint8_t* ex, int t, SynthState& st) const {
    const int KS = g.K + 4;                               // padded row stride: rows of different fields hit different banks
    int* staged = reinterpret_cast<int*>(ex + (g.F + 1) * KS * 4);
    if (t == 0) *staged = -1;                             // sample whose rows are staged
    for (int e = t; e < KS; e += 256) reinterpret_cast<float*>(ex)[g.F * KS + e] = 0.f;   // zero row for padded pairs
    st.sample = -1;
    const int kk = un.m_tile * kRows + t;               // this thread's channel row of the 256-row stage
    const int pr = kk >> 2, dh = (kk >> 1) & 1, dw = kk & 1;
    const bool live = pr < g.P;
    st.oi_off = (live ? pair_i[pr] : g.F) * KS + dh;
    st.oj_off = (live ? pair_j[pr] : g.F) * KS + dw;
  }
  __device__ void synth_a(uint8_t* sA, Unit un, int kc_in, int t256, const uint8_t* ex_c, SynthState& st) const {
    uint8_t* ex = const_cast<uint8_t*>(ex_c);
    const int kc = sk.base(kc_in);
    const bool lo_part = sk.a_lo(kc_in);
    const int t = t256 & 127, tile = t256 >> 7;          // row inside its 128-row tile; which of the two A tiles
    const int KS = g.K + 4;
    float* o = reinterpret_cast<float*>(ex);
    int* staged = reinterpret_cast<int*>(ex + (g.F + 1) * KS * 4);
    const int m0 = (un.z * chunks_per_split + kc) * BK;   // 64 positions: 4 h-rows x 16 w of one sample
    const int b = m0 >> 8, hb = (m0 >> 4) & 15;
    if (*staged != b) {                                   // uniform across the 256 producer threads
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int K4 = g.K / 4;
      const float4* src = reinterpret_cast<const float4*>(rows + (int64_t)b * g.F * g.K);
      for (int e = t256; e < g.F * K4; e += 256) {
        const int f = e / K4, c4 = e - f * K4;
        reinterpret_cast<float4*>(o + f * KS)[c4] = b < g.B ? __ldg(src + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (t256 == 0) *staged = b;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    if (st.sample != b) {
      st.sample = b;
#pragma unroll
      for (int wv = 0; wv < 16; ++wv) st.oj[wv] = o[st.oj_off + 2 * wv];
    }
    const float* oi = o + st.oi_off + 2 * hb;
    uint8_t* sT = sA + tile * A_STAGE_BYTES;
#pragma unroll
    for (int c = 0; c < 8; ++c) {                         // chunk c: 8 positions, h = hb + c/2, w = (c&1)*8 .. +7
      const float a = oi[2 * (c >> 1)];
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        pk[e] = (SPLIT && lo_part) ? pack2_lo(a * st.oj[(c & 1) * 8 + 2 * e], a * st.oj[(c & 1) * 8 + 2 * e + 1])
                        : pack2(a * st.oj[(c & 1) * 8 + 2 * e], a * st.oj[(c & 1) * 8 + 2 * e + 1]);
      *reinterpret_cast<uint4*>(sT + sw128_offset(t, c)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
  struct Epilogue {
    const ConvWgradTC& p; int row;
    __device__ Epilogue(const ConvWgradTC& p_, uint8_t*, int row_, int) : p(p_), row(row_) {}
    __device__ void begin(Unit) {}
    __device__ void chunk(Unit un, int c0, const float (&v)[32]) {
      const int at = c0 / p.g.BN, cc = c0 - at * p.g.BN;   // accumulator of A tile `at`
      float4* dst = reinterpret_cast<float4*>(p.partial + ((int64_t)un.z * 4 * p.g.Pp + un.m_tile * kRows + at * BM + row) * p.g.Pp +
                                              un.n_tile * p.g.BN + cc);
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) dst[q4] = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
    }
    __device__ void end(Unit) {}
    __device__ void finish() {}
  };
};

// =================================================================================================
// Small kernels around the GEMMs
// =================================================================================================
// bf16 operand copies of the fp32 master filters W_l[(tap,p), q] (HWIO, CFFM.py:376-377):
//   Wt[q][kk] (forward B operand) and Wd[kk][q] (data-gradient B operand), zero padded, with
//   kk = tap*Pp + p for l >= 1 and kk = p*4 + tap for layer 0 (the order the cube is synthesised in).
struct PrepLayers { const float* W[kMaxConv]; bf16 *Wt[kMaxConv], *Wd[kMaxConv], *Wt_lo[kMaxConv], *Wd_lo[kMaxConv]; int l0[kMaxConv]; };
__global__ void k_prep_weights(const PrepLayers a, int P, int Pp) {
  const int ly = blockIdx.y;                                 // all layers in one launch
  const float* __restrict__ W = a.W[ly];
  bf16* __restrict__ Wt = a.Wt[ly]; bf16* __restrict__ Wd = a.Wd[ly];
  bf16* __restrict__ Wt_lo = a.Wt_lo[ly]; bf16* __restrict__ Wd_lo = a.Wd_lo[ly];
  const int l0 = a.l0[ly];
  const int64_t total = (int64_t)4 * Pp * Pp;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(e / Pp), q = (int)(e - (int64_t)kk * Pp);
    int tap, pr;
    if (l0) { pr = kk >> 2; tap = kk & 3; } else { tap = kk / Pp; pr = kk - tap * Pp; }
    const float v = (pr < P && q < P) ? W[((int64_t)tap * P + pr) * P + q] : 0.f;
    const bf16 b = __float2bfloat16_rn(v);
    Wd[e] = b;
    Wt[(int64_t)q * 4 * Pp + kk] = b;
    if (Wt_lo) {   // split mode: lo = bf16(v - hi)
      const bf16 bl = __float2bfloat16_rn(v - __bfloat162float(b));
      Wd_lo[e] = bl;
      Wt_lo[(int64_t)q * 4 * Pp + kk] = bl;
    }
  }
}

// dW[(tap,p), q] = sum_split partial[split][kk][q]
__global__ void k_wgrad_reduce(const float* __restrict__ partial, int n_split, int P, int Pp, int l0, float* __restrict__ gW) {
  const int64_t total = (int64_t)4 * P * P;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / P), q = (int)(e - (int64_t)r * P);
    const int tap = r / P, pr = r - tap * P;
    const int kk = l0 ? pr * 4 + tap : tap * Pp + pr;
    float s = 0.f;
    for (int z = 0; z < n_split; ++z) s += partial[((int64_t)z * 4 * Pp + kk) * Pp + q];
    gW[e] = s;
  }
}

// d b_l[q] = sum_m dY_l[m, q]: chunk partials over the rows, then the chunks in order.
// A thread owns 8 adjacent channels (one 16-byte load per row) and every R-th row of the chunk; the R
// row lanes are combined through shared memory in a fixed order.  blockDim = (Pp/8) * R.
__global__ void k_colsum_bf16(const bf16* __restrict__ X, const bf16* __restrict__ Xlo, int64_t rows, int Pp, int n,
                              float* __restrict__ partial, int C) {
  extern __shared__ float red[];               // [R][Pp]
  const int CG = Pp >> 3;
  const int R = blockDim.x / CG;
  const int cg = threadIdx.x % CG, rl = threadIdx.x / CG;
  const int c = blockIdx.x;
  const int64_t rpc = (rows + C - 1) / C;
  const int64_t r0 = (int64_t)c * rpc;
  const int64_t r1 = r0 + rpc < rows ? r0 + rpc : rows;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (rl < R) {
    for (int64_t r = r0 + rl; r < r1; r += R) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(X + r * Pp) + cg);
      const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[2 * j] += __uint_as_float(wv[j] << 16);
        acc[2 * j + 1] += __uint_as_float(wv[j] & 0xFFFF0000u);
      }
      if (Xlo) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(Xlo + r * Pp) + cg);
        const uint32_t uv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[2 * j] += __uint_as_float(uv[j] << 16);
          acc[2 * j + 1] += __uint_as_float(uv[j] & 0xFFFF0000u);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[rl * Pp + cg * 8 + j] = acc[j];
  }
  __syncthreads();
  for (int col = threadIdx.x; col < n; col += blockDim.x) {
    float t = 0.f;
    for (int q = 0; q < R; ++q) t += red[q * Pp + col];
    partial[(int64_t)c * n + col] = t;
  }
}
// All layers' bias gradients in one launch pair (blockIdx.y = layer): the per-layer pair costs two launches per layer,
// which is most of their time at the reference's dataset sizes.
struct ColsumLayers {
  const bf16* X[kMaxConv]; const bf16* Xlo[kMaxConv]; int64_t rows[kMaxConv]; int C[kMaxConv]; int64_t out_off[kMaxConv];
  int n_layers, Pp, n; float* partial; int64_t partial_stride;
};
__global__ void k_colsum_bf16_layers(const ColsumLayers a) {
  extern __shared__ float red[];               // [R][Pp]
  const int l = blockIdx.y;
  const int C = a.C[l];
  const int c = blockIdx.x;
  if (c >= C) return;
  const bf16* __restrict__ X = a.X[l];
  const bf16* __restrict__ Xlo = a.Xlo[l];
  const int Pp = a.Pp;
  const int CG = Pp >> 3;
  const int R = blockDim.x / CG;
  const int cg = threadIdx.x % CG, rl = threadIdx.x / CG;
  const int64_t rows = a.rows[l];
  const int64_t rpc = (rows + C - 1) / C;
  const int64_t r0 = (int64_t)c * rpc;
  const int64_t r1 = r0 + rpc < rows ? r0 + rpc : rows;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (rl < R) {
    for (int64_t r = r0 + rl; r < r1; r += R) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(X + r * Pp) + cg);
      const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[2 * j] += __uint_as_float(wv[j] << 16);
        acc[2 * j + 1] += __uint_as_float(wv[j] & 0xFFFF0000u);
      }
      if (Xlo) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(Xlo + r * Pp) + cg);
        const uint32_t uv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[2 * j] += __uint_as_float(uv[j] << 16);
          acc[2 * j + 1] += __uint_as_float(uv[j] & 0xFFFF0000u);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[rl * Pp + cg * 8 + j] = acc[j];
  }
  __syncthreads();
  float* partial = a.partial + (int64_t)l * a.partial_stride;
  for (int col = threadIdx.x; col < a.n; col += blockDim.x) {
    float t = 0.f;
    for (int q = 0; q < R; ++q) t += red[q * Pp + col];
    partial[(int64_t)c * a.n + col] = t;
  }
}
__global__ void k_sum_chunks_layers(const ColsumLayers a, float* __restrict__ g) {
  const int l = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.n) return;
  const float* partial = a.partial + (int64_t)l * a.partial_stride;
  float s = 0.f;
  for (int c = 0; c < a.C[l]; ++c) s += partial[(int64_t)c * a.n + j];
  g[a.out_off[l] + j] = s;
}
__global__ void k_sum_chunks(const float* __restrict__ partial, int n, int C, float* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += partial[(int64_t)c * n + j];
  out[j] = s;
}

// dY of the last live layer: X_{d-1} feeds sum_pooling[d-1] only
template <int ACT>
__global__ void k_dy_top_bf16(const bf16* __restrict__ X, const float* __restrict__ gout, const float* __restrict__ v_lvl,
                              int H, int Pp, int64_t total, bf16* __restrict__ dY, bf16* __restrict__ dY_lo) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int64_t pix = e / Pp;
  const int64_t bh = pix / H;
  const int h = (int)(bh % H);
  const int64_t b = bh / H;
  // (gelu: X points at the stored derivative phi'(Y))
  const float m = ACT == CFFM_ACT_GELU ? __bfloat162float(X[e]) : (__bfloat162float(X[e]) > 0.f ? phi_scale<ACT>() : 0.f);
  const float v = gout[b] * v_lvl[h] * m;
  const bf16 hi = __float2bfloat16_rn(v);
  dY[e] = hi;
  if (dY_lo) dY_lo[e] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

#include "conv0_fact.cuh"
#include "conv0_dfact.cuh"
#include "conv0_wfact.cuh"

}  // namespace tc

// =================================================================================================
// Host side
// =================================================================================================
using namespace tc;

struct TCState {
  int Pp = 0, BN = 0;
  bf16* X[kMaxConv + 1] = {};    // X[l], l = 1..n_live: phi(Y_{l-1}), NHWC [B, K>>l, K>>l, Pp]
  bf16* dY[kMaxConv] = {};       // dY[l], l = 0..n_live-1
  bf16* Wt[kMaxConv] = {};       // [Pp][4Pp]
  bf16* Wd[kMaxConv] = {};       // [4Pp][Pp]
  // split mode (CFFM_PREC_BF16X3): the lo halves of every operand tensor above
  bool split = false;
  bf16* Xlo[kMaxConv + 1] = {};
  bf16* dYlo[kMaxConv] = {};
  bf16* Wtlo[kMaxConv] = {};
  bf16* Wdlo[kMaxConv] = {};
  float* wg_partial = nullptr;   // weight-gradient split scratch
  int64_t wg_partial_floats = 0;
  float* bg_partial = nullptr;   // bias-gradient chunk scratch [64][P]
  float* pool_part = nullptr;    // forward layers >= 1: pooled sums per N tile [tiles_n][B * K/4]
  bf16* Dphi[kMaxConv + 1] = {}; // gelu, training: phi'(Y_{l-1}) in the layout of X[l]
  bf16* Wf0 = nullptr;           // layer-0 filters of the factorised forward: [Q16][KA][nblk*64] (conv0_fact.cuh)
  bf16* Wf0lo = nullptr;         // split mode: their lo halves
  bf16* X1i = nullptr;           // split mode: X_1 as the factorised forward writes it (hi / lo interleaved per 8 channels)
  bool dfact = false;            // the layer-0 data gradient runs in factorised form (conv0_dfact.cuh)
  float* df_bpart = nullptr;     // [tiles][4][Q16] column sums of dY0 collected by the factorised data gradient
  float2* pterm0 = nullptr;      // [B][F] pooling terms of the layer-0 data gradient
  float* wf_part = nullptr;      // [W0_SPLIT_MAX][Q16][KA][KA] partial sums of the factorised layer-0 weight gradient
  bf16* A8 = nullptr;            // [B8*16][nblk*64] bf16 rows a_{b,h} (A tiles of the factorised weight gradient)
  bf16* A8lo = nullptr;          // split mode: their lo halves
  int KA = 0, nblk = 0, Q16 = 0;
  int fact_min_batch = 512;      // the factorised kernels have a fixed cost (they win from a few hundred samples on)
  TmaEncoder enc;
};

static int ilog2i(int x) { int l = 0; while ((1 << (l + 1)) <= x) ++l; return l; }

static int pick_bn(int Pp) {  // largest multiple of 64 (<= 256) dividing Pp
  for (int bn = 256; bn > 64; bn -= 64) if (Pp % bn == 0) return bn;
  return 64;
}

template <class T>
static int tcmalloc(Model* m, T** p, int64_t n) {
  cudaError_t e = dev_malloc((void**)p, sizeof(T) * (size_t)(n > 0 ? n : 1));
  if (e != cudaSuccess) { m->err = std::string("cudaMalloc (bf16 path): ") + cudaGetErrorString(e); *p = nullptr; return CFFM_ERR_NOMEM; }
  return CFFM_OK;
}
#define TCTRY(x) do { int _r = (x); if (_r != CFFM_OK) return _r; } while (0)

// Scoring (forward only) runs on the tensor cores for outer_dims 16 / 32 / 64 and every activation (the sweep of
// BASELINE.json configs[4]); training needs outer_dims == 32; gelu, whose derivative does not follow from the stored
// post-activation, trains in bf16 mode with its derivative stored beside the activations (tc_train_supported).
int tc_supported(Model* m) {
  if (!m->cfg.outer_conv) return CFFM_OK;
  if (m->Ko != 16 && m->Ko != 32 && m->Ko != 64) { m->err = "tensor-core precisions need outer_dims 16, 32 or 64 (other sizes run in fp32)"; return CFFM_ERR_UNSUPPORTED; }
  if (m->F > 48) { m->err = "tensor-core precisions need num_field <= 48"; return CFFM_ERR_UNSUPPORTED; }  // dgrad0 smem: 11*F*K*4 B
  // layer-0 forward: the rows of a tile's samples + the pair table share 14 KB of producer scratch
  const int S = m->Ko == 16 ? 2 : 1, Pp = (m->P + 63) & ~63;
  if (S * (m->F + 1) * m->Ko * 4 + Pp * 4 > 14 * 1024) { m->err = "tensor-core precisions: num_field too large for this outer_dims"; return CFFM_ERR_UNSUPPORTED; }
  return CFFM_OK;
}
int tc_train_supported(Model* m) {
  if (!m->cfg.outer_conv) return CFFM_OK;
  if (m->Ko != 32) { m->err = "training in a tensor-core precision needs outer_dims == 32 (16 / 64: scoring only; train in fp32)"; return CFFM_ERR_UNSUPPORTED; }
  if (m->cfg.activation == CFFM_ACT_GELU && m->cfg.precision != CFFM_PREC_BF16) {
    m->err = "gelu trains in fp32 or bf16 (bf16 stores its derivative in bf16 beside the activations; the split mode does not)";
    return CFFM_ERR_UNSUPPORTED;
  }
  return CFFM_OK;
}

int tc_alloc(Model* m, bool train) {
  if (!m->cfg.outer_conv) return CFFM_OK;
  TCState* st = reinterpret_cast<TCState*>(m->tcs);
  if (!st) {
    st = new TCState();
    m->tcs = st;
    if (!st->enc.init()) { m->err = "cuTensorMapEncodeTiled is not available"; return CFFM_ERR_CUDA; }
    st->Pp = (m->P + 63) & ~63;   // one 64-element swizzle atom is the unit of every operand box
    st->BN = pick_bn(st->Pp);
    st->split = m->cfg.precision == CFFM_PREC_BF16X3;
    const int64_t B = m->max_batch, Pp = st->Pp;
    for (int l = 1; l <= m->n_live; ++l) { const int64_t H = m->Ko >> l; TCTRY(tcmalloc(m, &st->X[l], B * H * H * Pp)); }
    for (int l = 0; l < m->n_live; ++l) { TCTRY(tcmalloc(m, &st->Wt[l], 4 * Pp * Pp)); TCTRY(tcmalloc(m, &st->Wd[l], 4 * Pp * Pp)); }
    if (st->split) {
      for (int l = 1; l <= m->n_live; ++l) { const int64_t H = m->Ko >> l; TCTRY(tcmalloc(m, &st->Xlo[l], B * H * H * Pp)); }
      for (int l = 0; l < m->n_live; ++l) { TCTRY(tcmalloc(m, &st->Wtlo[l], 4 * Pp * Pp)); TCTRY(tcmalloc(m, &st->Wdlo[l], 4 * Pp * Pp)); }
    }
    TCTRY(tcmalloc(m, &st->pool_part, (Pp / st->BN) * B * (m->Ko >> 2)));
    const char* f0 = getenv("CFFM_FWD0");
    const char* mb = getenv("CFFM_FACT_MIN_BATCH");
    const char* mf = getenv("CFFM_FACT_MIN_FIELDS");
    if (mb) st->fact_min_batch = atoi(mb);
    // factorised layer-0 kernels: worthwhile when the direct form is big (P = F(F-1)/2 channels) and the batch is not tiny
    // (split mode: the kernels split their intermediates (Z, E^T, E) into hi + lo as well)
    // (gelu keeps layer 0 in the direct form: only the k_tc epilogues store the derivative it trains with)
    if (m->Ko == 32 && m->cfg.activation != CFFM_ACT_GELU && 2 * m->F <= F0_KA_MAX && m->F >= (mf ? atoi(mf) : 16) &&
        !(f0 && !strcmp(f0, "direct"))) {
      st->KA = (2 * m->F + 15) & ~15; st->nblk = st->KA > 64 ? 2 : 1; st->Q16 = (m->P + 15) & ~15;
      const int64_t n = (int64_t)st->Q16 * st->KA * st->nblk * 64;
      TCTRY(tcmalloc(m, &st->Wf0, n));
      CFFM_CUDA_OK(m, cudaMemset(st->Wf0, 0, sizeof(bf16) * (size_t)n));
      // channels Q16..Pp-1 of X1 are never written by that kernel
      CFFM_CUDA_OK(m, cudaMemset(st->X[1], 0, sizeof(bf16) * (size_t)(B * (m->Ko >> 1) * (m->Ko >> 1) * Pp)));
      if (st->split) {
        TCTRY(tcmalloc(m, &st->Wf0lo, n));
        CFFM_CUDA_OK(m, cudaMemset(st->Wf0lo, 0, sizeof(bf16) * (size_t)n));
        const int64_t ni = 2 * B * (m->Ko >> 1) * (m->Ko >> 1) * Pp;
        TCTRY(tcmalloc(m, &st->X1i, ni));
        CFFM_CUDA_OK(m, cudaMemset(st->X1i, 0, sizeof(bf16) * (size_t)ni));   // channels Q16..Pp-1 are never written
        CFFM_CUDA_OK(m, cudaMemset(st->Xlo[1], 0, sizeof(bf16) * (size_t)(B * (m->Ko >> 1) * (m->Ko >> 1) * Pp)));
      }
    }
  }
  if (train && !st->dY[0]) {
    TCTRY(tc_train_supported(m));
    const int64_t B = m->max_batch, Pp = st->Pp;
    for (int l = 0; l < m->n_live; ++l) { const int64_t H = m->Ko >> (l + 1); TCTRY(tcmalloc(m, &st->dY[l], B * H * H * Pp)); }
    if (st->split)
      for (int l = 0; l < m->n_live; ++l) { const int64_t H = m->Ko >> (l + 1); TCTRY(tcmalloc(m, &st->dYlo[l], B * H * H * Pp)); }
    if (m->cfg.activation == CFFM_ACT_GELU)
      for (int l = 1; l <= m->n_live; ++l) { const int64_t H = m->Ko >> l; TCTRY(tcmalloc(m, &st->Dphi[l], B * H * H * Pp)); }
    // split factor of the largest weight gradient decides the scratch size
    const int tiles = (4 * st->Pp / BM) * (st->Pp / st->BN);
    int max_split = (2 * 148 + tiles - 1) / tiles; if (max_split < 1) max_split = 1;
    st->wg_partial_floats = (int64_t)max_split * 4 * Pp * Pp;
    TCTRY(tcmalloc(m, &st->wg_partial, st->wg_partial_floats));
    TCTRY(tcmalloc(m, &st->bg_partial, (int64_t)kMaxConv * 2 * 148 * (int64_t)m->P));
    const char* d0 = getenv("CFFM_DGRAD0");
    if (st->Wf0 && !(d0 && !strcmp(d0, "direct"))) {   // factorised layer-0 data gradient
      st->dfact = true;
      TCTRY(tcmalloc(m, &st->pterm0, B * m->F));
      TCTRY(tcmalloc(m, &st->df_bpart, ((B + 7) / 8) * (st->split ? 8 : 4) * (int64_t)st->Q16));
    }
    const char* w0 = getenv("CFFM_WGRAD0");
    if (st->Wf0 && !(w0 && !strcmp(w0, "direct")))     // factorised layer-0 weight gradient
    {
      TCTRY(tcmalloc(m, &st->wf_part, (int64_t)W0_SPLIT_MAX * st->Q16 * st->KA * st->KA));
      TCTRY(tcmalloc(m, &st->A8, ((B + 7) / 8 * 8) * 16 * st->nblk * 64));
      if (st->split) TCTRY(tcmalloc(m, &st->A8lo, ((B + 7) / 8 * 8) * 16 * st->nblk * 64));
    }
  }
  return CFFM_OK;
}

void tc_free(Model* m) {
  TCState* st = reinterpret_cast<TCState*>(m->tcs);
  if (!st) return;
  for (int l = 0; l <= kMaxConv; ++l) { if (st->X[l]) dev_free(st->X[l]); if (st->Xlo[l]) dev_free(st->Xlo[l]); if (st->Dphi[l]) dev_free(st->Dphi[l]); }
  for (int l = 0; l < kMaxConv; ++l) {
    if (st->dY[l]) dev_free(st->dY[l]); if (st->Wt[l]) dev_free(st->Wt[l]); if (st->Wd[l]) dev_free(st->Wd[l]);
    if (st->dYlo[l]) dev_free(st->dYlo[l]); if (st->Wtlo[l]) dev_free(st->Wtlo[l]); if (st->Wdlo[l]) dev_free(st->Wdlo[l]);
  }
  if (st->wg_partial) dev_free(st->wg_partial);
  if (st->bg_partial) dev_free(st->bg_partial);
  if (st->pool_part) dev_free(st->pool_part);
  if (st->Wf0) dev_free(st->Wf0);
  if (st->Wf0lo) dev_free(st->Wf0lo);
  if (st->X1i) dev_free(st->X1i);
  if (st->pterm0) dev_free(st->pterm0);
  if (st->df_bpart) dev_free(st->df_bpart);
  if (st->wf_part) dev_free(st->wf_part);
  if (st->A8) dev_free(st->A8);
  if (st->A8lo) dev_free(st->A8lo);
  delete st;
  m->tcs = nullptr;
}

static Geom make_geom(const Model* m, const TCState* st, int B, int l) {
  Geom g;
  g.B = B; g.P = m->P; g.Pp = st->Pp; g.F = m->F; g.K = m->Ko;
  g.Hin = m->Ko >> l; g.Ho = g.Hin >> 1; g.lgHo = ilog2i(g.Ho);
  g.M = B * g.Ho * g.Ho; g.BN = st->BN; g.tiles_n = st->Pp / st->BN;
  return g;
}

// 5-D im2col view of an NHWC bf16 tensor [B, Hin, Hin, Pp]: (c = (dw,p), w, dh, h, b); box = 64 channels x `rows` positions
static bool im2col_map(const TCState* st, CUtensorMap* map, const bf16* X, int B, int Hin, int Pp, int rows) {
  const int Ho = Hin / 2;
  const uint64_t dims[5] = {(uint64_t)2 * Pp, (uint64_t)Ho, 2, (uint64_t)Ho, (uint64_t)B};
  const uint64_t str[4] = {(uint64_t)2 * Pp * 2, (uint64_t)Hin * Pp * 2, (uint64_t)2 * Hin * Pp * 2, (uint64_t)Hin * Hin * Pp * 2};
  int bw = Ho, bh = rows / bw; if (bh > Ho) bh = Ho; if (bh < 1) bh = 1;
  int bb = rows / (bw * bh); if (bb < 1) bb = 1;
  const uint32_t box[5] = {64, (uint32_t)bw, 1, (uint32_t)bh, (uint32_t)bb};
  return st->enc.encode_bf16(map, const_cast<bf16*>(X), 5, dims, str, box);
}
static bool mat_map(const TCState* st, CUtensorMap* map, const bf16* X, int64_t rows, int cols, int box_rows, int box_cols) {
  const uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  const uint64_t str[1] = {(uint64_t)cols * 2};
  const uint32_t box[2] = {(uint32_t)box_cols, (uint32_t)box_rows};
  return st->enc.encode_bf16(map, const_cast<bf16*>(X), 2, dims, str, box);
}

template <class Pol>
static int launch_tc(Model* m, const Pol& p, int units_hint, cudaStream_t s) {
  static PerDeviceOnce attr_once;
  bool& attr_done = attr_once();
  if (!attr_done) {
    static_assert(smem_bytes<Pol>() <= (size_t)SMEM_LIMIT, "policy exceeds the shared memory of an SM");
    CFFM_CUDA_OK(m, cudaFuncSetAttribute(k_tc<Pol>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<Pol>()));
    attr_done = true;
  }
  int grid = units_hint < 148 ? units_hint : 148;
  if (grid < 1) grid = 1;
  k_tc<Pol><<<grid, block_threads<Pol>(), smem_bytes<Pol>(), s>>>(p);
  m->launches++;
  CFFM_CUDA_OK(m, cudaGetLastError());
  return CFFM_OK;
}
#define TC_MAP_OK(m, ok) do { if (!(ok)) { (m)->err = "cuTensorMapEncodeTiled failed"; return CFFM_ERR_CUDA; } } while (0)

template <int ACT, bool SPLIT>
static int fwd0_fact_launch(Model* m, TCState* st, int B, int sp_off, cudaStream_t s) {
  Fwd0FactParams p;
  memset(&p.mapW, 0, sizeof(p.mapW)); memset(&p.mapW2, 0, sizeof(p.mapW2)); memset(&p.mapX2, 0, sizeof(p.mapX2));
  TC_MAP_OK(m, mat_map(st, &p.mapW, st->Wf0, (int64_t)st->Q16 * st->KA, st->nblk * 64, st->KA, 64));
  if (SPLIT) TC_MAP_OK(m, mat_map(st, &p.mapW2, st->Wf0lo, (int64_t)st->Q16 * st->KA, st->nblk * 64, st->KA, 64));
  bf16* xdst = SPLIT ? st->X1i : st->X[1];
  {  // X1 seen as (q, row = b*16+h, w) for the epilogue's dense (16, 32, 8) TMA stores; split mode: the interleaved scratch,
     // a pixel row is 2*Pp elements and a box covers the 16 elements [8 hi | 8 lo] of an 8-channel group
    const uint64_t rowel = (uint64_t)st->Pp * (SPLIT ? 2 : 1);
    const uint64_t dims[3] = {rowel, (uint64_t)B * 16, 16};
    const uint64_t str[2] = {(uint64_t)16 * rowel * 2, rowel * 2};
    const uint32_t box[3] = {16, 32, 8};
    const char* e = getenv("CFFM_F0_TMASTORE");
    p.tma_store = !(e && !strcmp(e, "0")) && st->enc.encode_bf16(&p.mapX, xdst, 3, dims, str, box, false) ? 1 : 0;
    if (!p.tma_store) { memset(&p.mapX, 0, sizeof(p.mapX)); memset(&p.mapX2, 0, sizeof(p.mapX2)); }
  }
  p.rows = m->outer_rows; p.bias = m->dense_w + m->lay.conv_b[0]; p.Xout = xdst; p.Xout_lo = st->Xlo[1];
  p.t1 = m->t1; p.t1_dim = m->t1_dim; p.sp_off = sp_off;
  p.B = B; p.F = m->F; p.P = m->P; p.Pp = st->Pp; p.KA = st->KA; p.nblk = st->nblk; p.Q16 = st->Q16;
  constexpr int SMEM = SPLIT ? F0S_SMEM : F0_SMEM;
  static PerDeviceOnce attr_once;
  bool& attr_done = attr_once();
  if (!attr_done) {
    CFFM_CUDA_OK(m, cudaFuncSetAttribute(k_fwd0_fact<ACT, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_done = true;
  }
  int grid = (B + 7) / 8; if (grid > 148) grid = 148;
  k_fwd0_fact<ACT, SPLIT><<<grid, F0_THREADS, SMEM, s>>>(p);
  m->launches++;
  if (SPLIT) {   // interleaved scratch -> the separate hi / lo tensors of the next layer's TMA boxes
    const int64_t units = (int64_t)B * 256 * st->Pp / 8;
    k_deinterleave_x1<<<148 * 16, 256, 0, s>>>(reinterpret_cast<const uint4*>(st->X1i), units, reinterpret_cast<uint4*>(st->X[1]),
                                              reinterpret_cast<uint4*>(st->Xlo[1]));
    m->launches++;
  }
  CFFM_CUDA_OK(m, cudaGetLastError());
  return CFFM_OK;
}

template <bool SPLIT>
static int dgrad0_fact_launch(Model* m, TCState* st, int B, cudaStream_t s) {
  Dgrad0FactParams p;
  memset(&p.mapW, 0, sizeof(p.mapW)); memset(&p.mapW2, 0, sizeof(p.mapW2));
  TC_MAP_OK(m, mat_map(st, &p.mapW, st->Wf0, (int64_t)st->Q16 * st->KA, st->nblk * 64, st->KA, 64));
  if (SPLIT) TC_MAP_OK(m, mat_map(st, &p.mapW2, st->Wf0lo, (int64_t)st->Q16 * st->KA, st->nblk * 64, st->KA, 64));
  p.dY = st->dY[0]; p.dYlo = st->dYlo[0]; p.rows = m->outer_rows; p.gout = m->gout; p.v_head = m->v_head; p.pterm = st->pterm0; p.g_rows = m->g_outer_rows;
  p.bpart = st->df_bpart;
  p.B = B; p.F = m->F; p.P = m->P; p.Pp = st->Pp; p.KA = st->KA; p.nblk = st->nblk; p.Q16 = st->Q16;
  static PerDeviceOnce attr_once;
  bool& attr_done = attr_once();
  if (!attr_done) {
    CFFM_CUDA_OK(m, cudaFuncSetAttribute(k_dgrad0_fact<SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, g0_smem(SPLIT)));
    attr_done = true;
  }
  k_pool_terms0<<<(B + 7) / 8, 256, 0, s>>>(m->outer_rows, m->v_head, B, m->F, st->pterm0);
  int grid = (B + 7) / 8; if (grid > 148) grid = 148;
  k_dgrad0_fact<SPLIT><<<grid, G0_THREADS, g0_smem(SPLIT), s>>>(p);
  // bias gradient of layer 0 from the column sums the builders collected (replaces the colsum pass over dY0)
  k_dfact_bias_reduce<<<ceil_div(m->P, 32), dim3(32, 32), 0, s>>>(st->df_bpart, ((B + 7) / 8) * (SPLIT ? 8 : 4), st->Q16, m->P, m->dense_g + m->lay.conv_b[0]);
  m->launches += 3;
  CFFM_CUDA_OK(m, cudaGetLastError());
  return CFFM_OK;
}

int tc_prep_weights(Model* m, int B, cudaStream_t s) {
  TCState* st = reinterpret_cast<TCState*>(m->tcs);
  CFFM_PROF(m, "prep_weights_bf16", s);
  // layer 0: the direct kernels' bf16 copies are not needed when every layer-0 kernel of this call is the factorised one
  const bool fact0 = st->Wf0 && B >= st->fact_min_batch && (!st->dY[0] || (st->dfact && st->wf_part));
  PrepLayers pa;
  memset(&pa, 0, sizeof(pa));
  int nl = 0;
  for (int l = fact0 ? 1 : 0; l < m->n_live; ++l, ++nl) {
    pa.W[nl] = m->dense_w + m->lay.conv_w[l]; pa.Wt[nl] = st->Wt[l]; pa.Wd[nl] = st->Wd[l];
    pa.Wt_lo[nl] = st->Wtlo[l]; pa.Wd_lo[nl] = st->Wdlo[l]; pa.l0[nl] = l == 0 ? 1 : 0;
  }
  if (nl > 0) {
    const int64_t total = 4ll * st->Pp * st->Pp;
    int blocks = (int)((total + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
    k_prep_weights<<<dim3(blocks, nl), 256, 0, s>>>(pa, m->P, st->Pp);
    m->launches++;
  }
  if (st->Wf0 && B >= st->fact_min_batch) {
    k_prep_w0_fact<<<148 * 8, 256, 0, s>>>(m->dense_w + m->lay.conv_w[0], m->pair_i, m->pair_j, m->P, st->KA, st->nblk * 64, st->Wf0,
                                           st->Wf0lo);
    m->launches++;
  }
  CFFM_CUDA_OK(m, cudaGetLastError());
  return CFFM_OK;
}

template <int ACT, bool SPLIT>
static int conv_forward_act(Model* m, int B, cudaStream_t s) {
  TCState* st = reinterpret_cast<TCState*>(m->tcs);
  const int K = m->Ko, Pp = st->Pp;
  int off = K;
  for (int l = 0; l < m->n_live; ++l) {
    const Geom g = make_geom(m, st, B, l);
    const std::string tag = "conv_fwd_l" + std::to_string(l);
    CFFM_PROF(m, tag.c_str(), s);
    const int m_tiles = (g.M + BM - 1) / BM;
    if (l == 0 && st->Wf0 && B >= st->fact_min_batch) {
      TCTRY((fwd0_fact_launch<ACT, SPLIT>(m, st, B, off, s)));
    } else if (l == 0) {
      ConvFwdTC<ACT, true, SPLIT> p;
      // two accumulators (one per N tile of a unit) + four A stages share the 512 TMEM columns: N <= 192.
      // An even number of N tiles covers Pp with the least overhang (weights beyond Pp: TMA zero fill).
      Geom g0 = g;
      int best = 1 << 30;
      for (int t = 1; t <= Pp / 32; ++t) {
        const int bn = ((Pp + 2 * t - 1) / (2 * t) + 31) / 32 * 32;
        if (bn > PAIR_BN_MAX) continue;
        if (2 * t * bn < best) { best = 2 * t * bn; g0.BN = bn; g0.tiles_n = 2 * t; }
      }
      p.g = g0; p.bias = m->dense_w + m->lay.conv_b[0]; p.Xout = st->X[1]; p.t1 = m->t1; p.t1_dim = m->t1_dim; p.sp_off = off;
      p.rows = m->outer_rows; p.pair_i = m->pair_i; p.pair_j = m->pair_j; p.pool_part = nullptr;
      p.sk.nparts = st->split ? 3 : 1; p.Xout_lo = st->Xlo[1]; p.Dphi = st->Dphi[1];
      memset(&p.mapA, 0, sizeof(p.mapA)); memset(&p.mapA2, 0, sizeof(p.mapA2)); memset(&p.mapB2, 0, sizeof(p.mapB2));
      TC_MAP_OK(m, mat_map(st, &p.mapB, st->Wt[0], Pp, 4 * Pp, g0.BN, 64));
      if (st->split) TC_MAP_OK(m, mat_map(st, &p.mapB2, st->Wtlo[0], Pp, 4 * Pp, g0.BN, 64));
      TCTRY(launch_tc(m, p, m_tiles, s));
    } else {
      ConvFwdTC<ACT, false, SPLIT> p;
      p.g = g; p.bias = m->dense_w + m->lay.conv_b[l]; p.Xout = st->X[l + 1]; p.t1 = m->t1; p.t1_dim = m->t1_dim; p.sp_off = off;
      p.rows = nullptr; p.pair_i = nullptr; p.pair_j = nullptr; p.pool_part = st->pool_part;
      p.sk.nparts = st->split ? 3 : 1; p.Xout_lo = st->Xlo[l + 1]; p.Dphi = st->Dphi[l + 1];
      memset(&p.mapA2, 0, sizeof(p.mapA2)); memset(&p.mapB2, 0, sizeof(p.mapB2));
      TC_MAP_OK(m, im2col_map(st, &p.mapA, st->X[l], B, g.Hin, Pp, BM));
      TC_MAP_OK(m, mat_map(st, &p.mapB, st->Wt[l], Pp, 4 * Pp, g.BN, 64));
      if (st->split) {
        TC_MAP_OK(m, im2col_map(st, &p.mapA2, st->Xlo[l], B, g.Hin, Pp, BM));
        TC_MAP_OK(m, mat_map(st, &p.mapB2, st->Wtlo[l], Pp, 4 * Pp, g.BN, 64));
      }
      TCTRY(launch_tc(m, p, m_tiles * g.tiles_n, s));
      if (g.tiles_n > 1) {
        k_pool_parts<<<(B * g.Ho + 255) / 256, 256, 0, s>>>(st->pool_part, g.tiles_n, B, g.Ho, m->t1, m->t1_dim, off);
        m->launches++;
      }
    }
    off += g.Ho;
  }
  return CFFM_OK;
}

int tc_conv_forward(Model* m, int B, cudaStream_t s) {
  int r = tc_prep_weights(m, B, s);
  if (r != CFFM_OK) return r;
  const bool split = reinterpret_cast<TCState*>(m->tcs)->split;
  CFFM_DISPATCH_PHI(m->cfg.activation, r = split ? conv_forward_act<ACT, true>(m, B, s) : conv_forward_act<ACT, false>(m, B, s));
  return r;
}

template <int ACT, bool SPLIT>
static int conv_backward_act(Model* m, int B, cudaStream_t s) {
  TCState* st = reinterpret_cast<TCState*>(m->tcs);
  const int K = m->Ko, Pp = st->Pp, P = m->P;
  float* g = m->dense_g;
  cudaStream_t wside = s;   // the side stream the weight gradients were put on, if any
  int lvl_off[kMaxConv + 1]; lvl_off[0] = 0;
  for (int l = 0; l < m->conv_depth; ++l) lvl_off[l + 1] = lvl_off[l] + (K >> l);
  {  // top of the stack
    const int l = m->n_live - 1;
    const int H = K >> (l + 1);
    const int64_t total = (int64_t)B * H * H * Pp;
    CFFM_PROF(m, "dy_top", s);
    k_dy_top_bf16<ACT><<<ceil_div(total, 256), 256, 0, s>>>(ACT == CFFM_ACT_GELU ? st->Dphi[l + 1] : st->X[l + 1], m->gout, m->v_head + lvl_off[l + 1], H, Pp, total, st->dY[l],
                                                            st->dYlo[l]);
    m->launches++;
  }
  // Bias gradients d b_l = column sums of dY_l.  Big batches: per layer, right when dY_l has been written (it is still in
  // L2).  Small batches (the reference's datasets): all layers in one launch pair after the loop -- at those sizes the
  // eight launches cost more than the sums.
  const bool colsum_late = small_step(m, B);
  // small batches: the column sums of all layers in one launch pair, once every dY_l is complete
  auto colsum_all_layers = [&](cudaStream_t cs) {
    CFFM_PROF(m, "colsum", cs);
    ColsumLayers a;
    memset(&a, 0, sizeof(a));
    a.Pp = Pp; a.n = P; a.partial = st->bg_partial; a.partial_stride = (int64_t)2 * 148 * P;
    int maxC = 1, nl = 0;
    for (int l = 0; l < m->n_live; ++l) {
      if (l == 0 && st->df_bpart && B >= st->fact_min_batch) continue;   // collected by k_dgrad0_fact
      const int64_t rows = (int64_t)B * (K >> (l + 1)) * (K >> (l + 1));
      a.X[nl] = st->dY[l]; a.Xlo[nl] = st->dYlo[l]; a.rows[nl] = rows;
      a.C[nl] = (int)std::min<int64_t>(64, std::max<int64_t>(1, (rows + 63) / 64));   // few chunks: the second stage walks them one by one
      a.out_off[nl] = m->lay.conv_b[l];
      maxC = std::max(maxC, a.C[nl]);
      ++nl;
    }
    a.n_layers = nl;
    if (nl > 0) {
      const int CG = Pp / 8;
      int R = 512 / CG; if (R < 1) R = 1; if (R > 32) R = 32;
      k_colsum_bf16_layers<<<dim3(maxC, nl), CG * R, sizeof(float) * R * Pp, cs>>>(a);
      k_sum_chunks_layers<<<dim3(ceil_div(P, 128), nl), 128, 0, cs>>>(a, g);
      m->launches += 2;
    }
  };
  for (int l = m->n_live - 1; l >= 0; --l) {
    const Geom gm = make_geom(m, st, B, l);
    const int64_t rows = gm.M;
    if (!colsum_late && !(l == 0 && st->df_bpart && B >= st->fact_min_batch)) {  // (layer 0, factorised: collected by k_dgrad0_fact)
      CFFM_PROF(m, "colsum", s);
      const int C = (int)std::min<int64_t>(2 * 148, std::max<int64_t>(1, (rows + 63) / 64));
      const int CG = Pp / 8;
      int R = 512 / CG; if (R < 1) R = 1; if (R > 32) R = 32;
      k_colsum_bf16<<<C, CG * R, sizeof(float) * R * Pp, s>>>(st->dY[l], st->dYlo[l], rows, Pp, P, st->bg_partial, C);
      k_sum_chunks<<<ceil_div(P, 128), 128, 0, s>>>(st->bg_partial, P, C, g + m->lay.conv_b[l]);
      m->launches += 2;
    }
    {  // weight gradient: beside the data-gradient chain (side stream 1: dY_l is complete on `s` at this point; the
       // weight gradients of the layers follow one another there, so they can share their split-reduction scratch)
      const std::string tag = "conv_wgrad_l" + std::to_string(l);
      const cudaStream_t ws = colsum_late ? side_fork(m, s, 1) : s;   // (colsum_late = small_step: see model.h)
      if (ws != s) wside = ws;
      CFFM_PROF(m, tag.c_str(), ws);
      const int rows_per_unit = l == 0 ? 2 * BM : BM;
      const int tiles = (4 * Pp / rows_per_unit) * gm.tiles_n;
      const int chunks_total = (int)((rows + BK - 1) / BK);
      // split factor: fill whole waves of 148 persistent CTAs (units = tiles * split)
      int want = 1; double best = 0.0;
      // (a split needs work to amortise: at least 32 reduction chunks each -- at the reference's dataset sizes a 74-way
      // split left 14 chunks per CTA and a partial-sum reduction that took twice as long as the GEMM)
      const int max_split = (int)std::min<int64_t>(st->wg_partial_floats / (4ll * Pp * Pp), std::max(1, chunks_total / 32));
      for (int sfac = 1; sfac <= max_split && sfac * tiles <= 4 * 148; ++sfac) {
        const int units = sfac * tiles;
        const double eff = (double)units / (148.0 * ((units + 147) / 148));
        const double score = eff * (units >= 148 ? 1.0 : (double)units / 148.0);
        if (score > best + 1e-9) { best = score; want = sfac; }
      }
      const int cps = (chunks_total + want - 1) / want;
      const int n_split = (chunks_total + cps - 1) / cps;
      if (l == 0 && st->wf_part && B >= st->fact_min_batch) {
        Wgrad0FactParams p;
        const int B8 = (B + 7) / 8 * 8, KP = st->nblk * 64;
        k_build_a8<<<148 * 4, 256, 0, ws>>>(m->outer_rows, B, B8, m->F, KP, st->A8, st->A8lo);
        memset(&p.mapA, 0, sizeof(p.mapA)); memset(&p.mapA2, 0, sizeof(p.mapA2));
        TC_MAP_OK(m, mat_map(st, &p.mapA, st->A8, (int64_t)B8 * 16, KP, BM, 64));
        if (SPLIT) TC_MAP_OK(m, mat_map(st, &p.mapA2, st->A8lo, (int64_t)B8 * 16, KP, BM, 64));
        p.dY = st->dY[0]; p.dYlo = st->dYlo[0]; p.part = st->wf_part;
        p.B = B; p.F = m->F; p.P = P; p.Pp = Pp; p.KA = st->KA; p.nblk = st->nblk; p.Q16 = st->Q16;
        p.nsplit = std::min(W0_SPLIT_MAX, (B + 7) / 8);
        static PerDeviceOnce attr_once;
        bool& attr_done = attr_once();
        if (!attr_done) {
          CFFM_CUDA_OK(m, cudaFuncSetAttribute(k_wgrad0_fact<SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, w0_smem(SPLIT)));
          attr_done = true;
        }
        const int units = (st->Q16 / W0_QS) * p.nsplit;
        k_wgrad0_fact<SPLIT><<<units < 148 ? units : 148, W0_THREADS, w0_smem(SPLIT), ws>>>(p);
        const int64_t tot = 4ll * P * P;
        int rb = (int)((tot + 255) / 256); if (rb > 148 * 8) rb = 148 * 8;
        k_wfact_reduce<<<rb, 256, 0, ws>>>(st->wf_part, m->pair_i, m->pair_j, P, st->KA, st->Q16, p.nsplit, g + m->lay.conv_w[0]);
        m->launches += 3;
        CFFM_CUDA_OK(m, cudaGetLastError());
      } else if (l == 0) {
        ConvWgradTC<ACT, true, SPLIT> p;
        p.g = gm; p.chunks_total = chunks_total; p.chunks_per_split = cps; p.n_split = n_split; p.partial = st->wg_partial;
        p.rows = m->outer_rows; p.pair_i = m->pair_i; p.pair_j = m->pair_j;
        p.sk.nparts = st->split ? 3 : 1;
        memset(&p.mapA, 0, sizeof(p.mapA)); memset(&p.mapA2, 0, sizeof(p.mapA2)); memset(&p.mapB2, 0, sizeof(p.mapB2));
        TC_MAP_OK(m, mat_map(st, &p.mapB, st->dY[0], rows, Pp, 64, 64));
        if (st->split) TC_MAP_OK(m, mat_map(st, &p.mapB2, st->dYlo[0], rows, Pp, 64, 64));
        TCTRY(launch_tc(m, p, tiles * n_split, ws));
      } else {
        ConvWgradTC<ACT, false, SPLIT> p;
        p.g = gm; p.chunks_total = chunks_total; p.chunks_per_split = cps; p.n_split = n_split; p.partial = st->wg_partial;
        p.rows = nullptr; p.pair_i = nullptr; p.pair_j = nullptr;
        p.sk.nparts = st->split ? 3 : 1;
        memset(&p.mapA2, 0, sizeof(p.mapA2)); memset(&p.mapB2, 0, sizeof(p.mapB2));
        TC_MAP_OK(m, im2col_map(st, &p.mapA, st->X[l], B, gm.Hin, Pp, 64));
        TC_MAP_OK(m, mat_map(st, &p.mapB, st->dY[l], rows, Pp, 64, 64));
        if (st->split) {
          TC_MAP_OK(m, im2col_map(st, &p.mapA2, st->Xlo[l], B, gm.Hin, Pp, 64));
          TC_MAP_OK(m, mat_map(st, &p.mapB2, st->dYlo[l], rows, Pp, 64, 64));
        }
        TCTRY(launch_tc(m, p, tiles * n_split, ws));
      }
      if (!(l == 0 && st->wf_part && B >= st->fact_min_batch)) {
        const int64_t total = 4ll * P * P;
        int blocks = (int)((total + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
        k_wgrad_reduce<<<blocks, 256, 0, ws>>>(st->wg_partial, n_split, P, Pp, l == 0 ? 1 : 0, g + m->lay.conv_w[l]);
        m->launches++;
      }
    }
    {  // data gradient
      const std::string tag = "conv_dgrad_l" + std::to_string(l);
      CFFM_PROF(m, tag.c_str(), s);
      Geom gd = gm; gd.tiles_n = 4 * Pp / gm.BN;
      if (l == 0 && st->dfact && B >= st->fact_min_batch) {
        TCTRY(dgrad0_fact_launch<SPLIT>(m, st, B, s));
      } else if (l == 0) {
        Conv0DgradTC p;
        p.g = gd; p.rows = m->outer_rows; p.gout = m->gout; p.v_head = m->v_head; p.pair_i = m->pair_i; p.pair_j = m->pair_j;
        p.g_rows = m->g_outer_rows;
        p.sk.nparts = st->split ? 3 : 1;
        memset(&p.mapA2, 0, sizeof(p.mapA2)); memset(&p.mapB2, 0, sizeof(p.mapB2));
        TC_MAP_OK(m, mat_map(st, &p.mapA, st->dY[0], rows, Pp, BM, 64));
        TC_MAP_OK(m, mat_map(st, &p.mapB, st->Wd[0], 4 * Pp, Pp, gd.BN, 64));
        if (st->split) {
          TC_MAP_OK(m, mat_map(st, &p.mapA2, st->dYlo[0], rows, Pp, BM, 64));
          TC_MAP_OK(m, mat_map(st, &p.mapB2, st->Wdlo[0], 4 * Pp, Pp, gd.BN, 64));
        }
        TCTRY(launch_tc(m, p, B, s));
      } else {
        ConvDgradTC<ACT, SPLIT> p;
        p.g = gd; p.X = ACT == CFFM_ACT_GELU ? st->Dphi[l] : st->X[l]; p.dYprev = st->dY[l - 1]; p.gout = m->gout; p.v_head = m->v_head; p.sp_off = lvl_off[l];
        p.sk.nparts = st->split ? 3 : 1; p.dYprev_lo = st->dYlo[l - 1];
        memset(&p.mapA2, 0, sizeof(p.mapA2)); memset(&p.mapB2, 0, sizeof(p.mapB2));
        TC_MAP_OK(m, mat_map(st, &p.mapA, st->dY[l], rows, Pp, BM, 64));
        TC_MAP_OK(m, mat_map(st, &p.mapB, st->Wd[l], 4 * Pp, Pp, gd.BN, 64));
        if (st->split) {
          TC_MAP_OK(m, mat_map(st, &p.mapA2, st->dYlo[l], rows, Pp, BM, 64));
          TC_MAP_OK(m, mat_map(st, &p.mapB2, st->Wdlo[l], 4 * Pp, Pp, gd.BN, 64));
        }
        TCTRY(launch_tc(m, p, ((gd.M + BM - 1) / BM) * gd.tiles_n, s));
      }
    }
  }
  // (on the side stream after layer 0's weight gradient they cost 14 us of a 190 us Frappe step; so did the filters'
  // operand copies beside the gather: a branch is worth it from a few tens of microseconds of work)
  if (colsum_late) colsum_all_layers(s);
  return side_join(m, wside, s, 1);
}

int tc_conv_backward(Model* m, int B, cudaStream_t s) {
  int r = CFFM_OK;
  const bool split = reinterpret_cast<TCState*>(m->tcs)->split;
  CFFM_DISPATCH_PHI(m->cfg.activation, r = split ? conv_backward_act<ACT, true>(m, B, s) : conv_backward_act<ACT, false>(m, B, s));
  return r;
}

// debug access for the parity tests: X_{l+1} = phi(Y_l) converted to fp32 [B,Ho,Ho,P]
__global__ void k_unpad_bf16(const bf16* __restrict__ X, const bf16* __restrict__ Xlo, int64_t rows, int P, int Pp, float* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * P) return;
  const int64_t r = e / P; const int c = (int)(e - r * P);
  out[e] = __bfloat162float(X[r * Pp + c]) + (Xlo ? __bfloat162float(Xlo[r * Pp + c]) : 0.f);
}
int tc_debug_fetch(Model* m, bool grad, int l, float* dev_out, int64_t rows) {
  TCState* st = reinterpret_cast<TCState*>(m->tcs);
  if (!st) return CFFM_ERR_INVALID;
  const bf16* src = grad ? st->dY[l] : st->X[l + 1];
  const bf16* src_lo = grad ? st->dYlo[l] : st->Xlo[l + 1];
  if (!src) return CFFM_ERR_INVALID;
  k_unpad_bf16<<<ceil_div(rows * m->P, 256), 256>>>(src, src_lo, rows, m->P, st->Pp, dev_out);
  return cudaDeviceSynchronize() == cudaSuccess ? CFFM_OK : CFFM_ERR_CUDA;
}

}  // namespace cffm
