// Layer-0 forward in factorised form (included by conv_tc.cu).
//
// The interaction cube is rank one per pair (CFFM.py:355-367: X0[a,c,p] = o_i[a] o_j[c]), so the 2x2/s2
// convolution over it (CFFM.py:420-438) is a bilinear form per output channel q:
//
//   Y0[b,h,w,q] = a_{b,h}^T  Wq  a_{b,w},     a_{b,h}[2i+dh] = o_i[b, 2h+dh]   (2F numbers)
//                                             Wq[2i+dh, 2j+dw] = W0[dh,dw,(i,j),q] for i < j, else 0
//
//   step 1   Z[(b,h), n] = sum_k A[(b,h), k] Wq[k, n]          tcgen05 SS MMA, M=128 (8 samples x 16 h), N=K=KA
//   step 2   D[(b,h), (b',w)] = sum_n Z[(b,h), n] A[(b',w), n]  tcgen05 TS MMA (Z as bf16 in TMEM), N=128, K=KA
//            Y0[b,h,w,q] = D[(b,h), (b,w)]  (the 16 columns of the row's own sample)
//
// Step 1 is 2*KA^2 flops per (row, q) instead of 2*16*4P for the sixteen w of that row in the direct form
// (7.5x fewer at F=39); step 2 wastes 7/8 of its columns (a tile holds 8 samples) and still costs less than
// step 1.  The A tile (128 x KA bf16, K-major SWIZZLE_128B) is both the A operand of step 1 and the B operand
// of step 2.
//
//   warp 0      TMA: one weight slab Wq^T (KA x KA, two 64-wide blocks when KA > 64) per q into a ring
//   warp 1      issues step 1, warp 14 issues step 2 (independent instruction streams, coupled by mbarriers only)
//   warps 2..5  converters: Z fp32 (TMEM) -> bf16, written over the same TMEM columns (A operand of step 2);
//               three Z buffers cover the latency of the step 1 -> convert -> step 2 ring
//   warps 6..13 epilogue: own columns of D (half of the sample's 16 each) + bias, activation, pooling sums; 16
//               channels are collected in registers so that every store is a full 32-byte sector of X1[b,h,w,:];
//               they also build the A tile
// (included inside namespace cffm::tc of conv_tc.cu, after pack2 / phi_f)
#pragma once

constexpr int F0_NST = 6;     // slab ring depth: a stage is busy for (TMA latency + its MMAs), so hiding ~1.8 us of latency behind 0.3 us of MMAs takes 6
constexpr int F0_KA_MAX = 80;
constexpr int F0_SLAB_BYTES = 2 * F0_KA_MAX * 128;
constexpr int F0_THREADS = 480;
constexpr int F0_NZ = 3;                  // Z buffers (step-1 accumulator, overwritten in place by its bf16 copy)
constexpr int F0_Z = 0, F0_Z_STRIDE = 80, F0_D2 = 256, F0_D2_STRIDE = 128;
constexpr int F0_BIAS_MAX = 1280;

struct F0Ctl {
  uint64_t full_b[F0_NST], empty_b[F0_NST];
  uint64_t z_full[F0_NZ], z_empty[F0_NZ], zb_full[F0_NZ], d2_full[2], d2_empty[2];
  uint64_t a_ready;
  uint32_t tmem_base, pad;
};
constexpr int F0_STAGE_BYTES = 8 * 32 * 32;   // per epilogue warp: [8 w][32 rows][16 channels] bf16
constexpr int F0_SMEM = 1024 + 2 * A_STAGE_BYTES + F0_NST * F0_SLAB_BYTES + 256 + F0_BIAS_MAX * 4 + 2 * BM * 4 + 8 * F0_STAGE_BYTES + 128;
static_assert(sizeof(F0Ctl) <= 256, "control block");
static_assert(F0_SMEM <= 227 * 1024, "factorised forward exceeds the shared memory of an SM");
// Split mode (CFFM_PREC_BF16X3): the A tile and every weight slab exist as hi and lo halves, Z is split into hi and
// lo when it is converted (the two bf16 copies fill exactly the columns of the fp32 Z they are made from) and both
// steps issue three MMAs per K step: hi*hi + lo*hi + hi*lo.  Shared memory: A tile 2 x 32 KB, slab ring of two stages
// of (hi, lo) pairs (a third changed nothing).  The epilogue collects 8 channels instead of 16 and writes them
// INTERLEAVED -- per 8 channels 16 bytes of hi, then 16 bytes of lo -- into a scratch tensor X1i, so that a (row, w) of
// a group is one full 32-byte sector; k_deinterleave_x1 then makes the separate hi / lo tensors the next layer's TMA
// boxes want.  With 16-byte rows straight into the hi / lo tensors the stores cost 7 of the kernel's 12.5 ms (ablation
// without stores: 5.3 ms) and wrote 11.7 GB for 6.4 GB of payload; reading the interleaved tensor with 16-byte TMA boxes
// instead doubled the time of both layer-1 kernels.  The extra pass moves 12.8 GB (~2.2 ms).
constexpr int F0S_NST = 2;
constexpr int F0S_SMEM = 1024 + 4 * A_STAGE_BYTES + F0S_NST * 2 * F0_SLAB_BYTES + 256 + F0_BIAS_MAX * 4 + 2 * BM * 4 + 8 * F0_STAGE_BYTES + 128;
static_assert(F0S_SMEM <= 227 * 1024, "split-mode factorised forward exceeds the shared memory of an SM");

struct Fwd0FactParams {
  CUtensorMap mapW;     // Wf0 viewed as [Q16*KA rows][nblk*64 cols] bf16, box (64, KA)
  CUtensorMap mapX;     // X1 as (q: Pp, row = b*16+h: B*16, w: 16), dense box (16, 32, 8) for the epilogue's TMA stores
  CUtensorMap mapW2, mapX2;   // split mode: the lo halves (mapX / mapX2 then have box (8, 32, 8))
  bf16* Xout_lo;
  int tma_store;        // 0: the tensor map could not be encoded, lanes store their sectors themselves
  const float* rows;    // outer rows [B][F][32]
  const float* bias;
  bf16* Xout;           // X1 [B][16][16][Pp]
  float* t1; int t1_dim, sp_off;
  int B, F, P, Pp, KA, nblk, Q16;
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&r)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// Wf0[q][n = 2j+dw][k = 2i+dh] = W0[dh][dw][p(i,j)][q]; entries with i >= j stay zero (set once at allocation)
__global__ void k_prep_w0_fact(const float* __restrict__ W0, const int* __restrict__ pair_i, const int* __restrict__ pair_j, int P,
                               int KA, int KP, bf16* __restrict__ out, bf16* __restrict__ out_lo) {
  const int64_t total = 4ll * P * P;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(e % P);
    const int64_t r = e / P;
    const int p = (int)(r % P), tap = (int)(r / P);
    const int dh = tap >> 1, dw = tap & 1;
    const int k = 2 * pair_i[p] + dh, n = 2 * pair_j[p] + dw;
    const bf16 v = __float2bfloat16(W0[e]);
    out[((int64_t)q * KA + n) * KP + k] = v;
    if (out_lo) out_lo[((int64_t)q * KA + n) * KP + k] = __float2bfloat16(W0[e] - __bfloat162float(v));
  }
}

// X1i [pixel][Pp / 8][hi8 | lo8] -> X1 hi [pixel][Pp], X1 lo [pixel][Pp]; a thread moves one 32-byte unit, neighbouring
// threads write neighbouring 16-byte pieces of both outputs (full sectors on both sides)
__global__ void k_deinterleave_x1(const uint4* __restrict__ Xi, int64_t units, uint4* __restrict__ Xhi, uint4* __restrict__ Xlo) {
  for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (int64_t)gridDim.x * blockDim.x) {
    const uint4 h = __ldcs(Xi + 2 * u), l = __ldcs(Xi + 2 * u + 1);   // read once: streaming loads
    Xhi[u] = h; Xlo[u] = l;
  }
}

template <int ACT, bool SPLIT>
__global__ void __launch_bounds__(F0_THREADS, 1) k_fwd0_fact(const __grid_constant__ Fwd0FactParams prm) {
  constexpr int NST = SPLIT ? F0S_NST : F0_NST;                         // slab ring depth
  constexpr int SLAB_STAGE = (SPLIT ? 2 : 1) * F0_SLAB_BYTES;           // one stage: the slab (hi) [+ its lo half]
  constexpr int AT_BYTES = (SPLIT ? 4 : 2) * A_STAGE_BYTES;             // A tile: [hi: nblk blocks][lo: nblk blocks]
  constexpr int QG = SPLIT ? 8 : 16;                                    // channels an epilogue thread collects per store
  constexpr int WARP_STAGE = F0_STAGE_BYTES;                            // staging tile of an epilogue warp
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sAt = smem;                                   // [nblk][128 rows][128 B] (split: hi at 0, lo at 2 * A_STAGE_BYTES)
  uint8_t* sW = sAt + AT_BYTES;                          // ring of weight slabs
  F0Ctl* ctl = reinterpret_cast<F0Ctl*>(sW + NST * SLAB_STAGE);
  float* sbias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctl) + 256);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int KA = prm.KA, nblk = prm.nblk, Q = prm.Q16;
  const int n_tiles = (prm.B + 7) >> 3;
  const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const uint32_t slab_bytes = (uint32_t)(nblk * KA * 128);

  if (warp == 0 && lane == 0) { prefetch_tmap(&prm.mapW); if (SPLIT) prefetch_tmap(&prm.mapW2); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&ctl->full_b[s], 1); mbar_init(&ctl->empty_b[s], 1); }
    for (int b = 0; b < F0_NZ; ++b) { mbar_init(&ctl->z_full[b], 1); mbar_init(&ctl->z_empty[b], 1); mbar_init(&ctl->zb_full[b], 4); }
    for (int b = 0; b < 2; ++b) { mbar_init(&ctl->d2_full[b], 1); mbar_init(&ctl->d2_empty[b], 8); }
    mbar_init(&ctl->a_ready, 8);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&ctl->tmem_base, 512);
  for (int e = threadIdx.x; e < Q; e += F0_THREADS) sbias[e] = e < prm.P ? __ldg(prm.bias + e) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA: weight slabs
    if (lane == 0) {
      uint32_t n = 0;
      for (int t = 0; t < my_tiles; ++t)
        for (int q = 0; q < Q; ++q, ++n) {
          const int s = n % NST; const uint32_t ph = (n / NST) & 1;
          mbar_wait(&ctl->empty_b[s], ph ^ 1);
          mbar_arrive_expect_tx(&ctl->full_b[s], (SPLIT ? 2u : 1u) * slab_bytes);
          for (int blk = 0; blk < nblk; ++blk) {
            tma_load_2d(sW + s * SLAB_STAGE + blk * KA * 128, &prm.mapW, &ctl->full_b[s], blk * 64, q * KA);
            if (SPLIT) tma_load_2d(sW + s * SLAB_STAGE + F0_SLAB_BYTES + blk * KA * 128, &prm.mapW2, &ctl->full_b[s], blk * 64, q * KA);
          }
        }
    }
  } else if (warp == 1 || warp == 14) {
    // ------------------------------------------------------------------ MMA issuers: warp 1 step 1, warp 14 step 2
    // Two independent instruction streams coupled by mbarriers only.  All lanes run the loops (uniform control
    // flow); the MMAs and commits of one q sit inside ONE elect.sync region, which ptxas turns into
    // straight-line UTCHMMA issue (a lane == 0 test costs a 16-instruction per-thread loop around every
    // tcgen05 instruction).  The D buffer = q & 1 is compile-time (Q % 4 == 0); the slab ring slot and the Z
    // buffer index are running counters.
    const uint32_t at_addr = smem_u32(sAt);
    const int ksteps = KA / UMMA_K;
    uint64_t adesc[F0_KA_MAX / UMMA_K];
#pragma unroll
    for (int k = 0; k < F0_KA_MAX / UMMA_K; ++k)
      adesc[k] = umma_desc_k_sw128(at_addr + (uint32_t)((k >> 2) * A_STAGE_BYTES)) + (uint64_t)((k & 3) * 2);
    constexpr uint64_t A_LO = (uint64_t)((2 * A_STAGE_BYTES) >> 4);     // descriptor offset of the lo half of the A tile
    int zb = 0; uint32_t zph = 0;     // Z buffer of the current q and its phase
    if (warp == 1) {
      const uint32_t idesc1 = umma_idesc_bf16(BM, KA);
      uint64_t wk[F0_KA_MAX / UMMA_K];
#pragma unroll
      for (int k = 0; k < F0_KA_MAX / UMMA_K; ++k) wk[k] = (uint64_t)((((k >> 2) * KA * 128) >> 4) + (k & 3) * 2);
      constexpr uint64_t W_LO = (uint64_t)(F0_SLAB_BYTES >> 4);
      int ws = 0; uint32_t wph = 0;     // ring slot of the current q and its phase
      const uint64_t wbase = umma_desc_k_sw128(smem_u32(sW));
      for (int t = 0; t < my_tiles; ++t) {
        mbar_wait(&ctl->a_ready, (uint32_t)(t & 1));
        tc_fence_after();
        for (int q = 0; q < Q; q += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            mbar_wait(&ctl->full_b[ws], wph);
            mbar_wait(&ctl->z_empty[zb], zph ^ 1);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t d1 = tmem_base + (uint32_t)(F0_Z + zb * F0_Z_STRIDE);
              const uint64_t wd = wbase + (uint64_t)(ws * (SLAB_STAGE >> 4));
#pragma unroll
              for (int k = 0; k < F0_KA_MAX / UMMA_K; ++k)
                if (k < ksteps) umma_bf16(d1, adesc[k], wd + wk[k], idesc1, k != 0);
              if constexpr (SPLIT) {
#pragma unroll
                for (int k = 0; k < F0_KA_MAX / UMMA_K; ++k)
                  if (k < ksteps) umma_bf16(d1, adesc[k] + A_LO, wd + wk[k], idesc1, true);            // lo * hi
#pragma unroll
                for (int k = 0; k < F0_KA_MAX / UMMA_K; ++k)
                  if (k < ksteps) umma_bf16(d1, adesc[k], wd + W_LO + wk[k], idesc1, true);            // hi * lo
              }
              umma_commit(&ctl->empty_b[ws]);
              umma_commit(&ctl->z_full[zb]);
            }
            __syncwarp();
            if (++zb == F0_NZ) { zb = 0; zph ^= 1; }
            if (++ws == NST) { ws = 0; wph ^= 1; }
          }
        }
      }
    } else {
      const uint32_t idesc2 = umma_idesc_bf16(BM, 128);
      const uint32_t z_lo = (uint32_t)(KA / 2);          // columns of the lo half of the converted Z
      for (int t = 0; t < my_tiles; ++t) {
        mbar_wait(&ctl->a_ready, (uint32_t)(t & 1));
        tc_fence_after();
        for (int q = 0; q < Q; q += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            mbar_wait(&ctl->zb_full[zb], zph);
            mbar_wait(&ctl->d2_empty[u & 1], (uint32_t)((u >> 1) ^ 1));
            tc_fence_after();
            if (elect_one()) {
              const uint32_t d2 = tmem_base + (uint32_t)(F0_D2 + (u & 1) * F0_D2_STRIDE);
              const uint32_t za = tmem_base + (uint32_t)(F0_Z + zb * F0_Z_STRIDE);
#pragma unroll
              for (int k = 0; k < F0_KA_MAX / UMMA_K; ++k)
                if (k < ksteps) umma_bf16_ts(d2, za + (uint32_t)(k * 8), adesc[k], idesc2, k != 0);
              if constexpr (SPLIT) {
#pragma unroll
                for (int k = 0; k < F0_KA_MAX / UMMA_K; ++k)
                  if (k < ksteps) umma_bf16_ts(d2, za + z_lo + (uint32_t)(k * 8), adesc[k], idesc2, true);     // Z lo * A hi
#pragma unroll
                for (int k = 0; k < F0_KA_MAX / UMMA_K; ++k)
                  if (k < ksteps) umma_bf16_ts(d2, za + (uint32_t)(k * 8), adesc[k] + A_LO, idesc2, true);     // Z hi * A lo
              }
              umma_commit(&ctl->z_empty[zb]);        // the Z buffer (fp32 and its bf16 overlay) is free again
              umma_commit(&ctl->d2_full[u & 1]);
            }
            __syncwarp();
            if (++zb == F0_NZ) { zb = 0; zph ^= 1; }
          }
        }
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ converters: Z fp32 -> bf16, in place
    // (the bf16 A operand of step 2 overwrites the first KA/2 columns of the fp32 Z it was made from; in split mode
    // the lo half takes the other KA/2 columns)
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    int zb = 0; uint32_t zph = 0;
    for (int t = 0; t < my_tiles; ++t)
      for (int q = 0; q < Q; ++q) {
        const uint32_t zaddr = tmem_base + lane_off + (uint32_t)(F0_Z + zb * F0_Z_STRIDE);
        mbar_wait(&ctl->z_full[zb], zph);
        tc_fence_after();
        float v[F0_KA_MAX / 16][16];
#pragma unroll
        for (int c = 0; c < F0_KA_MAX / 16; ++c)
          if (c * 16 < KA) tmem_ld16(zaddr + (uint32_t)(c * 16), v[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < F0_KA_MAX / 16; ++c)
          if (c * 16 < KA) {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) pk[j] = pack2(v[c][2 * j], v[c][2 * j + 1]);
            tmem_st8(zaddr + (uint32_t)(c * 8), pk);
            if constexpr (SPLIT) {
#pragma unroll
              for (int j = 0; j < 8; ++j) pk[j] = pack2_lo(v[c][2 * j], v[c][2 * j + 1]);
              tmem_st8(zaddr + (uint32_t)(KA / 2 + c * 8), pk);
            }
          }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->zb_full[zb]);
        if (++zb == F0_NZ) { zb = 0; zph ^= 1; }
      }
  } else {
    // ------------------------------------------------------------------ epilogue (+ A tile builder), 8 warps
    // Two warps per TMEM lane quarter: group 0 takes w = 0..7 of every channel, group 1 takes w = 8..15, so a
    // thread keeps 8 x 16 results (64 registers) and still stores whole 32-byte sectors.
    const int ew = warp - 6, grp = ew >> 2;
    const int qd = warp & 3;
    const int r = qd * 32 + lane;                 // tile row: sample r>>4 of the tile, h = r&15
    const int h = r & 15;
    const bool hi = (lane & 16) != 0;             // second sample of this warp: its block starts 16 columns later
    const uint32_t d2_addr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(F0_D2 + qd * 32 + grp * 8);
    float* xch = sbias + F0_BIAS_MAX;             // pooling sums of group 1, [2][128]
    // staging tile of this warp for the TMA store (128-byte aligned, after xch); split: hi tile, then lo tile
    uint8_t* stage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(xch + 2 * BM) + 127) & ~(uintptr_t)127) + ew * WARP_STAGE;
    auto build_a = [&](int tile) {
      const int b = tile * 8 + (r >> 4);
      const float* src = prm.rows + ((int64_t)(b < prm.B ? b : 0) * prm.F) * 32 + 2 * h;
      for (int c = grp; c < KA / 8; c += 2) {     // 16-byte chunk c = fields 4c .. 4c+3
        uint32_t wv[4], wl[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = 4 * c + e;
          float2 o = make_float2(0.f, 0.f);
          if (i < prm.F && b < prm.B) o = __ldg(reinterpret_cast<const float2*>(src + i * 32));
          wv[e] = pack2(o.x, o.y);
          if (SPLIT) wl[e] = pack2_lo(o.x, o.y);
        }
        *reinterpret_cast<uint4*>(sAt + (c >> 3) * A_STAGE_BYTES + sw128_offset(r, c & 7)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        if (SPLIT)
          *reinterpret_cast<uint4*>(sAt + (2 + (c >> 3)) * A_STAGE_BYTES + sw128_offset(r, c & 7)) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->a_ready);
    };
    if (my_tiles > 0) build_a((int)blockIdx.x);
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = (int)blockIdx.x + t * (int)gridDim.x;
      const int b = tile * 8 + (r >> 4);
      float rowsum = 0.f;
      for (int q0 = 0; q0 < Q; q0 += QG) {
        uint32_t acc[8][QG / 2];
        uint32_t accl[SPLIT ? 8 : 1][QG / 2];
        float prev[8];
#pragma unroll
        for (int qq = 0; qq < QG; ++qq) {         // q & 1 == qq & 1, (q >> 1) & 1 == (qq >> 1) & 1
          mbar_wait(&ctl->d2_full[qq & 1], (uint32_t)((qq >> 1) & 1));
          tc_fence_after();
          float lo[8], up[8];
          tmem_ld8(d2_addr + (uint32_t)((qq & 1) * F0_D2_STRIDE), lo);
          tmem_ld8(d2_addr + (uint32_t)((qq & 1) * F0_D2_STRIDE + 16), up);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl->d2_empty[qq & 1]);
          const float bq = sbias[q0 + qq];
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            const float x = phi_f<ACT>((hi ? up[w] : lo[w]) + bq);
            rowsum += x;
            if (qq & 1) {
              acc[w][qq >> 1] = pack2(prev[w], x);
              if constexpr (SPLIT) accl[w][qq >> 1] = pack2_lo(prev[w], x);
            } else prev[w] = x;
          }
        }
        if (prm.tma_store) {
          // 32 rows x 8 w x QG channels of this warp -> staging tile -> one TMA store (rows beyond the batch are
          // clipped by the tensor map); the copy engine does the scattered 32-byte writes, not the LSU
          if (lane == 0) tma_store_wait_read();     // the previous store has finished reading the tile
          __syncwarp();
          if constexpr (SPLIT) {
            // interleaved scratch X1i: [8 hi | 8 lo] of this 8-channel group = one 32-byte sector per (row, w)
#pragma unroll
            for (int w = 0; w < 8; ++w) {
              uint4* d = reinterpret_cast<uint4*>(stage + (w * 32 + lane) * 32);
              d[0] = make_uint4(acc[w][0], acc[w][1], acc[w][2], acc[w][3]);
              d[1] = make_uint4(accl[w][0], accl[w][1], accl[w][2], accl[w][3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tma_store_3d(&prm.mapX, stage, 2 * q0, tile * BM + qd * 32, grp * 8);
          } else {
#pragma unroll
            for (int w = 0; w < 8; ++w) {
              uint4* d = reinterpret_cast<uint4*>(stage + (w * 32 + lane) * 32);
              d[0] = make_uint4(acc[w][0], acc[w][1], acc[w][2], acc[w][3]);
              d[1] = make_uint4(acc[w][QG / 2 - 4], acc[w][QG / 2 - 3], acc[w][QG / 2 - 2], acc[w][QG / 2 - 1]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tma_store_3d(&prm.mapX, stage, q0, tile * BM + qd * 32, grp * 8);
          }
        } else if (b < prm.B) {
          const int64_t off = (((int64_t)b * 16 + h) * 16 + grp * 8) * prm.Pp + q0;
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            if constexpr (SPLIT) {   // Xout = the interleaved scratch X1i
              const uint32_t r8[8] = {acc[w][0], acc[w][1], acc[w][2], acc[w][3], accl[w][0], accl[w][1], accl[w][2], accl[w][3]};
              st_global_256(prm.Xout + 2 * (off + (int64_t)w * prm.Pp), r8);   // (off counts plain elements; q0 is a multiple of 8)
            } else {
              uint32_t r8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) r8[e] = acc[w][e % (QG / 2)];
              st_global_256(prm.Xout + off + (int64_t)w * prm.Pp, r8);   // one full 32-byte sector per lane
            }
          }
        }
      }
      // pooling sum of the row: the two groups meet in shared memory (double-buffered on the tile parity)
      float* slot = xch + (t & 1) * BM + r;
      if (grp == 1) *slot = rowsum;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (grp == 0 && b < prm.B) prm.t1[(int64_t)b * prm.t1_dim + prm.sp_off + h] = rowsum + *slot;
      // every MMA of this tile has retired (the last step-2 result was read above): the A tile may be rebuilt
      if (t + 1 < my_tiles) build_a(tile + (int)gridDim.x);
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}
