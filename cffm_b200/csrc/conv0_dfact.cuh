// Layer-0 data gradient in factorised form (included inside namespace cffm::tc of conv_tc.cu, after conv0_fact.cuh).
//
// With a_{b,x}[2i+d] = o_i[b, 2x+d] and Y0[b,h,w,q] = a_{b,h}^T Wq a_{b,w} (conv0_fact.cuh), the gradient of the
// outer rows is
//   d o_i[b, 2x+d] = sum_{w,q} dY0[b,x,w,q] (Wq   a_{b,w})[2i+d]        term 1 (x is the h of the output position)
//                  + sum_{h,q} dY0[b,h,x,q] (Wq^T a_{b,h})[2i+d]        term 2 (x is the w)
//                  + pooling term (sum_pooling[0] sees the cube itself)
// Both terms are evaluated "dY first": for a tile of 8 samples (128 rows = (sample, x)) and one channel q
//   E1[(b,h), n] = sum_{(b',w)} Dq[(b,h),(b',w)] A[(b',w), n]       Dq = block-diagonal: Dq[(b,h),(b,w)] = dY0[b,h,w,q]
//   E2[(b,w), k] = sum_{(b',h)} Dq[(b',h),(b,w)] A[(b',h), k]       the same buffer read as an MN-major A operand
//   D[(b,x), c] += sum_n E1[(b,x), n] Wq[c, n] + sum_k E2[(b,x), k] Wq[k, c]       one accumulator for the whole tile
// so the 3 GB cube gradient never exists, the result needs no cross-lane reduction, and the tensor work is
// 2 x (8 + 5) small MMAs per (tile, q): 3x fewer tensor cycles than the direct form (whose epilogue was the limit).
//
// Both terms read ONE filter slab Wf0[q] (rows n, k contiguous): term 2 as a K-major B operand, term 1 as an MN-major
// one (K = n runs over the slab's rows) -- no transposed copy of the filters, half the slab traffic.
// Split mode (bf16x3): dY0, the A tile and the slab come as hi + lo, every product is hi*hi + lo*hi + hi*lo into the
// same fp32 accumulator (three MMAs), E is split chunk by chunk (16 columns: hi words | lo words) in its own TMEM columns.  The operands
// double in shared memory: one Dq buffer and two slab stages instead of two and three.
//
//   warp 0       TMA: the filter slab of q (split mode: hi and lo)
//   warp 1       issues the E MMAs (SS; A = Dq K-major / MN-major, B = the A tile as an MN-major operand)
//   warp 2       issues the accumulating MMAs (TS: E as bf16 in TMEM)
//   warp 3       TMEM allocation
//   warps 4..7   convert E1 fp32 -> bf16 in place; at the end of a tile: D + pooling term -> g_rows
//   warps 8..11  convert E2
//   warps 12..19 two builder groups: dY0 (8 channels per load) -> diagonal blocks of Dq; group 0 also builds the A tile
//                (the groups take alternate 8-channel sets; split mode: every set, group 0 the hi blocks, group 1 the lo)
// Measured (B = 8192, F = 39): 4.3 ms in bf16; 20 ms in split mode against 12 ms of MMA work (E MMAs are bound by
// their shared-memory operands, ~68 clk each; tensor pipe 48 % active).  With one Dq buffer the chain
// E MMAs done -> builders store + fence -> E MMAs of the next channel is serial, and shared memory (209 of 227 KB) has
// no room for a second one.  Tried without gain: alternating 4-channel sets with 8-byte loads (26 ms), an L2 prefetch
// of the next set, the next set prefetched into registers (setmaxnreg 40 / 56 / 160 for MMA / converter / builder
// warps: 21 ms).  Switching the builders' loads off looked 7 ms faster, but that run multiplied zeros (the step runs
// under the power cap, 1.63 of 1.97 GHz) -- the loads are not what the kernel waits for.
#pragma once

constexpr int G0_THREADS = 640;
// ring depths: E buffers in tensor memory, Dq buffers, filter-slab stages.  A slab stage is busy for (TMA latency + its
// MMAs) ~ 1.5 + 0.55 us, so two stages fed the tensor pipe one channel per ~1 us; three stages (and two Dq buffers,
// which only have to cover the builders' store + fence) fit the same shared memory.
constexpr int G0_NE = 4, G0_ND_MAX = 2, G0_NW_MAX = 3;
__host__ __device__ constexpr int g0_nd(bool split) { return split ? 1 : 2; }
__host__ __device__ constexpr int g0_nw(bool split) { return split ? 2 : 3; }
constexpr int G0_E = 0, G0_E_STRIDE = 80, G0_D = 320;
constexpr int G0_DQ_BYTES = 2 * A_STAGE_BYTES;            // 128 rows x 128 columns bf16

struct G0Ctl {
  uint64_t w_full[G0_NW_MAX], w_empty[G0_NW_MAX], dq_full[G0_ND_MAX], dq_empty[G0_ND_MAX];
  uint64_t dqw_full[4], dqw_empty[4];   // split mode: the one Dq buffer is handed over per builder warp (two samples)
  uint64_t e_full[G0_NE], e_conv[G0_NE], e_empty[G0_NE];
  uint64_t a_ready, a_free, d_full, d_empty, grp_done[2];
  uint32_t tmem_base, pad;
};
static_assert(sizeof(G0Ctl) <= 512, "control block");
constexpr int g0_smem(bool split) {
  return 1024 + (split ? 2 : 1) * (2 * A_STAGE_BYTES + g0_nd(split) * G0_DQ_BYTES + g0_nw(split) * F0_SLAB_BYTES) + 512;
}
static_assert(g0_smem(false) <= 227 * 1024 && g0_smem(true) <= 227 * 1024, "factorised data gradient exceeds the shared memory of an SM");

struct Dgrad0FactParams {
  CUtensorMap mapW, mapW2;   // Wf0 [q][n][k] (split mode: and its lo half) viewed as [Q16*KA rows][nblk*64 cols], box (64, KA)
  const bf16* dY;            // dY0 [B][16][16][Pp]
  const bf16* dYlo;          // split mode: its lo half
  const float* rows;         // outer rows [B][F][32]
  const float* gout;         // [B]
  const float* v_head;       // pooling weights of level 0: v[0..31]
  const float2* pterm;       // [B][F]: (sum_{j>f} S_j, sum_{i<f} T_i), S_j = sum_c o_j[c], T_i = sum_a v[a] o_i[a]
  float* g_rows;             // [B][F][32]
  float* bpart;              // [tiles][4 warps][Q16]: column sums of dY0 per builder warp (bias gradient of layer 0);
                             // split mode: [tiles][2 parts][4 warps][Q16], the sums of the hi and of the lo tensor
  int B, F, P, Pp, KA, nblk, Q16;
};

// d b_0[q] = sum over tiles and builder warps (fixed order) of the column sums collected by k_dgrad0_fact.
// Block = 32 channels x 32 row slices (slice s takes rows s, s + 32, ...), the slices meet in shared memory in slice order.
__global__ void __launch_bounds__(1024) k_dfact_bias_reduce(const float* __restrict__ bpart, int rows, int Q16, int P, float* __restrict__ out) {
  __shared__ float part[32][33];
  const int q = blockIdx.x * 32 + threadIdx.x, sl = threadIdx.y;
  float s = 0.f;
  if (q < P)
    for (int r = sl; r < rows; r += 32) s += bpart[(int64_t)r * Q16 + q];
  part[sl][threadIdx.x] = s;
  __syncthreads();
  if (sl == 0 && q < P) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += part[i][threadIdx.x];
    out[q] = t;
  }
}

// pooling term of the layer-0 data gradient, per sample and field (see Dgrad0FactParams::pterm)
__global__ void k_pool_terms0(const float* __restrict__ rows, const float* __restrict__ v, int B, int F, float2* __restrict__ out) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const float vl = v[lane];
  // lane holds S_f, T_f of fields f = lane and f = lane + 32 (F <= 48)
  float S0 = 0.f, T0 = 0.f, S1 = 0.f, T1 = 0.f;
  for (int f = 0; f < F; ++f) {
    const float x = rows[((int64_t)b * F + f) * 32 + lane];
    const float sv = warp_sum(x), tv = warp_sum(x * vl);
    if ((f & 31) == lane) { if (f < 32) { S0 = sv; T0 = tv; } else { S1 = sv; T1 = tv; } }
  }
  // inclusive scans over the lanes, then exclusive prefix of T and exclusive suffix of S over the fields
  float iS0 = S0, iT0 = T0, iS1 = S1, iT1 = T1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float a0 = __shfl_up_sync(0xffffffffu, iS0, o), a1 = __shfl_up_sync(0xffffffffu, iT0, o);
    const float a2 = __shfl_up_sync(0xffffffffu, iS1, o), a3 = __shfl_up_sync(0xffffffffu, iT1, o);
    if (lane >= o) { iS0 += a0; iT0 += a1; iS1 += a2; iT1 += a3; }
  }
  const float totS0 = __shfl_sync(0xffffffffu, iS0, 31), totT0 = __shfl_sync(0xffffffffu, iT0, 31);
  const float totS1 = __shfl_sync(0xffffffffu, iS1, 31);
  float2* o = out + (int64_t)b * F;
  if (lane < F) o[lane] = make_float2(totS1 + (totS0 - iS0), iT0 - T0);
  if (lane + 32 < F) o[lane + 32] = make_float2(totS1 - iS1, totT0 + (iT1 - T1));
}

template <bool SPLIT>
__global__ void __launch_bounds__(G0_THREADS, 1) k_dgrad0_fact(const __grid_constant__ Dgrad0FactParams prm) {
  constexpr int NP = SPLIT ? 2 : 1;                      // precision parts of an operand: hi (, lo)
  constexpr int G0_ND = g0_nd(SPLIT), G0_NW = g0_nw(SPLIT);
  constexpr int ABYTES = 2 * A_STAGE_BYTES;              // one part of the A tile
  constexpr int DBUF = NP * G0_DQ_BYTES, WSTAGE = NP * F0_SLAB_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sAt = smem;                                   // A tile [2 blocks][128 rows][128 B] (split: hi tile, lo tile)
  uint8_t* sDq = sAt + NP * ABYTES;                      // G0_ND buffers of [2 blocks][128 rows][128 B] (split: hi, lo)
  uint8_t* sW = sDq + G0_ND * DBUF;                      // G0_NW stages of the Wf0 slab (split: hi, lo)
  G0Ctl* ctl = reinterpret_cast<G0Ctl*>(sW + G0_NW * WSTAGE);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int KA = prm.KA, nblk = prm.nblk, Q = prm.Q16;
  const int ksteps = KA / UMMA_K;
  const int n_tiles = (prm.B + 7) >> 3;
  const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const uint32_t slab_bytes = (uint32_t)(nblk * KA * 128);

  if (warp == 0 && lane == 0) { prefetch_tmap(&prm.mapW); if (SPLIT) prefetch_tmap(&prm.mapW2); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < G0_NW; ++s) { mbar_init(&ctl->w_full[s], 1); mbar_init(&ctl->w_empty[s], 1); }
    for (int d = 0; d < G0_ND; ++d) { mbar_init(&ctl->dq_full[d], 4 * NP); mbar_init(&ctl->dq_empty[d], 1); }
    for (int w = 0; w < 4; ++w) { mbar_init(&ctl->dqw_full[w], 2); mbar_init(&ctl->dqw_empty[w], 1); }
    for (int e = 0; e < G0_NE; ++e) { mbar_init(&ctl->e_full[e], 1); mbar_init(&ctl->e_conv[e], 4); mbar_init(&ctl->e_empty[e], 1); }
    mbar_init(&ctl->a_ready, 4); mbar_init(&ctl->a_free, 1); mbar_init(&ctl->d_full, 1); mbar_init(&ctl->d_empty, 4);
    mbar_init(&ctl->grp_done[0], 4); mbar_init(&ctl->grp_done[1], 4);
    fence_barrier_init();
  }
  if (warp == 3) tmem_alloc(&ctl->tmem_base, 512);
  // the off-diagonal part of the Dq buffers is zero for ever
  for (int e = threadIdx.x; e < G0_ND * DBUF / 16; e += G0_THREADS) reinterpret_cast<uint4*>(sDq)[e] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA: filter slabs
    uint32_t n = 0;
    for (int t = 0; t < my_tiles; ++t)
      for (int q = 0; q < Q; ++q, ++n) {
        const int s = n % G0_NW; const uint32_t ph = (n / G0_NW) & 1;
        mbar_wait(&ctl->w_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&ctl->w_full[s], NP * slab_bytes);
          for (int blk = 0; blk < nblk; ++blk) {
            tma_load_2d(sW + s * WSTAGE + blk * KA * 128, &prm.mapW, &ctl->w_full[s], blk * 64, q * KA);
            if (SPLIT) tma_load_2d(sW + s * WSTAGE + F0_SLAB_BYTES + blk * KA * 128, &prm.mapW2, &ctl->w_full[s], blk * 64, q * KA);
          }
        }
        __syncwarp();
      }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ E MMAs: M = 128, N = KA, K = 128
    const uint32_t at_addr = smem_u32(sAt), dq_addr = smem_u32(sDq);
    const uint32_t idesc_k = umma_idesc_bf16(BM, KA, false, true), idesc_mn = umma_idesc_bf16(BM, KA, true, true);
    uint64_t bdesc[8];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) bdesc[ks] = umma_desc_mn_sw128(at_addr + (uint32_t)(ks * 2048), A_STAGE_BYTES, 1024);
    constexpr uint64_t B_LO = (uint64_t)(ABYTES >> 4), D_LO = (uint64_t)(G0_DQ_BYTES >> 4);   // lo parts, in descriptor units
    uint32_t n = 0; int d = 0; uint32_t dph = 0;
    for (int t = 0; t < my_tiles; ++t) {
      mbar_wait(&ctl->a_ready, (uint32_t)(t & 1));
      tc_fence_after();
      for (int q = 0; q < Q; ++q, ++n) {
        const int e1 = (n & 1) * 2; const uint32_t eph = (n >> 1) & 1;
        const uint32_t dq = dq_addr + (uint32_t)(d * DBUF);
        if constexpr (SPLIT) {
          // One Dq buffer: K step ks of both terms reads only what builder warp ks / 2 wrote (the rows and columns of
          // samples 2w, 2w + 1; everything else in those columns and rows is zero for ever), so the buffer changes hands
          // warp by warp -- the builders rewrite the first samples of the next channel while the MMAs of the last
          // samples of this one run, instead of after all 48.
          mbar_wait(&ctl->e_empty[e1], eph ^ 1);
          mbar_wait(&ctl->e_empty[e1 + 1], eph ^ 1);
          const uint32_t et1 = tmem_base + (uint32_t)(G0_E + e1 * G0_E_STRIDE), et2 = et1 + (uint32_t)G0_E_STRIDE;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            mbar_wait(&ctl->dqw_full[w], n & 1);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int ks = 2 * w; ks < 2 * w + 2; ++ks) {
                const uint64_t ak = umma_desc_k_sw128(dq + (uint32_t)((ks >> 2) * A_STAGE_BYTES)) + (uint64_t)((ks & 3) * 2);
                const uint64_t am = umma_desc_mn_sw128(dq + (uint32_t)(ks * 2048), A_STAGE_BYTES, 1024);
                umma_bf16(et1, ak, bdesc[ks], idesc_k, ks != 0);
                umma_bf16(et1, ak + D_LO, bdesc[ks], idesc_k, true);
                umma_bf16(et1, ak, bdesc[ks] + B_LO, idesc_k, true);
                umma_bf16(et2, am, bdesc[ks], idesc_mn, ks != 0);
                umma_bf16(et2, am + D_LO, bdesc[ks], idesc_mn, true);
                umma_bf16(et2, am, bdesc[ks] + B_LO, idesc_mn, true);
              }
              umma_commit(&ctl->dqw_empty[w]);
              if (w == 3) {
                umma_commit(&ctl->e_full[e1]);
                umma_commit(&ctl->e_full[e1 + 1]);
                if (q == Q - 1) umma_commit(&ctl->a_free);
              }
            }
            __syncwarp();
          }
          continue;
        }
        mbar_wait(&ctl->dq_full[d], dph);
        mbar_wait(&ctl->e_empty[e1], eph ^ 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t et = tmem_base + (uint32_t)(G0_E + e1 * G0_E_STRIDE);
            const uint64_t ad = umma_desc_k_sw128(dq + (uint32_t)((ks >> 2) * A_STAGE_BYTES)) + (uint64_t)((ks & 3) * 2);
            umma_bf16(et, ad, bdesc[ks], idesc_k, ks != 0);
            if (SPLIT) { umma_bf16(et, ad + D_LO, bdesc[ks], idesc_k, true); umma_bf16(et, ad, bdesc[ks] + B_LO, idesc_k, true); }
          }
          umma_commit(&ctl->e_full[e1]);
        }
        __syncwarp();
        mbar_wait(&ctl->e_empty[e1 + 1], eph ^ 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t et = tmem_base + (uint32_t)(G0_E + (e1 + 1) * G0_E_STRIDE);
            const uint64_t ad = umma_desc_mn_sw128(dq + (uint32_t)(ks * 2048), A_STAGE_BYTES, 1024);
            umma_bf16(et, ad, bdesc[ks], idesc_mn, ks != 0);
            if (SPLIT) { umma_bf16(et, ad + D_LO, bdesc[ks], idesc_mn, true); umma_bf16(et, ad, bdesc[ks] + B_LO, idesc_mn, true); }
          }
          umma_commit(&ctl->e_full[e1 + 1]);
          umma_commit(&ctl->dq_empty[d]);
          if (q == Q - 1) umma_commit(&ctl->a_free);
        }
        __syncwarp();
        if (++d == G0_ND) { d = 0; dph ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ accumulating MMAs (TS): M = 128, N = KA, K = KA
    // term 2: B[N = c][K = k] = Wq[k, c] = slab row c, K-major; term 1: B[N = c][K = n] = Wq[c, n] = slab row n, column c:
    // MN-major, a K step is 16 slab rows (2048 bytes), the two 64-column blocks are KA * 128 bytes apart
    const uint32_t idesc = umma_idesc_bf16(BM, KA), idesc_t1 = umma_idesc_bf16(BM, KA, false, true);
    uint64_t wk[F0_KA_MAX / UMMA_K];
#pragma unroll
    for (int k = 0; k < F0_KA_MAX / UMMA_K; ++k) wk[k] = (uint64_t)((((k >> 2) * KA * 128) >> 4) + (k & 3) * 2);
    constexpr uint64_t W_LO = (uint64_t)(F0_SLAB_BYTES >> 4);
    constexpr uint32_t e_lo = 8;                           // split mode: a K step of E = 16 columns, hi words | lo words
    constexpr uint32_t e_step = SPLIT ? 16 : 8;
    const uint32_t d_tmem = tmem_base + (uint32_t)G0_D;
    uint32_t n = 0;
    for (int t = 0; t < my_tiles; ++t) {
      mbar_wait(&ctl->d_empty, (uint32_t)((t & 1) ^ 1));
      for (int q = 0; q < Q; ++q, ++n) {
        const int s = n % G0_NW; const uint32_t wph = (n / G0_NW) & 1;
        const int e1 = (n & 1) * 2; const uint32_t eph = (n >> 1) & 1;
        const uint64_t wT = umma_desc_mn_sw128(smem_u32(sW + s * WSTAGE), (uint32_t)(KA * 128), 1024), wN = umma_desc_k_sw128(smem_u32(sW + s * WSTAGE));
        mbar_wait(&ctl->w_full[s], wph);
        mbar_wait(&ctl->e_conv[e1], eph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < F0_KA_MAX / UMMA_K; ++k)
            if (k < ksteps) {
              const uint32_t ea = tmem_base + (uint32_t)(G0_E + e1 * G0_E_STRIDE) + (uint32_t)k * e_step;
              const uint64_t wd = wT + (uint64_t)(k * (2048 >> 4));
              umma_bf16_ts(d_tmem, ea, wd, idesc_t1, (q | k) != 0);
              if (SPLIT) { umma_bf16_ts(d_tmem, ea + e_lo, wd, idesc_t1, true); umma_bf16_ts(d_tmem, ea, wd + W_LO, idesc_t1, true); }
            }
          umma_commit(&ctl->e_empty[e1]);
        }
        __syncwarp();
        mbar_wait(&ctl->e_conv[e1 + 1], eph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < F0_KA_MAX / UMMA_K; ++k)
            if (k < ksteps) {
              const uint32_t ea = tmem_base + (uint32_t)(G0_E + (e1 + 1) * G0_E_STRIDE) + (uint32_t)k * e_step;
              umma_bf16_ts(d_tmem, ea, wN + wk[k], idesc, true);
              if (SPLIT) { umma_bf16_ts(d_tmem, ea + e_lo, wN + wk[k], idesc, true); umma_bf16_ts(d_tmem, ea, wN + wk[k] + W_LO, idesc, true); }
            }
          umma_commit(&ctl->e_empty[e1 + 1]);
          umma_commit(&ctl->w_empty[s]);
          if (q == Q - 1) umma_commit(&ctl->d_full);
        }
        __syncwarp();
      }
    }
  } else if (warp == 3) {
    // (TMEM allocation only)
  } else if (warp < 12) {
    // ------------------------------------------------------------------ converters: E fp32 -> bf16 in place
    const int term = (warp - 4) >> 2;                     // set 0: E1 (even buffers), set 1: E2
    const int qd = warp & 3;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    uint32_t n = 0;
    for (int t = 0; t < my_tiles; ++t) {
      for (int q = 0; q < Q; ++q, ++n) {
        const int e = (n & 1) * 2 + term; const uint32_t eph = (n >> 1) & 1;
        const uint32_t ea = tmem_base + lane_off + (uint32_t)(G0_E + e * G0_E_STRIDE);
        mbar_wait(&ctl->e_full[e], eph);
        tc_fence_after();
        if constexpr (SPLIT) {
          // every 16-column chunk in place: its hi words -> the chunk's columns 0..7, its lo words -> 8..15 (a chunk is
          // one K step of the accumulating MMAs; nothing has to wait in registers for other columns to be read)
#pragma unroll
          for (int pass = 0; pass < (F0_KA_MAX / 16 + 1) / 2; ++pass) {
            float v[2][16];
#pragma unroll
            for (int c = 0; c < 2; ++c)
              if ((2 * pass + c) * 16 < KA && 2 * pass + c < F0_KA_MAX / 16) tmem_ld16(ea + (uint32_t)((2 * pass + c) * 16), v[c]);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 2; ++c)
              if ((2 * pass + c) * 16 < KA && 2 * pass + c < F0_KA_MAX / 16) {
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) pk[j] = pack2(v[c][2 * j], v[c][2 * j + 1]);
                tmem_st8(ea + (uint32_t)((2 * pass + c) * 16), pk);
#pragma unroll
                for (int j = 0; j < 8; ++j) pk[j] = pack2_lo(v[c][2 * j], v[c][2 * j + 1]);
                tmem_st8(ea + (uint32_t)((2 * pass + c) * 16 + 8), pk);
              }
          }
        } else
        // two passes (48 + 32 columns) keep the register count inside the 640-thread budget; the bf16 words of
        // the first pass land in columns 0..23, which pass two has no need to read (it reads 48..79)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float v[3][16];
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if ((half * 3 + c) * 16 < KA && (half == 0 || c < 2)) tmem_ld16(ea + (uint32_t)((half * 3 + c) * 16), v[c]);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if ((half * 3 + c) * 16 < KA && (half == 0 || c < 2)) {
              uint32_t pk[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) pk[j] = pack2(v[c][2 * j], v[c][2 * j + 1]);
              tmem_st8(ea + (uint32_t)((half * 3 + c) * 8), pk);
            }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->e_conv[e]);
      }
      if (term == 0) {
        // ---- tile result: D + pooling term -> g_rows[b][f][2x .. 2x+1]
        const int tile = (int)blockIdx.x + t * (int)gridDim.x;
        const int r = qd * 32 + lane, x = r & 15;
        const int b = tile * 8 + (r >> 4);
        mbar_wait(&ctl->d_full, (uint32_t)(t & 1));
        tc_fence_after();
        const bool ok = b < prm.B;
        const float gb = ok ? __ldg(prm.gout + b) : 0.f;
        const float v0 = __ldg(prm.v_head + 2 * x), v1 = __ldg(prm.v_head + 2 * x + 1);
#pragma unroll
        for (int c = 0; c < F0_KA_MAX / 16; ++c)
          if (c * 16 < KA) {
            float v[16];
            tmem_ld16(tmem_base + lane_off + (uint32_t)(G0_D + c * 16), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int f = c * 8 + j;
              if (ok && f < prm.F) {
                const float2 pt = __ldg(prm.pterm + (int64_t)b * prm.F + f);
                float2 o;
                o.x = v[2 * j] + gb * fmaf(v0, pt.x, pt.y);
                o.y = v[2 * j + 1] + gb * fmaf(v1, pt.x, pt.y);
                *reinterpret_cast<float2*>(prm.g_rows + ((int64_t)b * prm.F + f) * 32 + 2 * x) = o;
              }
            }
          }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->d_empty);
      }
    }
  } else {
    // ------------------------------------------------------------------ builders: dY0 -> diagonal blocks of Dq
    const int grp = (warp - 12) >> 2;                     // q groups of 8: group g takes those with (q >> 3) & 1 == g
                                                          // (split mode: every group of 8; g = 0 the hi part, g = 1 the lo)
    const int r = (warp & 3) * 32 + lane;                 // tile row = (sample r >> 4, h = r & 15)
    const int bl = r >> 4, h = r & 15;
    const uint32_t part_off = SPLIT ? (uint32_t)(grp * G0_DQ_BYTES) : 0u;
    const uint32_t dq_off = part_off + (uint32_t)((bl >> 2) * A_STAGE_BYTES) + sw128_offset(r, (bl & 3) * 2);
    const uint32_t dq_off2 = part_off + (uint32_t)((bl >> 2) * A_STAGE_BYTES) + sw128_offset(r, (bl & 3) * 2 + 1);
    const bf16* dYsrc = SPLIT && grp == 1 ? prm.dYlo : prm.dY;
    uint32_t n = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = (int)blockIdx.x + t * (int)gridDim.x;
      const int b = tile * 8 + bl;
      if (grp == 0) {
        // A tile of this tile's samples (every E MMA of the previous tile has retired: a_free)
        if (t > 0) mbar_wait(&ctl->a_free, (uint32_t)((t - 1) & 1));
        const float* src = prm.rows + ((int64_t)(b < prm.B ? b : 0) * prm.F) * 32 + 2 * h;
        for (int c = 0; c < KA / 8; ++c) {
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
            *reinterpret_cast<uint4*>(sAt + ABYTES + (c >> 3) * A_STAGE_BYTES + sw128_offset(r, c & 7)) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->a_ready);
      }
      const bf16* src = dYsrc + (((int64_t)(b < prm.B ? b : 0) * 16 + h) * 16) * prm.Pp;
      for (int q0 = 0; q0 < Q; q0 += 8, n += 8) {
        if (!SPLIT && ((q0 >> 3) & 1) != grp) continue;
        uint4 dv[16];
#pragma unroll
        for (int w = 0; w < 16; ++w)
          dv[w] = b < prm.B ? __ldg(reinterpret_cast<const uint4*>(src + (int64_t)w * prm.Pp + q0)) : make_uint4(0u, 0u, 0u, 0u);
        {  // d b_0[q] = sum of dY0 over all positions: this warp's 32 rows x 16 w of the 8 channels (the data is here anyway)
          float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int w = 0; w < 16; ++w) {
            const uint32_t wv[4] = {dv[w].x, dv[w].y, dv[w].z, dv[w].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { cs[2 * j] += __uint_as_float(wv[j] << 16); cs[2 * j + 1] += __uint_as_float(wv[j] & 0xFFFF0000u); }
          }
          float mine = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float v = warp_sum(cs[j]); if (lane == j) mine = v; }
          if (lane < 8) prm.bpart[(((int64_t)tile * (SPLIT ? 2 : 1) + (SPLIT ? grp : 0)) * 4 + (warp & 3)) * Q + q0 + lane] = mine;
        }
        // The groups take the 8-channel groups alternately and must fill the Dq ring in channel order (a parity
        // wait cannot tell one ring revolution from the next): wait until the other group has finished the
        // preceding 8 channels.  n >> 3 = index of this 8-channel group over the whole kernel.
        if (!SPLIT && n >= 8) mbar_wait(&ctl->grp_done[grp ^ 1], (uint32_t)((((n >> 3) - 1) >> 1) & 1));
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) {
          const uint32_t m = n + qq;
          const int d = m % G0_ND; const uint32_t dph = (m / G0_ND) & 1;
          uint32_t pk[8];
#pragma unroll
          for (int w = 0; w < 16; w += 2) {
            const uint32_t a = (qq >> 1) == 0 ? dv[w].x : (qq >> 1) == 1 ? dv[w].y : (qq >> 1) == 2 ? dv[w].z : dv[w].w;
            const uint32_t c = (qq >> 1) == 0 ? dv[w + 1].x : (qq >> 1) == 1 ? dv[w + 1].y : (qq >> 1) == 2 ? dv[w + 1].z : dv[w + 1].w;
            pk[w >> 1] = __byte_perm(a, c, (qq & 1) ? 0x7632 : 0x5410);
          }
          mbar_wait(SPLIT ? &ctl->dqw_empty[warp & 3] : &ctl->dq_empty[d], dph ^ 1);
          uint8_t* base = sDq + d * DBUF;
          *reinterpret_cast<uint4*>(base + dq_off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(base + dq_off2) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(SPLIT ? &ctl->dqw_full[warp & 3] : &ctl->dq_full[d]);
        }
        if (!SPLIT && lane == 0) mbar_arrive(&ctl->grp_done[grp]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem_base, 512);
}
