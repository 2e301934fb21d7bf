// Layer-0 weight gradient in factorised form (included inside namespace cffm::tc of conv_tc.cu, after conv0_dfact.cuh).
//
//   dWq[k, n] = sum_{b,h,w} dY0[b,h,w,q] a_{b,h}[k] a_{b,w}[n]          (Wq, a: conv0_fact.cuh)
//             = sum_{(b,h)} A[(b,h), k] E_q[(b,h), n],   E_q[(b,h), n] = sum_w dY0[b,h,w,q] a_{b,w}[n]
//
//   E step   Et[n, (b,h)] = sum_w A[(b,w), n] dY0[b,h,w,q]                  per sample one SS MMA, M = n (A tile rows of
//                                                                        the sample as MN-major A operand), N = 16, K = 16
//   W step   dWq^T[n, k] += sum_{(b,h)} Et[n,(b,h)] A[(b,h), k]          TS MMA: Et converted to bf16 in place in TMEM,
//                                                                        B = the A tile as MN-major B operand, K = 128
// (E is produced transposed so that the contraction index of the second step runs along TMEM columns.)
// The E step touches only the 16 x 16 block of a sample, so the dY0 blocks of a channel are kept compact: two 2 KB
// regions of 16 rows (h) x 128 bytes, sample s of the tile in region s / 4 at byte columns 32 (s % 4) -- 4 KB per
// channel instead of a 32 KB block-diagonal matrix.
// Split mode (bf16x3): A8 and dY0 come as hi + lo; every product is hi*hi + lo*hi + hi*lo into the same fp32
// accumulator (three MMAs), E^T is split sample by sample (16 columns: hi words | lo words) in place, one builder group
// places the hi blocks and the other the lo blocks of every tile.
// The accumulators of 4 channels stay in TMEM while the CTA walks its share of the batch, so a unit is
// (4 channels) x (a third of the tiles); CTAs that run at the same time hold neighbouring channel sets and read
// the same 32-byte sectors of dY0 (L2 hits).  Partial sums per split are reduced in fixed order by k_wfact_reduce.
// A unit spends only 4 channels on a tile, so the per-tile work has to be cheap: the A tiles are rows of a bf16
// matrix A8[(b,h)][k] made once per step (k_build_a8) and arrive by TMA; builders only place the dY0 blocks.
//
//   warp 0       TMA: A tile of the next batch tile (double-buffered)
//   warp 1       issues the E MMAs          warp 2   issues the W MMAs          warp 3   TMEM allocation
//   warps 4..11  two converter sets (alternate E buffers): Et fp32 -> bf16 in place; set 0 also writes the
//                accumulators of a finished unit to the partial sums
//   warps 12..19 two builder groups, alternate tiles: dY0 -> diagonal blocks of Dq
#pragma once

constexpr int W0_THREADS = 640;
constexpr int W0_QS = 4;                                   // channels per unit: one aligned 8-byte load per position;
                                                           // 4 x 80 accumulator columns + 3 x 64 for E^T = 512 TMEM columns
constexpr int W0_SPLIT_MAX = 3;                            // batch splits (fewer when the batch has fewer tiles)
constexpr int W0_ND = 3, W0_NE = 3;                        // Dq buffers; E buffers (an item = half a tile = 4 samples = 64 columns)
constexpr int W0_DW = 0, W0_DW_STRIDE = 80, W0_E = 320, W0_E_STRIDE = 64;

struct W0Ctl {
  uint64_t a_ready[2], a_free[2], dq_full[W0_ND], dq_empty[W0_ND];
  uint64_t e_full[W0_NE], e_conv[W0_NE], e_empty[W0_NE];
  uint64_t dw_full, dw_empty, grp_done[2];
  uint32_t tmem_base, pad;
};
static_assert(sizeof(W0Ctl) <= 256, "control block");
constexpr int W0_DQC_BYTES = 4096;                          // compact dY0 blocks of one channel (8 samples x 16 x 16 bf16)
constexpr int w0_smem(bool split) { return 1024 + (split ? 2 : 1) * (2 * G0_DQ_BYTES + W0_ND * W0_DQC_BYTES) + 256; }
static_assert(w0_smem(true) <= 227 * 1024, "factorised weight gradient exceeds the shared memory of an SM");

struct Wgrad0FactParams {
  CUtensorMap mapA;          // A8 [B8*16 rows][nblk*64 cols] bf16 (B8 = batch rounded up to 8), box (64, 128)
  CUtensorMap mapA2;         // split mode: its lo half
  const bf16* dY;            // dY0 [B][16][16][Pp]
  const bf16* dYlo;          // split mode: its lo half
  float* part;               // [nsplit][Q16][KA (n)][KA (k)] fp32
  int B, F, P, Pp, KA, nblk, Q16, nsplit;
};

// A8[(b,h)][2i+d] = o_i[b, 2h+d] in bf16, zero beyond 2F and for padded samples: the A tile of every 8-sample tile
__global__ void k_build_a8(const float* __restrict__ rows, int B, int B8, int F, int KP, bf16* __restrict__ out, bf16* __restrict__ out_lo) {
  const int64_t total = (int64_t)B8 * 16 * (KP / 2);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e % (KP / 2));
    const int64_t bh = e / (KP / 2);
    const int h = (int)(bh & 15); const int64_t b = bh >> 4;
    float2 o = make_float2(0.f, 0.f);
    if (i < F && b < B) o = *reinterpret_cast<const float2*>(rows + (b * F + i) * 32 + 2 * h);
    reinterpret_cast<uint32_t*>(out)[e] = pack2(o.x, o.y);
    if (out_lo) reinterpret_cast<uint32_t*>(out_lo)[e] = pack2_lo(o.x, o.y);
  }
}

// g[dh][dw][p][q] = sum over splits of part[s][q][n = 2 j_p + dw][k = 2 i_p + dh]
__global__ void k_wfact_reduce(const float* __restrict__ part, const int* __restrict__ pair_i, const int* __restrict__ pair_j, int P,
                               int KA, int Q16, int nsplit, float* __restrict__ g) {
  const int64_t total = 4ll * P * P;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(e % P);
    const int64_t r = e / P;
    const int p = (int)(r % P), tap = (int)(r / P);
    const int k = 2 * pair_i[p] + (tap >> 1), n = 2 * pair_j[p] + (tap & 1);
    float s = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) s += part[(((int64_t)sp * Q16 + q) * KA + n) * KA + k];
    g[e] = s;
  }
}

template <bool SPLIT>
__global__ void __launch_bounds__(W0_THREADS, 1) k_wgrad0_fact(const __grid_constant__ Wgrad0FactParams prm) {
  constexpr int NP = SPLIT ? 2 : 1;                      // precision parts of an operand: hi (, lo)
  constexpr int ABUF = NP * G0_DQ_BYTES, DBUF = NP * W0_DQC_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sAt = smem;                                   // 2 A tiles (split: hi tile, lo tile)
  uint8_t* sDq = sAt + 2 * ABUF;                         // W0_ND buffers of compact dY0 blocks (split: hi, lo)
  W0Ctl* ctl = reinterpret_cast<W0Ctl*>(sDq + W0_ND * DBUF);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int KA = prm.KA;
  const int n_tiles = (prm.B + 7) >> 3;
  const int n_qs = (prm.Q16 + W0_QS - 1) / W0_QS;
  const int n_units = n_qs * prm.nsplit;
  const int my_units = (int)blockIdx.x < n_units ? (n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  // unit u: channel set u % n_qs, tiles [t0, t1) of split u / n_qs
  auto unit_tiles = [&](int u, int& t0, int& t1) { const int sp = u / n_qs; t0 = sp * n_tiles / prm.nsplit; t1 = (sp + 1) * n_tiles / prm.nsplit; };

  if (warp == 1 && lane == 0) {
    for (int b = 0; b < 2; ++b) { mbar_init(&ctl->a_ready[b], 1); mbar_init(&ctl->a_free[b], 1); mbar_init(&ctl->grp_done[b], 4); }
    for (int d = 0; d < W0_ND; ++d) { mbar_init(&ctl->dq_full[d], 4 * NP); mbar_init(&ctl->dq_empty[d], 1); }
    for (int e = 0; e < W0_NE; ++e) {
      mbar_init(&ctl->e_full[e], 1); mbar_init(&ctl->e_conv[e], 4); mbar_init(&ctl->e_empty[e], 1);
    }
    mbar_init(&ctl->dw_full, 1); mbar_init(&ctl->dw_empty, 4);
    fence_barrier_init();
  }
  if (warp == 3) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA: A tiles
    prefetch_tmap(&prm.mapA);
    if (SPLIT) prefetch_tmap(&prm.mapA2);
    uint32_t T = 0;
    for (int it = 0; it < my_units; ++it) {
      int t0, t1; unit_tiles((int)blockIdx.x + it * (int)gridDim.x, t0, t1);
      for (int t = t0; t < t1; ++t, ++T) {
        const int ab = T & 1;
        mbar_wait(&ctl->a_free[ab], ((T >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&ctl->a_ready[ab], (uint32_t)(NP * prm.nblk * A_STAGE_BYTES));
          for (int blk = 0; blk < prm.nblk; ++blk) {
            tma_load_2d(sAt + ab * ABUF + blk * A_STAGE_BYTES, &prm.mapA, &ctl->a_ready[ab], blk * 64, t * BM);
            if (SPLIT) tma_load_2d(sAt + ab * ABUF + G0_DQ_BYTES + blk * A_STAGE_BYTES, &prm.mapA2, &ctl->a_ready[ab], blk * 64, t * BM);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ E MMAs
    const uint32_t at_addr = smem_u32(sAt), dq_addr = smem_u32(sDq);
    // One M=128 (n) x N=16 (h) x K=16 (w) MMA per sample: the sample's 16 x 16 block of dY0 and the 16 rows of the A
    // tile that belong to it, no accumulation over samples (split mode: three MMAs into the same columns).
    const uint32_t idesc = umma_idesc_bf16(BM, 16, true, false);
    uint32_t T = 0, n = 0; int d = 0; uint32_t dph = 0;
    for (int it = 0; it < my_units; ++it) {
      int t0, t1; unit_tiles((int)blockIdx.x + it * (int)gridDim.x, t0, t1);
      for (int t = t0; t < t1; ++t, ++T) {
        const int ab = T & 1;
        mbar_wait(&ctl->a_ready[ab], (T >> 1) & 1);
        tc_fence_after();
        // descriptors: base of the tile / buffer + a constant per sample (start-address field counts 16-byte units)
        const uint64_t a_base = umma_desc_mn_sw128(at_addr + (uint32_t)(ab * ABUF), A_STAGE_BYTES, 1024);
        const uint64_t a_lo = umma_desc_mn_sw128(at_addr + (uint32_t)(ab * ABUF + G0_DQ_BYTES), A_STAGE_BYTES, 1024);
        for (int j = 0; j < W0_QS; ++j) {
          const uint64_t b_base = umma_desc_k_sw128(dq_addr + (uint32_t)(d * DBUF));
          const uint64_t b_lo = umma_desc_k_sw128(dq_addr + (uint32_t)(d * DBUF + W0_DQC_BYTES));
          mbar_wait(&ctl->dq_full[d], dph);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf, ++n) {   // an item is half a tile: 4 samples, 64 columns of E^T
            const int e = n % W0_NE; const uint32_t eph = (n / W0_NE) & 1;
            const uint32_t et = tmem_base + (uint32_t)(W0_E + e * W0_E_STRIDE);
            mbar_wait(&ctl->e_empty[e], eph ^ 1);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int sb = 0; sb < 4; ++sb) {
                const int ks = hf * 4 + sb;         // sample of the tile
                const uint64_t ao = (uint64_t)(ks * (2048 >> 4)), bo = (uint64_t)((((ks >> 2) * 2048) >> 4) + (ks & 3) * 2);
                umma_bf16(et + (uint32_t)(sb * 16), a_base + ao, b_base + bo, idesc, false);
                if (SPLIT) {
                  umma_bf16(et + (uint32_t)(sb * 16), a_lo + ao, b_base + bo, idesc, true);
                  umma_bf16(et + (uint32_t)(sb * 16), a_base + ao, b_lo + bo, idesc, true);
                }
              }
              umma_commit(&ctl->e_full[e]);
              if (hf == 1) umma_commit(&ctl->dq_empty[d]);
            }
            __syncwarp();
          }
          if (++d == W0_ND) { d = 0; dph ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ W MMAs (TS): M = 128 (n), N = KA (k), K = 128 (b,h)
    const uint32_t at_addr = smem_u32(sAt);
    const uint32_t idesc = umma_idesc_bf16(BM, KA, false, true);
    uint32_t T = 0, n = 0;
    for (int it = 0; it < my_units; ++it) {
      int t0, t1; unit_tiles((int)blockIdx.x + it * (int)gridDim.x, t0, t1);
      mbar_wait(&ctl->dw_empty, (uint32_t)((it & 1) ^ 1));
      tc_fence_after();
      for (int t = t0; t < t1; ++t, ++T) {
        const int ab = T & 1;
        const uint64_t b_base = umma_desc_mn_sw128(at_addr + (uint32_t)(ab * ABUF), A_STAGE_BYTES, 1024);
        const uint64_t b_lo = umma_desc_mn_sw128(at_addr + (uint32_t)(ab * ABUF + G0_DQ_BYTES), A_STAGE_BYTES, 1024);
        const bool acc0 = t != t0;
#pragma unroll
        for (int j = 0; j < W0_QS; ++j) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf, ++n) {
            const int e = n % W0_NE; const uint32_t eph = (n / W0_NE) & 1;
            const uint32_t et = tmem_base + (uint32_t)(W0_E + e * W0_E_STRIDE);
            mbar_wait(&ctl->e_conv[e], eph);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int sb = 0; sb < 4; ++sb) {
                const uint32_t dw = tmem_base + (uint32_t)(W0_DW + j * W0_DW_STRIDE);
                const uint64_t bo = (uint64_t)((hf * 4 + sb) * (2048 >> 4));
                const uint32_t ea = et + (uint32_t)(sb * (SPLIT ? 16 : 8));   // split: a sample's 16 columns = hi words | lo words
                umma_bf16_ts(dw, ea, b_base + bo, idesc, acc0 || hf != 0 || sb != 0);
                if (SPLIT) {
                  umma_bf16_ts(dw, ea + 8, b_base + bo, idesc, true);
                  umma_bf16_ts(dw, ea, b_lo + bo, idesc, true);
                }
              }
              umma_commit(&ctl->e_empty[e]);
              if (j == W0_QS - 1 && hf == 1) umma_commit(&ctl->a_free[ab]);
              if (j == W0_QS - 1 && hf == 1 && t == t1 - 1) umma_commit(&ctl->dw_full);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 3) {
    // (TMEM allocation only)
  } else if (warp < 12) {
    // ------------------------------------------------------------------ converters: Et fp32 -> bf16 in place
    const int set = (warp - 4) >> 2;                      // two sets, alternate items
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    uint32_t n = 0;
    for (int it = 0; it < my_units; ++it) {
      const int u = (int)blockIdx.x + it * (int)gridDim.x;
      int t0, t1; unit_tiles(u, t0, t1);
      for (int t = t0; t < t1; ++t)
        for (int j = 0; j < 2 * W0_QS; ++j, ++n) {        // items: (channel, half tile)
          if ((int)(n & 1) != set) continue;              // the sets take alternate items
          const int e = n % W0_NE; const uint32_t eph = (n / W0_NE) & 1;
          const uint32_t ea = tmem_base + lane_off + (uint32_t)(W0_E + e * W0_E_STRIDE);
          mbar_wait(&ctl->e_full[e], eph);
          tc_fence_after();
          if constexpr (SPLIT) {
            // sample by sample in place: the 16 columns of a sample -> its hi words (columns 0..7) | lo words (8..15)
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
              float v[2][16];
              tmem_ld16(ea + (uint32_t)(pass * 32), v[0]);
              tmem_ld16(ea + (uint32_t)(pass * 32 + 16), v[1]);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                uint32_t pk[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) pk[jj] = pack2(v[c][2 * jj], v[c][2 * jj + 1]);
                tmem_st8(ea + (uint32_t)(pass * 32 + c * 16), pk);
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) pk[jj] = pack2_lo(v[c][2 * jj], v[c][2 * jj + 1]);
                tmem_st8(ea + (uint32_t)(pass * 32 + c * 16 + 8), pk);
              }
            }
          } else
          // two passes of 32 columns; pass p writes bf16 columns 16p.., whose fp32 content has been read
#pragma unroll
          for (int pass = 0; pass < 2; ++pass) {
            float v[2][16];
            tmem_ld16(ea + (uint32_t)(pass * 32), v[0]);
            tmem_ld16(ea + (uint32_t)(pass * 32 + 16), v[1]);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t pk[8];
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) pk[jj] = pack2(v[c][2 * jj], v[c][2 * jj + 1]);
              tmem_st8(ea + (uint32_t)(pass * 16 + c * 8), pk);
            }
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl->e_conv[e]);
        }
      if (set != 0) continue;
      // ---- unit result: 4 accumulators [k = row][n] -> partial sums of this split
      mbar_wait(&ctl->dw_full, (uint32_t)(it & 1));
      tc_fence_after();
      const int qs = u % n_qs, sp = u / n_qs;
      for (int j = 0; j < W0_QS; ++j) {
        const int q = qs * W0_QS + j;
        float* dst = prm.part + ((((int64_t)sp * prm.Q16 + q) * KA + r) * KA);   // row n = r, columns k
#pragma unroll
        for (int c = 0; c < F0_KA_MAX / 16; ++c)
          if (c * 16 < KA) {
            float v[16];
            tmem_ld16(tmem_base + lane_off + (uint32_t)(W0_DW + j * W0_DW_STRIDE + c * 16), v);
            tmem_ld_wait();
            if (r < KA && q < prm.Q16) {
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4)
                reinterpret_cast<float4*>(dst + c * 16)[q4] = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
            }
          }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->dw_empty);
    }
  } else {
    // ------------------------------------------------------------------ builders: groups alternate tiles (split mode:
    // every tile, group 0 the hi blocks, group 1 the lo blocks)
    const int grp = (warp - 12) >> 2;
    const int r = (warp & 3) * 32 + lane;
    const int bl = r >> 4, h = r & 15;
    const uint32_t dq_off = (uint32_t)((SPLIT ? grp * W0_DQC_BYTES : 0) + (bl >> 2) * 2048) + sw128_offset(h, (bl & 3) * 2);
    const uint32_t dq_off2 = (uint32_t)((SPLIT ? grp * W0_DQC_BYTES : 0) + (bl >> 2) * 2048) + sw128_offset(h, (bl & 3) * 2 + 1);
    const bf16* dYsrc = SPLIT && grp == 1 ? prm.dYlo : prm.dY;
    constexpr uint32_t TSTEP = SPLIT ? 1 : 2;
    // T = index of a (unit, tile) pair in this CTA's walk; group g takes the pairs with T & 1 == g (split mode: all)
    auto decode = [&](uint32_t T, int& t, int& q0) -> bool {
      uint32_t acc = 0;
      for (int it = 0; it < my_units; ++it) {
        const int u = (int)blockIdx.x + it * (int)gridDim.x;
        int t0, t1; unit_tiles(u, t0, t1);
        if (T < acc + (uint32_t)(t1 - t0)) { t = t0 + (int)(T - acc); q0 = (u % n_qs) * W0_QS; return true; }
        acc += (uint32_t)(t1 - t0);
      }
      return false;
    };
    // channels q0 .. q0+3 (q0 = 4 * set index): one aligned 8-byte load per position
    uint2 raw[16];
    auto issue = [&](int t, int q0) {
      const int b = t * 8 + bl;
      const bool ok = b < prm.B;
      const bf16* src = dYsrc + (((int64_t)(ok ? b : 0) * 16 + h) * 16) * prm.Pp + q0;
#pragma unroll
      for (int w = 0; w < 16; ++w) raw[w] = ok ? __ldg(reinterpret_cast<const uint2*>(src + (int64_t)w * prm.Pp)) : make_uint2(0u, 0u);
    };
    uint32_t T = SPLIT ? 0u : (uint32_t)grp;
    int t, q0;
    bool have = decode(T, t, q0);
    if (have) issue(t, q0);
    while (have) {
      // the loads of this tile were issued one round ago: compress them to the three channels
      uint32_t cur[W0_QS][8];
#pragma unroll
      for (int j = 0; j < W0_QS; ++j) {
        const uint32_t sel = (j & 1) ? 0x7632 : 0x5410;
#pragma unroll
        for (int w = 0; w < 16; w += 2)
          cur[j][w >> 1] = __byte_perm((j >> 1) ? raw[w].y : raw[w].x, (j >> 1) ? raw[w + 1].y : raw[w + 1].x, sel);
      }
      // next tile of this group: loads in flight while this tile's blocks are placed
      int tn, qn;
      const bool have_next = decode(T + TSTEP, tn, qn);
      if (have_next) issue(tn, qn);
      // ring order: the other group must have finished the previous tile's channels
      if (!SPLIT && T >= 1) mbar_wait(&ctl->grp_done[grp ^ 1], (uint32_t)((((T - 1) >> 1)) & 1));
#pragma unroll
      for (int j = 0; j < W0_QS; ++j) {
        const uint32_t m = T * W0_QS + j;
        const int d = m % W0_ND; const uint32_t dph = (m / W0_ND) & 1;
        mbar_wait(&ctl->dq_empty[d], dph ^ 1);
        uint8_t* base = sDq + d * DBUF;
        *reinterpret_cast<uint4*>(base + dq_off) = make_uint4(cur[j][0], cur[j][1], cur[j][2], cur[j][3]);
        *reinterpret_cast<uint4*>(base + dq_off2) = make_uint4(cur[j][4], cur[j][5], cur[j][6], cur[j][7]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->dq_full[d]);
      }
      if (!SPLIT && lane == 0) mbar_arrive(&ctl->grp_done[grp]);
      T += TSTEP; t = tn; q0 = qn; have = have_next;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) tmem_dealloc(tmem_base, 512);
}
