// Warp-specialised tcgen05 pipeline shared by every tensor-core kernel of the conv stack.
//
//   warp 0        : TMA producer (one elected lane): operand tiles -> shared memory stages
//   warp 1        : MMA issuer (lane 0): tcgen05.mma over the stages, accumulators in TMEM
//   warps 2..5(9) : epilogue: tcgen05.ld TMEM -> registers -> policy epilogue (bias, activation,
//                   pooling sums, gradient contractions ...) -> global memory
//   next 8 warps  : (policies with kSynthA) build the A stage in shared memory themselves, in the
//                   UMMA SWIZZLE_128B layout -- used for the interaction cube, which never exists
//                   in HBM (CFFM.py:355-367)
//
// Three mbarrier pipelines: smem full/empty (producers <-> MMA), TMEM full/empty (MMA <-> epilogue,
// two accumulator buffers), and the persistent unit loop every role walks in the same order.
#pragma once
#include "tc05.cuh"

namespace cffm {
namespace tc {

constexpr int MAX_STAGES = 4;
constexpr int MAX_BN = 256;
constexpr int B_STAGE_BYTES_MAX = MAX_BN * BK * 2;
constexpr int SYNTH_WARPS = 8;
template <class P>
constexpr int block_threads() { return 64 + 32 * P::kEpiWarps + (P::kSynthA ? 32 * SYNTH_WARPS : 0); }
constexpr int SMEM_LIMIT = 227 * 1024;
// kBPair: two N tiles (<= 192 wide each) share every A stage; the B stage holds both
constexpr int PAIR_BN_MAX = 192;
template <class P>
__host__ __device__ constexpr int b_stage_bytes() { return P::kBPair ? 2 * PAIR_BN_MAX * BK * 2 : B_STAGE_BYTES_MAX; }

struct Ctl {
  uint64_t full[MAX_STAGES], empty[MAX_STAGES], tfull[2], tempty[2];
  uint32_t tmem_base;
  uint32_t pad;
};
constexpr int CTL_BYTES = (sizeof(Ctl) + 63) & ~63;

// dynamic shared memory of a policy: 1 KB alignment slack + stages + control block + policy scratch
template <class P>
constexpr size_t smem_bytes() {
  return 1024 + (size_t)P::kStages * ((P::kATmem ? 0 : P::kATiles * A_STAGE_BYTES) + b_stage_bytes<P>()) + CTL_BYTES + P::kExtraBytes;
}

__device__ __forceinline__ uint32_t tmem_cols_for(int cols) {
  uint32_t c = 32;
  while ((int)c < cols) c <<= 1;
  return c;
}

struct Unit { int m_tile, n_tile, z; };

// Operand descriptors of one UMMA (16 reduction elements) inside a stage.
struct KMajorA {   // rows x 64 bf16, reduction index contiguous: step = 32 bytes
  __device__ uint64_t a_desc(uint32_t addr, int k) const { return umma_desc_k_sw128(addr) + (uint64_t)(k * 2); }
};
struct KMajorB {
  __device__ uint64_t b_desc(uint32_t addr, int k) const { return umma_desc_k_sw128(addr) + (uint64_t)(k * 2); }
};
struct MNMajorA {  // [64 reduction rows x 64 elements] blocks of 8 KB: step = 16 rows = 2048 bytes
  __device__ uint64_t a_desc(uint32_t addr, int k) const { return umma_desc_mn_sw128(addr + k * 2048, 8192, 1024); }
};
struct MNMajorB {
  __device__ uint64_t b_desc(uint32_t addr, int k) const { return umma_desc_mn_sw128(addr + k * 2048, 8192, 1024); }
};

// Policy interface (all __device__):
//   static constexpr bool kSynthA
//   static constexpr int kStages, kExtraBytes (policy scratch; the second half belongs to the synth warps)
//   static constexpr int kEpiWarps (4, or 8: two warps per TMEM lane quarter taking alternate 32-column chunks),
//   static constexpr bool kATmem (kSynthA only: the producers write the A stage into tensor memory with tcgen05.st
//                        and the MMA reads it from there -- no shared-memory traffic for A; 32 columns per stage),
//   static constexpr int kATiles (128-row A tiles that share one B stage: accumulators side by side in TMEM),
//                        kAccBufs (1 or 2 accumulator buffers; kAccBufs * kATiles * bn() <= 512 columns)
//   static constexpr bool kBPair (needs kAccBufs == 2, kEpiWarps == 8): a unit covers N tiles 2*n_tile and
//                        2*n_tile+1; both are multiplied with the same A stage (the two accumulator buffers hold the
//                        two tiles, each drained by its own four epilogue warps) -- halves the A work per flop
//   static constexpr bool kEpiPrefetch: the epilogue object has prefetch(Unit) (called before the accumulator is
//                        complete) and chunk_i(Unit, ci, c0, v) for its ci-th chunk instead of chunk()
//   int n_iters(cta, ncta) const; Unit unit(cta, ncta, it) const  -- the unit sequence of one CTA
//   int k_chunks(Unit) const (>= 1); int bn() const  (UMMA N of this launch)
//   void load_a(uint8_t* sA, uint64_t* bar, Unit, int kc) const      -- TMA for the A stage (!kSynthA)
//   void load_b(uint8_t* sB, uint64_t* bar, Unit, int kc) const      -- TMA for the B stage
//   uint32_t tx_bytes() const                                        -- bytes the TMA loads deliver per stage
//   struct SynthState; void synth_begin(Unit, uint8_t* extra, int t, SynthState&) const  -- kSynthA: per-unit staging
//   void synth_a(uint8_t* sA, Unit, int kc, int t, const uint8_t* extra, SynthState&) const  (256 threads)
//   uint64_t a_desc(addr, k), b_desc(addr, k); uint32_t idesc()      -- UMMA descriptors (see KMajorA ...)
//   epilogue object: see each policy
template <class P>
__global__ void __launch_bounds__(block_threads<P>(), 1) k_tc(const __grid_constant__ P prm) {
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment by pointer arithmetic on the __shared__ array (an integer round trip would make
  // every access below a generic LD/ST instead of LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  constexpr int STAGES = P::kStages;
  constexpr int A_BYTES = P::kATmem ? 0 : P::kATiles * A_STAGE_BYTES;
  constexpr int A_TMEM_COLS = BK / 2;        // one stage of A in tensor memory: 64 bf16 per row = 32 columns
  uint8_t* sB = sA + STAGES * A_BYTES;
  constexpr int B_BYTES = b_stage_bytes<P>();
  Ctl* ctl = reinterpret_cast<Ctl*>(sB + STAGES * B_BYTES);
  uint8_t* extra = reinterpret_cast<uint8_t*>(ctl) + CTL_BYTES;
  uint8_t* extra_synth = extra + P::kExtraBytes / 2;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int BN = prm.bn();
  const int ACC_COLS = P::kATiles * BN;      // TMEM columns of one accumulator buffer
  const int A_TMEM_BASE = P::kAccBufs * ACC_COLS;   // first column of the A stages (kATmem)
  const uint32_t ncols = tmem_cols_for(P::kAccBufs * ACC_COLS + (P::kATmem ? STAGES * A_TMEM_COLS : 0));
  const int n_iters = prm.n_iters((int)blockIdx.x, (int)gridDim.x);

  if (warp == 0 && lane == 0) prm.prefetch();
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&ctl->full[s], 1 + (P::kSynthA ? (P::kSynthAlternate ? SYNTH_WARPS / 2 : SYNTH_WARPS) : 0)); mbar_init(&ctl->empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&ctl->tfull[b], 1); mbar_init(&ctl->tempty[b], P::kBPair ? 4 : P::kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&ctl->tmem_base, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  // Both issuing warps run their loops with all 32 lanes (uniform control flow) and put the TMA / tcgen05
  // instructions of one stage inside ONE elect.sync region: ptxas then emits them back to back.  A `lane == 0`
  // test instead costs a ~16-instruction per-thread loop (VOTEU / ELECT / BRA.U.ANY) around every UTCHMMA,
  // about 170 cycles per MMA -- more than a 128x256x16 MMA takes on the tensor pipe.  elect.sync with a full
  // mask picks the same lane every time, which tcgen05.commit relies on.
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < n_iters; ++it) {
      const Unit un = prm.unit((int)blockIdx.x, (int)gridDim.x, it);
      const int KC = prm.k_chunks(un);
      for (int kc = 0; kc < KC; ++kc) {
        mbar_wait(&ctl->empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&ctl->full[stage], prm.tx_bytes());
          if constexpr (!P::kSynthA) prm.load_a(sA + stage * A_BYTES, &ctl->full[stage], un, kc);
          if constexpr (P::kBPair) {
            prm.load_b(sB + stage * B_BYTES, &ctl->full[stage], Unit{un.m_tile, 2 * un.n_tile, un.z}, kc);
            prm.load_b(sB + stage * B_BYTES + BN * BK * 2, &ctl->full[stage], Unit{un.m_tile, 2 * un.n_tile + 1, un.z}, kc);
          } else {
            prm.load_b(sB + stage * B_BYTES, &ctl->full[stage], un, kc);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = prm.idesc();
    int stage = 0; uint32_t phase = 0; int buf = 0; uint32_t bphase = 0;
    for (int it = 0; it < n_iters; ++it) {
      const int KC = prm.k_chunks(prm.unit((int)blockIdx.x, (int)gridDim.x, it));
      if constexpr (P::kBPair) {
        static_assert(!P::kBPair || (P::kATmem && P::kAccBufs == 2 && P::kEpiWarps == 8 && P::kATiles == 1), "kBPair layout");
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(&ctl->full[stage], phase);
          if (kc == 0) mbar_wait(&ctl->tempty[0], bphase ^ 1);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(sB + stage * B_BYTES);
          const uint32_t a_tmem = tmem_base + (uint32_t)(A_TMEM_BASE + stage * A_TMEM_COLS);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16_ts(tmem_base, a_tmem + (uint32_t)(k * (UMMA_K / 2)), prm.b_desc(b_addr, k), idesc, (kc | k) != 0);
            if (kc == KC - 1) umma_commit(&ctl->tfull[0]);
          }
          __syncwarp();
          if (kc == 0) { mbar_wait(&ctl->tempty[1], bphase ^ 1); tc_fence_after(); }
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16_ts(tmem_base + (uint32_t)BN, a_tmem + (uint32_t)(k * (UMMA_K / 2)),
                           prm.b_desc(b_addr + (uint32_t)(BN * BK * 2), k), idesc, (kc | k) != 0);
            if (kc == KC - 1) umma_commit(&ctl->tfull[1]);
            umma_commit(&ctl->empty[stage]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        bphase ^= 1;
        continue;
      }
      mbar_wait(&ctl->tempty[buf], bphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * ACC_COLS);
      for (int kc = 0; kc < KC; ++kc) {
        mbar_wait(&ctl->full[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + stage * A_BYTES);
        const uint32_t b_addr = smem_u32(sB + stage * B_BYTES);
        if (elect_one()) {
          if constexpr (P::kATmem) {
            const uint32_t a_tmem = tmem_base + (uint32_t)(A_TMEM_BASE + stage * A_TMEM_COLS);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16_ts(d_tmem, a_tmem + (uint32_t)(k * (UMMA_K / 2)), prm.b_desc(b_addr, k), idesc, (kc | k) != 0);
          } else {
#pragma unroll
            for (int at = 0; at < P::kATiles; ++at)
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                umma_bf16(d_tmem + (uint32_t)(at * BN), prm.a_desc(a_addr + at * A_STAGE_BYTES, k), prm.b_desc(b_addr, k), idesc,
                          (kc | k) != 0);
          }
          umma_commit(&ctl->empty[stage]);              // frees the smem stage when these MMAs retire
          if (kc == KC - 1) umma_commit(&ctl->tfull[buf]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (P::kAccBufs == 2) { buf ^= 1; if (buf == 0) bphase ^= 1; } else { bphase ^= 1; }
    }
  } else if (warp < 2 + P::kEpiWarps) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;          // row inside the 128-row tile
    const int sub = (warp - 2) >> 2;        // kEpiWarps == 8: which of the two warps of this quarter
    typename P::Epilogue epi(prm, extra, row, warp - 2);
    int buf = 0; uint32_t bphase = 0;
    if constexpr (P::kBPair) {
      // warps of group `sub` own accumulator `sub` = N tile 2*n_tile + sub of every unit
      for (int it = 0; it < n_iters; ++it) {
        const Unit un0 = prm.unit((int)blockIdx.x, (int)gridDim.x, it);
        const Unit un = {un0.m_tile, 2 * un0.n_tile + sub, un0.z};
        mbar_wait(&ctl->tfull[sub], bphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t)(sub * BN) + ((uint32_t)(q * 32) << 16);
        epi.begin(un);
        // the drain is not hidden behind the next unit's MMAs here: keep the next chunk's TMEM load in flight
        float v[2][32];
        tmem_ld32(taddr, v[0]);
#pragma unroll
        for (int c = 0; c < PAIR_BN_MAX / 32; ++c) {
          if (c * 32 < BN) {
            tmem_ld_wait();
            if ((c + 1) * 32 < BN) tmem_ld32(taddr + (uint32_t)((c + 1) * 32), v[(c + 1) & 1]);
            epi.chunk(un, c * 32, v[c & 1]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->tempty[sub]);
        epi.end(un);
        bphase ^= 1;
      }
    } else if constexpr (P::kEpiPrefetch) {
      // eight warps, chunk (2*ci + sub) of the accumulator belongs to warp group `sub`: the policy issues its
      // global loads for the whole unit before the accumulator is complete (they overlap the unit's MMAs),
      // and the TMEM load of the next chunk stays in flight behind the current one
      static_assert(!P::kEpiPrefetch || (P::kEpiWarps == 8 && P::kATiles == 1), "kEpiPrefetch layout");
      for (int it = 0; it < n_iters; ++it) {
        const Unit un = prm.unit((int)blockIdx.x, (int)gridDim.x, it);
        epi.prefetch(un);
        mbar_wait(&ctl->tfull[buf], bphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t)(buf * ACC_COLS) + ((uint32_t)(q * 32) << 16);
        epi.begin(un);
        float v[2][32];
        if (sub * 32 < ACC_COLS) tmem_ld32(taddr + (uint32_t)(sub * 32), v[0]);
#pragma unroll
        for (int ci = 0; ci < MAX_BN / 64; ++ci) {
          const int c0 = (2 * ci + sub) * 32;
          if (c0 < ACC_COLS) {
            tmem_ld_wait();
            if (c0 + 64 < ACC_COLS) tmem_ld32(taddr + (uint32_t)(c0 + 64), v[(ci + 1) & 1]);
            epi.chunk_i(un, ci, c0, v[ci & 1]);   // ci is a constant after unrolling
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->tempty[buf]);
        epi.end(un);
        if (P::kAccBufs == 2) { buf ^= 1; if (buf == 0) bphase ^= 1; } else { bphase ^= 1; }
      }
    } else
    for (int it = 0; it < n_iters; ++it) {
      const Unit un = prm.unit((int)blockIdx.x, (int)gridDim.x, it);
      mbar_wait(&ctl->tfull[buf], bphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(buf * ACC_COLS) + ((uint32_t)(q * 32) << 16);
      epi.begin(un);
      for (int c0 = 0; c0 < ACC_COLS; c0 += 32) {
        if (P::kEpiWarps == 8 && ((c0 >> 5) & 1) != sub) continue;
        float v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
        epi.chunk(un, c0, v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->tempty[buf]);   // accumulator buffer may be overwritten
      epi.end(un);
      if (P::kAccBufs == 2) { buf ^= 1; if (buf == 0) bphase ^= 1; } else { bphase ^= 1; }
    }
    epi.finish();
  } else {
    // ------------------------------------------------------------------ A synthesis (kSynthA only)
    if constexpr (P::kSynthA) {
      // 256 producer threads: row = TMEM lane quarter of this warp * 32 + lane; grp = which group of four warps.
      // Normal mode: both groups build one stage together (grp = which half of the 64-wide row).
      // kSynthAlternate: the groups take alternate stages (a thread builds its whole row), so the latency
      // chain of one stage (LDS -> math -> st -> wait) overlaps with the other group's stage.
      const int grp = (warp - 2 - P::kEpiWarps) >> 2;
      const int row = (warp & 3) * 32 + lane;
      const int t = (grp << 7) + row;
      typename P::SynthState sst;            // per-thread producer state that lives across stages
      int stage = 0; uint32_t phase = 0; int seq = 0;
      for (int it = 0; it < n_iters; ++it) {
        const Unit un = prm.unit((int)blockIdx.x, (int)gridDim.x, it);
        const int KC = prm.k_chunks(un);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        prm.synth_begin(un, extra_synth, t, sst);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int kc = 0; kc < KC; ++kc, ++seq) {
          if (!P::kSynthAlternate || (seq & 1) == grp) {
            mbar_wait(&ctl->empty[stage], phase ^ 1);
            if constexpr (P::kATmem) {
              const uint32_t ta = tmem_base + (uint32_t)(A_TMEM_BASE + stage * A_TMEM_COLS) + ((uint32_t)((warp & 3) * 32) << 16);
              if constexpr (P::kSynthAlternate) {
                prm.synth_a_tmem(ta, un, kc, row, extra_synth, sst);
                prm.synth_a_tmem(ta, un, kc, 128 + row, extra_synth, sst);
              } else {
                prm.synth_a_tmem(ta, un, kc, t, extra_synth, sst);
              }
              tmem_st_wait();
              tc_fence_before();
            } else {
              prm.synth_a(sA + stage * A_BYTES, un, kc, t, extra_synth, sst);
              fence_proxy_async_smem();        // generic-proxy stores -> visible to the tensor core
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl->full[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, ncols);
}

}  // namespace tc
}  // namespace cffm
