// Tensor-core (tcgen05) kernels, part 1: TMA descriptor encoding and the plain bf16 GEMM operator
// C[M,N] (fp32) = A[M,K] . B[N,K]^T used to validate the pipeline (cffm_op_gemm_bf16_dev).
#include <stdio.h>

#include <string>

#include "common.cuh"
#include "tc_kernel.cuh"

namespace cffm {

bool TmaEncoder::init() {
  if (fn) return true;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p ||
      q != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return false;
  }
  fn = reinterpret_cast<EncodeFn>(p);
  return true;
}

bool TmaEncoder::encode_bf16(CUtensorMap* out, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                             const uint32_t* box, bool swizzle128) const {
  if (!fn) return false;
  cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

namespace tc {

// ---- plain GEMM policy ------------------------------------------------------------------------
struct PlainGemm : KMajorA, KMajorB {
  static constexpr bool kSynthA = false;
  __device__ uint32_t idesc() const { return umma_idesc_bf16(BM, BN); }
  static constexpr int kStages = 4, kExtraBytes = 0, kATiles = 1, kAccBufs = 2, kEpiWarps = 4;
  static constexpr bool kATmem = false, kSynthAlternate = false, kBPair = false, kEpiPrefetch = false;
  CUtensorMap mapA, mapB;   // A: dims (K, M) box (64, 128); B: dims (K, N) box (64, BN)
  float* C; int M, N, K, BN, tiles_n;
  __device__ int bn() const { return BN; }
  __device__ int n_units() const { return ((M + BM - 1) / BM) * tiles_n; }
  __device__ int n_iters(int cta, int ncta) const { const int n = n_units(); return cta < n ? (n - cta + ncta - 1) / ncta : 0; }
  __device__ Unit unit(int cta, int ncta, int it) const { const int u = cta + it * ncta; return {u / tiles_n, u % tiles_n, 0}; }
  __device__ int k_chunks(Unit) const { return (K + BK - 1) / BK; }
  __device__ uint32_t tx_bytes() const { return (uint32_t)(A_STAGE_BYTES + BN * BK * 2); }
  __device__ void prefetch() const { prefetch_tmap(&mapA); prefetch_tmap(&mapB); }
  __device__ void load_a(uint8_t* s, uint64_t* bar, Unit un, int kc) const { tma_load_2d(s, &mapA, bar, kc * BK, un.m_tile * BM); }
  __device__ void load_b(uint8_t* s, uint64_t* bar, Unit un, int kc) const { tma_load_2d(s, &mapB, bar, kc * BK, un.n_tile * BN); }
  struct SynthState {};
  __device__ void synth_begin(Unit, uint8_t*, int, SynthState&) const {}
  __device__ void synth_a(uint8_t*, Unit, int, int, const uint8_t*, SynthState&) const {}
  struct Epilogue {
    const PlainGemm& p; int row;
    __device__ Epilogue(const PlainGemm& p_, uint8_t*, int row_, int) : p(p_), row(row_) {}
    __device__ void begin(Unit) {}
    __device__ void chunk(Unit un, int c0, const float (&v)[32]) {
      const int m = un.m_tile * BM + row;
      if (m >= p.M) return;
      float* dst = p.C + (int64_t)m * p.N + un.n_tile * p.BN + c0;
      const int nleft = p.N - (un.n_tile * p.BN + c0);
#pragma unroll
      for (int j = 0; j < 32; ++j) if (j < nleft) dst[j] = v[j];
    }
    __device__ void end(Unit) {}
    __device__ void finish() {}
  };
};

// ---- TN GEMM policy: C[M,N] = sum_r A[r,M] B[r,N] (both operands MN-major) ---------------------
struct PlainGemmTN : MNMajorA, MNMajorB {
  static constexpr bool kSynthA = false;
  static constexpr int kStages = 4, kExtraBytes = 0, kATiles = 1, kAccBufs = 2, kEpiWarps = 4;
  static constexpr bool kATmem = false, kSynthAlternate = false, kBPair = false, kEpiPrefetch = false;
  CUtensorMap mapA, mapB;   // A: dims (M, R) box (64, 64); B: dims (N, R) box (64, 64)
  float* C; int M, N, R, BN, tiles_n;
  __device__ uint32_t idesc() const { return umma_idesc_bf16(BM, BN, true, true); }
  __device__ int bn() const { return BN; }
  __device__ int n_units() const { return ((M + BM - 1) / BM) * tiles_n; }
  __device__ int n_iters(int cta, int ncta) const { const int n = n_units(); return cta < n ? (n - cta + ncta - 1) / ncta : 0; }
  __device__ Unit unit(int cta, int ncta, int it) const { const int u = cta + it * ncta; return {u / tiles_n, u % tiles_n, 0}; }
  __device__ int k_chunks(Unit) const { return (R + BK - 1) / BK; }
  __device__ uint32_t tx_bytes() const { return (uint32_t)((BM + BN) * BK * 2); }
  __device__ void prefetch() const { prefetch_tmap(&mapA); prefetch_tmap(&mapB); }
  __device__ void load_a(uint8_t* s, uint64_t* bar, Unit un, int kc) const {
    for (int i = 0; i < BM / 64; ++i) tma_load_2d(s + i * 8192, &mapA, bar, un.m_tile * BM + i * 64, kc * BK);
  }
  __device__ void load_b(uint8_t* s, uint64_t* bar, Unit un, int kc) const {
    for (int i = 0; i < BN / 64; ++i) tma_load_2d(s + i * 8192, &mapB, bar, un.n_tile * BN + i * 64, kc * BK);
  }
  struct SynthState {};
  __device__ void synth_begin(Unit, uint8_t*, int, SynthState&) const {}
  __device__ void synth_a(uint8_t*, Unit, int, int, const uint8_t*, SynthState&) const {}
  struct Epilogue {
    const PlainGemmTN& p; int row;
    __device__ Epilogue(const PlainGemmTN& p_, uint8_t*, int row_, int) : p(p_), row(row_) {}
    __device__ void begin(Unit) {}
    __device__ void chunk(Unit un, int c0, const float (&v)[32]) {
      const int m = un.m_tile * BM + row;
      if (m >= p.M) return;
      float* dst = p.C + (int64_t)m * p.N + un.n_tile * p.BN + c0;
      const int nleft = p.N - (un.n_tile * p.BN + c0);
#pragma unroll
      for (int j = 0; j < 32; ++j) if (j < nleft) dst[j] = v[j];
    }
    __device__ void end(Unit) {}
    __device__ void finish() {}
  };
};

}  // namespace tc

static TmaEncoder g_enc;
static thread_local std::string g_tc_err;
const char* tc_last_error() { return g_tc_err.c_str(); }

int tc_gemm_bf16(const void* A, const void* B, float* C, int M, int N, int K, cudaStream_t s) {
  using namespace tc;
  if (!g_enc.init()) { g_tc_err = "cuTensorMapEncodeTiled is not available"; return CFFM_ERR_CUDA; }
  if (K % 8 != 0) { g_tc_err = "K must be a multiple of 8 (16-byte rows for TMA)"; return CFFM_ERR_INVALID; }
  PlainGemm p;
  int bn = N >= 256 ? 256 : ((N + 15) / 16) * 16;
  p.BN = bn; p.tiles_n = (N + bn - 1) / bn; p.C = C; p.M = M; p.N = N; p.K = K;
  const uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[1] = {(uint64_t)K * 2};
  const uint32_t bA[2] = {BK, BM};
  const uint64_t dB[2] = {(uint64_t)K, (uint64_t)N}, sB[1] = {(uint64_t)K * 2};
  const uint32_t bB[2] = {BK, (uint32_t)bn};
  if (!g_enc.encode_bf16(&p.mapA, const_cast<void*>(A), 2, dA, sA, bA) ||
      !g_enc.encode_bf16(&p.mapB, const_cast<void*>(B), 2, dB, sB, bB)) {
    g_tc_err = "cuTensorMapEncodeTiled failed";
    return CFFM_ERR_CUDA;
  }
  static PerDeviceOnce attr_once;
  bool& attr_done = attr_once();
  if (!attr_done) {
    if (cudaFuncSetAttribute(k_tc<PlainGemm>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<PlainGemm>()) != cudaSuccess) {
      g_tc_err = "cudaFuncSetAttribute(smem) failed"; return CFFM_ERR_CUDA;
    }
    attr_done = true;
  }
  const int units = ((M + BM - 1) / BM) * p.tiles_n;
  int grid = units < 148 ? units : 148;
  if (grid < 1) grid = 1;
  k_tc<PlainGemm><<<grid, block_threads<PlainGemm>(), smem_bytes<PlainGemm>(), s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_tc_err = cudaGetErrorString(e); return CFFM_ERR_CUDA; }
  return CFFM_OK;
}

int tc_gemm_bf16_tn(const void* A, const void* B, float* C, int M, int N, int R, cudaStream_t s) {
  using namespace tc;
  if (!g_enc.init()) { g_tc_err = "cuTensorMapEncodeTiled is not available"; return CFFM_ERR_CUDA; }
  if (M % 8 != 0 || N % 64 != 0) { g_tc_err = "TN GEMM: M % 8 == 0 and N % 64 == 0 required"; return CFFM_ERR_INVALID; }
  PlainGemmTN p;
  const int bn = N >= 256 ? 256 : N;
  p.BN = bn; p.tiles_n = (N + bn - 1) / bn; p.C = C; p.M = M; p.N = N; p.R = R;
  const uint64_t dA[2] = {(uint64_t)M, (uint64_t)R}, sA[1] = {(uint64_t)M * 2};
  const uint64_t dB[2] = {(uint64_t)N, (uint64_t)R}, sB[1] = {(uint64_t)N * 2};
  const uint32_t bx[2] = {64, 64};
  if (!g_enc.encode_bf16(&p.mapA, const_cast<void*>(A), 2, dA, sA, bx) ||
      !g_enc.encode_bf16(&p.mapB, const_cast<void*>(B), 2, dB, sB, bx)) {
    g_tc_err = "cuTensorMapEncodeTiled failed";
    return CFFM_ERR_CUDA;
  }
  static PerDeviceOnce attr_once;
  bool& attr_done = attr_once();
  if (!attr_done) {
    if (cudaFuncSetAttribute(k_tc<PlainGemmTN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<PlainGemmTN>()) != cudaSuccess) {
      g_tc_err = "cudaFuncSetAttribute(smem) failed"; return CFFM_ERR_CUDA;
    }
    attr_done = true;
  }
  const int units = ((M + BM - 1) / BM) * p.tiles_n;
  int grid = units < 148 ? units : 148;
  k_tc<PlainGemmTN><<<grid, block_threads<PlainGemmTN>(), smem_bytes<PlainGemmTN>(), s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_tc_err = cudaGetErrorString(e); return CFFM_ERR_CUDA; }
  return CFFM_OK;
}

}  // namespace cffm

extern "C" int cffm_op_gemm_bf16_tn_dev(const void* a_dev, const void* b_dev, float* c_dev, int32_t M, int32_t N, int32_t R,
                                        void* stream) {
  if (!a_dev || !b_dev || !c_dev || M < 1 || N < 1 || R < 1) return CFFM_ERR_INVALID;
  return cffm::tc_gemm_bf16_tn(a_dev, b_dev, c_dev, M, N, R, (cudaStream_t)stream);
}

extern "C" int cffm_op_gemm_bf16_dev(const void* a_dev, const void* b_dev, float* c_dev, int32_t M, int32_t N, int32_t K,
                                     void* stream) {
  if (!a_dev || !b_dev || !c_dev || M < 1 || N < 1 || K < 1) return CFFM_ERR_INVALID;
  return cffm::tc_gemm_bf16(a_dev, b_dev, c_dev, M, N, K, (cudaStream_t)stream);
}
extern "C" const char* cffm_tc_last_error(void) { return cffm::tc_last_error(); }
