// Persistent, warp-specialised bf16 GEMM for sm_100a.
//
//   C[M,N] = epilogue( A_op[M,K] * B_op[N,K]^T ),  fp32 accumulation in TMEM.
//
// One CTA per SM, 320 threads:
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, 4 stages x 48 KB)
//   warp 1      : MMA issuer     (one lane issues tcgen05.mma 128x256x16; tcgen05.commit frees smem stages)
//   warps 2..9  : epilogue       (tcgen05.ld -> registers -> smem transpose -> fused, coalesced 128-bit global I/O)
// The 512 TMEM columns hold two 128x256 fp32 accumulators, so the epilogue of tile i overlaps the MMAs of
// tile i+1.  Operands may be K-major (row-major [rows][K]) or MN-major ([K][rows], e.g. activations in a weight
// gradient), which lets forward, dgrad and wgrad all read the tensors PyTorch already holds, with no transposes.
//
// Replaces: every nn.Linear forward/backward on the LaViLa dual-encoder hot path
// (lavila/models/timesformer.py:53-56,110,142; lavila/models/openai_model.py:186-192; models.py:146,160).
#include <mutex>

#include "../../include/lavila_b200.h"
#include "host_common.h"
#include "gemm_epilogue.cuh"
#include "ptx.cuh"

namespace lv {

namespace gemm {
constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle atom
constexpr int STAGES = 4;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int ATOM_BYTES = 64 * BK * 2;  // one 64(MN) x 64(K) MN-major sub-tile = 8 KB
constexpr int EPI_WARPS = 8;  // two warps per TMEM lane quarter, each owning half of the 256 columns
constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_PITCH * 4;
constexpr int NUM_BARS = 2 * STAGES + 4;
constexpr int SMEM_BYTES = 1024 /*alignment slack*/ + STAGES * STAGE_BYTES + EPI_BYTES + NUM_BARS * 8 + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int TMEM_COLS = 512;



template <int A_MN, int B_MN, int CT_FLAGS>  // CT_FLAGS < 0: epilogue flags are read at run time
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args g) {
  extern __shared__ uint8_t smem_raw[];
  // align by OFFSET arithmetic on the shared array so the compiler keeps the shared address space (LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  float* sEpi = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_tiles = g.num_m_blks * g.num_n_blks;
  const int total_items = n_tiles * g.k_splits;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int split = item / n_tiles;
        const int tile = item - split * n_tiles;
        const int m_blk = tile / g.num_n_blks;
        const int n_blk = tile - m_blk * g.num_n_blks;
        const int kb0 = split * g.kb_per_split;
        const int kb1 = min(g.num_kb, kb0 + g.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          const int k0 = kb * BK;
          uint8_t* a_dst = sA + stage * A_STAGE_BYTES;
          uint8_t* b_dst = sB + stage * B_STAGE_BYTES;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(a_dst + j * ATOM_BYTES, &tmA, &full_bar[stage], m_blk * BM + j * 64, k0);
          } else {
            tma_load_2d(a_dst, &tmA, &full_bar[stage], k0, m_blk * BM);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(b_dst + j * ATOM_BYTES, &tmB, &full_bar[stage], n_blk * BN + j * 64, k0);
          } else {
            tma_load_2d(b_dst, &tmB, &full_bar[stage], k0, n_blk * BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
        const int split = item / n_tiles;
        const int kb0 = split * g.kb_per_split;
        const int kb1 = min(g.num_kb, kb0 + g.kb_per_split);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + stage * A_STAGE_BYTES);
          const uint32_t b_base = smem_u32(sB + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = A_MN ? make_smem_desc_sw128(a_base + k * UMMA_K * 128, ATOM_BYTES, 1024)
                                        : make_smem_desc_sw128(a_base + k * UMMA_K * 2, 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc_sw128(b_base + k * UMMA_K * 128, ATOM_BYTES, 1024)
                                        : make_smem_desc_sw128(b_base + k * UMMA_K * 2, 16, 1024);
            tc_mma_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);  // smem stage reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull_bar[as]);  // accumulator complete
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;        // TMEM lane quarter this warp may access
    const int e = warp - 2;        // 0..7
    const int half = e >> 2;       // which 128 of the 256 accumulator columns
    float* buf = sEpi + e * 32 * EPI_PITCH;
    const int flags = CT_FLAGS >= 0 ? CT_FLAGS : g.flags;
    float scale = 1.0f;
    if (flags & LV_EPI_SCALE) {
      scale = __ldg(g.scale_ptr);
      if (flags & LV_EPI_SCALE_TANH) scale = tanhf(scale);
    }
    int it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int split = item / n_tiles;
      const int tile = item - split * n_tiles;
      const int m_blk = tile / g.num_n_blks;
      const int n_blk = tile - m_blk * g.num_n_blks;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      epilogue_tile<CT_FLAGS>(g, flags, scale, buf, &tfull_bar[as], aphase, tmem_base + as * BN, m_blk * BM + q * 32, n_blk * BN,
                              half, q, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int A_MN, int B_MN, int CT_FLAGS>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const Args& g, cudaStream_t stream) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, []() {
    attr_err = cudaFuncSetAttribute(gemm_bf16_kernel<A_MN, B_MN, CT_FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return set_error((int)attr_err, "gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  const int total = g.num_m_blks * g.num_n_blks * g.k_splits;
  const int grid = total < sm_count() ? total : sm_count();
  gemm_bf16_kernel<A_MN, B_MN, CT_FLAGS><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmA, tmB, g);
  return check_launch("lv_gemm_bf16");
}

template <int A_MN, int B_MN, int F>
static int try_spec(bool& hit, const CUtensorMap& tmA, const CUtensorMap& tmB, const Args& g, cudaStream_t st) {
  if constexpr (is_specialised(A_MN, B_MN, F)) {
    if (!hit && g.flags == F) { hit = true; return launch<A_MN, B_MN, F>(tmA, tmB, g, st); }
  }
  return 0;
}

template <int A_MN, int B_MN>
static int dispatch(const CUtensorMap& tmA, const CUtensorMap& tmB, const Args& g, cudaStream_t st) {
  bool hit = false;
  int rc = 0;
#define LV_TRY(F) if (!hit) rc = try_spec<A_MN, B_MN, (F)>(hit, tmA, tmB, g, st);
  LV_TRY(0)
  LV_TRY(LV_EPI_BIAS)
  LV_TRY(LV_EPI_BIAS | LV_EPI_QUICKGELU)
  LV_TRY(LV_EPI_OUT_F32)
  LV_TRY(LV_EPI_BIAS | LV_EPI_RESID | LV_EPI_OUT_F32)
  LV_TRY(LV_EPI_BIAS | LV_EPI_SCALE | LV_EPI_SCALE_TANH | LV_EPI_RESID | LV_EPI_OUT_F32)
  LV_TRY(LV_EPI_DQUICKGELU)
  LV_TRY(LV_EPI_ATOMIC | LV_EPI_OUT_F32)
#undef LV_TRY
  if (hit) return rc;
  return launch<A_MN, B_MN, -1>(tmA, tmB, g, st);
}

}  // namespace gemm
}  // namespace lv

extern "C" int lv_gemm_bf16(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, int64_t M,
                            int64_t N, int64_t K, int k_splits, const LvGemmEpilogue* epi, void* stream) {
  using namespace lv;
  using namespace lv::gemm;
  LV_REQUIRE(A && B && epi && epi->out, "lv_gemm_bf16: null pointer");
  LV_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "lv_gemm_bf16: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  LV_REQUIRE((N & 3) == 0 && (epi->ldo & 3) == 0, "lv_gemm_bf16: N and ldo must be multiples of 4");
  const int flags = epi->flags;
  LV_REQUIRE(!(flags & LV_EPI_BIAS) || epi->bias, "lv_gemm_bf16: LV_EPI_BIAS without bias");
  LV_REQUIRE(!(flags & LV_EPI_RESID) || (epi->resid && (epi->ldr & 3) == 0), "lv_gemm_bf16: LV_EPI_RESID without resid / odd ldr");
  LV_REQUIRE(!(flags & (LV_EPI_QUICKGELU | LV_EPI_COPY_BF16)) || (epi->out2 && (epi->ldo2 & 3) == 0), "lv_gemm_bf16: out2 required");
  LV_REQUIRE(!((flags & LV_EPI_QUICKGELU) && (flags & LV_EPI_COPY_BF16)), "lv_gemm_bf16: QUICKGELU and COPY_BF16 both use out2");
  LV_REQUIRE(!(flags & LV_EPI_DQUICKGELU) || (epi->aux && (epi->ldaux & 3) == 0), "lv_gemm_bf16: aux required");
  LV_REQUIRE(!(flags & LV_EPI_SCALE) || epi->scale_ptr, "lv_gemm_bf16: scale_ptr required");
  if (k_splits < 1) k_splits = 1;
  LV_REQUIRE(k_splits == 1 || (flags & LV_EPI_ATOMIC), "lv_gemm_bf16: k_splits > 1 requires LV_EPI_ATOMIC");

  CUtensorMap tmA, tmB;
  int rc;
  if (a_mn) rc = make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK);
  else      rc = make_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM);
  if (rc) return rc;
  if (b_mn) rc = make_tmap_2d_bf16(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK);
  else      rc = make_tmap_2d_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, BN);
  if (rc) return rc;

  Args g;
  g.M = (int)M; g.N = (int)N; g.K = (int)K;
  g.num_m_blks = (int)((M + BM - 1) / BM);
  g.num_n_blks = (int)((N + BN - 1) / BN);
  g.num_kb = (int)((K + BK - 1) / BK);
  if (k_splits > g.num_kb) k_splits = g.num_kb;
  g.kb_per_split = (g.num_kb + k_splits - 1) / k_splits;
  g.k_splits = (g.num_kb + g.kb_per_split - 1) / g.kb_per_split;
  g.sk_total_kb = 0;
  g.sk_kb_per_cta = 0;
  g.flags = flags | ((flags & LV_EPI_ATOMIC) ? LV_EPI_OUT_F32 : 0);
  g.out = epi->out; g.ldo = epi->ldo;
  g.out2 = epi->out2; g.ldo2 = epi->ldo2;
  g.bias = epi->bias;
  g.resid = epi->resid; g.ldr = epi->ldr;
  g.aux = reinterpret_cast<const __nv_bfloat16*>(epi->aux); g.ldaux = epi->ldaux;
  g.scale_ptr = epi->scale_ptr;

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!a_mn && !b_mn) return dispatch<0, 0>(tmA, tmB, g, st);
  if (!a_mn && b_mn) return dispatch<0, 1>(tmA, tmB, g, st);
  if (a_mn && b_mn) return dispatch<1, 1>(tmA, tmB, g, st);
  return dispatch<1, 0>(tmA, tmB, g, st);
}
