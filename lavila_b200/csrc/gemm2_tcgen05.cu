// 2-CTA ("cta_group::2") variant of the persistent bf16 GEMM: a cluster of two CTAs on one TPC computes a 256 x 256
// output tile.  Each CTA TMA-loads its own 128 rows of A and HALF of the B tile (128 of the 256 N rows); one
// tcgen05.mma.cta_group::2 issued by the leader CTA drives both SMs' tensor cores (M = 256), each reading the other
// CTA's B half through the pair.  Per SM and k-block this moves 32 KB into shared memory and reads 32 KB out of it
// (1-CTA 128x256 tiles: 48 + 48 KB) -- the shared-memory pipe, not the tensor pipe, bounds the 1-CTA kernel
// (profiles/ncu_r01_fwd_b16_summary.txt).
//
// Synchronisation (per CTA unless noted):
//   full[s]   (leader only) : leader arrives with expect_tx(2 x 32 KB); BOTH CTAs' TMA loads complete_tx on it
//                             (barrier address with the peer bit cleared, `.cta_group::2` TMA)
//   empty[s]                : tcgen05.commit.cta_group::2 ... multicast -> arrives in both CTAs when the MMAs that read
//                             stage s have retired
//   tfull[a]                : multicast commit after the last k-block of a tile
//   tempty[a] (leader only) : the 2 x 8 epilogue warps of the pair arrive remotely (mapa + mbarrier.arrive.shared::cluster)
// Same operand layouts, epilogues and C ABI as gemm_tcgen05.cu.
#include <mutex>

#include "../../include/lavila_b200.h"
#include "gemm_epilogue.cuh"
#include "gemm_epilogue_rows.cuh"
#include <cstdlib>
#include "host_common.h"
#include "ptx.cuh"

namespace lv {
namespace gemm2 {

using gemm::Args;
using gemm::BN;
using gemm::EPI_PITCH;

constexpr int BM_CTA = 128;   // rows per CTA; the pair covers 256
constexpr int BN_CTA = 128;   // B rows (N) loaded per CTA
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM_CTA * BK * 2;   // 16 KB
constexpr int B_STAGE_BYTES = BN_CTA * BK * 2;   // 16 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int ATOM_BYTES = 64 * BK * 2;
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int TMEM_COLS = 512;
// Epilogue modes (compile time):
//   0  transposing epilogue of gemm_epilogue.cuh (per-thread global loads / stores; every flag set), 6 operand stages
//   1  row-per-thread epilogue with TMA stores (gemm_epilogue_rows.cuh), output-only flag sets, 5 stages + 32 KB staging
//   2  the same + TMA prefetch of the saved pre-activation slab (dQuickGELU), 4 stages + 32 KB staging + 64 KB slabs
//   3  fp32 output + fp32 residual (proj / fc2 forward): residual boxes in by TMA, added in place, out by TMA; 4 stages + 64 KB
__host__ __device__ constexpr int stages_of(int mode) { return mode == 0 ? 6 : (mode == 1 ? 5 : 4); }
__host__ __device__ constexpr int epi_bytes_of(int mode) {
  return mode == 0 ? EPI_WARPS * 32 * EPI_PITCH * 4
       : mode == 3 ? EPI_WARPS * gemm::ROWS_F32_BYTES
                   : EPI_WARPS * (gemm::ROWS_STAGE_BYTES + (mode == 2 ? gemm::ROWS_AUX_BYTES : 0));
}
__host__ __device__ constexpr int num_bars_of(int mode) {
  return 2 * stages_of(mode) + 4 + (mode == 2 ? EPI_WARPS : 0) + (mode == 3 ? 2 * EPI_WARPS : 0);
}
__host__ __device__ constexpr int smem_bytes_of(int mode) {
  return 1024 + stages_of(mode) * STAGE_BYTES + epi_bytes_of(mode) + num_bars_of(mode) * 8 + 16;
}
static_assert(smem_bytes_of(0) <= 227 * 1024 && smem_bytes_of(1) <= 227 * 1024 && smem_bytes_of(2) <= 227 * 1024 &&
              smem_bytes_of(3) <= 227 * 1024, "shared memory budget");
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load executed by either CTA of the pair; completes transaction bytes on the LEADER's barrier.
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the MMAs issued so far have retired) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}

template <int A_MN, int B_MN, int CT_FLAGS, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2,
                  const __grid_constant__ CUtensorMap tmAux, const Args g) {
  constexpr int STAGES = stages_of(MODE);
  constexpr int EPI_BYTES = epi_bytes_of(MODE);
  constexpr int NUM_BARS = num_bars_of(MODE);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
  uint8_t* sEpiRaw = smem + STAGES * STAGE_BYTES;            // 1024-byte aligned (stage sizes are multiples of 1024)
  float* sEpi = reinterpret_cast<float*>(sEpiRaw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint64_t* aux_bar = bars + 2 * STAGES + 4;                 // [EPI_WARPS], MODE 2 only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (MODE >= 1) {
      prefetch_tmap(&tmO);
      if (CT_FLAGS & LV_EPI_QUICKGELU) prefetch_tmap(&tmO2);
    }
    if (MODE == 2) {
      prefetch_tmap(&tmAux);
      for (int w = 0; w < EPI_WARPS; ++w) mbar_init(&aux_bar[w], 1);
    }
    if (MODE == 3) {
      prefetch_tmap(&tmAux);     // the residual's tensor map travels in the aux slot
      for (int w = 0; w < 2 * EPI_WARPS; ++w) mbar_init(&aux_bar[w], 1);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 2 * EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc2(tmem_slot, TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of both CTAs initialised, TMEM of both allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;     // num_m_blks counts 256-row tiles

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      gemm::ItemIter iter(g, cluster_id, num_clusters);
      int tile, kb0, kb1;
      while (iter.next(tile, kb0, kb1)) {
        const int m_blk = tile / g.num_n_blks;
        const int n_blk = tile - m_blk * g.num_n_blks;
        const int m0 = m_blk * 256 + rank * BM_CTA;
        const int n0 = n_blk * BN + rank * BN_CTA;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          const int k0 = kb * BK;
          uint8_t* a_dst = sA + stage * A_STAGE_BYTES;
          uint8_t* b_dst = sB + stage * B_STAGE_BYTES;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM_CTA / 64; ++j) tma_load_2d_2sm(a_dst + j * ATOM_BYTES, &tmA, &full_bar[stage], m0 + j * 64, k0);
          } else {
            tma_load_2d_2sm(a_dst, &tmA, &full_bar[stage], k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN_CTA / 64; ++j) tma_load_2d_2sm(b_dst + j * ATOM_BYTES, &tmB, &full_bar[stage], n0 + j * 64, k0);
          } else {
            tma_load_2d_2sm(b_dst, &tmB, &full_bar[stage], k0, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      gemm::ItemIter iter(g, cluster_id, num_clusters);
      int tile, kb0, kb1;
      for (int it = 0; iter.next(tile, kb0, kb1); ++it) {
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
            tc_mma2_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit2(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit2(&tfull_bar[as]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (both CTAs, own 128 rows)
    const int q = warp & 3;
    const int e = warp - 2;
    const int half = e >> 2;
    const int flags = CT_FLAGS >= 0 ? CT_FLAGS : g.flags;
    gemm::ItemIter iter(g, cluster_id, num_clusters);
    int tile, kb0, kb1;
    if constexpr (MODE == 0) {
      float* buf = sEpi + e * 32 * EPI_PITCH;
      float scale = 1.0f;
      if (flags & LV_EPI_SCALE) {
        scale = __ldg(g.scale_ptr);
        if (flags & LV_EPI_SCALE_TANH) scale = tanhf(scale);
      }
      for (int it = 0; iter.next(tile, kb0, kb1); ++it) {
        const int m_blk = tile / g.num_n_blks;
        const int n_blk = tile - m_blk * g.num_n_blks;
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        gemm::epilogue_tile<CT_FLAGS>(g, flags, scale, buf, &tfull_bar[as], aphase, tmem_base + as * BN,
                                      m_blk * 256 + (int)rank * BM_CTA + q * 32, n_blk * BN, half, q, lane);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(&tempty_bar[as], 0);   // the leader's MMA warp owns the accumulator hand-off
      }
    } else if constexpr (MODE == 3) {
      uint8_t* stage = sEpiRaw + e * gemm::ROWS_F32_BYTES;
      uint32_t ph[2] = {0u, 0u};
      float scale = 1.0f;
      if (flags & LV_EPI_SCALE) {
        scale = __ldg(g.scale_ptr);
        if (flags & LV_EPI_SCALE_TANH) scale = tanhf(scale);
      }
      for (int it = 0; iter.next(tile, kb0, kb1); ++it) {
        const int m_blk = tile / g.num_n_blks;
        const int n_blk = tile - m_blk * g.num_n_blks;
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        uint64_t* tempty = &tempty_bar[as];
        gemm::epilogue_rows_f32_tile<CT_FLAGS>(g, scale, stage, &aux_bar[2 * e], ph, &tmO, &tmAux, &tfull_bar[as], aphase,
                                               tmem_base + as * BN, m_blk * 256 + (int)rank * BM_CTA + q * 32, n_blk * BN, half, q,
                                               lane, [&]() { if (lane == 0) mbar_arrive_remote(tempty, 0); });
      }
      if (lane == 0) tma_store_wait_all();
    } else {
      uint8_t* stage = sEpiRaw + e * gemm::ROWS_STAGE_BYTES;
      uint8_t* aux = sEpiRaw + EPI_WARPS * gemm::ROWS_STAGE_BYTES + e * gemm::ROWS_AUX_BYTES;
      // the pre-activation slab of a tile: 4 boxes of [32 rows x 32 columns] bf16 for this warp's rows / column half
      auto request_aux = [&](int t) {
        const int m_blk = t / g.num_n_blks;
        const int n_blk = t - m_blk * g.num_n_blks;
        const int m_base = m_blk * 256 + (int)rank * BM_CTA + q * 32;
        const int nb = n_blk * BN + half * 128;
        int boxes = 0;
        for (int cc = 0; cc < 4; ++cc) boxes += (nb + cc * 32 < g.N) ? 1 : 0;
        if (boxes == 0) { mbar_arrive(&aux_bar[e]); return; }
        mbar_arrive_expect_tx(&aux_bar[e], boxes * 2048);
        for (int cc = 0; cc < boxes; ++cc) tma_load_2d(aux + cc * 2048, &tmAux, &aux_bar[e], nb + cc * 32, m_base);
      };
      bool have = iter.next(tile, kb0, kb1);
      if (MODE == 2 && have && lane == 0) request_aux(tile);
      for (int it = 0; have; ++it) {
        const int m_blk = tile / g.num_n_blks;
        const int n_blk = tile - m_blk * g.num_n_blks;
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        uint64_t* tempty = &tempty_bar[as];
        gemm::epilogue_rows_tile<CT_FLAGS>(g, stage, aux, &aux_bar[e], (uint32_t)(it & 1), &tmO, &tmO2, &tfull_bar[as], aphase,
                                           tmem_base + as * BN, m_blk * 256 + (int)rank * BM_CTA + q * 32, n_blk * BN, half, q,
                                           lane, [&]() { if (lane == 0) mbar_arrive_remote(tempty, 0); });
        have = iter.next(tile, kb0, kb1);
        if (MODE == 2) {
          __syncwarp();                                       // every lane has finished reading this tile's slab
          if (have && lane == 0) request_aux(tile);
        }
      }
      if (lane == 0) tma_store_wait_all();                    // the staging tiles must outlive the stores that read them
    }
  }

  tc_fence_before();
  cluster_sync_all();   // neither CTA may exit (or free TMEM) while its peer can still touch its smem / barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, TMEM_COLS);
  }
}

struct Maps {
  CUtensorMap A, B, O, O2, Aux;
};

template <int A_MN, int B_MN, int CT_FLAGS, int MODE>
static int launch(const Maps& tm, const Args& g, cudaStream_t stream) {
  constexpr int SMEM = smem_bytes_of(MODE);
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, []() {
    attr_err = cudaFuncSetAttribute(gemm2_bf16_kernel<A_MN, B_MN, CT_FLAGS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  });
  if (attr_err != cudaSuccess) return set_error((int)attr_err, "gemm2: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  const int total = g.num_m_blks * g.num_n_blks * g.k_splits;
  int clusters = sm_count() / 2;
  if (g.sk_kb_per_cta > 0) clusters = (int)((g.sk_total_kb + g.sk_kb_per_cta - 1) / g.sk_kb_per_cta);
  else if (total < clusters) clusters = total;
  gemm2_bf16_kernel<A_MN, B_MN, CT_FLAGS, MODE><<<2 * clusters, NUM_THREADS, SMEM, stream>>>(tm.A, tm.B, tm.O, tm.O2, tm.Aux, g);
  return check_launch("lv_gemm_bf16 (2-CTA)");
}

// rows_mode: 0 = transposing epilogue; 1 / 2 = row-per-thread TMA-store epilogue available for this call (maps built)
template <int A_MN, int B_MN, int F>
static int try_spec(bool& hit, const Maps& tm, const Args& g, int rows_mode, cudaStream_t st) {
  if constexpr (gemm::is_specialised(A_MN, B_MN, F)) {
    if (!hit && g.flags == F) {
      hit = true;
      if constexpr (gemm::rows_supported(F) && !A_MN) {
        constexpr int MODE = gemm::rows_needs_aux(F) ? 2 : 1;
        if (rows_mode == MODE) return launch<A_MN, B_MN, F, MODE>(tm, g, st);
      }
      if constexpr (gemm::rows_f32_resid(F) && !A_MN && !B_MN) {
        if (rows_mode == 3) return launch<A_MN, B_MN, F, 3>(tm, g, st);
      }
      return launch<A_MN, B_MN, F, 0>(tm, g, st);
    }
  }
  return 0;
}

template <int A_MN, int B_MN>
static int dispatch(const Maps& tm, const Args& g, int rows_mode, cudaStream_t st) {
  bool hit = false;
  int rc = 0;
#define LV_TRY(F) if (!hit) rc = try_spec<A_MN, B_MN, (F)>(hit, tm, g, rows_mode, st);
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
  return launch<A_MN, B_MN, -1, 0>(tm, g, st);
}

// LAVILA_B200_GEMM_ROWS_EPI=0 keeps the transposing epilogue everywhere (A/B runs)
static bool rows_epilogue_enabled() {
  static const bool on = []() { const char* e = std::getenv("LAVILA_B200_GEMM_ROWS_EPI"); return !(e && e[0] == '0'); }();
  return on;
}

}  // namespace gemm2
}  // namespace lv

// Same contract as lv_gemm_bf16 (include/lavila_b200.h); 256 x 256 tiles computed by CTA pairs.
extern "C" int lv_gemm_bf16_2cta(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, int64_t M,
                                 int64_t N, int64_t K, int k_splits, const LvGemmEpilogue* epi, void* stream) {
  using namespace lv;
  using namespace lv::gemm2;
  LV_REQUIRE(A && B && epi && epi->out, "lv_gemm_bf16_2cta: null pointer");
  LV_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "lv_gemm_bf16_2cta: bad shape");
  LV_REQUIRE((N & 3) == 0 && (epi->ldo & 3) == 0, "lv_gemm_bf16_2cta: N and ldo must be multiples of 4");
  const int flags = epi->flags;
  LV_REQUIRE(!(flags & LV_EPI_BIAS) || epi->bias, "lv_gemm_bf16_2cta: LV_EPI_BIAS without bias");
  LV_REQUIRE(!(flags & LV_EPI_RESID) || (epi->resid && (epi->ldr & 3) == 0), "lv_gemm_bf16_2cta: bad resid");
  LV_REQUIRE(!(flags & (LV_EPI_QUICKGELU | LV_EPI_COPY_BF16)) || (epi->out2 && (epi->ldo2 & 3) == 0), "lv_gemm_bf16_2cta: out2 required");
  LV_REQUIRE(!(flags & LV_EPI_DQUICKGELU) || (epi->aux && (epi->ldaux & 3) == 0), "lv_gemm_bf16_2cta: aux required");
  LV_REQUIRE(!(flags & LV_EPI_SCALE) || epi->scale_ptr, "lv_gemm_bf16_2cta: scale_ptr required");
  if (k_splits < 1) k_splits = 1;
  LV_REQUIRE(k_splits == 1 || (flags & LV_EPI_ATOMIC), "lv_gemm_bf16_2cta: k_splits > 1 requires LV_EPI_ATOMIC");

  Maps tm{};
  int rc;
  if (a_mn) rc = make_tmap_2d_bf16(&tm.A, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK);
  else      rc = make_tmap_2d_bf16(&tm.A, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM_CTA);
  if (rc) return rc;
  if (b_mn) rc = make_tmap_2d_bf16(&tm.B, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK);
  else      rc = make_tmap_2d_bf16(&tm.B, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, BN_CTA);
  if (rc) return rc;
  // Row-per-thread TMA-store epilogue: output-only flag sets (and dQuickGELU) with bf16 outputs whose base / pitch TMA accepts
  int rows_mode = 0;
  auto tma_ok = [](const void* p, long long ld_elems) { return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld_elems & 7) == 0; };
  if (rows_epilogue_enabled() && !a_mn && gemm::rows_supported(flags) && tma_ok(epi->out, epi->ldo) &&
      (!(flags & LV_EPI_QUICKGELU) || tma_ok(epi->out2, epi->ldo2)) && (!(flags & LV_EPI_DQUICKGELU) || tma_ok(epi->aux, epi->ldaux))) {
    rows_mode = gemm::rows_needs_aux(flags) ? 2 : 1;
    rc = make_tmap_2d(&tm.O, epi->out, 2, (uint64_t)N, (uint64_t)M, (uint64_t)epi->ldo, 32, 32, 64);
    if (!rc && (flags & LV_EPI_QUICKGELU)) rc = make_tmap_2d(&tm.O2, epi->out2, 2, (uint64_t)N, (uint64_t)M, (uint64_t)epi->ldo2, 32, 32, 64);
    if (!rc && (flags & LV_EPI_DQUICKGELU)) rc = make_tmap_2d(&tm.Aux, epi->aux, 2, (uint64_t)N, (uint64_t)M, (uint64_t)epi->ldaux, 32, 32, 64);
    if (rc) return rc;
  } else if (rows_epilogue_enabled() && !a_mn && !b_mn && gemm::rows_f32_resid(flags) && K < 2048 &&   // deep K keeps 6 operand stages
             (reinterpret_cast<uintptr_t>(epi->out) & 15) == 0 &&
             (epi->ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(epi->resid) & 15) == 0 && (epi->ldr & 3) == 0) {
    rows_mode = 3;
    rc = make_tmap_2d(&tm.O, epi->out, 4, (uint64_t)N, (uint64_t)M, (uint64_t)epi->ldo, 32, 32, 128);
    if (!rc) rc = make_tmap_2d(&tm.Aux, epi->resid, 4, (uint64_t)N, (uint64_t)M, (uint64_t)epi->ldr, 32, 32, 128);
    if (rc) return rc;
  }

  gemm::Args g;
  g.M = (int)M; g.N = (int)N; g.K = (int)K;
  g.num_m_blks = (int)((M + 255) / 256);
  g.num_n_blks = (int)((N + BN - 1) / BN);
  g.num_kb = (int)((K + BK - 1) / BK);
  if (k_splits > g.num_kb) k_splits = g.num_kb;
  g.kb_per_split = (g.num_kb + k_splits - 1) / k_splits;
  g.k_splits = (g.num_kb + g.kb_per_split - 1) / g.kb_per_split;
  g.sk_total_kb = 0;
  g.sk_kb_per_cta = 0;
  // stream-K (gemm::ItemIter) is implemented but NOT enabled: measured on the TSF-B weight gradients it is 10 % slower than
  // classic split-K, because CTA pairs working on the same tile row at different k offsets no longer share operand panels
  // in L2 (DRAM traffic x4.5).  Wave quantisation is handled on the host instead (ops.wgrad_splits picks a split count whose
  // item count fills whole waves).
  g.flags = flags | ((flags & LV_EPI_ATOMIC) ? LV_EPI_OUT_F32 : 0);
  g.out = epi->out; g.ldo = epi->ldo;
  g.out2 = epi->out2; g.ldo2 = epi->ldo2;
  g.bias = epi->bias;
  g.resid = epi->resid; g.ldr = epi->ldr;
  g.aux = reinterpret_cast<const __nv_bfloat16*>(epi->aux); g.ldaux = epi->ldaux;
  g.scale_ptr = epi->scale_ptr;

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!a_mn && !b_mn) return dispatch<0, 0>(tm, g, rows_mode, st);
  if (!a_mn && b_mn) return dispatch<0, 1>(tm, g, rows_mode, st);
  if (a_mn && b_mn) return dispatch<1, 1>(tm, g, rows_mode, st);
  return dispatch<1, 0>(tm, g, rows_mode, st);
}
