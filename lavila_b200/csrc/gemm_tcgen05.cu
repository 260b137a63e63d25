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
#include "ptx.cuh"

namespace lv {

namespace gemm {
constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle atom
constexpr int STAGES = 4;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int ATOM_BYTES = 64 * BK * 2;  // one 64(MN) x 64(K) MN-major sub-tile = 8 KB
constexpr int EPI_PITCH = 16;  // floats: a 32-row x 16-column half chunk per warp, 16-byte units XOR-swizzled
constexpr int EPI_WARPS = 8;  // two warps per TMEM lane quarter, each owning half of the 256 columns
constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_PITCH * 4;
constexpr int NUM_BARS = 2 * STAGES + 4;
constexpr int SMEM_BYTES = 1024 /*alignment slack*/ + STAGES * STAGE_BYTES + EPI_BYTES + NUM_BARS * 8 + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int TMEM_COLS = 512;

struct Args {
  int M, N, K;
  int num_m_blks, num_n_blks, k_splits, kb_per_split, num_kb;
  int flags;
  void* out;
  long long ldo;
  void* out2;
  long long ldo2;
  const float* bias;
  const float* resid;
  long long ldr;
  const __nv_bfloat16* aux;
  long long ldaux;
  const float* scale_ptr;
};

// sigmoid(x) = 0.5 * tanh(0.5 x) + 0.5 : one SFU op (tanh.approx.f32, rel. error ~2^-11) instead of ex2 + rcp.
// The GELU epilogues are SFU-bound otherwise (128x256 outputs x 2 MUFU / 16 per clk = 4096 cycles per tile).
__device__ __forceinline__ float sigmoidf_fast(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}

// Flag sets with a dedicated compile-time specialisation (everything else runs the runtime-flag kernel).
__host__ __device__ constexpr bool is_specialised(int a_mn, int b_mn, int f) {
  if (!a_mn && !b_mn)
    return f == 0 || f == LV_EPI_BIAS || f == (LV_EPI_BIAS | LV_EPI_QUICKGELU) || f == LV_EPI_OUT_F32 ||
           f == (LV_EPI_BIAS | LV_EPI_RESID | LV_EPI_OUT_F32) ||
           f == (LV_EPI_BIAS | LV_EPI_SCALE | LV_EPI_SCALE_TANH | LV_EPI_RESID | LV_EPI_OUT_F32);
  if (!a_mn && b_mn) return f == 0 || f == LV_EPI_DQUICKGELU || f == LV_EPI_OUT_F32;
  if (a_mn && b_mn) return f == (LV_EPI_ATOMIC | LV_EPI_OUT_F32);
  return false;
}

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
    const int r8 = lane >> 2;        // row within a group of 8
    const int ch = lane & 3;         // this lane's 16-byte unit (4 columns) inside a 16-column half chunk
    int it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int split = item / n_tiles;
      const int tile = item - split * n_tiles;
      const int m_blk = tile / g.num_n_blks;
      const int n_blk = tile - m_blk * g.num_n_blks;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m_base = m_blk * BM + q * 32;
      const int n0 = n_blk * BN;
      constexpr int CPW = BN / 64;  // 32-column chunks per warp (4)
      // ---- everything this warp needs from global memory is requested BEFORE waiting on the accumulator:
      //      the whole tile's bias, and the first chunk's residual / aux operands (then one chunk ahead).
      float4 bias_r[CPW][2];
#pragma unroll
      for (int cc = 0; cc < CPW; ++cc)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int n = n0 + (half * CPW + cc) * 32 + hh * 16 + ch * 4;
          bias_r[cc][hh] = make_float4(0.f, 0.f, 0.f, 0.f);
          if ((flags & LV_EPI_BIAS) && !(flags & LV_EPI_ROWBIAS) && n < g.N)
            bias_r[cc][hh] = __ldg(reinterpret_cast<const float4*>(g.bias + n));
        }
      // residual / aux operands are fetched one 16-column half chunk ahead of their use
      float4 nxt_res[4];
      uint2 nxt_aux[4];
      auto prefetch = [&](int u) {
        if (flags & (LV_EPI_RESID | LV_EPI_DQUICKGELU)) {
          const int n = n0 + (half * CPW + (u >> 1)) * 32 + (u & 1) * 16 + ch * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int m = m_base + 8 * i + r8;
            if (m < g.M && n < g.N) {
              if (flags & LV_EPI_RESID) nxt_res[i] = __ldg(reinterpret_cast<const float4*>(g.resid + (long long)m * g.ldr + n));
              if (flags & LV_EPI_DQUICKGELU) nxt_aux[i] = __ldg(reinterpret_cast<const uint2*>(g.aux + (long long)m * g.ldaux + n));
            }
          }
        }
      };
      prefetch(0);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < CPW; ++cc) {
        const int c = half * CPW + cc;
        const bool chunk_ok = n0 + c * 32 < g.N;   // warp-uniform
        uint32_t r[32];
        if (chunk_ok) {
          tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + as * BN + c * 32, r);
          tmem_ld_wait();
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float4 pre_res[4];
          uint2 pre_aux[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { pre_res[i] = nxt_res[i]; pre_aux[i] = nxt_aux[i]; }
          if (cc * 2 + hh + 1 < 2 * CPW) prefetch(cc * 2 + hh + 1);
          if (!chunk_ok) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(buf + lane * EPI_PITCH + ((j ^ ((lane >> 1) & 3)) << 2)) =
                make_float4(__uint_as_float(r[hh * 16 + 4 * j]), __uint_as_float(r[hh * 16 + 4 * j + 1]),
                            __uint_as_float(r[hh * 16 + 4 * j + 2]), __uint_as_float(r[hh * 16 + 4 * j + 3]));
          __syncwarp();
          const int n = n0 + c * 32 + hh * 16 + ch * 4;
          const bool n_ok = n < g.N;
          float4 bv = bias_r[cc][hh];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row_l = 8 * i + r8;
            const int m = m_base + row_l;
            float4 v = *reinterpret_cast<const float4*>(buf + row_l * EPI_PITCH + ((ch ^ ((row_l >> 1) & 3)) << 2));
            if (m < g.M && n_ok) {
              if (flags & LV_EPI_ROWBIAS) { const float b = __ldg(g.bias + m); bv = make_float4(b, b, b, b); }
              if (flags & LV_EPI_BIAS) { v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w; }
              if (flags & LV_EPI_QUICKGELU) {
                uint2 hb = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.out2) + (long long)m * g.ldo2 + n) = hb;
                const float2 h0 = unpack_bf16x2(hb.x), h1 = unpack_bf16x2(hb.y);
                v.x = h0.x * sigmoidf_fast(1.702f * h0.x);
                v.y = h0.y * sigmoidf_fast(1.702f * h0.y);
                v.z = h1.x * sigmoidf_fast(1.702f * h1.x);
                v.w = h1.y * sigmoidf_fast(1.702f * h1.y);
              }
              if (flags & LV_EPI_GELU_TANH) {
                float* vv[4] = {&v.x, &v.y, &v.z, &v.w};
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                  const float xx = *vv[e2];
                  float th;
                  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.7978845608028654f * (xx + 0.044715f * xx * xx * xx)));
                  *vv[e2] = 0.5f * xx * (1.0f + th);
                }
              }
              if (flags & LV_EPI_SQRELU) {
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                v.x *= v.x; v.y *= v.y; v.z *= v.z; v.w *= v.w;
              }
              if (flags & LV_EPI_DQUICKGELU) {
                const uint2 hb = pre_aux[i];
                const float2 h0 = unpack_bf16x2(hb.x), h1 = unpack_bf16x2(hb.y);
                const float s0 = sigmoidf_fast(1.702f * h0.x), s1 = sigmoidf_fast(1.702f * h0.y);
                const float s2 = sigmoidf_fast(1.702f * h1.x), s3 = sigmoidf_fast(1.702f * h1.y);
                v.x *= s0 * (1.0f + 1.702f * h0.x * (1.0f - s0));
                v.y *= s1 * (1.0f + 1.702f * h0.y * (1.0f - s1));
                v.z *= s2 * (1.0f + 1.702f * h1.x * (1.0f - s2));
                v.w *= s3 * (1.0f + 1.702f * h1.y * (1.0f - s3));
              }
              if (flags & LV_EPI_SCALE) { v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale; }
              if (flags & LV_EPI_RESID) {
                const float4 rr = pre_res[i];
                v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
              }
              if (flags & LV_EPI_ATOMIC) {
                float* o = reinterpret_cast<float*>(g.out) + (long long)m * g.ldo + n;
                red_add_f32(o, v.x); red_add_f32(o + 1, v.y); red_add_f32(o + 2, v.z); red_add_f32(o + 3, v.w);
              } else if (flags & LV_EPI_OUT_F32) {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(g.out) + (long long)m * g.ldo + n) = v;
              } else {
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.out) + (long long)m * g.ldo + n) =
                    make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
              }
              if (flags & LV_EPI_COPY_BF16) {
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(g.out2) + (long long)m * g.ldo2 + n) =
                    make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
              }
            }
          }
          __syncwarp();
        }
      }
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
