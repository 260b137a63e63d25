// Skinny GEMM for KV-cached decoding:  C[M, N] = epilogue( A[M, K] x W ),  W stored [K][ldb] with N contiguous (the HF Conv1D
// layout, lavila/models/gpt2_gated.py:47 `Conv1D`: y = x @ W + b), M = number of sequences being decoded (32 ... 320).
//
// At M <= a few hundred rows a decoder GEMM is a pure weight-streaming problem (GPT-2 XL: 15 MB of bf16 weights per GEMM,
// 48 x 6-8 GEMMs per generated token): the roofline is HBM bandwidth, and what matters is that ALL SMs stream W.  The
// 128 x 256-tile tcgen05 kernel launches N/256 CTAs for such a shape (7 for N = 1600) and reaches ~0.3 TB/s (47 us per GEMM
// measured).  Here the grid is (N/64 column tiles) x (K splits) x (M/64 row chunks), sized to ~2 CTAs per SM:
//   phase 1  skinny_gemm_kernel: each CTA streams its [K slice] x [64 columns] block of W through a double-buffered
//            cp.async ring (128-byte rows, XOR swizzle), multiplies with bf16 mma.sync.m16n8k16 (fp32 accumulate) and
//            writes its partial tile to a fp32 workspace [split][M][N] with plain stores;
//   phase 2  skinny_epilogue_kernel: sums the splits IN ORDER (deterministic, unlike atomics: sampling must be reproducible),
//            applies bias / GELU-tanh / squared-ReLU / tanh-gate / residual exactly as gemm_epilogue.cuh and writes bf16 or fp32.
// The MMA fragment code is the P.V pattern of flash_attn.cu (A fragments by ldmatrix from a row-major tile, B fragments by
// ldmatrix.trans from the [k][n] tile).
#include "../../include/lavila_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace lv {
namespace skinny {

constexpr int BM = 64, BN = 64, BKK = 64, ROW_BYTES = 128, WARPS = 4, THREADS = WARPS * 32;
constexpr int TILE = 64 * ROW_BYTES;

struct Params {
  const __nv_bfloat16* A; long long lda;
  const __nv_bfloat16* W; long long ldb;
  float* ws;                 // [splits][M][N]
  int M, N, K;
  int splits, num_kb;        // num_kb = K / 64
};

__device__ __forceinline__ uint32_t swz(int row, int chunk) { return row * ROW_BYTES + ((chunk ^ (row & 7)) << 4); }
__device__ __forceinline__ void cp_async16(uint32_t s, const void* g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void zero16(uint32_t a) { asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0u) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A tile: 64 rows (m0 + r, zero beyond M) x 64 k;  W tile: 64 k rows x 64 columns.
__device__ __forceinline__ void load_tiles(const Params& p, uint32_t sA, uint32_t sW, int m0, int n0, int k0, int tid) {
  for (int idx = tid; idx < 64 * 8; idx += THREADS) {
    const int r = idx >> 3, c = idx & 7;
    if (m0 + r < p.M) cp_async16(sA + swz(r, c), p.A + (long long)(m0 + r) * p.lda + k0 + c * 8);
    else zero16(sA + swz(r, c));
    cp_async16(sW + swz(r, c), p.W + (long long)(k0 + r) * p.ldb + n0 + c * 8);
  }
}

__global__ void __launch_bounds__(THREADS)
skinny_gemm_kernel(const Params p) {
  __shared__ __align__(128) uint8_t smem[4 * TILE];   // A0 A1 | W0 W1
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * BN, split = blockIdx.y, m0 = blockIdx.z * BM;
  const uint32_t sA = smem_u32(smem), sW = sA + 2 * TILE;
  const int kb0 = (int)((long long)split * p.num_kb / p.splits), kb1 = (int)((long long)(split + 1) * p.num_kb / p.splits);
  const int g = lane >> 2, t = lane & 3;
  float acc[8][4];
#pragma unroll
  for (int d = 0; d < 8; ++d) acc[d][0] = acc[d][1] = acc[d][2] = acc[d][3] = 0.f;

  if (kb0 < kb1) {
    load_tiles(p, sA, sW, m0, n0, kb0 * BKK, tid);
    cp_commit();
  }
  for (int kb = kb0; kb < kb1; ++kb) {
    const int stage = (kb - kb0) & 1;
    if (kb + 1 < kb1) {
      load_tiles(p, sA + (stage ^ 1) * TILE, sW + (stage ^ 1) * TILE, m0, n0, (kb + 1) * BKK, tid);
      cp_commit();
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
    const uint32_t tA = sA + stage * TILE, tW = sW + stage * TILE;
    uint32_t af[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(af[ks], tA + swz(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4)));
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t wf[4];
        ldsm_x4_t(wf, tW + swz(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * dp + (lane >> 4)));
        mma16816(acc[2 * dp], af[kk], wf[0], wf[1]);
        mma16816(acc[2 * dp + 1], af[kk], wf[2], wf[3]);
      }
    }
    __syncthreads();   // the stage is refilled in the next iteration
  }
  // partial tile -> workspace [split][M][N]; C fragment: rows g / g+8 of the warp's 16, columns d*8 + 2t, +1
  const int r0 = m0 + warp * 16 + g, r1 = r0 + 8;
  float* base = p.ws + (long long)split * p.M * p.N;
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    const int col = n0 + d * 8 + 2 * t;
    if (r0 < p.M) *reinterpret_cast<float2*>(base + (long long)r0 * p.N + col) = make_float2(acc[d][0], acc[d][1]);
    if (r1 < p.M) *reinterpret_cast<float2*>(base + (long long)r1 * p.N + col) = make_float2(acc[d][2], acc[d][3]);
  }
}

// out[m, n] = epilogue( sum_s ws[s][m][n] ), 4 columns per thread.  Same operation order as gemm::epilogue_tile.
__global__ void __launch_bounds__(256)
skinny_epilogue_kernel(const float* __restrict__ ws, int splits, int M, int N, int flags, const float* __restrict__ bias,
                       const float* __restrict__ scale_ptr, const float* __restrict__ resid, long long ldr, void* __restrict__ out,
                       long long ldo) {
  const long long total = (long long)M * (N / 4);
  float scale = 1.0f;
  if (flags & LV_EPI_SCALE) {
    scale = __ldg(scale_ptr);
    if (flags & LV_EPI_SCALE_TANH) scale = tanhf(scale);
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / (N / 4)), n = (int)(i % (N / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
      const float4 x = __ldcs(reinterpret_cast<const float4*>(ws + ((long long)s * M + m) * N + n));
      v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
    }
    if (flags & LV_EPI_BIAS) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n));
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (flags & LV_EPI_GELU_TANH) {
      float* vv[4] = {&v.x, &v.y, &v.z, &v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float xx = *vv[e];
        float th;
        asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.7978845608028654f * (xx + 0.044715f * xx * xx * xx)));
        *vv[e] = 0.5f * xx * (1.0f + th);
      }
    }
    if (flags & LV_EPI_SQRELU) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      v.x *= v.x; v.y *= v.y; v.z *= v.z; v.w *= v.w;
    }
    if (flags & LV_EPI_SCALE) { v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale; }
    if (flags & LV_EPI_RESID) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(resid + (long long)m * ldr + n));
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (flags & LV_EPI_OUT_F32) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (long long)m * ldo + n) = v;
    } else {
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + (long long)m * ldo + n) =
          make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
  }
}

}  // namespace skinny
}  // namespace lv

using namespace lv;

extern "C" int lv_gemm_skinny_splits(int64_t M, int64_t N, int64_t K) {
  if (M <= 0 || N <= 0 || K <= 0 || N % 64 || K % 64) return 0;
  const long long ctas = (N / 64) * ((M + 63) / 64);
  const long long num_kb = K / 64;
  long long s = (2ll * sm_count() + ctas - 1) / ctas;
  if (s < 1) s = 1;
  if (s > num_kb) s = num_kb;
  if (s > 32) s = 32;
  return (int)s;
}

extern "C" int lv_gemm_skinny_bf16(const void* A, int64_t lda, const void* W, int64_t ldb, int64_t M, int64_t N, int64_t K,
                                   float* workspace, int splits, const LvGemmEpilogue* epi, void* stream) {
  LV_REQUIRE(A && W && workspace && epi && epi->out, "lv_gemm_skinny_bf16: null pointer");
  LV_REQUIRE(M > 0 && N > 0 && K > 0 && N % 64 == 0 && K % 64 == 0, "lv_gemm_skinny_bf16: needs N %% 64 == 0 and K %% 64 == 0 (M=%lld N=%lld K=%lld)",
             (long long)M, (long long)N, (long long)K);
  LV_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0, "lv_gemm_skinny_bf16: operands must be 16-byte aligned");
  LV_REQUIRE(splits >= 1 && splits <= K / 64, "lv_gemm_skinny_bf16: bad split count %d", splits);
  const int supported = LV_EPI_BIAS | LV_EPI_GELU_TANH | LV_EPI_SQRELU | LV_EPI_SCALE | LV_EPI_SCALE_TANH | LV_EPI_RESID | LV_EPI_OUT_F32;
  LV_REQUIRE((epi->flags & ~supported) == 0, "lv_gemm_skinny_bf16: unsupported epilogue flags 0x%x", epi->flags);
  LV_REQUIRE(!(epi->flags & LV_EPI_BIAS) || epi->bias, "lv_gemm_skinny_bf16: bias missing");
  LV_REQUIRE(!(epi->flags & LV_EPI_RESID) || epi->resid, "lv_gemm_skinny_bf16: resid missing");
  LV_REQUIRE(!(epi->flags & LV_EPI_SCALE) || epi->scale_ptr, "lv_gemm_skinny_bf16: scale_ptr missing");
  LV_REQUIRE(epi->ldo % 4 == 0 && (!(epi->flags & LV_EPI_RESID) || epi->ldr % 4 == 0), "lv_gemm_skinny_bf16: ldo / ldr must be multiples of 4");
  skinny::Params p{};
  p.A = (const __nv_bfloat16*)A; p.lda = lda; p.W = (const __nv_bfloat16*)W; p.ldb = ldb;
  p.ws = workspace; p.M = (int)M; p.N = (int)N; p.K = (int)K; p.splits = splits; p.num_kb = (int)(K / 64);
  dim3 grid((unsigned)(N / 64), (unsigned)splits, (unsigned)((M + 63) / 64));
  skinny::skinny_gemm_kernel<<<grid, skinny::THREADS, 0, (cudaStream_t)stream>>>(p);
  int rc = check_launch("lv_gemm_skinny_bf16");
  if (rc) return rc;
  const long long total = M * (N / 4);
  const int blocks = (int)((total + 255) / 256 < 4ll * sm_count() ? (total + 255) / 256 : 4ll * sm_count());
  skinny::skinny_epilogue_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(workspace, splits, (int)M, (int)N, epi->flags, epi->bias,
                                                                          epi->scale_ptr, epi->resid, epi->ldr, epi->out, epi->ldo);
  return check_launch("lv_gemm_skinny_bf16(epilogue)");
}
