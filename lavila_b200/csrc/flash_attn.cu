// General attention forward for the narrator's inference path: cross-attention over the 256 pooled video tokens
// (lavila/models/gpt2_gated.py:320-360, _attn :206-238), CoCa attention pooling with ONE shared 64-d key/value head
// (multi-query, lavila/models/coca.py:100-125) and GPT-2 causal self-attention (:464-477).
//
// Flash-style: one CTA = 64 query rows of one (batch, head); K/V walk in 64-key blocks through a double-buffered
// cp.async ring; S = Q K^T and O += P V on bf16 tensor cores (mma.sync m16n8k16, fp32 accumulate), online softmax in
// fp32.  head_dim = 64.  Inputs are addressed in place: element (b, h, i, d) of Q lives at
// q[(b*q_rows + i)*ld_q + h*64 + d]; K/V at k[(b*kv_rows + j)*ld_kv + h*kv_head_stride + d] (kv_head_stride = 0 for
// multi-query).  Causal mask (optional): key j visible to query i iff j <= i + (Lk - Lq)   (HF GPT-2 semantics).
#include "../../include/lavila_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace lv {
namespace flash {

constexpr int HD = 64, BQ = 64, BK = 64, ROW_BYTES = 128, WARPS = 4;
constexpr float LOG2E = 1.4426950408889634f;

struct Params {
  const __nv_bfloat16 *q, *k, *v;
  __nv_bfloat16* out;
  long long ld_q, ld_kv, ld_out;
  long long q_rows, kv_rows;   // rows per batch element
  int kv_head_stride;
  int Lq, Lk, causal;
  float scale;
  const int* lk_ptr;   // optional: the key count lives in device memory (CUDA-graph replay of a decoding step)
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
__device__ __forceinline__ float qmax(float v) { v = fmaxf(v, __shfl_xor_sync(~0u, v, 1)); return fmaxf(v, __shfl_xor_sync(~0u, v, 2)); }
__device__ __forceinline__ float qsum(float v) { v += __shfl_xor_sync(~0u, v, 1); return v + __shfl_xor_sync(~0u, v, 2); }

__device__ __forceinline__ void load_rows(uint32_t tile, const __nv_bfloat16* base, long long ld, int row0, int nvalid, int tid) {
  for (int idx = tid; idx < BK * 8; idx += WARPS * 32) {
    const int r = idx >> 3, c = idx & 7;
    if (r < nvalid) cp_async16(tile + swz(r, c), base + (long long)(row0 + r) * ld + c * 8);
    else zero16(tile + swz(r, c));
  }
}

__global__ void __launch_bounds__(WARPS * 32)
flash_fwd_kernel(const Params p) {
  __shared__ __align__(128) uint8_t smem[(BQ + 4 * BK) * ROW_BYTES];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const uint32_t sQ = smem_u32(smem), sK = sQ + BQ * ROW_BYTES, sV = sK + 2 * BK * ROW_BYTES;
  const int q0 = qb * BQ;
  const __nv_bfloat16* qbase = p.q + (long long)b * p.q_rows * p.ld_q + h * HD;
  const __nv_bfloat16* kbase = p.k + (long long)b * p.kv_rows * p.ld_kv + h * p.kv_head_stride;
  const __nv_bfloat16* vbase = p.v + (long long)b * p.kv_rows * p.ld_kv + h * p.kv_head_stride;
  const int nq = min(BQ, p.Lq - q0);
  const int Lk = p.lk_ptr ? __ldg(p.lk_ptr) : p.Lk;
  int Lk_eff = Lk;   // causal: keys beyond the last query row of this block are never visible
  if (p.causal) Lk_eff = min(Lk, q0 + nq + (Lk - p.Lq));
  const int nblocks = (Lk_eff + BK - 1) / BK;

  load_rows(sQ, qbase, p.ld_q, q0, nq, tid);
  load_rows(sK, kbase, p.ld_kv, 0, min(BK, Lk_eff), tid);
  load_rows(sV, vbase, p.ld_kv, 0, min(BK, Lk_eff), tid);
  cp_commit();

  const int g = lane >> 2, t = lane & 3;
  const int r0 = warp * 16 + g, r1 = r0 + 8;          // rows inside the query block
  const int off = Lk - p.Lq;
  const float sl2 = p.scale * LOG2E;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float o[8][4];
#pragma unroll
  for (int d = 0; d < 8; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
  uint32_t qf[4][4];

  for (int kb = 0; kb < nblocks; ++kb) {
    const int stage = kb & 1;
    if (kb + 1 < nblocks) {   // prefetch the next K/V block into the other stage
      const int k0n = (kb + 1) * BK;
      load_rows(sK + (stage ^ 1) * BK * ROW_BYTES, kbase, p.ld_kv, k0n, min(BK, Lk_eff - k0n), tid);
      load_rows(sV + (stage ^ 1) * BK * ROW_BYTES, vbase, p.ld_kv, k0n, min(BK, Lk_eff - k0n), tid);
      cp_commit();
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
    if (kb == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ldsm_x4(qf[ks], sQ + swz(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4)));
    }
    const uint32_t tK = sK + stage * BK * ROW_BYTES, tV = sV + stage * BK * ROW_BYTES;
    const int k0 = kb * BK;
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t kf[4];
        ldsm_x4(kf, tK + swz(nt * 8 + (lane & 7), 4 * kp + (lane >> 3)));
        mma16816(s[nt], qf[2 * kp], kf[0], kf[1]);
        mma16816(s[nt], qf[2 * kp + 1], kf[2], kf[3]);
      }
    }
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = k0 + nt * 8 + 2 * t + e;
        const bool in = col < Lk;
        const bool v0 = in && (!p.causal || col <= q0 + r0 + off), v1 = in && (!p.causal || col <= q0 + r1 + off);
        s[nt][e] = v0 ? s[nt][e] : -INFINITY;
        s[nt][2 + e] = v1 ? s[nt][2 + e] : -INFINITY;
        bm0 = fmaxf(bm0, s[nt][e]);
        bm1 = fmaxf(bm1, s[nt][2 + e]);
      }
    bm0 = qmax(bm0);
    bm1 = qmax(bm1);
    const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);
    // rows with nothing visible yet keep m = -inf; guard the (-inf) - (-inf) case
    const float c0 = (mn0 == -INFINITY) ? 1.f : exp2f((m0 - mn0) * sl2), c1 = (mn1 == -INFINITY) ? 1.f : exp2f((m1 - mn1) * sl2);
    const float b0 = (mn0 == -INFINITY) ? 0.f : mn0 * sl2, b1 = (mn1 == -INFINITY) ? 0.f : mn1 * sl2;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[nt][e] = exp2f(fmaf(s[nt][e], sl2, -b0));
        s[nt][2 + e] = exp2f(fmaf(s[nt][2 + e], sl2, -b1));
        rs0 += s[nt][e];
        rs1 += s[nt][2 + e];
      }
    l0 = l0 * c0 + qsum(rs0);
    l1 = l1 * c1 + qsum(rs1);
    m0 = mn0;
    m1 = mn1;
#pragma unroll
    for (int d = 0; d < 8; ++d) { o[d][0] *= c0; o[d][1] *= c0; o[d][2] *= c1; o[d][3] *= c1; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4] = {pack_bf16x2(s[2 * kk][0], s[2 * kk][1]), pack_bf16x2(s[2 * kk][2], s[2 * kk][3]),
                       pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t vf[4];
        ldsm_x4_t(vf, tV + swz(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * dp + (lane >> 4)));
        mma16816(o[2 * dp], a, vf[0], vf[1]);
        mma16816(o[2 * dp + 1], a, vf[2], vf[3]);
      }
    }
    __syncthreads();   // everyone is done with this stage before it is refilled two iterations later
  }
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
  // stage O in the (now free) Q tile, then coalesced 128-byte row stores
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(sQ + swz(r0, d) + 4 * t), "r"(pack_bf16x2(o[d][0] * i0, o[d][1] * i0)) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(sQ + swz(r1, d) + 4 * t), "r"(pack_bf16x2(o[d][2] * i1, o[d][3] * i1)) : "memory");
  }
  __syncwarp();
  __nv_bfloat16* obase = p.out + (long long)b * p.q_rows * p.ld_out + h * HD;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = it * 32 + lane, r = warp * 16 + (idx >> 3), c = idx & 7;
    if (r < nq) {
      uint4 v;
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sQ + swz(r, c)));
      *reinterpret_cast<uint4*>(obase + (long long)(q0 + r) * p.ld_out + c * 8) = v;
    }
  }
}

}  // namespace flash
}  // namespace lv

using namespace lv;

static int flash_launch(const void* q, int64_t ld_q, int64_t q_rows, const void* k, const void* v, int64_t ld_kv, int64_t kv_rows,
                        int kv_head_stride, void* out, int64_t ld_out, int B, int H, int Lq, int Lk, const int32_t* lk_dev,
                        int causal, float scale, void* stream, const char* what) {
  LV_REQUIRE(q && k && v && out && B > 0 && H > 0 && Lq > 0 && (Lk > 0 || lk_dev), "%s: bad arguments", what);
  LV_REQUIRE(ld_q % 8 == 0 && ld_kv % 8 == 0 && ld_out % 8 == 0 && kv_head_stride % 8 == 0, "%s: strides must be multiples of 8 elements", what);
  LV_REQUIRE(((uintptr_t)q & 15) == 0 && ((uintptr_t)k & 15) == 0 && ((uintptr_t)v & 15) == 0 && ((uintptr_t)out & 15) == 0, "%s: pointers must be 16-byte aligned", what);
  // >= 64 query rows per (batch, head) -- CoCa pooling (256 x 1025), teacher-forced cross-attention (77 x 256) -- run the tcgen05
  // key-tiled kernel (flash_tc.cu); causal self-attention from 128 rows (at 77 x 77 the 128-row tile is mostly masked: measured
  // 0.27 ms vs 0.16 ms here).  Decoding steps (1 query row per sequence: a matrix-vector product per head, HBM-bound) and the
  // device-resident key count of the CUDA-graph replay stay on the mma.sync kernel below.
  if ((Lq >= 128 || (Lq >= 64 && !causal)) && !lk_dev && flash_tc_enabled())
    return flash_attn_fwd_tc(q, ld_q, q_rows, H * flash::HD, k, v, ld_kv, kv_rows, (H - 1) * kv_head_stride + flash::HD, kv_head_stride,
                             out, ld_out, nullptr, B, 1, H, Lq, Lk, 0, 0, 0, 0, 0, causal, scale, (cudaStream_t)stream);
  flash::Params p{};
  p.q = (const __nv_bfloat16*)q; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v; p.out = (__nv_bfloat16*)out;
  p.ld_q = ld_q; p.ld_kv = ld_kv; p.ld_out = ld_out; p.q_rows = q_rows; p.kv_rows = kv_rows;
  p.kv_head_stride = kv_head_stride; p.Lq = Lq; p.Lk = Lk; p.causal = causal; p.scale = scale; p.lk_ptr = lk_dev;
  dim3 grid((Lq + flash::BQ - 1) / flash::BQ, H, B);
  flash::flash_fwd_kernel<<<grid, flash::WARPS * 32, 0, (cudaStream_t)stream>>>(p);
  return check_launch(what);
}

extern "C" int lv_flash_attn_fwd(const void* q, int64_t ld_q, int64_t q_rows, const void* k, const void* v, int64_t ld_kv,
                                 int64_t kv_rows, int kv_head_stride, void* out, int64_t ld_out, int B, int H, int Lq, int Lk,
                                 int causal, float scale, void* stream) {
  return flash_launch(q, ld_q, q_rows, k, v, ld_kv, kv_rows, kv_head_stride, out, ld_out, B, H, Lq, Lk, nullptr, causal, scale,
                      stream, "lv_flash_attn_fwd");
}

extern "C" int lv_flash_attn_fwd_dyn(const void* q, int64_t ld_q, int64_t q_rows, const void* k, const void* v, int64_t ld_kv,
                                     int64_t kv_rows, int kv_head_stride, void* out, int64_t ld_out, int B, int H, int Lq,
                                     const int32_t* lk_dev, int causal, float scale, void* stream) {
  LV_REQUIRE(lk_dev, "lv_flash_attn_fwd_dyn: lk_dev is NULL");
  return flash_launch(q, ld_q, q_rows, k, v, ld_kv, kv_rows, kv_head_stride, out, ld_out, B, H, Lq, 0, lk_dev, causal, scale,
                      stream, "lv_flash_attn_fwd_dyn");
}
