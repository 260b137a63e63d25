// Fused GEMM epilogue shared by the 1-CTA and 2-CTA tcgen05 kernels: one call drains this warp's share (32 rows x 128
// columns) of a 128 x 256 fp32 accumulator from TMEM and applies bias / activation / gate / residual / casts.
#pragma once
#include "../../include/lavila_b200.h"
#include "ptx.cuh"

namespace lv {
namespace gemm {

constexpr int BN = 256;
constexpr int EPI_PITCH = 16;  // floats: a 32-row x 16-column half chunk per warp, 16-byte units XOR-swizzled

struct Args {
  int M, N, K;
  int num_m_blks, num_n_blks, k_splits, kb_per_split, num_kb;
  // stream-K (split-K GEMMs with the atomic epilogue): the tiles x k-blocks iteration space is cut into one contiguous,
  // equally long span per CTA (pair); a span covers the tail of one tile and the head of the next.  0 = classic split-K.
  long long sk_total_kb, sk_kb_per_cta;
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


// Work items of one CTA (pair): (tile, k-block range).  Every warp role walks the same sequence.
struct ItemIter {
  bool streamk;
  int item, stride, total_items, n_tiles, kb_per_split, num_kb;
  long long gk, g1;
  __device__ __forceinline__ ItemIter(const Args& g, int cta_id, int num_ctas) {
    streamk = g.sk_kb_per_cta > 0;
    n_tiles = g.num_m_blks * g.num_n_blks;
    num_kb = g.num_kb;
    kb_per_split = g.kb_per_split;
    item = cta_id;
    stride = num_ctas;
    total_items = n_tiles * g.k_splits;
    gk = (long long)cta_id * g.sk_kb_per_cta;
    g1 = gk + g.sk_kb_per_cta;
    if (g1 > g.sk_total_kb) g1 = g.sk_total_kb;
  }
  __device__ __forceinline__ bool next(int& tile, int& kb0, int& kb1) {
    if (streamk) {
      if (gk >= g1) return false;
      tile = (int)(gk / num_kb);
      kb0 = (int)(gk - (long long)tile * num_kb);
      const long long left = g1 - gk;
      kb1 = (left < (long long)(num_kb - kb0)) ? kb0 + (int)left : num_kb;
      gk += kb1 - kb0;
      return true;
    }
    if (item >= total_items) return false;
    const int split = item / n_tiles;
    tile = item - split * n_tiles;
    kb0 = split * kb_per_split;
    kb1 = min(num_kb, kb0 + kb_per_split);
    item += stride;
    return true;
  }
};

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

// q: TMEM lane quarter of this warp, half: which 128 of the 256 columns, buf: this warp's 2 KB staging buffer,
// tmem_acc: TMEM address of the accumulator stage (lane 0, first column), m_base: first of this warp's 32 rows.
template <int CT_FLAGS>
__device__ __forceinline__ void epilogue_tile(const Args& g, const int flags, const float scale, float* buf, uint64_t* tfull,
                                              const uint32_t aphase, const uint32_t tmem_acc, const int m_base, const int n0,
                                              const int half, const int q, const int lane) {
  const int r8 = lane >> 2;        // row within a group of 8
  const int ch = lane & 3;         // this lane's 16-byte unit (4 columns) inside a 16-column half chunk
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
  mbar_wait(tfull, aphase);
  tc_fence_after();
#pragma unroll
  for (int cc = 0; cc < CPW; ++cc) {
    const int c = half * CPW + cc;
    const bool chunk_ok = n0 + c * 32 < g.N;   // warp-uniform
    uint32_t r[32];
    if (chunk_ok) {
      tmem_ld_32x32(tmem_acc + (uint32_t(q * 32) << 16) + c * 32, r);
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
            red_add_v4_f32(o, v);          // one 16-byte vector reduction (n and ldo are multiples of 4)
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
}

}  // namespace gemm
}  // namespace lv
