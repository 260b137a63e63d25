// Divided space-time attention (and the text tower's causal attention) for LaViLa, forward + backward.
//
// The reference (lavila/models/timesformer.py:107-144) materialises per-group q/k/v copies with einops, repeats and
// concatenates the CLS key/value into every group, and runs bmm -> softmax -> bmm with `sim` in HBM.  Here a "group"
// (one (clip, head, frame) for space, one (clip, head, spatial position) for time, one (caption, head) for text) is
// addressed in place inside the packed projection output qkv[rows, 3*D] through (base_row, row_stride); the CLS
// key/value is appended as one extra key; scores never leave the SM.
//
//   group kernels : bf16 tensor-core MMAs (mma.sync m16n8k16, fp32 accumulate), whole group resident in shared
//                   memory (128-byte rows, XOR-swizzled 16-byte units), fp32 softmax via quad shuffles.
//   cls kernels   : the CLS query attends to all N tokens (:119) -- a bandwidth-bound GEMV-like pass.
//
// Layouts: qkv bf16 [rows, 3*D] = [q | k | v], each D = H * 64 (head h at columns h*64..h*64+63);
//          out bf16 [rows, D]; lse fp32 [rows, H] (log-sum-exp of the scaled scores, natural log).
// head_dim is fixed at 64 (TSF-B/L, CLIP text towers, GPT-2 XL all use 64).
#include "../../include/lavila_b200.h"
#include "host_common.h"
#include "ptx.cuh"
#include <stdlib.h>

namespace lv {
namespace attn {

constexpr int HD = 64;            // head dim
constexpr int ROW_BYTES = 128;    // 64 bf16
constexpr float LOG2E = 1.4426950408889634f;

struct Params {
  const __nv_bfloat16* qkv;   // [rows, 3D]
  long long ld_qkv;
  __nv_bfloat16* out;         // fwd: output; bwd: forward output (for delta)
  long long ld_out;
  float* lse;                 // [rows, H]
  int H, D;
  int Lq;                     // queries (= non-CLS keys) per group
  int has_cls, causal;
  long long row_stride;       // token stride between consecutive group rows
  int inner;                  // groups per (b): space: T frames; time: n positions; text: 1
  long long clip_rows;        // rows per clip / caption (N or L)
  long long inner_stride;     // row offset between consecutive inner groups (space: n, time: 1)
  int first;                  // first patch row inside a clip (1 when a CLS token leads, else 0)
  long long num_groups;       // B * H * inner
  int qt;                     // query tiles of 16 per group
  int wg;                     // warps per group
  int groups_per_cta;
  float scale;
  // backward only
  const __nv_bfloat16* dout;  // [rows, D]
  long long ld_dout;
  __nv_bfloat16* dqkv;        // [rows, 3D]
  long long ld_dqkv;
  float* dcls_kv;             // [B, H, 2, 64] fp32 (CLS key/value gradient accumulator)
  int accumulate_kv;          // add to the k/v gradient already in dqkv (written by the CLS-query backward)
  int staged;                 // small groups: O and the old k/v gradient are staged in shared memory by cp.async
  int group_bytes;            // shared-memory bytes per group (backward)
};

__device__ __forceinline__ uint32_t swz(int row, int chunk) { return row * ROW_BYTES + ((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void st_shared_zero16(uint32_t addr) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

struct GroupCoord {
  int b, h;
  long long base_row;  // first group row (global row index)
  long long cls_row;
};
// CTA -> groups.  A CTA owns `groups_per_cta` consecutive inner positions of ONE (clip, head); CTAs are ordered
// head-fastest, so CTAs that run at the same time read different heads of the same token rows (contiguous DRAM runs),
// while the groups inside a CTA share (clip, head) and their CLS-key gradients can be reduced in shared memory.
__device__ __forceinline__ bool decode_group(const Params& p, long long cta, int g_local, GroupCoord& c) {
  const int jb_count = (p.inner + p.groups_per_cta - 1) / p.groups_per_cta;
  c.h = (int)(cta % p.H);
  const long long t = cta / p.H;
  const int jb = (int)(t % jb_count);
  c.b = (int)(t / jb_count);
  const int in = jb * p.groups_per_cta + g_local;
  c.cls_row = (long long)c.b * p.clip_rows;
  c.base_row = c.cls_row + p.first + (long long)in * p.inner_stride;
  return in < p.inner;
}

// Cooperative load of `rows_pad` rows x 64 bf16 into a swizzled tile.  Row r < n_rows comes from
// src + (base_row + r*row_stride)*ld ; row == n_rows (if extra) from extra_ptr ; other rows are zero-filled.
__device__ __forceinline__ void load_tile(uint32_t tile, int rows_pad, int n_rows, const __nv_bfloat16* src,
                                          long long ld, long long base_row, long long row_stride,
                                          const __nv_bfloat16* extra_ptr, int tid, int nthr, bool active) {
  for (int idx = tid; idx < rows_pad * 8; idx += nthr) {
    const int r = idx >> 3, c = idx & 7;
    const uint32_t dst = tile + swz(r, c);
    if (active && r < n_rows) cp_async16(dst, src + (base_row + (long long)r * row_stride) * ld + c * 8);
    else if (active && r == n_rows && extra_ptr) cp_async16(dst, extra_ptr + c * 8);
    else st_shared_zero16(dst);
  }
}

// ================================================================================================ forward
template <int NT>  // NT = key tiles of 8 held in registers (keys <= 8*NT, NT even)
__global__ void __launch_bounds__(256)
group_attn_fwd_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g_local = warp / p.wg, w_in_g = warp - g_local * p.wg;
  const int q_rows = p.qt * 16, k_rows = NT * 8;
  const uint32_t gs = smem_u32(smem) + g_local * (q_rows + 2 * k_rows) * ROW_BYTES;
  const uint32_t sQ = gs, sK = gs + q_rows * ROW_BYTES, sV = sK + k_rows * ROW_BYTES;
  GroupCoord gc;
  const bool active = decode_group(p, blockIdx.x, g_local, gc);
  const int Lk = p.Lq + (p.has_cls ? 1 : 0);
  {
    const int tid = w_in_g * 32 + lane, nthr = p.wg * 32;
    const __nv_bfloat16* qb = p.qkv + gc.h * HD;
    const __nv_bfloat16* cls_k = p.has_cls ? qb + gc.cls_row * p.ld_qkv + p.D : nullptr;
    const __nv_bfloat16* cls_v = p.has_cls ? qb + gc.cls_row * p.ld_qkv + 2 * p.D : nullptr;
    load_tile(sQ, q_rows, p.Lq, qb, p.ld_qkv, gc.base_row, p.row_stride, nullptr, tid, nthr, active);
    load_tile(sK, k_rows, p.Lq, qb + p.D, p.ld_qkv, gc.base_row, p.row_stride, cls_k, tid, nthr, active);
    load_tile(sV, k_rows, p.Lq, qb + 2 * p.D, p.ld_qkv, gc.base_row, p.row_stride, cls_v, tid, nthr, active);
  }
  cp_async_wait_all();
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  const float sl2 = p.scale * LOG2E;
  const int nt_used = (Lk + 7) >> 3;
  for (int qt = w_in_g; qt < p.qt; qt += p.wg) {
    uint32_t qf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      ldsm_x4(qf[ks], sQ + swz(qt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4)));
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      if (nt < nt_used) {
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
          uint32_t kf[4];
          ldsm_x4(kf, sK + swz(nt * 8 + (lane & 7), 4 * kp + (lane >> 3)));
          mma16816(s[nt], qf[2 * kp], kf[0], kf[1]);
          mma16816(s[nt], qf[2 * kp + 1], kf[2], kf[3]);
        }
      }
    }
    // ---- softmax over keys (fp32), rows r0 = qt*16+g and r1 = r0+8
    const int r0 = qt * 16 + g, r1 = r0 + 8;
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = nt * 8 + 2 * t + e;
        const bool in = col < Lk;
        const bool v0 = in && (!p.causal || col <= r0), v1 = in && (!p.causal || col <= r1);
        s[nt][e] = v0 ? s[nt][e] : -INFINITY;
        s[nt][2 + e] = v1 ? s[nt][2 + e] : -INFINITY;
        m0 = fmaxf(m0, s[nt][e]);
        m1 = fmaxf(m1, s[nt][2 + e]);
      }
    }
    m0 = quad_max(m0);
    m1 = quad_max(m1);
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[nt][e] = exp2f((s[nt][e] - m0) * sl2);
        s[nt][2 + e] = exp2f((s[nt][2 + e] - m1) * sl2);
        sum0 += s[nt][e];
        sum1 += s[nt][2 + e];
      }
    }
    sum0 = quad_sum(sum0);
    sum1 = quad_sum(sum1);
    // ---- O = P V
    float o[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
      if (kk * 16 < Lk) {
        uint32_t a[4];
        a[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        a[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        a[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        a[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t vf[4];
          ldsm_x4_t(vf, sV + swz(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * dp + (lane >> 4)));
          mma16816(o[2 * dp], a, vf[0], vf[1]);
          mma16816(o[2 * dp + 1], a, vf[2], vf[3]);
        }
      }
    }
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
    // ---- stage O (bf16) in this warp's Q tile, then 128-byte-per-row coalesced stores
    __syncwarp();
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      st_shared_u32(sQ + swz(r0, dt) + 4 * t, pack_bf16x2(o[dt][0] * inv0, o[dt][1] * inv0));
      st_shared_u32(sQ + swz(r1, dt) + 4 * t, pack_bf16x2(o[dt][2] * inv1, o[dt][3] * inv1));
    }
    __syncwarp();
    if (active) {
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int idx = it * 32 + lane, r = idx >> 3, c = idx & 7;
        const int qrow = qt * 16 + r;
        if (qrow < p.Lq) {
          const uint4 v = ld_shared_v4(sQ + swz(qrow, c));
          *reinterpret_cast<uint4*>(p.out + (gc.base_row + (long long)qrow * p.row_stride) * p.ld_out + gc.h * HD + c * 8) = v;
        }
      }
      if (t == 0) {
        if (r0 < p.Lq) p.lse[(gc.base_row + (long long)r0 * p.row_stride) * p.H + gc.h] = m0 * p.scale + logf(sum0);
        if (r1 < p.Lq) p.lse[(gc.base_row + (long long)r1 * p.row_stride) * p.H + gc.h] = m1 * p.scale + logf(sum1);
      }
    }
  }
}

// ================================================================================================ backward
// Shared memory per group: Q, dO [q_rows][64]; K, V [k_rows][64]; lse*log2e, delta [q_rows] fp32;
// dS^T [k_rows][q_rows + 8] bf16 (scaled by `scale`, so dQ = dS K and dK = dS^T Q need no further factor).
__device__ __forceinline__ int bwd_group_bytes(int q_rows, int k_rows) {
  return 2 * q_rows * ROW_BYTES + 2 * k_rows * ROW_BYTES + 2 * q_rows * 4 + k_rows * (q_rows + 8) * 2;
}

// Write one key tile's dK or dV (fp32 registers, C-fragment layout) to dqkv: bf16 staging in the warp's private
// V-tile rows, then 128-byte-per-row stores (optionally adding the CLS-query contribution already in dqkv);
// the CLS key row goes to the fp32 accumulator with atomics at full precision.
__device__ __forceinline__ void write_kv_grad(const Params& p, const GroupCoord& gc, float (&acc)[8][4], int part,
                                              int cls_slot, uint32_t sV, int kt, int key0, int key1, int lane, int t,
                                              bool active, uint32_t sOld, float* cls_red) {
  __syncwarp();
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    st_shared_u32(sV + swz(key0, dt) + 4 * t, pack_bf16x2(acc[dt][0], acc[dt][1]));
    st_shared_u32(sV + swz(key1, dt) + 4 * t, pack_bf16x2(acc[dt][2], acc[dt][3]));
  }
  if (active && p.has_cls && p.dcls_kv) {
    // CLS key row: full-precision fp32; staged per group in shared memory, reduced per CTA before the global atomics
    float* base = cls_red + cls_slot * HD;
    if (key0 == p.Lq) {
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) { base[dt * 8 + 2 * t] = acc[dt][0]; base[dt * 8 + 2 * t + 1] = acc[dt][1]; }
    }
    if (key1 == p.Lq) {
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) { base[dt * 8 + 2 * t] = acc[dt][2]; base[dt * 8 + 2 * t + 1] = acc[dt][3]; }
    }
  }
  __syncwarp();
  if (active) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = it * 32 + lane, r = idx >> 3, c = idx & 7;
      const int key = kt * 16 + r;
      if (key < p.Lq) {
        uint4 v = ld_shared_v4(sV + swz(key, c));
        __nv_bfloat16* dst = p.dqkv + (gc.base_row + (long long)key * p.row_stride) * p.ld_dqkv + part * p.D + gc.h * HD + c * 8;
        if (p.accumulate_kv) {
          const uint4 old = sOld ? ld_shared_v4(sOld + swz(key, c)) : *reinterpret_cast<const uint4*>(dst);
          uint32_t vw[4] = {v.x, v.y, v.z, v.w};
          const uint32_t ow[4] = {old.x, old.y, old.z, old.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 x = unpack_bf16x2(vw[j]), y = unpack_bf16x2(ow[j]);
            vw[j] = pack_bf16x2(x.x + y.x, x.y + y.y);
          }
          v = make_uint4(vw[0], vw[1], vw[2], vw[3]);
        }
        *reinterpret_cast<uint4*>(dst) = v;
      }
    }
  }
}

__global__ void __launch_bounds__(256)
group_attn_bwd_kernel(const Params p, const int k_rows) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g_local = warp / p.wg, w_in_g = warp - g_local * p.wg;
  const int q_rows = p.qt * 16;
  const int ds_pitch = (q_rows + 8) * 2;  // bytes
  const uint32_t gs = smem_u32(smem) + g_local * p.group_bytes;
  const uint32_t sQ = gs, sdO = sQ + q_rows * ROW_BYTES, sK = sdO + q_rows * ROW_BYTES, sV = sK + k_rows * ROW_BYTES;
  const uint32_t sLse = sV + k_rows * ROW_BYTES, sDelta = sLse + q_rows * 4, sdS = sDelta + q_rows * 4;
  // small groups (time attention): O and the k/v gradient to accumulate onto are staged by the same cp.async batch,
  // so no global-load latency is exposed after the MMAs
  const bool staged = p.staged != 0;
  const uint32_t sO = sdS + k_rows * ds_pitch;
  const uint32_t sOldK = sO + q_rows * ROW_BYTES, sOldV = sOldK + k_rows * ROW_BYTES;
  float* cls_red = reinterpret_cast<float*>(smem + p.groups_per_cta * p.group_bytes) + g_local * 2 * HD;
  float* lse_s = reinterpret_cast<float*>(smem + (sLse - smem_u32(smem)));
  float* delta_s = reinterpret_cast<float*>(smem + (sDelta - smem_u32(smem)));
  GroupCoord gc;
  const bool active = decode_group(p, blockIdx.x, g_local, gc);
  const int Lk = p.Lq + (p.has_cls ? 1 : 0);
  const int tid = w_in_g * 32 + lane, nthr = p.wg * 32;
  {
    const __nv_bfloat16* qb = p.qkv + gc.h * HD;
    const __nv_bfloat16* cls_k = p.has_cls ? qb + gc.cls_row * p.ld_qkv + p.D : nullptr;
    const __nv_bfloat16* cls_v = p.has_cls ? qb + gc.cls_row * p.ld_qkv + 2 * p.D : nullptr;
    load_tile(sQ, q_rows, p.Lq, qb, p.ld_qkv, gc.base_row, p.row_stride, nullptr, tid, nthr, active);
    load_tile(sdO, q_rows, p.Lq, p.dout + gc.h * HD, p.ld_dout, gc.base_row, p.row_stride, nullptr, tid, nthr, active);
    load_tile(sK, k_rows, p.Lq, qb + p.D, p.ld_qkv, gc.base_row, p.row_stride, cls_k, tid, nthr, active);
    load_tile(sV, k_rows, p.Lq, qb + 2 * p.D, p.ld_qkv, gc.base_row, p.row_stride, cls_v, tid, nthr, active);
    if (staged) {
      load_tile(sO, q_rows, p.Lq, p.out + gc.h * HD, p.ld_out, gc.base_row, p.row_stride, nullptr, tid, nthr, active);
      if (p.accumulate_kv) {
        const __nv_bfloat16* db = p.dqkv + gc.h * HD;
        load_tile(sOldK, k_rows, p.Lq, db + p.D, p.ld_dqkv, gc.base_row, p.row_stride, nullptr, tid, nthr, active);
        load_tile(sOldV, k_rows, p.Lq, db + 2 * p.D, p.ld_dqkv, gc.base_row, p.row_stride, nullptr, tid, nthr, active);
      }
      for (int r = tid; r < q_rows; r += nthr)
        lse_s[r] = (active && r < p.Lq) ? p.lse[(gc.base_row + (long long)r * p.row_stride) * p.H + gc.h] * LOG2E : 0.f;
    } else {
      // delta_q = sum_d dO[q,d] * O[q,d]  (8 consecutive lanes share a row) and lse, straight from global memory
      for (int idx = tid; idx < q_rows * 8; idx += nthr) {
        const int r = idx >> 3, c = idx & 7;
        float part = 0.f;
        if (active && r < p.Lq) {
          const long long grow = gc.base_row + (long long)r * p.row_stride;
          const uint4 a = __ldg(reinterpret_cast<const uint4*>(p.dout + grow * p.ld_dout + gc.h * HD + c * 8));
          const uint4 b = __ldg(reinterpret_cast<const uint4*>(p.out + grow * p.ld_out + gc.h * HD + c * 8));
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 x = unpack_bf16x2(aw[j]), y = unpack_bf16x2(bw[j]);
            part += x.x * y.x + x.y * y.y;
          }
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        part += __shfl_xor_sync(0xffffffffu, part, 4);
        if (c == 0) {
          delta_s[r] = part;
          lse_s[r] = (active && r < p.Lq) ? p.lse[(gc.base_row + (long long)r * p.row_stride) * p.H + gc.h] * LOG2E : 0.f;
        }
      }
    }
  }
  cp_async_wait_all();
  __syncthreads();
  if (staged) {   // delta from the staged O and dO tiles
    for (int idx = tid; idx < q_rows * 8; idx += nthr) {
      const int r = idx >> 3, c = idx & 7;
      const uint4 a = ld_shared_v4(sdO + swz(r, c)), b = ld_shared_v4(sO + swz(r, c));
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 x = unpack_bf16x2(aw[j]), y = unpack_bf16x2(bw[j]);
        part += x.x * y.x + x.y * y.y;
      }
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      if (c == 0) delta_s[r] = part;
    }
    __syncthreads();
  }

  const int g = lane >> 2, t = lane & 3;
  const float sl2 = p.scale * LOG2E;
  // ------------------------------------------------------------ phase 1: key tiles -> dK, dV, dS^T
  for (int kt = w_in_g; kt * 16 < k_rows; kt += p.wg) {
    if (kt * 16 >= Lk) {  // pure padding tile: dS^T rows must still be zero for phase 2
      for (int idx = lane; idx < 16 * (q_rows / 2); idx += 32) {
        const int r = idx / (q_rows / 2), c2 = idx - r * (q_rows / 2);
        st_shared_u32(sdS + (kt * 16 + r) * ds_pitch + c2 * 4, 0u);
      }
      continue;
    }
    uint32_t kf[4][4], vf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t off = swz(kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4));
      ldsm_x4(kf[ks], sK + off);
      ldsm_x4(vf[ks], sV + off);
    }
    float dv[8][4], dk[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      dv[dt][0] = dv[dt][1] = dv[dt][2] = dv[dt][3] = 0.f;
      dk[dt][0] = dk[dt][1] = dk[dt][2] = dk[dt][3] = 0.f;
    }
    const int key0 = kt * 16 + g, key1 = key0 + 8;
    for (int qb = 0; qb < p.qt; ++qb) {
      const int q0 = qb * 16;
      float st[2][4], dp[2][4];
#pragma unroll
      for (int nq = 0; nq < 2; ++nq) {
        st[nq][0] = st[nq][1] = st[nq][2] = st[nq][3] = 0.f;
        dp[nq][0] = dp[nq][1] = dp[nq][2] = dp[nq][3] = 0.f;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
          uint32_t bq[4], bo[4];
          const uint32_t off = swz(q0 + nq * 8 + (lane & 7), 4 * kp + (lane >> 3));
          ldsm_x4(bq, sQ + off);
          ldsm_x4(bo, sdO + off);
          mma16816(st[nq], kf[2 * kp], bq[0], bq[1]);
          mma16816(st[nq], kf[2 * kp + 1], bq[2], bq[3]);
          mma16816(dp[nq], vf[2 * kp], bo[0], bo[1]);
          mma16816(dp[nq], vf[2 * kp + 1], bo[2], bo[3]);
        }
      }
      // P^T and dS^T for keys (key0, key1) x queries (q0 + nq*8 + 2t + e)
      uint32_t pa[4], da[4];
#pragma unroll
      for (int nq = 0; nq < 2; ++nq) {
        float pv[4], dsv[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int qi = q0 + nq * 8 + 2 * t + e;
          const float l2 = lse_s[qi], dl = delta_s[qi];
          const bool qok = qi < p.Lq;
          const bool ok0 = qok && key0 < Lk && (!p.causal || key0 <= qi);
          const bool ok1 = qok && key1 < Lk && (!p.causal || key1 <= qi);
          const float p0 = ok0 ? exp2f(st[nq][e] * sl2 - l2) : 0.f;
          const float p1 = ok1 ? exp2f(st[nq][2 + e] * sl2 - l2) : 0.f;
          pv[e] = p0;
          pv[2 + e] = p1;
          dsv[e] = p0 * (dp[nq][e] - dl) * p.scale;
          dsv[2 + e] = p1 * (dp[nq][2 + e] - dl) * p.scale;
        }
        pa[2 * nq] = pack_bf16x2(pv[0], pv[1]);
        pa[2 * nq + 1] = pack_bf16x2(pv[2], pv[3]);
        da[2 * nq] = pack_bf16x2(dsv[0], dsv[1]);
        da[2 * nq + 1] = pack_bf16x2(dsv[2], dsv[3]);
        st_shared_u32(sdS + key0 * ds_pitch + (q0 + nq * 8 + 2 * t) * 2, da[2 * nq]);
        st_shared_u32(sdS + key1 * ds_pitch + (q0 + nq * 8 + 2 * t) * 2, da[2 * nq + 1]);
      }
      // dV += P^T dO ; dK += dS^T Q   (A from registers: k index = query)
#pragma unroll
      for (int dpi = 0; dpi < 4; ++dpi) {
        uint32_t bo[4], bq[4];
        const uint32_t off = swz(q0 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * dpi + (lane >> 4));
        ldsm_x4_t(bo, sdO + off);
        ldsm_x4_t(bq, sQ + off);
        mma16816(dv[2 * dpi], pa, bo[0], bo[1]);
        mma16816(dv[2 * dpi + 1], pa, bo[2], bo[3]);
        mma16816(dk[2 * dpi], da, bq[0], bq[1]);
        mma16816(dk[2 * dpi + 1], da, bq[2], bq[3]);
      }
    }
    // ---- write dV then dK for this key tile; staging buffer = this warp's (private) V tile rows
    write_kv_grad(p, gc, dv, /*part=*/2, /*cls_slot=*/1, sV, kt, key0, key1, lane, t, active, (staged && p.accumulate_kv) ? sOldV : 0u, cls_red);
    write_kv_grad(p, gc, dk, /*part=*/1, /*cls_slot=*/0, sV, kt, key0, key1, lane, t, active, (staged && p.accumulate_kv) ? sOldK : 0u, cls_red);
  }
  __syncthreads();
  // ------------------------------------------------------------ phase 2: query tiles -> dQ = dS K
  for (int qt = w_in_g; qt < p.qt; qt += p.wg) {
    float dq[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) dq[dt][0] = dq[dt][1] = dq[dt][2] = dq[dt][3] = 0.f;
    for (int kk = 0; kk * 16 < k_rows; ++kk) {
      if (kk * 16 >= Lk) break;
      uint32_t a[4];
      ldsm_x4_t(a, sdS + (kk * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * ds_pitch + (qt * 16 + ((lane >> 3) & 1) * 8) * 2);
#pragma unroll
      for (int dpi = 0; dpi < 4; ++dpi) {
        uint32_t bk[4];
        ldsm_x4_t(bk, sK + swz(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * dpi + (lane >> 4)));
        mma16816(dq[2 * dpi], a, bk[0], bk[1]);
        mma16816(dq[2 * dpi + 1], a, bk[2], bk[3]);
      }
    }
    const int r0 = qt * 16 + g, r1 = r0 + 8;
    __syncwarp();
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      st_shared_u32(sQ + swz(r0, dt) + 4 * t, pack_bf16x2(dq[dt][0], dq[dt][1]));
      st_shared_u32(sQ + swz(r1, dt) + 4 * t, pack_bf16x2(dq[dt][2], dq[dt][3]));
    }
    __syncwarp();
    if (active) {
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int idx = it * 32 + lane, r = idx >> 3, c = idx & 7;
        const int qrow = qt * 16 + r;
        if (qrow < p.Lq) {
          const uint4 v = ld_shared_v4(sQ + swz(qrow, c));
          *reinterpret_cast<uint4*>(p.dqkv + (gc.base_row + (long long)qrow * p.row_stride) * p.ld_dqkv + gc.h * HD + c * 8) = v;
        }
      }
    }
  }
  // ------------------------------------------------------------ CLS key/value gradient: per-CTA reduction, then atomics
  if (p.has_cls && p.dcls_kv) {
    __syncthreads();
    const float* red_all = reinterpret_cast<const float*>(smem + p.groups_per_cta * p.group_bytes);
    const int jb_count = (p.inner + p.groups_per_cta - 1) / p.groups_per_cta;
    const int jb = (int)((blockIdx.x / p.H) % jb_count);
    int n_active = p.inner - jb * p.groups_per_cta;
    if (n_active > p.groups_per_cta) n_active = p.groups_per_cta;
    const long long bh = (long long)(blockIdx.x / ((long long)p.H * jb_count)) * p.H + (blockIdx.x % p.H);
    for (int e = threadIdx.x; e < 2 * HD; e += blockDim.x) {
      float run = 0.f;
      for (int gl = 0; gl < n_active; ++gl) run += red_all[gl * 2 * HD + e];
      atomicAdd(p.dcls_kv + bh * 2 * HD + e, run);
    }
  }
}

// ================================================================================================ CLS query attention
// One CTA per (clip, head).  8 lanes share a key row (one 16-byte unit = 8 dims each) so every global access is a
// coalesced 128-byte row; a warp walks 4 keys per step, the CTA 32.  Single pass over K and V:
//   forward : online softmax per 8-lane group, partial (m, l, acc) merged through shared memory;
//   backward: p_j from the saved lse, ds_j = p_j (dO.v_j - delta); dk_j, dv_j written in the same pass; dq reduced.
constexpr int CLS_THREADS = 256;
constexpr int CLS_GROUPS = CLS_THREADS / 8;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ float oct_sum(float v) {  // sum over the 8 lanes that share a key row
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v + __shfl_xor_sync(0xffffffffu, v, 4);
}

// Addressing (elements): query of clip b at q + b*q_bs; key/value j of clip b at k|v + (b*N + j)*ld_kv;
// output at out + b*o_bs; lse at lse + b*l_bs  (+ h*64 / + h for the head).
struct ClsAddr {
  const __nv_bfloat16 *q, *k, *v;
  long long q_bs, ld_kv;
  __nv_bfloat16* out;
  long long o_bs;
  float* lse;
  long long l_bs;
  // backward
  const __nv_bfloat16* dout;
  long long do_bs;
  __nv_bfloat16 *dq, *dk, *dv;
  long long dq_bs, ld_dkv;
  float* dcls_kv;       // when non-null the CLS key/value row (j = 0) goes here in fp32 instead of dk/dv row 0
  int accumulate;       // add to what is already in dk / dv / dcls_kv (the group backward ran first)
};

__global__ void __launch_bounds__(CLS_THREADS)
cls_attn_fwd_kernel(const ClsAddr a, int H, int N, float scale) {
  __shared__ float s_m[CLS_GROUPS], s_l[CLS_GROUPS], s_acc[CLS_GROUPS][HD];
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const long long row0 = (long long)b * N;
  const int tid = threadIdx.x, grp = tid >> 3, c = tid & 7;
  float q[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(a.q + b * a.q_bs + h * HD + c * 8)), q);
  float m = -INFINITY, l = 0.f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  constexpr int UNR = 4;   // keys in flight per thread: the pass is HBM-latency-bound otherwise
  const int iters = (N + CLS_GROUPS * UNR - 1) / (CLS_GROUPS * UNR);
  for (int it = 0; it < iters; ++it) {
    uint4 kr[UNR], vr[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = (it * UNR + u) * CLS_GROUPS + grp;
      const long long roff = (row0 + (j < N ? j : 0)) * a.ld_kv + h * HD + c * 8;
      kr[u] = __ldg(reinterpret_cast<const uint4*>(a.k + roff));
      vr[u] = __ldg(reinterpret_cast<const uint4*>(a.v + roff));
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = (it * UNR + u) * CLS_GROUPS + grp;
      float k[8], v[8];
      unpack8(kr[u], k);
      unpack8(vr[u], v);
      float sp = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) sp += q[e] * k[e];
      const float sc = oct_sum(sp) * scale;
      if (j < N) {
        const float mn = fmaxf(m, sc);
        const float corr = __expf(m - mn), pj = __expf(sc - mn);
        l = l * corr + pj;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = acc[e] * corr + pj * v[e];
        m = mn;
      }
    }
  }
  if (c == 0) { s_m[grp] = m; s_l[grp] = l; }
#pragma unroll
  for (int e = 0; e < 8; ++e) s_acc[grp][c * 8 + e] = acc[e];
  __syncthreads();
  if (tid < HD) {
    float M = -INFINITY;
    for (int g2 = 0; g2 < CLS_GROUPS; ++g2) M = fmaxf(M, s_m[g2]);
    float Lsum = 0.f, o = 0.f;
    for (int g2 = 0; g2 < CLS_GROUPS; ++g2) {
      const float w = __expf(s_m[g2] - M);
      Lsum += s_l[g2] * w;
      o += s_acc[g2][tid] * w;
    }
    a.out[b * a.o_bs + h * HD + tid] = __float2bfloat16_rn(o / Lsum);
    if (tid == 0) a.lse[b * a.l_bs + h] = M + logf(Lsum);
  }
}

// Backward of the CLS query attention.  Writes: dqkv[cls row, q part]; dqkv[rows 1..N-1, k and v parts] (this query's
// contribution; the group backward adds its own on top); dcls_kv[b,h] (fp32) = contribution to the CLS key/value.
__global__ void __launch_bounds__(CLS_THREADS)
cls_attn_bwd_kernel(const ClsAddr a, int H, int N, float scale) {
  __shared__ float s_dq[CLS_GROUPS][HD];
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const long long row0 = (long long)b * N;
  const int tid = threadIdx.x, grp = tid >> 3, c = tid & 7;
  float q[8], dO[8], o[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(a.q + b * a.q_bs + h * HD + c * 8)), q);
  unpack8(__ldg(reinterpret_cast<const uint4*>(a.dout + b * a.do_bs + h * HD + c * 8)), dO);
  unpack8(__ldg(reinterpret_cast<const uint4*>(a.out + b * a.o_bs + h * HD + c * 8)), o);
  float dpart = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) dpart += dO[e] * o[e];
  const float delta = oct_sum(dpart);
  const float L = a.lse[b * a.l_bs + h];
  float dq[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  constexpr int UNR = 4;
  const int iters = (N + CLS_GROUPS * UNR - 1) / (CLS_GROUPS * UNR);
  for (int it = 0; it < iters; ++it) {
    uint4 kr[UNR], vr[UNR], okr[UNR], ovr[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = (it * UNR + u) * CLS_GROUPS + grp;
      const long long jj = row0 + (j < N ? j : 0);
      const long long roff = jj * a.ld_kv + h * HD + c * 8;
      kr[u] = __ldg(reinterpret_cast<const uint4*>(a.k + roff));
      vr[u] = __ldg(reinterpret_cast<const uint4*>(a.v + roff));
      if (a.accumulate) {
        const long long doff = jj * a.ld_dkv + h * HD + c * 8;
        okr[u] = *reinterpret_cast<const uint4*>(a.dk + doff);
        ovr[u] = *reinterpret_cast<const uint4*>(a.dv + doff);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int j = (it * UNR + u) * CLS_GROUPS + grp;
      float k[8], v[8];
      unpack8(kr[u], k);
      unpack8(vr[u], v);
      float sp = 0.f, dpp = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) { sp += q[e] * k[e]; dpp += dO[e] * v[e]; }
      const float sc = oct_sum(sp) * scale;
      const float dpj = oct_sum(dpp);
      if (j < N) {
        const float pj = __expf(sc - L);
        const float dsj = pj * (dpj - delta) * scale;   // d loss / d (q.k_j)
        float dk[8], dv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          dq[e] += dsj * k[e];
          dk[e] = dsj * q[e];
          dv[e] = pj * dO[e];
        }
        if (j == 0 && a.dcls_kv) {
          float* kb = a.dcls_kv + ((long long)b * H + h) * 2 * HD + c * 8;
          if (a.accumulate) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { kb[e] += dk[e]; kb[HD + e] += dv[e]; }
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) { kb[e] = dk[e]; kb[HD + e] = dv[e]; }
          }
        } else {
          const long long doff = (row0 + j) * a.ld_dkv + h * HD + c * 8;
          if (a.accumulate) {   // streaming read-modify-write: the group backward has already written its part
            float ok_[8], ov_[8];
            unpack8(okr[u], ok_);
            unpack8(ovr[u], ov_);
#pragma unroll
            for (int e = 0; e < 8; ++e) { dk[e] += ok_[e]; dv[e] += ov_[e]; }
          }
          *reinterpret_cast<uint4*>(a.dk + doff) = make_uint4(pack_bf16x2(dk[0], dk[1]), pack_bf16x2(dk[2], dk[3]),
                                                              pack_bf16x2(dk[4], dk[5]), pack_bf16x2(dk[6], dk[7]));
          *reinterpret_cast<uint4*>(a.dv + doff) = make_uint4(pack_bf16x2(dv[0], dv[1]), pack_bf16x2(dv[2], dv[3]),
                                                              pack_bf16x2(dv[4], dv[5]), pack_bf16x2(dv[6], dv[7]));
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) s_dq[grp][c * 8 + e] = dq[e];
  __syncthreads();
  if (tid < HD) {
    float t = 0.f;
    for (int g2 = 0; g2 < CLS_GROUPS; ++g2) t += s_dq[g2][tid];
    a.dq[b * a.dq_bs + h * HD + tid] = __float2bfloat16_rn(t);
  }
}

// dqkv[cls row, k|v part] = bf16(dcls_kv)
__global__ void cls_kv_finalize_kernel(const float* __restrict__ dcls_kv, __nv_bfloat16* __restrict__ dqkv, long long lddq,
                                       int H, int D, int N) {
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int tid = threadIdx.x;  // 128: [k 64 | v 64]
  const int part = tid >> 6, d = tid & 63;
  dqkv[(long long)b * N * lddq + (1 + part) * D + h * HD + d] =
      __float2bfloat16_rn(dcls_kv[(((long long)b * H + h) * 2 + part) * HD + d]);
}

}  // namespace attn
}  // namespace lv

using namespace lv;

namespace {

// LV_TIME_ATTN_GENERIC=1 routes time attention through the generic group kernels (A/B measurements only)
bool generic_time_attn() {
  static const bool v = [] { const char* e = getenv("LV_TIME_ATTN_GENERIC"); return e && e[0] == '1'; }();
  return v;
}

// mode: 0 = space, 1 = time, 2 = causal text
int fill_params(attn::Params& p, int mode, int B, int H, int T, int n, int L) {
  p.H = H;
  p.D = H * attn::HD;
  if (mode == 0) { p.Lq = n; p.has_cls = 1; p.causal = 0; p.row_stride = 1; p.inner = T; p.inner_stride = n; p.first = 1; p.clip_rows = 1 + (long long)T * n; }
  else if (mode == 1) { p.Lq = T; p.has_cls = 1; p.causal = 0; p.row_stride = n; p.inner = n; p.inner_stride = 1; p.first = 1; p.clip_rows = 1 + (long long)T * n; }
  else if (mode == 2) { p.Lq = L; p.has_cls = 0; p.causal = 1; p.row_stride = 1; p.inner = 1; p.inner_stride = 0; p.first = 0; p.clip_rows = L; }
  else return set_error(-1, "attention: unknown mode %d", mode);
  p.num_groups = (long long)B * H * p.inner;
  p.qt = (p.Lq + 15) / 16;
  p.scale = 1.0f / 8.0f;  // head_dim ** -0.5 (timesformer.py:94, nn.MultiheadAttention)
  return 0;
}

template <int NT>
int launch_fwd(attn::Params& p, cudaStream_t st) {
  const int k_rows = NT * 8;
  const int per_group = (p.qt * 16 + 2 * k_rows) * attn::ROW_BYTES;
  int wg = (p.qt + 1) / 2;  // two query tiles per warp
  if (wg > 8) wg = 8;
  if (wg < 1) wg = 1;
  int gpc = 8 / wg;
  while (gpc > 1 && gpc * per_group > 100 * 1024) gpc >>= 1;
  p.wg = wg;
  p.groups_per_cta = gpc;
  const int smem = gpc * per_group;
  static int configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(attn::group_attn_fwd_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error((int)e, "attention fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = 227 * 1024;
  }
  LV_REQUIRE(smem <= 227 * 1024, "attention fwd: group does not fit in shared memory (%d bytes)", smem);
  const long long grid = (long long)(p.num_groups / p.inner) * ((p.inner + gpc - 1) / gpc);   // B*H CTAs per inner block
  attn::group_attn_fwd_kernel<NT><<<(unsigned)grid, gpc * wg * 32, smem, st>>>(p);
  return check_launch("lv_group_attn_fwd");
}

}  // namespace

extern "C" int lv_group_attn_fwd(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* lse, int mode, int B,
                                 int H, int T, int n, int L, void* stream) {
  LV_REQUIRE(qkv && out && lse && B > 0 && H > 0, "lv_group_attn_fwd: bad arguments");
  LV_REQUIRE(ld_qkv % 8 == 0 && ld_out % 8 == 0, "lv_group_attn_fwd: leading dimensions must be multiples of 8");
  if (mode == 1 && T <= 16 && !generic_time_attn())
    return time_attn_small_fwd(qkv, ld_qkv, out, ld_out, lse, B, H, T, n, (cudaStream_t)stream);
  attn::Params p{};
  int rc = fill_params(p, mode, B, H, T, n, L);
  if (rc) return rc;
  p.qkv = (const __nv_bfloat16*)qkv; p.ld_qkv = ld_qkv;
  p.out = (__nv_bfloat16*)out; p.ld_out = ld_out;
  p.lse = lse;
  const int Lk = p.Lq + p.has_cls;
  cudaStream_t st = (cudaStream_t)stream;
  if (Lk <= 32) return launch_fwd<4>(p, st);
  if (Lk <= 80) return launch_fwd<10>(p, st);
  if (Lk <= 208) return launch_fwd<26>(p, st);
  LV_REQUIRE(mode != 2, "lv_group_attn_fwd: causal groups of %d keys not supported (max 208)", Lk);
  return big_group_attn_fwd(qkv, ld_qkv, out, ld_out, lse, mode, B, H, T, n, st);   // TSF-L/14: 257 / 577 keys
}

extern "C" int lv_group_attn_bwd(const void* qkv, int64_t ld_qkv, const void* out, int64_t ld_out, const float* lse,
                                 const void* dout, int64_t ld_dout, void* dqkv, int64_t ld_dqkv, float* dcls_kv,
                                 int accumulate_kv, int mode, int B, int H, int T, int n, int L, void* stream) {
  LV_REQUIRE(qkv && out && lse && dout && dqkv && B > 0 && H > 0, "lv_group_attn_bwd: bad arguments");
  LV_REQUIRE(ld_qkv % 8 == 0 && ld_out % 8 == 0 && ld_dout % 8 == 0 && ld_dqkv % 8 == 0, "lv_group_attn_bwd: leading dimensions must be multiples of 8");
  if (mode == 1 && T <= 16 && !accumulate_kv && !generic_time_attn()) {
    LV_REQUIRE(dcls_kv, "lv_group_attn_bwd: dcls_kv required for CLS modes");
    return time_attn_small_bwd(qkv, ld_qkv, out, ld_out, lse, dout, ld_dout, dqkv, ld_dqkv, dcls_kv, B, H, T, n, (cudaStream_t)stream);
  }
  attn::Params p{};
  int rc = fill_params(p, mode, B, H, T, n, L);
  if (rc) return rc;
  LV_REQUIRE(!p.has_cls || dcls_kv, "lv_group_attn_bwd: dcls_kv required for CLS modes");
  p.qkv = (const __nv_bfloat16*)qkv; p.ld_qkv = ld_qkv;
  p.out = (__nv_bfloat16*)out; p.ld_out = ld_out;
  p.lse = (float*)lse;
  p.dout = (const __nv_bfloat16*)dout; p.ld_dout = ld_dout;
  p.dqkv = (__nv_bfloat16*)dqkv; p.ld_dqkv = ld_dqkv;
  p.dcls_kv = dcls_kv;
  p.accumulate_kv = accumulate_kv;
  const int Lk = p.Lq + p.has_cls;
  const int k_rows = (Lk + 15) / 16 * 16;
  const int q_rows = p.qt * 16;
  int per_group = 2 * q_rows * attn::ROW_BYTES + 2 * k_rows * attn::ROW_BYTES + 2 * q_rows * 4 + k_rows * (q_rows + 8) * 2;
  p.staged = q_rows <= 32 ? 1 : 0;
  if (p.staged) per_group += q_rows * attn::ROW_BYTES + 2 * k_rows * attn::ROW_BYTES;
  p.group_bytes = per_group;
  if (per_group + 512 > 227 * 1024) {   // TSF-L/14 groups (257 / 577 keys): key-tiled kernels
    LV_REQUIRE(mode != 2 && !accumulate_kv, "lv_group_attn_bwd: groups of %d keys need mode 0/1 and accumulate_kv = 0", Lk);
    return big_group_attn_bwd(qkv, ld_qkv, out, ld_out, lse, dout, ld_dout, dqkv, ld_dqkv, dcls_kv, mode, B, H, T, n,
                              (cudaStream_t)stream);
  }
  int wg = (k_rows / 16 + 1) / 2;
  if (wg > 8) wg = 8;
  if (wg < 1) wg = 1;
  int gpc = 8 / wg;
  while (gpc > 1 && gpc * (per_group + 512) > 113 * 1024) gpc >>= 1;
  p.wg = wg;
  p.groups_per_cta = gpc;
  const int smem = gpc * (per_group + 2 * attn::HD * 4);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn::group_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error((int)e, "attention bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const long long grid = (long long)(p.num_groups / p.inner) * ((p.inner + gpc - 1) / gpc);
  attn::group_attn_bwd_kernel<<<(unsigned)grid, gpc * wg * 32, smem, (cudaStream_t)stream>>>(p, k_rows);
  return check_launch("lv_group_attn_bwd");
}

extern "C" int lv_cls_attn_fwd(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* lse, int B, int H, int N,
                               void* stream) {
  LV_REQUIRE(qkv && out && lse && B > 0 && H > 0 && N > 0, "lv_cls_attn_fwd: bad arguments");
  LV_REQUIRE(ld_qkv % 8 == 0, "lv_cls_attn_fwd: ld_qkv must be a multiple of 8");
  attn::ClsAddr a{};
  const __nv_bfloat16* base = (const __nv_bfloat16*)qkv;
  const long long D = (long long)H * attn::HD;
  a.q = base; a.k = base + D; a.v = base + 2 * D; a.q_bs = (long long)N * ld_qkv; a.ld_kv = ld_qkv;
  a.out = (__nv_bfloat16*)out; a.o_bs = (long long)N * ld_out; a.lse = lse; a.l_bs = (long long)N * H;
  attn::cls_attn_fwd_kernel<<<B * H, attn::CLS_THREADS, 0, (cudaStream_t)stream>>>(a, H, N, 0.125f);
  return check_launch("lv_cls_attn_fwd");
}

extern "C" int lv_cls_attn_bwd(const void* qkv, int64_t ld_qkv, const void* out, int64_t ld_out, const void* dout,
                               int64_t ld_dout, const float* lse, void* dqkv, int64_t ld_dqkv, float* dcls_kv, int accumulate,
                               int B, int H, int N, void* stream) {
  LV_REQUIRE(qkv && out && dout && lse && dqkv && dcls_kv && B > 0 && H > 0 && N > 0, "lv_cls_attn_bwd: bad arguments");
  LV_REQUIRE(ld_qkv % 8 == 0 && ld_dqkv % 8 == 0, "lv_cls_attn_bwd: leading dimensions must be multiples of 8");
  attn::ClsAddr a{};
  const __nv_bfloat16* base = (const __nv_bfloat16*)qkv;
  __nv_bfloat16* dbase = (__nv_bfloat16*)dqkv;
  const long long D = (long long)H * attn::HD;
  a.q = base; a.k = base + D; a.v = base + 2 * D; a.q_bs = (long long)N * ld_qkv; a.ld_kv = ld_qkv;
  a.out = (__nv_bfloat16*)out; a.o_bs = (long long)N * ld_out; a.lse = (float*)lse; a.l_bs = (long long)N * H;
  a.dout = (const __nv_bfloat16*)dout; a.do_bs = (long long)N * ld_dout;
  a.dq = dbase; a.dq_bs = (long long)N * ld_dqkv; a.dk = dbase + D; a.dv = dbase + 2 * D; a.ld_dkv = ld_dqkv;
  a.dcls_kv = dcls_kv;
  a.accumulate = accumulate;
  attn::cls_attn_bwd_kernel<<<B * H, attn::CLS_THREADS, 0, (cudaStream_t)stream>>>(a, H, N, 0.125f);
  return check_launch("lv_cls_attn_bwd");
}

extern "C" int lv_cls_kv_finalize(const float* dcls_kv, void* dqkv, int64_t ld_dqkv, int B, int H, int N, void* stream) {
  LV_REQUIRE(dcls_kv && dqkv && B > 0 && H > 0, "lv_cls_kv_finalize: bad arguments");
  attn::cls_kv_finalize_kernel<<<B * H, 128, 0, (cudaStream_t)stream>>>(dcls_kv, (__nv_bfloat16*)dqkv, ld_dqkv, H, H * attn::HD, N);
  return check_launch("lv_cls_kv_finalize");
}

// CLS query attention with separate q [B, D], packed kv [B*N, 2D] = [k | v] (the last block's CLS-only tail: only the
// CLS row of the final SpaceTimeBlock is consumed by norm(x)[:, 0], lavila/models/timesformer.py:376-378).
extern "C" int lv_cls_query_attn_fwd(const void* q, const void* kv, void* out, float* lse, int B, int H, int N, void* stream) {
  LV_REQUIRE(q && kv && out && lse && B > 0 && H > 0 && N > 0, "lv_cls_query_attn_fwd: bad arguments");
  attn::ClsAddr a{};
  const long long D = (long long)H * attn::HD;
  a.q = (const __nv_bfloat16*)q; a.q_bs = D;
  a.k = (const __nv_bfloat16*)kv; a.v = a.k + D; a.ld_kv = 2 * D;
  a.out = (__nv_bfloat16*)out; a.o_bs = D; a.lse = lse; a.l_bs = H;
  attn::cls_attn_fwd_kernel<<<B * H, attn::CLS_THREADS, 0, (cudaStream_t)stream>>>(a, H, N, 0.125f);
  return check_launch("lv_cls_query_attn_fwd");
}

extern "C" int lv_cls_query_attn_bwd(const void* q, const void* kv, const void* out, const void* dout, const float* lse,
                                     void* dq, void* dkv, int B, int H, int N, void* stream) {
  LV_REQUIRE(q && kv && out && dout && lse && dq && dkv && B > 0 && H > 0 && N > 0, "lv_cls_query_attn_bwd: bad arguments");
  attn::ClsAddr a{};
  const long long D = (long long)H * attn::HD;
  a.q = (const __nv_bfloat16*)q; a.q_bs = D;
  a.k = (const __nv_bfloat16*)kv; a.v = a.k + D; a.ld_kv = 2 * D;
  a.out = (__nv_bfloat16*)out; a.o_bs = D; a.lse = (float*)lse; a.l_bs = H;
  a.dout = (const __nv_bfloat16*)dout; a.do_bs = D;
  a.dq = (__nv_bfloat16*)dq; a.dq_bs = D;
  a.dk = (__nv_bfloat16*)dkv; a.dv = a.dk + D; a.ld_dkv = 2 * D;
  a.dcls_kv = nullptr;
  attn::cls_attn_bwd_kernel<<<B * H, attn::CLS_THREADS, 0, (cudaStream_t)stream>>>(a, H, N, 0.125f);
  return check_launch("lv_cls_query_attn_bwd");
}
