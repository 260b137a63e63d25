// Time attention of the divided space-time block (lavila/models/timesformer.py:126-133), specialised for small groups:
// one group = the <= 16 frames of one (clip, head, spatial position) plus the clip's CLS key/value (<= 17 keys).
//
// The generic group kernel (attention.cu) gives such a group a whole warp and a load -> compute -> store life with
// nothing in flight while it computes; at 8-12 resident warps per SM that is latency-bound (0.33-0.53 of the HBM peak).
// Here a CTA is persistent and owns a contiguous range of (clip, position) units; warp w of the CTA is head w, so the
// CTA reads whole 3*D-wide token rows (q|k|v of all its heads are contiguous) and every warp runs its own two-stage
// cp.async pipeline: the tiles of unit u+1 are in flight while unit u is computed and stored.  Scores live in
// registers (queries are the MMA M dimension: 16 queries x 24 key slots); P^T / dS^T operands are produced with
// movmatrix, so the backward needs no shared-memory round trip.  The CLS key/value gradient of a head is accumulated
// in registers across the units of a clip and flushed with one set of atomics per (clip, head, CTA).
//
// Key slots: 0..Lq-1 = the group's own tokens, slot 16 = CLS; slots Lq..15 and 17..23 are masked.
#include "../../include/lavila_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace lv {
namespace tattn {

constexpr int HD = 64;
constexpr int ROW_BYTES = 128;
constexpr float LOG2E = 1.4426950408889634f;
constexpr int MAX_WARPS = 12;
constexpr int Q_TILE = 16 * ROW_BYTES;         // 16 query rows
constexpr int KV_TILE = 17 * ROW_BYTES;        // 16 token rows + the CLS row
constexpr int FWD_STAGE = Q_TILE + 2 * KV_TILE;
constexpr int BWD_STAGE = 2 * Q_TILE + 2 * KV_TILE;
constexpr int CLS_SCRATCH = 256;                      // q_cls [64] bf16 | dO_cls [64] bf16 of the warp's (clip, head) (fused CLS query)
constexpr int FWD_WARP_BYTES = 2 * FWD_STAGE + CLS_SCRATCH;
constexpr int BWD_WARP_BYTES = 2 * BWD_STAGE + 128 + CLS_SCRATCH;   // + delta[16] fp32 (padded)

struct Params {
  const __nv_bfloat16* qkv; long long ld_qkv;
  __nv_bfloat16* out; long long ld_out;          // fwd: written; bwd: the forward output (read)
  float* lse;
  const __nv_bfloat16* dout; long long ld_dout;
  __nv_bfloat16* dqkv; long long ld_dqkv;
  float* dcls_kv;
  // CLS-query fusion (optional).  The clip's CLS query attends to every token (timesformer.py:116-119); its attention over the
  // <= 16 keys of a unit (the CLS key counted in the unit of spatial position 0 only) is evaluated next to the group:
  //   forward : partial (max, sum, unnormalised output[64]) per (clip, head, position) -> cls_part [B*H*n][66], merged by
  //             time_cls_combine_kernel;
  //   backward: the CLS query's contribution is added to the unit's dV / dK fragments before they are stored, dQ_cls is
  //             accumulated in registers over the units of a clip and flushed into dcls_q fp32 [B][H][64].
  float* cls_part;
  float* dcls_q;
  int H, D, Lq, n, hc, hchunks;
  long long clip_rows;
  int units;                                     // B * hchunks * n
  float scale;
};

__device__ __forceinline__ uint32_t swz(int row, int chunk) { return row * ROW_BYTES + ((chunk ^ (row & 7)) << 4); }
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void st_shared_zero16(uint32_t addr) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
}
__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
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
// transpose an 8x8 bf16 matrix held one row-pair per lane (the C-fragment packing) -> A-fragment of the transpose
__device__ __forceinline__ uint32_t movm_t(uint32_t v) {
  uint32_t r;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ int kv_row(int r) { return r > 16 ? 16 : r; }   // key slots past the CLS row alias it (masked)

struct Unit {
  int bh;                 // clip * hchunks + head chunk (CLS accumulators are per bh)
  int h;                  // this warp's head
  long long base_row;     // first token row of the group
  long long cls_row;
};
__device__ __forceinline__ Unit decode(const Params& p, int u, int warp) {
  Unit c;
  c.bh = u / p.n;
  const int j = u - c.bh * p.n;
  const int b = c.bh / p.hchunks, hck = c.bh - b * p.hchunks;
  c.h = hck * p.hc + warp;
  c.cls_row = (long long)b * p.clip_rows;
  c.base_row = c.cls_row + 1 + j;
  return c;
}

// 16 (or 17 with the CLS row) rows x 128 bytes -> swizzled tile, 16 bytes per cp.async
__device__ __forceinline__ void load_rows(uint32_t tile, int Lq, const __nv_bfloat16* src, long long ld, long long base_row,
                                          long long row_stride, const __nv_bfloat16* cls_ptr, int lane) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = it * 32 + lane, r = idx >> 3, c = idx & 7;
    if (r < Lq) cp_async16(tile + swz(r, c), src + (base_row + (long long)r * row_stride) * ld + c * 8);
  }
  if (cls_ptr && lane < 8) cp_async16(tile + swz(16, lane), cls_ptr + lane * 8);
}

// lane = (key = lane & 15, half = lane >> 4): partial dot of tile row `key` with a 64-element bf16 vector over its 32-element half
__device__ __forceinline__ float half_row_dot(uint32_t tile, int key, int half, uint32_t vec) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 a = ld_shared_v4(tile + swz(key, half * 4 + c));
    const uint4 b = ld_shared_v4(vec + (half * 4 + c) * 16);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = unpack_bf16x2(aw[j]), y = unpack_bf16x2(bw[j]);
      acc += x.x * y.x + x.y * y.y;
    }
  }
  return acc + __shfl_xor_sync(0xffffffffu, acc, 16);
}
__device__ __forceinline__ float warp_max16(float v) {   // over the 16 lanes that share (lane >> 4)
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum16(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float2 ld_shared_bf16x2(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return unpack_bf16x2(v);
}

// C-fragment tile (16 rows x 64 cols fp32) -> bf16 staging tile -> global rows (128 bytes per row, coalesced)
__device__ __forceinline__ void stage_and_store(uint32_t tile, float (&acc)[8][4], int Lq, __nv_bfloat16* dst, long long ld,
                                                long long base_row, long long row_stride, int lane) {
  const int g = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    st_shared_u32(tile + swz(g, dt) + 4 * t, pack_bf16x2(acc[dt][0], acc[dt][1]));
    st_shared_u32(tile + swz(g + 8, dt) + 4 * t, pack_bf16x2(acc[dt][2], acc[dt][3]));
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = it * 32 + lane, r = idx >> 3, c = idx & 7;
    if (r < Lq) {
      const uint4 v = ld_shared_v4(tile + swz(r, c));
      *reinterpret_cast<uint4*>(dst + (base_row + (long long)r * row_stride) * ld + c * 8) = v;
    }
  }
}

// ================================================================================================ forward
__global__ void __launch_bounds__(MAX_WARPS * 32, 1)
time_attn_fwd_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t ws = smem_u32(smem) + warp * FWD_WARP_BYTES;
  const int u0 = (int)((long long)blockIdx.x * p.units / gridDim.x), u1 = (int)((long long)(blockIdx.x + 1) * p.units / gridDim.x);
  if (p.Lq < 16) {   // rows that no load ever touches must be finite (zero)
    for (int st = 0; st < 2; ++st)
      for (int idx = lane; idx < 16 * 8; idx += 32) {
        const int r = idx >> 3, c = idx & 7;
        if (r >= p.Lq) {
          st_shared_zero16(ws + st * FWD_STAGE + swz(r, c));
          st_shared_zero16(ws + st * FWD_STAGE + Q_TILE + swz(r, c));
          st_shared_zero16(ws + st * FWD_STAGE + Q_TILE + KV_TILE + swz(r, c));
        }
      }
    __syncwarp();
  }
  const long long rs = p.n;
  const uint32_t sCls = ws + 2 * FWD_STAGE;     // q_cls of the current (clip, head)
  int cls_bh = -1;
  auto issue = [&](int u, int st) {
    const Unit c = decode(p, u, warp);
    const uint32_t sQ = ws + st * FWD_STAGE, sK = sQ + Q_TILE, sV = sK + KV_TILE;
    const __nv_bfloat16* qb = p.qkv + c.h * HD;
    load_rows(sQ, p.Lq, qb, p.ld_qkv, c.base_row, rs, nullptr, lane);
    load_rows(sK, p.Lq, qb + p.D, p.ld_qkv, c.base_row, rs, qb + c.cls_row * p.ld_qkv + p.D, lane);
    load_rows(sV, p.Lq, qb + 2 * p.D, p.ld_qkv, c.base_row, rs, qb + c.cls_row * p.ld_qkv + 2 * p.D, lane);
  };
  if (u0 < u1) issue(u0, 0);
  cp_async_commit();
  const float sl2 = p.scale * LOG2E;
  int st = 0;
  for (int u = u0; u < u1; ++u, st ^= 1) {
    if (u + 1 < u1) issue(u + 1, st ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    const Unit c = decode(p, u, warp);
    const uint32_t sQ = ws + st * FWD_STAGE, sK = sQ + Q_TILE, sV = sK + KV_TILE;
    uint32_t qf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(qf[ks], sQ + swz((lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4)));
    float s[3][4];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t kf[4];
        ldsm_x4(kf, sK + swz(kv_row(nt * 8 + (lane & 7)), 4 * kp + (lane >> 3)));
        mma16816(s[nt], qf[2 * kp], kf[0], kf[1]);
        mma16816(s[nt], qf[2 * kp + 1], kf[2], kf[3]);
      }
    }
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = nt * 8 + 2 * t + e;
        const bool ok = key < p.Lq || key == 16;
        s[nt][e] = ok ? s[nt][e] : -INFINITY;
        s[nt][2 + e] = ok ? s[nt][2 + e] : -INFINITY;
        m0 = fmaxf(m0, s[nt][e]);
        m1 = fmaxf(m1, s[nt][2 + e]);
      }
    m0 = quad_max(m0);
    m1 = quad_max(m1);
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[nt][e] = exp2f((s[nt][e] - m0) * sl2);
        s[nt][2 + e] = exp2f((s[nt][2 + e] - m1) * sl2);
        sum0 += s[nt][e];
        sum1 += s[nt][2 + e];
      }
    sum0 = quad_sum(sum0);
    sum1 = quad_sum(sum1);
    float o[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
    {
      const uint32_t a0[4] = {pack_bf16x2(s[0][0], s[0][1]), pack_bf16x2(s[0][2], s[0][3]),
                              pack_bf16x2(s[1][0], s[1][1]), pack_bf16x2(s[1][2], s[1][3])};
      const uint32_t a1[4] = {pack_bf16x2(s[2][0], s[2][1]), pack_bf16x2(s[2][2], s[2][3]), 0u, 0u};
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t vf[4];
        ldsm_x4_t(vf, sV + swz((lane & 7) + ((lane >> 3) & 1) * 8, 2 * dp + (lane >> 4)));
        mma16816(o[2 * dp], a0, vf[0], vf[1]);
        mma16816(o[2 * dp + 1], a0, vf[2], vf[3]);
        ldsm_x4_t(vf, sV + swz(kv_row(16 + (lane & 7) + ((lane >> 3) & 1) * 8), 2 * dp + (lane >> 4)));
        mma16816(o[2 * dp], a1, vf[0], vf[1]);
        mma16816(o[2 * dp + 1], a1, vf[2], vf[3]);
      }
    }
    const float inv0 = g < p.Lq ? 1.f / sum0 : 0.f, inv1 = g + 8 < p.Lq ? 1.f / sum1 : 0.f;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) { o[dt][0] *= inv0; o[dt][1] *= inv0; o[dt][2] *= inv1; o[dt][3] *= inv1; }
    if (p.cls_part) {
      // ---- fused CLS query: partial attention of q_cls over this unit's keys (+ the CLS key in the unit of position 0)
      if (c.bh != cls_bh) {
        cls_bh = c.bh;
        __syncwarp();
        if (lane < 8) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.qkv + c.cls_row * p.ld_qkv + c.h * HD + lane * 8));
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sCls + lane * 16), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
        __syncwarp();
      }
      const int key = lane & 15, half = lane >> 4;
      const bool first = (c.base_row - c.cls_row) == 1;                   // spatial position 0: this unit also owns the CLS key
      float sc = half_row_dot(sK, key, half, sCls);
      const float scc = first ? half_row_dot(sK, 16, half, sCls) : 0.f;   // q_cls . k_cls (warp-uniform branch)
      sc = key < p.Lq ? sc : -INFINITY;
      float mc = warp_max16(sc);
      if (first) mc = fmaxf(mc, scc);
      const float pc = exp2f((sc - mc) * sl2);                            // exp2(-inf) = 0 for the masked slots
      const float pcc = first ? exp2f((scc - mc) * sl2) : 0.f;
      const float lc = warp_sum16(pc) + pcc;
      float oc0 = 0.f, oc1 = 0.f;                                         // columns 2 * lane, 2 * lane + 1
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float pj = __shfl_sync(0xffffffffu, pc, j);
        const float2 v = ld_shared_bf16x2(sV + swz(j, lane >> 2) + (lane & 3) * 4);
        oc0 = fmaf(pj, v.x, oc0);
        oc1 = fmaf(pj, v.y, oc1);
      }
      {
        const float2 v = ld_shared_bf16x2(sV + swz(16, lane >> 2) + (lane & 3) * 4);
        oc0 = fmaf(pcc, v.x, oc0);
        oc1 = fmaf(pcc, v.y, oc1);
      }
      const long long pos = (c.base_row - c.cls_row - 1);
      float* dst = p.cls_part + (((c.cls_row / p.clip_rows) * p.H + c.h) * (long long)p.n + pos) * 66;
      if (lane == 0) { dst[0] = mc; dst[1] = lc; }
      *reinterpret_cast<float2*>(dst + 2 + 2 * lane) = make_float2(oc0, oc1);
    }
    stage_and_store(sQ, o, p.Lq, p.out + c.h * HD, p.ld_out, c.base_row, rs, lane);
    if (t == 0) {
      if (g < p.Lq) p.lse[(c.base_row + (long long)g * rs) * p.H + c.h] = m0 * p.scale + logf(sum0);
      if (g + 8 < p.Lq) p.lse[(c.base_row + (long long)(g + 8) * rs) * p.H + c.h] = m1 * p.scale + logf(sum1);
    }
    __syncwarp();   // every lane is done with this stage before the load two units ahead overwrites it
  }
}

// ================================================================================================ backward
__global__ void __launch_bounds__(MAX_WARPS * 32, 1)
time_attn_bwd_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t ws = smem_u32(smem) + warp * BWD_WARP_BYTES;
  const uint32_t sDelta = ws + 2 * BWD_STAGE;
  const int u0 = (int)((long long)blockIdx.x * p.units / gridDim.x), u1 = (int)((long long)(blockIdx.x + 1) * p.units / gridDim.x);
  if (p.Lq < 16) {
    for (int st = 0; st < 2; ++st)
      for (int idx = lane; idx < 16 * 8; idx += 32) {
        const int r = idx >> 3, c = idx & 7;
        if (r >= p.Lq) {
          const uint32_t b = ws + st * BWD_STAGE;
          st_shared_zero16(b + swz(r, c));
          st_shared_zero16(b + Q_TILE + swz(r, c));
          st_shared_zero16(b + 2 * Q_TILE + swz(r, c));
          st_shared_zero16(b + 2 * Q_TILE + KV_TILE + swz(r, c));
        }
      }
    __syncwarp();
  }
  const long long rs = p.n;
  auto issue = [&](const Unit& c, int st) {
    const uint32_t sQ = ws + st * BWD_STAGE, sdO = sQ + Q_TILE, sK = sdO + Q_TILE, sV = sK + KV_TILE;
    const __nv_bfloat16* qb = p.qkv + c.h * HD;
    load_rows(sQ, p.Lq, qb, p.ld_qkv, c.base_row, rs, nullptr, lane);
    load_rows(sdO, p.Lq, p.dout + c.h * HD, p.ld_dout, c.base_row, rs, nullptr, lane);
    load_rows(sK, p.Lq, qb + p.D, p.ld_qkv, c.base_row, rs, qb + c.cls_row * p.ld_qkv + p.D, lane);
    load_rows(sV, p.Lq, qb + 2 * p.D, p.ld_qkv, c.base_row, rs, qb + c.cls_row * p.ld_qkv + 2 * p.D, lane);
  };
  // the forward output (for delta) and lse travel through registers, one unit ahead like the tiles
  uint4 o_next[4];
  float lse_next[2];
  auto fetch_o = [&](const Unit& c) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = it * 32 + lane, r = idx >> 3, cc = idx & 7;
      o_next[it] = r < p.Lq ? __ldg(reinterpret_cast<const uint4*>(p.out + (c.base_row + (long long)r * rs) * p.ld_out + c.h * HD + cc * 8))
                            : make_uint4(0u, 0u, 0u, 0u);
    }
    lse_next[0] = g < p.Lq ? __ldg(p.lse + (c.base_row + (long long)g * rs) * p.H + c.h) * LOG2E : 0.f;
    lse_next[1] = g + 8 < p.Lq ? __ldg(p.lse + (c.base_row + (long long)(g + 8) * rs) * p.H + c.h) * LOG2E : 0.f;
  };
  // fused CLS query (p.dcls_q): q_cls | dO_cls of the current (clip, head) in the warp's scratch, lse / delta of the CLS row,
  // this lane's two columns of dQ_cls accumulated over the units of the clip
  const uint32_t sClsQ = ws + 2 * BWD_STAGE + 128, sClsDo = sClsQ + 128;
  int cls_bh = -1;
  float lse_c = 0.f, delta_c = 0.f, dqc0 = 0.f, dqc1 = 0.f;
  auto flush_dq = [&](const Unit& c) {
    if (p.dcls_q) {
      float* base = p.dcls_q + ((long long)(c.bh / p.hchunks) * p.H + c.h) * HD;
      atomicAdd(base + 2 * lane, dqc0);
      atomicAdd(base + 2 * lane + 1, dqc1);
      dqc0 = dqc1 = 0.f;
    }
  };
  float clsk[8][2], clsv[8][2];   // CLS key/value gradient of this warp's head, rows of lanes 0-3 only
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) clsk[dt][0] = clsk[dt][1] = clsv[dt][0] = clsv[dt][1] = 0.f;
  auto flush_cls = [&](const Unit& c) {
    if (g == 0) {
      float* base = p.dcls_kv + ((long long)(c.bh / p.hchunks) * p.H + c.h) * 2 * HD;
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        atomicAdd(base + dt * 8 + 2 * t, clsk[dt][0]);
        atomicAdd(base + dt * 8 + 2 * t + 1, clsk[dt][1]);
        atomicAdd(base + HD + dt * 8 + 2 * t, clsv[dt][0]);
        atomicAdd(base + HD + dt * 8 + 2 * t + 1, clsv[dt][1]);
      }
    }
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) clsk[dt][0] = clsk[dt][1] = clsv[dt][0] = clsv[dt][1] = 0.f;
  };

  Unit cur{};
  if (u0 < u1) {
    cur = decode(p, u0, warp);
    issue(cur, 0);
    fetch_o(cur);
  }
  cp_async_commit();
  const float sl2 = p.scale * LOG2E;
  int st = 0;
  for (int u = u0; u < u1; ++u, st ^= 1) {
    Unit nxt = cur;
    const float l0 = lse_next[0], l1 = lse_next[1];
    if (u + 1 < u1) {
      nxt = decode(p, u + 1, warp);
      issue(nxt, st ^ 1);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    const uint32_t sQ = ws + st * BWD_STAGE, sdO = sQ + Q_TILE, sK = sdO + Q_TILE, sV = sK + KV_TILE;
    // ---- delta_q = sum_d dO[q,d] O[q,d]
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = it * 32 + lane, r = idx >> 3, cc = idx & 7;
      const uint4 a = ld_shared_v4(sdO + swz(r, cc));
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {o_next[it].x, o_next[it].y, o_next[it].z, o_next[it].w};
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 x = unpack_bf16x2(aw[j]), y = unpack_bf16x2(bw[j]);
        part += x.x * y.x + x.y * y.y;
      }
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      if (cc == 0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sDelta + r * 4), "f"(part) : "memory");
    }
    __syncwarp();
    float d0, d1;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(d0) : "r"(sDelta + g * 4));
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(d1) : "r"(sDelta + (g + 8) * 4));
    // ---- S = Q K^T and dP = dO V^T  (16 queries x 24 key slots)
    float s[3][4], dp[3][4];
    {
      uint32_t qf[4][4], dof[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t off = swz((lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4));
        ldsm_x4(qf[ks], sQ + off);
        ldsm_x4(dof[ks], sdO + off);
      }
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
          uint32_t kf[4], vf[4];
          const uint32_t off = swz(kv_row(nt * 8 + (lane & 7)), 4 * kp + (lane >> 3));
          ldsm_x4(kf, sK + off);
          ldsm_x4(vf, sV + off);
          mma16816(s[nt], qf[2 * kp], kf[0], kf[1]);
          mma16816(s[nt], qf[2 * kp + 1], kf[2], kf[3]);
          mma16816(dp[nt], dof[2 * kp], vf[0], vf[1]);
          mma16816(dp[nt], dof[2 * kp + 1], vf[2], vf[3]);
        }
      }
    }
    // ---- P and dS (scaled) as packed bf16 pairs: [nt][0] = query row g, [nt][1] = query row g+8
    uint32_t pp[3][2], dsp[3][2];
    {
      const bool q0ok = g < p.Lq, q1ok = g + 8 < p.Lq;
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        float pv[4], dsv[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = nt * 8 + 2 * t + e;
          const bool ok = key < p.Lq || key == 16;
          const float p0 = (ok && q0ok) ? exp2f(s[nt][e] * sl2 - l0) : 0.f;
          const float p1 = (ok && q1ok) ? exp2f(s[nt][2 + e] * sl2 - l1) : 0.f;
          pv[e] = p0;
          pv[2 + e] = p1;
          dsv[e] = p0 * (dp[nt][e] - d0) * p.scale;
          dsv[2 + e] = p1 * (dp[nt][2 + e] - d1) * p.scale;
        }
        pp[nt][0] = pack_bf16x2(pv[0], pv[1]);
        pp[nt][1] = pack_bf16x2(pv[2], pv[3]);
        dsp[nt][0] = pack_bf16x2(dsv[0], dsv[1]);
        dsp[nt][1] = pack_bf16x2(dsv[2], dsv[3]);
      }
    }
    if (u + 1 < u1) fetch_o(nxt);   // in flight during the gradient MMAs and stores below
    float Pc = 0.f, dSc = 0.f, Pcc = 0.f, dScc = 0.f;     // CLS query x (key lane & 15 | the CLS key)
    if (p.dcls_q) {
      if (cur.bh != cls_bh) {
        cls_bh = cur.bh;
        __syncwarp();
        if (lane < 16) {
          const int part = lane >> 3, ch = lane & 7;
          const __nv_bfloat16* src = part ? p.dout + cur.cls_row * p.ld_dout + cur.h * HD + ch * 8
                                          : p.qkv + cur.cls_row * p.ld_qkv + cur.h * HD + ch * 8;
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"((part ? sClsDo : sClsQ) + ch * 16), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
        __syncwarp();
        // delta of the CLS row = dO_cls . O_cls ; lse in log2 units
        const float2 oc = unpack_bf16x2(__ldg(reinterpret_cast<const unsigned int*>(p.out + cur.cls_row * p.ld_out + cur.h * HD + 2 * lane)));
        const float2 dc = ld_shared_bf16x2(sClsDo + lane * 4);
        float part = oc.x * dc.x + oc.y * dc.y;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        delta_c = part;
        lse_c = __ldg(p.lse + cur.cls_row * p.H + cur.h) * LOG2E;
      }
      const int key = lane & 15, half = lane >> 4;
      const bool first = (cur.base_row - cur.cls_row) == 1;
      const float sc = half_row_dot(sK, key, half, sClsQ), dpc = half_row_dot(sV, key, half, sClsDo);
      const float scc = first ? half_row_dot(sK, 16, half, sClsQ) : 0.f;          // the CLS key pairs with the CLS query once per
      const float dpcc = first ? half_row_dot(sV, 16, half, sClsDo) : 0.f;        // clip: in the unit of position 0 (warp-uniform)
      Pc = key < p.Lq ? exp2f(sc * sl2 - lse_c) : 0.f;
      dSc = Pc * (dpc - delta_c) * p.scale;
      Pcc = first ? exp2f(scc * sl2 - lse_c) : 0.f;
      dScc = Pcc * (dpcc - delta_c) * p.scale;
      // dQ_cls += sum_j dS_j k_j (this lane: columns 2 * lane, 2 * lane + 1); K stays intact until the end of the unit
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float dsj = __shfl_sync(0xffffffffu, dSc, j);
        const float2 kk = ld_shared_bf16x2(sK + swz(j, lane >> 2) + (lane & 3) * 4);
        dqc0 = fmaf(dsj, kk.x, dqc0);
        dqc1 = fmaf(dsj, kk.y, dqc1);
      }
      {
        const float2 kk = ld_shared_bf16x2(sK + swz(16, lane >> 2) + (lane & 3) * 4);
        dqc0 = fmaf(dScc, kk.x, dqc0);
        dqc1 = fmaf(dScc, kk.y, dqc1);
      }
    }
    // rank-1 update of a gradient fragment (rows g, g + 8 = keys) with the CLS query's term: acc += w[key] * vec[col]
    auto add_cls_term = [&](float (&a)[8][4], float (&ck)[8][2], float w, float wcc, uint32_t vec) {
      const float w0 = __shfl_sync(0xffffffffu, w, g), w1 = __shfl_sync(0xffffffffu, w, g + 8);
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        const float2 x = ld_shared_bf16x2(vec + (dt * 8 + 2 * t) * 2);
        a[dt][0] = fmaf(w0, x.x, a[dt][0]);
        a[dt][1] = fmaf(w0, x.y, a[dt][1]);
        a[dt][2] = fmaf(w1, x.x, a[dt][2]);
        a[dt][3] = fmaf(w1, x.y, a[dt][3]);
        if (g == 0) {                  // the CLS key / value row lives in lanes 0-3
          ck[dt][0] = fmaf(wcc, x.x, ck[dt][0]);
          ck[dt][1] = fmaf(wcc, x.y, ck[dt][1]);
        }
      }
    };
    float acc[8][4];
    // ---- dV = P^T dO  (keys are the M dimension; the CLS slot is row 0 of a second, otherwise empty, key tile)
    {
      const uint32_t a[4] = {movm_t(pp[0][0]), movm_t(pp[1][0]), movm_t(pp[0][1]), movm_t(pp[1][1])};
      const uint32_t ac[4] = {movm_t(pp[2][0]), 0u, movm_t(pp[2][1]), 0u};
#pragma unroll
      for (int dpi = 0; dpi < 4; ++dpi) {
        uint32_t bo[4];
        ldsm_x4_t(bo, sdO + swz((lane & 7) + ((lane >> 3) & 1) * 8, 2 * dpi + (lane >> 4)));
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float* c4 = acc[2 * dpi + hh];
          c4[0] = c4[1] = c4[2] = c4[3] = 0.f;
          mma16816(acc[2 * dpi + hh], a, bo[2 * hh], bo[2 * hh + 1]);
          float tmp[4] = {0.f, 0.f, 0.f, 0.f};
          mma16816(tmp, ac, bo[2 * hh], bo[2 * hh + 1]);
          clsv[2 * dpi + hh][0] += tmp[0];
          clsv[2 * dpi + hh][1] += tmp[1];
        }
      }
    }
    if (p.dcls_q) add_cls_term(acc, clsv, Pc, Pcc, sClsDo);        // dV_j += P_cls[j] dO_cls
    stage_and_store(sV, acc, p.Lq, p.dqkv + 2 * p.D + cur.h * HD, p.ld_dqkv, cur.base_row, rs, lane);
    // ---- dK = dS^T Q
    {
      const uint32_t a[4] = {movm_t(dsp[0][0]), movm_t(dsp[1][0]), movm_t(dsp[0][1]), movm_t(dsp[1][1])};
      const uint32_t ac[4] = {movm_t(dsp[2][0]), 0u, movm_t(dsp[2][1]), 0u};
#pragma unroll
      for (int dpi = 0; dpi < 4; ++dpi) {
        uint32_t bq[4];
        ldsm_x4_t(bq, sQ + swz((lane & 7) + ((lane >> 3) & 1) * 8, 2 * dpi + (lane >> 4)));
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float* c4 = acc[2 * dpi + hh];
          c4[0] = c4[1] = c4[2] = c4[3] = 0.f;
          mma16816(acc[2 * dpi + hh], a, bq[2 * hh], bq[2 * hh + 1]);
          float tmp[4] = {0.f, 0.f, 0.f, 0.f};
          mma16816(tmp, ac, bq[2 * hh], bq[2 * hh + 1]);
          clsk[2 * dpi + hh][0] += tmp[0];
          clsk[2 * dpi + hh][1] += tmp[1];
        }
      }
    }
    if (p.dcls_q) add_cls_term(acc, clsk, dSc, dScc, sClsQ);       // dK_j += dS_cls[j] q_cls
    stage_and_store(sdO, acc, p.Lq, p.dqkv + p.D + cur.h * HD, p.ld_dqkv, cur.base_row, rs, lane);
    // ---- dQ = dS K
    {
      const uint32_t a0[4] = {dsp[0][0], dsp[0][1], dsp[1][0], dsp[1][1]};
      const uint32_t a1[4] = {dsp[2][0], dsp[2][1], 0u, 0u};
#pragma unroll
      for (int dpi = 0; dpi < 4; ++dpi) {
        uint32_t bk[4];
        ldsm_x4_t(bk, sK + swz((lane & 7) + ((lane >> 3) & 1) * 8, 2 * dpi + (lane >> 4)));
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float* c4 = acc[2 * dpi + hh];
          c4[0] = c4[1] = c4[2] = c4[3] = 0.f;
          mma16816(acc[2 * dpi + hh], a0, bk[2 * hh], bk[2 * hh + 1]);
        }
        ldsm_x4_t(bk, sK + swz(kv_row(16 + (lane & 7) + ((lane >> 3) & 1) * 8), 2 * dpi + (lane >> 4)));
        mma16816(acc[2 * dpi], a1, bk[0], bk[1]);
        mma16816(acc[2 * dpi + 1], a1, bk[2], bk[3]);
      }
    }
    stage_and_store(sQ, acc, p.Lq, p.dqkv + cur.h * HD, p.ld_dqkv, cur.base_row, rs, lane);
    if (u + 1 == u1 || nxt.bh != cur.bh) { flush_cls(cur); flush_dq(cur); }
    cur = nxt;
    __syncwarp();
  }
}

}  // namespace tattn

namespace tattn {
// CLS output row of a (clip, head) from its P partials (max, sum, o[64]): o = sum_j w_j o_j / sum_j w_j l_j
__global__ void __launch_bounds__(64)
time_cls_combine_kernel(const float* __restrict__ part, __nv_bfloat16* __restrict__ out, long long ld_out, float* __restrict__ lse,
                        int H, int P, long long clip_rows, float scale) {
  const int b = blockIdx.x / H, h = blockIdx.x - b * H, d = threadIdx.x;
  const float* pp = part + ((long long)b * H + h) * P * 66;
  float M = -INFINITY;
  for (int j = 0; j < P; ++j) M = fmaxf(M, pp[(long long)j * 66]);
  const float sl2 = scale * LOG2E;
  float L = 0.f, o = 0.f;
  for (int j = 0; j < P; ++j) {
    const float w = exp2f((pp[(long long)j * 66] - M) * sl2);
    L += w * pp[(long long)j * 66 + 1];
    o += w * pp[(long long)j * 66 + 2 + d];
  }
  const long long row = (long long)b * clip_rows;
  out[row * ld_out + h * HD + d] = __float2bfloat16_rn(o / L);
  if (d == 0) lse[row * H + h] = M * scale + logf(L);
}
// dqkv[cls row] = bf16([dq_cls | dk_cls | dv_cls]) from the fp32 accumulators
__global__ void time_cls_finalize_kernel(const float* __restrict__ dcls_kv, const float* __restrict__ dcls_q,
                                         __nv_bfloat16* __restrict__ dqkv, long long lddq, int H, int D, long long N) {
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int part = threadIdx.x >> 6, d = threadIdx.x & 63;
  const float v = part == 0 ? dcls_q[((long long)b * H + h) * HD + d] : dcls_kv[(((long long)b * H + h) * 2 + (part - 1)) * HD + d];
  dqkv[(long long)b * N * lddq + part * D + h * HD + d] = __float2bfloat16_rn(v);
}
}  // namespace tattn

// ---- launchers (called from lv_group_attn_fwd / lv_group_attn_bwd for mode 1 with <= 16 frames)
static int heads_per_cta(int H) {
  for (int hc = tattn::MAX_WARPS; hc >= 1; --hc)
    if (H % hc == 0) return hc;
  return 1;
}

static void fill(tattn::Params& p, int B, int H, int T, int n) {
  p.H = H; p.D = H * tattn::HD; p.Lq = T; p.n = n;
  p.hc = heads_per_cta(H); p.hchunks = H / p.hc;
  p.clip_rows = 1 + (long long)T * n;
  p.units = B * p.hchunks * n;
  p.scale = 0.125f;
}

static int time_fwd_launch(const void* qkv, long long ld_qkv, void* out, long long ld_out, float* lse, float* cls_part, int B, int H,
                           int T, int n, cudaStream_t st) {
  tattn::Params p{};
  fill(p, B, H, T, n);
  p.qkv = (const __nv_bfloat16*)qkv; p.ld_qkv = ld_qkv;
  p.out = (__nv_bfloat16*)out; p.ld_out = ld_out; p.lse = lse;
  p.cls_part = cls_part;
  const int smem = p.hc * tattn::FWD_WARP_BYTES;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tattn::time_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error((int)e, "time attention fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  int grid = sm_count();
  if (grid > p.units) grid = p.units;
  tattn::time_attn_fwd_kernel<<<grid, p.hc * 32, smem, st>>>(p);
  int rc = check_launch("lv_group_attn_fwd(time)");
  if (rc || !cls_part) return rc;
  tattn::time_cls_combine_kernel<<<B * H, 64, 0, st>>>(cls_part, (__nv_bfloat16*)out, ld_out, lse, H, n, p.clip_rows, p.scale);
  return check_launch("lv_time_attn_fwd_cls(combine)");
}

int time_attn_small_fwd(const void* qkv, long long ld_qkv, void* out, long long ld_out, float* lse, int B, int H, int T, int n,
                        cudaStream_t st) {
  return time_fwd_launch(qkv, ld_qkv, out, ld_out, lse, nullptr, B, H, T, n, st);
}

static int time_bwd_launch(const void* qkv, long long ld_qkv, const void* out, long long ld_out, const float* lse, const void* dout,
                           long long ld_dout, void* dqkv, long long ld_dqkv, float* dcls_kv, float* dcls_q, int B, int H, int T, int n,
                           cudaStream_t st);

int time_attn_small_bwd(const void* qkv, long long ld_qkv, const void* out, long long ld_out, const float* lse, const void* dout,
                        long long ld_dout, void* dqkv, long long ld_dqkv, float* dcls_kv, int B, int H, int T, int n,
                        cudaStream_t st) {
  return time_bwd_launch(qkv, ld_qkv, out, ld_out, lse, dout, ld_dout, dqkv, ld_dqkv, dcls_kv, nullptr, B, H, T, n, st);
}

static int time_bwd_launch(const void* qkv, long long ld_qkv, const void* out, long long ld_out, const float* lse, const void* dout,
                           long long ld_dout, void* dqkv, long long ld_dqkv, float* dcls_kv, float* dcls_q, int B, int H, int T, int n,
                           cudaStream_t st) {
  tattn::Params p{};
  fill(p, B, H, T, n);
  p.dcls_q = dcls_q;
  p.qkv = (const __nv_bfloat16*)qkv; p.ld_qkv = ld_qkv;
  p.out = (__nv_bfloat16*)out; p.ld_out = ld_out; p.lse = (float*)lse;
  p.dout = (const __nv_bfloat16*)dout; p.ld_dout = ld_dout;
  p.dqkv = (__nv_bfloat16*)dqkv; p.ld_dqkv = ld_dqkv; p.dcls_kv = dcls_kv;
  const int smem = p.hc * tattn::BWD_WARP_BYTES;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tattn::time_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error((int)e, "time attention bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  int grid = sm_count();
  if (grid > p.units) grid = p.units;
  tattn::time_attn_bwd_kernel<<<grid, p.hc * 32, smem, st>>>(p);
  int rc = check_launch("lv_group_attn_bwd(time)");
  if (rc || !dcls_q) return rc;
  tattn::time_cls_finalize_kernel<<<B * H, 192, 0, st>>>(dcls_kv, dcls_q, (__nv_bfloat16*)dqkv, ld_dqkv, H, p.D, p.clip_rows);
  return check_launch("lv_time_attn_bwd_cls(finalize)");
}

}  // namespace lv

// Time attention INCLUDING the CLS query row (pairs with lv_space_attn_*_tc_cls; <= 16 frames): fwd cls_part = fp32 scratch
// [B*H*n*66]; bwd dcls_kv fp32 [B][H][2][64] and dcls_q fp32 [B][H][64] zeroed by the caller.  Every row of out / lse / dqkv is
// written.  Replace lv_group_attn_fwd(mode 1) + lv_cls_attn_fwd and lv_group_attn_bwd(mode 1) + lv_cls_attn_bwd + lv_cls_kv_finalize.
extern "C" int lv_time_attn_fwd_cls(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* lse, float* cls_part, int B,
                                    int H, int T, int n, void* stream) {
  LV_REQUIRE(qkv && out && lse && cls_part && B > 0 && H > 0 && T > 0 && T <= 16 && n > 0, "lv_time_attn_fwd_cls: bad arguments (T <= 16)");
  LV_REQUIRE(ld_qkv % 8 == 0 && ld_out % 8 == 0, "lv_time_attn_fwd_cls: leading dimensions must be multiples of 8");
  return lv::time_fwd_launch(qkv, ld_qkv, out, ld_out, lse, cls_part, B, H, T, n, (cudaStream_t)stream);
}
extern "C" int lv_time_attn_bwd_cls(const void* qkv, int64_t ld_qkv, const void* out, int64_t ld_out, const float* lse,
                                    const void* dout, int64_t ld_dout, void* dqkv, int64_t ld_dqkv, float* dcls_kv, float* dcls_q,
                                    int B, int H, int T, int n, void* stream) {
  LV_REQUIRE(qkv && out && lse && dout && dqkv && dcls_kv && dcls_q && B > 0 && H > 0 && T > 0 && T <= 16 && n > 0,
             "lv_time_attn_bwd_cls: bad arguments (T <= 16)");
  LV_REQUIRE(ld_qkv % 8 == 0 && ld_out % 8 == 0 && ld_dout % 8 == 0 && ld_dqkv % 8 == 0, "lv_time_attn_bwd_cls: leading dimensions must be multiples of 8");
  return lv::time_bwd_launch(qkv, ld_qkv, out, ld_out, lse, dout, ld_dout, dqkv, ld_dqkv, dcls_kv, dcls_q, B, H, T, n, (cudaStream_t)stream);
}

namespace lv {

}  // namespace lv
