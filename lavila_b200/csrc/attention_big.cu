// Divided space-time attention for LARGE groups (more than 208 keys per group): the TSF-L/14 geometries of
// lavila/models/models.py:438-481 (336 px -> n = 576 patches per frame) and :1021-1070 (224 px -> n = 256), where one
// group's Q/K/V/P no longer fit in one CTA's shared memory (attention.cu) or one TMEM allocation (attention_tc.cu).
// Same contract as lv_group_attn_fwd / lv_group_attn_bwd (modes 0 and 1): qkv is addressed in place, a group is Lq token
// rows with a row stride plus the clip's CLS key/value as one extra key (timesformer.py:121-133), the softmax is fp32.
//
// Key-tiled ("flash") schedule, bf16 mma.sync.m16n8k16 with fp32 accumulation, head_dim 64:
//   forward : CTA = 64 queries of one group; K/V walk in 64-key blocks through a double-buffered cp.async ring; online
//             softmax; writes O and lse = max*scale + ln(sum).
//   backward: two kernels, no atomics on token rows and no workspace:
//     dKV   : CTA = 64 keys; Q/dO/O walk in 64-query blocks (double buffered); recomputes P^T = exp(S^T*scale - lse) and
//             dP^T = V dO^T, delta = rowsum(dO o O) from the staged tiles; dV += P^T dO, dK += dS^T Q in registers;
//             the CLS key's gradient goes to the fp32 accumulator dcls_kv with atomics (one row per group).
//     dQ    : CTA = 64 queries; K/V walk in 64-key blocks; dQ += dS K.
// The fragment code is the one validated in attention.cu (group kernels) and flash_attn.cu.
#include "../../include/lavila_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace lv {
namespace battn {

constexpr int HD = 64, BQ = 64, BK = 64, ROW_BYTES = 128, WARPS = 4, THREADS = WARPS * 32;
constexpr int TILE = 64 * ROW_BYTES;   // one 64-row x 64-dim bf16 tile
constexpr float LOG2E = 1.4426950408889634f;

struct Params {
  const __nv_bfloat16* qkv; long long ld_qkv;
  __nv_bfloat16* out; long long ld_out;         // fwd: written; bwd: the forward output (read)
  float* lse;                                   // [rows, H]
  const __nv_bfloat16* dout; long long ld_dout;
  __nv_bfloat16* dqkv; long long ld_dqkv;
  float* dcls_kv;                               // [B, H, 2, 64] fp32
  int H, D, Lq, has_cls;
  long long row_stride, clip_rows, inner_stride;
  int inner, first;
  int blocks;                                   // CTAs per group (query blocks or key blocks)
  float scale;
};

struct Group {
  int b, h;
  long long base_row, cls_row;
};

// CTA index -> (block inside the group, head, inner group, clip); block fastest, then head: CTAs running at the same
// time share a group's K/V (L2) and neighbouring heads of the same token rows (contiguous DRAM).
__device__ __forceinline__ int decode(const Params& p, long long cta, Group& g) {
  const int blk = (int)(cta % p.blocks);
  long long t = cta / p.blocks;
  g.h = (int)(t % p.H);
  t /= p.H;
  const int in = (int)(t % p.inner);
  g.b = (int)(t / p.inner);
  g.cls_row = (long long)g.b * p.clip_rows;
  g.base_row = g.cls_row + p.first + (long long)in * p.inner_stride;
  return blk;
}

__device__ __forceinline__ uint32_t swz(int row, int chunk) { return row * ROW_BYTES + ((chunk ^ (row & 7)) << 4); }
__device__ __forceinline__ void cp_async16(uint32_t s, const void* g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void zero16(uint32_t a) { asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0u) : "memory"); }
__device__ __forceinline__ void st_shared_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
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

// 64 rows of one group into a swizzled tile.  Tile row r holds group row j = j0 + r:
//   j < Lq            -> src[(base_row + j*row_stride)*ld ...]   (src already points at the head's 64 columns)
//   j == Lq and extra -> the CLS row (keys/values only)
//   otherwise         -> zeros
__device__ __forceinline__ void load_rows(uint32_t tile, const __nv_bfloat16* src, long long ld, long long base_row,
                                          long long row_stride, int j0, int Lq, const __nv_bfloat16* extra, int tid) {
  for (int idx = tid; idx < 64 * 8; idx += THREADS) {
    const int r = idx >> 3, c = idx & 7, j = j0 + r;
    const uint32_t dst = tile + swz(r, c);
    if (j < Lq) cp_async16(dst, src + (base_row + (long long)j * row_stride) * ld + c * 8);
    else if (j == Lq && extra) cp_async16(dst, extra + c * 8);
    else zero16(dst);
  }
}

// ================================================================================================ forward
__global__ void __launch_bounds__(THREADS)
big_attn_fwd_kernel(const Params p) {
  __shared__ __align__(128) uint8_t smem[5 * TILE];   // Q | K0 K1 | V0 V1
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  Group gc;
  const int qb = decode(p, blockIdx.x, gc);
  const uint32_t sQ = smem_u32(smem), sK = sQ + TILE, sV = sK + 2 * TILE;
  const int q0 = qb * BQ;
  const int Lk = p.Lq + p.has_cls;
  const int nblocks = (Lk + BK - 1) / BK;
  const __nv_bfloat16* qsrc = p.qkv + gc.h * HD;
  const __nv_bfloat16* ksrc = qsrc + p.D;
  const __nv_bfloat16* vsrc = qsrc + 2 * p.D;
  const __nv_bfloat16* cls_k = p.has_cls ? ksrc + gc.cls_row * p.ld_qkv : nullptr;
  const __nv_bfloat16* cls_v = p.has_cls ? vsrc + gc.cls_row * p.ld_qkv : nullptr;

  load_rows(sQ, qsrc, p.ld_qkv, gc.base_row, p.row_stride, q0, p.Lq, nullptr, tid);
  load_rows(sK, ksrc, p.ld_qkv, gc.base_row, p.row_stride, 0, p.Lq, cls_k, tid);
  load_rows(sV, vsrc, p.ld_qkv, gc.base_row, p.row_stride, 0, p.Lq, cls_v, tid);
  cp_commit();

  const int g = lane >> 2, t = lane & 3;
  const int r0 = warp * 16 + g, r1 = r0 + 8;   // rows inside the query block
  const float sl2 = p.scale * LOG2E;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float o[8][4];
#pragma unroll
  for (int d = 0; d < 8; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
  uint32_t qf[4][4];

  for (int kb = 0; kb < nblocks; ++kb) {
    const int stage = kb & 1;
    if (kb + 1 < nblocks) {
      load_rows(sK + (stage ^ 1) * TILE, ksrc, p.ld_qkv, gc.base_row, p.row_stride, (kb + 1) * BK, p.Lq, cls_k, tid);
      load_rows(sV + (stage ^ 1) * TILE, vsrc, p.ld_qkv, gc.base_row, p.row_stride, (kb + 1) * BK, p.Lq, cls_v, tid);
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
    const uint32_t tK = sK + stage * TILE, tV = sV + stage * TILE;
    const int k0 = kb * BK;
    const int kleft = Lk - k0;   // real keys in this block (the last block of a 257- or 577-key group holds ONE)
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      if (nt * 8 < kleft) {
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
          uint32_t kf[4];
          ldsm_x4(kf, tK + swz(nt * 8 + (lane & 7), 4 * kp + (lane >> 3)));
          mma16816(s[nt], qf[2 * kp], kf[0], kf[1]);
          mma16816(s[nt], qf[2 * kp + 1], kf[2], kf[3]);
        }
      }
    }
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool in = k0 + nt * 8 + 2 * t + e < Lk;
        s[nt][e] = in ? s[nt][e] : -INFINITY;
        s[nt][2 + e] = in ? s[nt][2 + e] : -INFINITY;
        bm0 = fmaxf(bm0, s[nt][e]);
        bm1 = fmaxf(bm1, s[nt][2 + e]);
      }
    bm0 = qmax(bm0);
    bm1 = qmax(bm1);
    // every key block holds at least one real key, so the running maxima are finite from the first block on
    const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);
    const float c0 = exp2f((m0 - mn0) * sl2), c1 = exp2f((m1 - mn1) * sl2);
    const float b0 = mn0 * sl2, b1 = mn1 * sl2;
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
      if (kk * 16 >= kleft) continue;
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
    __syncthreads();   // everyone is done with this stage before it is refilled
  }
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  // stage O in the warp's own rows of the Q tile (its fragments are in registers), then 128-byte row stores
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    st_shared_u32(sQ + swz(r0, d) + 4 * t, pack_bf16x2(o[d][0] * i0, o[d][1] * i0));
    st_shared_u32(sQ + swz(r1, d) + 4 * t, pack_bf16x2(o[d][2] * i1, o[d][3] * i1));
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = it * 32 + lane, r = warp * 16 + (idx >> 3), c = idx & 7;
    if (q0 + r < p.Lq) {
      const uint4 v = ld_shared_v4(sQ + swz(r, c));
      *reinterpret_cast<uint4*>(p.out + (gc.base_row + (long long)(q0 + r) * p.row_stride) * p.ld_out + gc.h * HD + c * 8) = v;
    }
  }
  if (t == 0) {
    if (q0 + r0 < p.Lq) p.lse[(gc.base_row + (long long)(q0 + r0) * p.row_stride) * p.H + gc.h] = m0 * p.scale + logf(l0);
    if (q0 + r1 < p.Lq) p.lse[(gc.base_row + (long long)(q0 + r1) * p.row_stride) * p.H + gc.h] = m1 * p.scale + logf(l1);
  }
}

// delta[r] = sum_d dO[r,d] * O[r,d] for the 64 rows of two staged tiles (8 lanes share a row); lse2[r] = lse * log2(e).
// Rows beyond the group get delta = 0 (their tiles are zero-filled) and lse2 = 0 (their P is masked by the caller).
__device__ __forceinline__ void delta_lse_rows(const Params& p, const Group& gc, uint32_t sdO, uint32_t sO, int q0, float* delta_s,
                                               float* lse_s, int row_begin, int row_count, int tid_local, int nthr) {
  for (int idx = tid_local; idx < row_count * 8; idx += nthr) {
    const int r = row_begin + (idx >> 3), c = idx & 7;
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
    if (c == 0) {
      delta_s[r] = part;
      lse_s[r] = (q0 + r < p.Lq) ? p.lse[(gc.base_row + (long long)(q0 + r) * p.row_stride) * p.H + gc.h] * LOG2E : 0.f;
    }
  }
}

// One warp's 16 x 64 dK or dV tile (C-fragment layout) -> dqkv; `part` 1 = k third, 2 = v third.
__device__ __forceinline__ void write_kv(const Params& p, const Group& gc, float (&acc)[8][4], int part, uint32_t sV, int warp,
                                         int lane, int k0) {
  const int g = lane >> 2, t = lane & 3;
  const int kl0 = warp * 16 + g, kl1 = kl0 + 8;
  __syncwarp();
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    st_shared_u32(sV + swz(kl0, dt) + 4 * t, pack_bf16x2(acc[dt][0], acc[dt][1]));
    st_shared_u32(sV + swz(kl1, dt) + 4 * t, pack_bf16x2(acc[dt][2], acc[dt][3]));
  }
  if (p.has_cls && p.dcls_kv) {
    float* base = p.dcls_kv + (((long long)gc.b * p.H + gc.h) * 2 + (part - 1)) * HD;
    if (k0 + kl0 == p.Lq) {
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) { atomicAdd(base + dt * 8 + 2 * t, acc[dt][0]); atomicAdd(base + dt * 8 + 2 * t + 1, acc[dt][1]); }
    }
    if (k0 + kl1 == p.Lq) {
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) { atomicAdd(base + dt * 8 + 2 * t, acc[dt][2]); atomicAdd(base + dt * 8 + 2 * t + 1, acc[dt][3]); }
    }
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = it * 32 + lane, r = warp * 16 + (idx >> 3), c = idx & 7;
    const int key = k0 + r;
    if (key < p.Lq) {
      const uint4 v = ld_shared_v4(sV + swz(r, c));
      *reinterpret_cast<uint4*>(p.dqkv + (gc.base_row + (long long)key * p.row_stride) * p.ld_dqkv + part * p.D + gc.h * HD + c * 8) = v;
    }
  }
}

// ================================================================================================ backward: dK, dV
// Shared memory: K | V | 2 stages of (Q | dO | O) | lse2[64] | delta[64]
constexpr int DKV_SMEM = 2 * TILE + 6 * TILE + 2 * 64 * 4;

__global__ void __launch_bounds__(THREADS)
big_attn_bwd_dkv_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  Group gc;
  const int kb = decode(p, blockIdx.x, gc);
  const uint32_t sK = smem_u32(smem), sV = sK + TILE, sStage = sV + TILE;
  float* lse_s = reinterpret_cast<float*>(smem + 8 * TILE);
  float* delta_s = lse_s + 64;
  const int Lk = p.Lq + p.has_cls;
  const int k0 = kb * BK;
  const int nq_blocks = (p.Lq + BQ - 1) / BQ;
  const __nv_bfloat16* qsrc = p.qkv + gc.h * HD;
  const __nv_bfloat16* cls_k = p.has_cls ? qsrc + p.D + gc.cls_row * p.ld_qkv : nullptr;
  const __nv_bfloat16* cls_v = p.has_cls ? qsrc + 2 * p.D + gc.cls_row * p.ld_qkv : nullptr;
  const __nv_bfloat16* dosrc = p.dout + gc.h * HD;
  const __nv_bfloat16* osrc = p.out + gc.h * HD;

  load_rows(sK, qsrc + p.D, p.ld_qkv, gc.base_row, p.row_stride, k0, p.Lq, cls_k, tid);
  load_rows(sV, qsrc + 2 * p.D, p.ld_qkv, gc.base_row, p.row_stride, k0, p.Lq, cls_v, tid);
  load_rows(sStage, qsrc, p.ld_qkv, gc.base_row, p.row_stride, 0, p.Lq, nullptr, tid);
  load_rows(sStage + TILE, dosrc, p.ld_dout, gc.base_row, p.row_stride, 0, p.Lq, nullptr, tid);
  load_rows(sStage + 2 * TILE, osrc, p.ld_out, gc.base_row, p.row_stride, 0, p.Lq, nullptr, tid);
  cp_commit();

  const int g = lane >> 2, t = lane & 3;
  const float sl2 = p.scale * LOG2E;
  const int kl0 = warp * 16 + g, kl1 = kl0 + 8;     // key rows inside the block
  const int key0 = k0 + kl0, key1 = k0 + kl1;       // key indices inside the group
  uint32_t kf[4][4], vf[4][4];
  float dv[8][4], dk[8][4];
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    dv[dt][0] = dv[dt][1] = dv[dt][2] = dv[dt][3] = 0.f;
    dk[dt][0] = dk[dt][1] = dk[dt][2] = dk[dt][3] = 0.f;
  }

  for (int qb = 0; qb < nq_blocks; ++qb) {
    const int stage = qb & 1;
    if (qb + 1 < nq_blocks) {
      const uint32_t nx = sStage + (stage ^ 1) * 3 * TILE;
      load_rows(nx, qsrc, p.ld_qkv, gc.base_row, p.row_stride, (qb + 1) * BQ, p.Lq, nullptr, tid);
      load_rows(nx + TILE, dosrc, p.ld_dout, gc.base_row, p.row_stride, (qb + 1) * BQ, p.Lq, nullptr, tid);
      load_rows(nx + 2 * TILE, osrc, p.ld_out, gc.base_row, p.row_stride, (qb + 1) * BQ, p.Lq, nullptr, tid);
      cp_commit();
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
    const uint32_t sQ = sStage + stage * 3 * TILE, sdO = sQ + TILE, sO = sdO + TILE;
    if (qb == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t off = swz(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4));
        ldsm_x4(kf[ks], sK + off);
        ldsm_x4(vf[ks], sV + off);
      }
    }
    delta_lse_rows(p, gc, sdO, sO, qb * BQ, delta_s, lse_s, 0, 64, tid, THREADS);
    __syncthreads();
#pragma unroll 1
    for (int sub = 0; sub < 4; ++sub) {
      const int ql0 = sub * 16;               // query rows inside the block
      if (qb * BQ + ql0 >= p.Lq || k0 + warp * 16 >= Lk) break;   // no queries left / this warp's 16 keys are all padding
      float st[2][4], dp[2][4];
#pragma unroll
      for (int nq = 0; nq < 2; ++nq) {
        st[nq][0] = st[nq][1] = st[nq][2] = st[nq][3] = 0.f;
        dp[nq][0] = dp[nq][1] = dp[nq][2] = dp[nq][3] = 0.f;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
          uint32_t bq[4], bo[4];
          const uint32_t off = swz(ql0 + nq * 8 + (lane & 7), 4 * kp + (lane >> 3));
          ldsm_x4(bq, sQ + off);
          ldsm_x4(bo, sdO + off);
          mma16816(st[nq], kf[2 * kp], bq[0], bq[1]);
          mma16816(st[nq], kf[2 * kp + 1], bq[2], bq[3]);
          mma16816(dp[nq], vf[2 * kp], bo[0], bo[1]);
          mma16816(dp[nq], vf[2 * kp + 1], bo[2], bo[3]);
        }
      }
      // P^T and dS^T for keys (key0, key1) x queries (ql0 + nq*8 + 2t + e)
      uint32_t pa[4], da[4];
#pragma unroll
      for (int nq = 0; nq < 2; ++nq) {
        float pv[4], dsv[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ql = ql0 + nq * 8 + 2 * t + e;
          const float l2 = lse_s[ql], dl = delta_s[ql];
          const bool qok = qb * BQ + ql < p.Lq;
          const float p0 = (qok && key0 < Lk) ? exp2f(st[nq][e] * sl2 - l2) : 0.f;
          const float p1 = (qok && key1 < Lk) ? exp2f(st[nq][2 + e] * sl2 - l2) : 0.f;
          pv[e] = p0;
          pv[2 + e] = p1;
          dsv[e] = p0 * (dp[nq][e] - dl) * p.scale;
          dsv[2 + e] = p1 * (dp[nq][2 + e] - dl) * p.scale;
        }
        pa[2 * nq] = pack_bf16x2(pv[0], pv[1]);
        pa[2 * nq + 1] = pack_bf16x2(pv[2], pv[3]);
        da[2 * nq] = pack_bf16x2(dsv[0], dsv[1]);
        da[2 * nq + 1] = pack_bf16x2(dsv[2], dsv[3]);
      }
      // dV += P^T dO ; dK += dS^T Q   (A from registers: the MMA k index is the query)
#pragma unroll
      for (int dpi = 0; dpi < 4; ++dpi) {
        uint32_t bo[4], bq[4];
        const uint32_t off = swz(ql0 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * dpi + (lane >> 4));
        ldsm_x4_t(bo, sdO + off);
        ldsm_x4_t(bq, sQ + off);
        mma16816(dv[2 * dpi], pa, bo[0], bo[1]);
        mma16816(dv[2 * dpi + 1], pa, bo[2], bo[3]);
        mma16816(dk[2 * dpi], da, bq[0], bq[1]);
        mma16816(dk[2 * dpi + 1], da, bq[2], bq[3]);
      }
    }
    __syncthreads();   // the stage (and lse/delta) may be overwritten from here on
  }

  // ---- write dV then dK: bf16 staging in this warp's own 16 rows of the V tile, then 128-byte row stores; the CLS key
  //      row (key == Lq) is accumulated at full precision into dcls_kv ([b][h][k|v][64]).
  write_kv(p, gc, dv, 2, sV, warp, lane, k0);
  write_kv(p, gc, dk, 1, sV, warp, lane, k0);
}

// ================================================================================================ backward: dQ
// Shared memory: Q | dO | O | K0 K1 | V0 V1 | lse2[64] | delta[64]
constexpr int DQ_SMEM = 7 * TILE + 2 * 64 * 4;

__global__ void __launch_bounds__(THREADS)
big_attn_bwd_dq_kernel(const Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  Group gc;
  const int qb = decode(p, blockIdx.x, gc);
  const uint32_t sQ = smem_u32(smem), sdO = sQ + TILE, sO = sdO + TILE, sK = sO + TILE, sV = sK + 2 * TILE;
  float* lse_s = reinterpret_cast<float*>(smem + 7 * TILE);
  float* delta_s = lse_s + 64;
  const int q0 = qb * BQ;
  const int Lk = p.Lq + p.has_cls;
  const int nblocks = (Lk + BK - 1) / BK;
  const __nv_bfloat16* qsrc = p.qkv + gc.h * HD;
  const __nv_bfloat16* ksrc = qsrc + p.D;
  const __nv_bfloat16* vsrc = qsrc + 2 * p.D;
  const __nv_bfloat16* cls_k = p.has_cls ? ksrc + gc.cls_row * p.ld_qkv : nullptr;
  const __nv_bfloat16* cls_v = p.has_cls ? vsrc + gc.cls_row * p.ld_qkv : nullptr;

  load_rows(sQ, qsrc, p.ld_qkv, gc.base_row, p.row_stride, q0, p.Lq, nullptr, tid);
  load_rows(sdO, p.dout + gc.h * HD, p.ld_dout, gc.base_row, p.row_stride, q0, p.Lq, nullptr, tid);
  load_rows(sO, p.out + gc.h * HD, p.ld_out, gc.base_row, p.row_stride, q0, p.Lq, nullptr, tid);
  load_rows(sK, ksrc, p.ld_qkv, gc.base_row, p.row_stride, 0, p.Lq, cls_k, tid);
  load_rows(sV, vsrc, p.ld_qkv, gc.base_row, p.row_stride, 0, p.Lq, cls_v, tid);
  cp_commit();

  const int g = lane >> 2, t = lane & 3;
  const int r0 = warp * 16 + g, r1 = r0 + 8;
  const float sl2 = p.scale * LOG2E;
  const bool ok0 = q0 + r0 < p.Lq, ok1 = q0 + r1 < p.Lq;
  float l2_0 = 0.f, l2_1 = 0.f, dl0 = 0.f, dl1 = 0.f;
  float dq[8][4];
#pragma unroll
  for (int d = 0; d < 8; ++d) dq[d][0] = dq[d][1] = dq[d][2] = dq[d][3] = 0.f;
  uint32_t qf[4][4], dof[4][4];

  for (int kb = 0; kb < nblocks; ++kb) {
    const int stage = kb & 1;
    if (kb + 1 < nblocks) {
      load_rows(sK + (stage ^ 1) * TILE, ksrc, p.ld_qkv, gc.base_row, p.row_stride, (kb + 1) * BK, p.Lq, cls_k, tid);
      load_rows(sV + (stage ^ 1) * TILE, vsrc, p.ld_qkv, gc.base_row, p.row_stride, (kb + 1) * BK, p.Lq, cls_v, tid);
      cp_commit();
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
    if (kb == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t off = swz(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4));
        ldsm_x4(qf[ks], sQ + off);
        ldsm_x4(dof[ks], sdO + off);
      }
      // each warp prepares lse/delta of its own 16 rows
      delta_lse_rows(p, gc, sdO, sO, q0, delta_s, lse_s, warp * 16, 16, lane, 32);
      __syncwarp();
      l2_0 = lse_s[r0]; l2_1 = lse_s[r1];
      dl0 = delta_s[r0]; dl1 = delta_s[r1];
    }
    const uint32_t tK = sK + stage * TILE, tV = sV + stage * TILE;
    const int k0 = kb * BK;
    const int kleft = Lk - k0;
    float s[8][4], dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
      if (nt * 8 >= kleft) continue;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t kf[4], vf[4];
        const uint32_t off = swz(nt * 8 + (lane & 7), 4 * kp + (lane >> 3));
        ldsm_x4(kf, tK + off);
        ldsm_x4(vf, tV + off);
        mma16816(s[nt], qf[2 * kp], kf[0], kf[1]);
        mma16816(s[nt], qf[2 * kp + 1], kf[2], kf[3]);
        mma16816(dp[nt], dof[2 * kp], vf[0], vf[1]);
        mma16816(dp[nt], dof[2 * kp + 1], vf[2], vf[3]);
      }
    }
    // dS = P o (dP - delta) * scale, packed as the A operand of dQ += dS K
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool in = k0 + nt * 8 + 2 * t + e < Lk;
        const float p0 = (in && ok0) ? exp2f(fmaf(s[nt][e], sl2, -l2_0)) : 0.f;
        const float p1 = (in && ok1) ? exp2f(fmaf(s[nt][2 + e], sl2, -l2_1)) : 0.f;
        s[nt][e] = p0 * (dp[nt][e] - dl0) * p.scale;
        s[nt][2 + e] = p1 * (dp[nt][2 + e] - dl1) * p.scale;
      }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (kk * 16 >= kleft) continue;
      uint32_t a[4] = {pack_bf16x2(s[2 * kk][0], s[2 * kk][1]), pack_bf16x2(s[2 * kk][2], s[2 * kk][3]),
                       pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
      for (int dpi = 0; dpi < 4; ++dpi) {
        uint32_t kf[4];
        ldsm_x4_t(kf, tK + swz(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, 2 * dpi + (lane >> 4)));
        mma16816(dq[2 * dpi], a, kf[0], kf[1]);
        mma16816(dq[2 * dpi + 1], a, kf[2], kf[3]);
      }
    }
    __syncthreads();
  }
  // stage dQ in the warp's own rows of the Q tile, then 128-byte row stores into the q third of dqkv
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    st_shared_u32(sQ + swz(r0, d) + 4 * t, pack_bf16x2(dq[d][0], dq[d][1]));
    st_shared_u32(sQ + swz(r1, d) + 4 * t, pack_bf16x2(dq[d][2], dq[d][3]));
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = it * 32 + lane, r = warp * 16 + (idx >> 3), c = idx & 7;
    if (q0 + r < p.Lq) {
      const uint4 v = ld_shared_v4(sQ + swz(r, c));
      *reinterpret_cast<uint4*>(p.dqkv + (gc.base_row + (long long)(q0 + r) * p.row_stride) * p.ld_dqkv + gc.h * HD + c * 8) = v;
    }
  }
}

}  // namespace battn

namespace {

int fill(battn::Params& p, int mode, int B, int H, int T, int n) {
  p.H = H;
  p.D = H * battn::HD;
  p.has_cls = 1;
  p.first = 1;
  p.clip_rows = 1 + (long long)T * n;
  if (mode == 0) { p.Lq = n; p.row_stride = 1; p.inner = T; p.inner_stride = n; }
  else if (mode == 1) { p.Lq = T; p.row_stride = n; p.inner = n; p.inner_stride = 1; }
  else return set_error(-1, "tiled attention: mode %d not supported (space = 0, time = 1)", mode);
  p.scale = 1.0f / 8.0f;   // head_dim ** -0.5 (timesformer.py:94)
  return 0;
}

}  // namespace

int big_group_attn_fwd(const void* qkv, long long ld_qkv, void* out, long long ld_out, float* lse, int mode, int B, int H, int T,
                       int n, cudaStream_t st) {
  if (mode == 0 && flash_tc_enabled() && (ld_qkv % 8) == 0 && (ld_out % 8) == 0 && ((uintptr_t)qkv & 15) == 0) {
    // space attention of TimeSformer-L/14 (257 / 577 keys per group): tcgen05 key-tiled kernel, the CLS key / value row of the
    // clip appended to the group's last key tile
    const int D = H * 64;
    const long long N = 1 + (long long)T * n;
    const __nv_bfloat16* base = (const __nv_bfloat16*)qkv;
    return flash_attn_fwd_tc(base, ld_qkv, N, D, base + D, base + 2 * D, ld_qkv, N, D, 64, out, ld_out, lse, B, T, H, n, n, 1, n, 1, n,
                             1, 0, 0.125f, st);
  }
  battn::Params p{};
  int rc = fill(p, mode, B, H, T, n);
  if (rc) return rc;
  p.qkv = (const __nv_bfloat16*)qkv; p.ld_qkv = ld_qkv;
  p.out = (__nv_bfloat16*)out; p.ld_out = ld_out;
  p.lse = lse;
  p.blocks = (p.Lq + battn::BQ - 1) / battn::BQ;
  const long long grid = (long long)B * p.inner * H * p.blocks;
  LV_REQUIRE(grid > 0 && grid < (1ll << 31), "tiled attention fwd: grid of %lld CTAs out of range", grid);
  battn::big_attn_fwd_kernel<<<(unsigned)grid, battn::THREADS, 0, st>>>(p);
  return check_launch("lv_group_attn_fwd(tiled)");
}

int big_group_attn_bwd(const void* qkv, long long ld_qkv, const void* out, long long ld_out, const float* lse, const void* dout,
                       long long ld_dout, void* dqkv, long long ld_dqkv, float* dcls_kv, int mode, int B, int H, int T, int n,
                       cudaStream_t st) {
  battn::Params p{};
  int rc = fill(p, mode, B, H, T, n);
  if (rc) return rc;
  p.qkv = (const __nv_bfloat16*)qkv; p.ld_qkv = ld_qkv;
  p.out = (__nv_bfloat16*)out; p.ld_out = ld_out;
  p.lse = (float*)lse;
  p.dout = (const __nv_bfloat16*)dout; p.ld_dout = ld_dout;
  p.dqkv = (__nv_bfloat16*)dqkv; p.ld_dqkv = ld_dqkv;
  p.dcls_kv = dcls_kv;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(battn::big_attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, battn::DKV_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(battn::big_attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, battn::DQ_SMEM);
    if (e != cudaSuccess) return set_error((int)e, "tiled attention bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int Lk = p.Lq + p.has_cls;
  p.blocks = (Lk + battn::BK - 1) / battn::BK;
  long long grid = (long long)B * p.inner * H * p.blocks;
  LV_REQUIRE(grid > 0 && grid < (1ll << 31), "tiled attention bwd: grid of %lld CTAs out of range", grid);
  battn::big_attn_bwd_dkv_kernel<<<(unsigned)grid, battn::THREADS, battn::DKV_SMEM, st>>>(p);
  rc = check_launch("lv_group_attn_bwd(tiled dKV)");
  if (rc) return rc;
  p.blocks = (p.Lq + battn::BQ - 1) / battn::BQ;
  grid = (long long)B * p.inner * H * p.blocks;
  battn::big_attn_bwd_dq_kernel<<<(unsigned)grid, battn::THREADS, battn::DQ_SMEM, st>>>(p);
  return check_launch("lv_group_attn_bwd(tiled dQ)");
}

}  // namespace lv
