// Space attention (one group = (clip, head, frame): n patch queries, n + CLS keys) on tcgen05 tensor cores.
//
//   forward :  S = Q K^T  (2 x [128 x 208] fp32 in TMEM)  ->  fp32 softmax by 8 warps (thread = row)
//              ->  P (bf16) staged in shared memory in the canonical K-major 128B-swizzled layout
//              ->  O = P V  ([128 x 64] fp32 in TMEM, V consumed MN-major straight from its TMA tile).
// Persistent CTAs (one per SM) walk the groups; Q, K, V tiles arrive by TMA directly from the packed projection
// output qkv[rows, 3D] (2D box = 64 columns x n rows); the CLS key/value row is appended by 16 threads.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one lane), warps 2..9 = softmax/epilogue
// (warps 2-5 own query tile 0, warps 6-9 own query tile 1; a warp may only touch TMEM lanes 32*(warp%4)..+31).
//
// Replaces the space half of VarAttention.forward (lavila/models/timesformer.py:121-134, attn() :35-39).
// Numerics: bf16 operands, fp32 accumulation and softmax -- same contract as the mma.sync kernels in attention.cu,
// which remain in use for the HBM-bound time attention (17 keys) and the small text tower.
#include <mutex>
#include <type_traits>

#include "../../include/lavila_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace lv {
namespace attn_tc {

constexpr int HD = 64;
constexpr int QROWS = 256;            // two M=128 query tiles (rows >= n are zero)
constexpr int KROWS = 208;            // UMMA N for S: keys padded to a multiple of 16
constexpr int SQ_BYTES = QROWS * 128;
constexpr int SK_BYTES = KROWS * 128;
constexpr int SP_BYTES = 4 * 128 * 128;  // P tile: 4 atoms of (128 rows x 64 keys)
constexpr int NTHREADS = 320;
constexpr int TM_S = 0;  // TMEM columns: S0 [0,208), S1 [208,416); O_t overwrites the first 64 columns of S_t
constexpr int TMEM_COLS = 512;
constexpr float LOG2E = 1.4426950408889634f;

struct Params {
  const __nv_bfloat16* qkv;
  long long ld_qkv;
  __nv_bfloat16* out;
  long long ld_out;
  float* lse;
  // CLS-query fusion (optional): the clip's CLS query rides as row n of the Q tile; its attention over THIS frame's keys
  // (the CLS key itself counted in frame 0 only) leaves the kernel as a partial (max, sum, unnormalised output[64]) per
  // (clip, head, frame) in cls_part [B*H*T][66]; cls_combine_kernel merges the T partials into the CLS output row / lse.
  float* cls_part;
  int H, D, n, T;
  long long clip_rows;
  long long num_groups;
  float scale;
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

struct Coord {
  int b, h, f;
  long long base_row, cls_row;
};
template <class P>
__device__ __forceinline__ Coord decode(const P& p, long long g) {
  Coord c;   // frame-fastest (head-fastest order was measured: no gain for the TMA gathers, slower read-modify-write)
  c.f = (int)(g % p.T);
  const long long bh = g / p.T;
  c.h = (int)(bh % p.H);
  c.b = (int)(bh / p.H);
  c.cls_row = (long long)c.b * p.clip_rows;
  c.base_row = c.cls_row + 1 + (long long)c.f * p.n;
  return c;
}

__global__ void __launch_bounds__(NTHREADS, 1)
space_attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + SQ_BYTES;
  uint8_t* sV = sK + SK_BYTES;
  uint8_t* sP = sV + SK_BYTES;  // two tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * SP_BYTES);
  uint64_t* bar_load = bars;       // Q + K landed
  uint64_t* bar_s = bars + 1;
  uint64_t* bar_p = bars + 2;      // [2]
  uint64_t* bar_o = bars + 4;      // [2]
  uint64_t* bar_ofree = bars + 6;  // [2]
  uint64_t* bar_loadv = bars + 8;  // V landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Lq = p.n, Lk = p.n + 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_qkv);
    mbar_init(bar_load, 1);
    mbar_init(bar_loadv, 1);
    mbar_init(bar_s, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&bar_p[t], 4);
      mbar_init(&bar_o[t], 1);
      mbar_init(&bar_ofree[t], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  // zero the padding rows once: they are never written again (TMA boxes cover rows [0, n) only)
  for (int idx = threadIdx.x; idx < (QROWS - Lq) * 8; idx += NTHREADS)
    st_shared_v4(smem_u32(sQ) + Lq * 128 + idx * 16, 0, 0, 0, 0);
  for (int idx = threadIdx.x; idx < (KROWS - Lq) * 8; idx += NTHREADS) {
    st_shared_v4(smem_u32(sK) + Lq * 128 + idx * 16, 0, 0, 0, 0);
    st_shared_v4(smem_u32(sV) + Lq * 128 + idx * 16, 0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    // Q and K of group i+1 are requested as soon as the S MMAs of group i have retired (they overlap the softmax),
    // V of group i+1 once the last P.V MMA of group i has read V.
    int it = 0;
    for (long long g = blockIdx.x; g < p.num_groups; g += gridDim.x, ++it) {
      const Coord c = decode(p, g);
      if (it > 0) {
        if (lane == 0) mbar_wait(bar_s, (it - 1) & 1);
        __syncwarp();
      }
      if (lane < 8) {  // CLS key row -> row Lq of the K tile (generic proxy, swizzled by hand)
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.qkv + c.cls_row * p.ld_qkv + p.D + c.h * HD + lane * 8));
        st_shared_v4(smem_u32(sK) + Lq * 128 + ((lane ^ (Lq & 7)) << 4), v.x, v.y, v.z, v.w);
      } else if (lane < 16 && p.cls_part) {  // CLS query row -> row Lq of the Q tile
        const int ch = lane - 8;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.qkv + c.cls_row * p.ld_qkv + c.h * HD + ch * 8));
        st_shared_v4(smem_u32(sQ) + Lq * 128 + ((ch ^ (Lq & 7)) << 4), v.x, v.y, v.z, v.w);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_expect_tx(bar_load, 2 * Lq * 128);
        tma_load_2d(sQ, &tm_qkv, bar_load, c.h * HD, (int)c.base_row);
        tma_load_2d(sK, &tm_qkv, bar_load, p.D + c.h * HD, (int)c.base_row);
      }
      if (it > 0) {
        if (lane == 0) mbar_wait(&bar_o[1], (it - 1) & 1);
        __syncwarp();
      }
      if (lane < 8) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.qkv + c.cls_row * p.ld_qkv + 2 * p.D + c.h * HD + lane * 8));
        st_shared_v4(smem_u32(sV) + Lq * 128 + ((lane ^ (Lq & 7)) << 4), v.x, v.y, v.z, v.w);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_expect_tx(bar_loadv, Lq * 128);
        tma_load_2d(sV, &tm_qkv, bar_loadv, 2 * p.D + c.h * HD, (int)c.base_row);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, KROWS, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, HD, 0, 1);
      const uint32_t q_base = smem_u32(sQ), k_base = smem_u32(sK), v_base = smem_u32(sV), p_base = smem_u32(sP);
      int it = 0;
      for (long long g = blockIdx.x; g < p.num_groups; g += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        mbar_wait(bar_load, ph);
        if (it > 0) {  // O_t lives in the first 64 columns of S_t: both must have been drained by the epilogue warps
          mbar_wait(&bar_ofree[0], (it - 1) & 1);
          mbar_wait(&bar_ofree[1], (it - 1) & 1);
        }
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc_mma_bf16(tmem_base + TM_S + t * KROWS, make_smem_desc_sw128(q_base + t * 16384 + ks * 32, 16, 1024),
                        make_smem_desc_sw128(k_base + ks * 32, 16, 1024), idesc_s, ks > 0);
        tc_commit(bar_s);
        mbar_wait(bar_loadv, ph);
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&bar_p[t], ph);   // P_t staged; the softmax warps have finished reading S_t
          tc_fence_after();
#pragma unroll
          for (int s = 0; s < KROWS / 16; ++s)
            tc_mma_bf16(tmem_base + TM_S + t * KROWS, make_smem_desc_sw128(p_base + t * SP_BYTES + (s >> 2) * 16384 + (s & 3) * 32, 16, 1024),
                        make_smem_desc_sw128(v_base + s * 2048, 8192, 1024), idesc_o, s > 0);
          tc_commit(&bar_o[t]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue warps
    const int e = warp - 2, t = e >> 2, q = warp & 3;
    const int row = q * 32 + lane;          // row inside the 128-row tile == TMEM lane
    const int qrow = t * 128 + row;
    const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
    const uint32_t p_tile = smem_u32(sP) + t * SP_BYTES + row * 128;
    const float sl2 = p.scale * LOG2E;
    int it = 0;
    for (long long g = blockIdx.x; g < p.num_groups; g += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const Coord c = decode(p, g);
      mbar_wait(bar_s, ph);
      tc_fence_after();
      const uint32_t ts = trow + TM_S + t * KROWS;
      // ---- pass 1: row maximum
      float m = -INFINITY;
#pragma unroll 1
      for (int cc = 0; cc < 6; ++cc) {
        uint32_t r[32];
        tmem_ld_32x32(ts + cc * 32, r);
        tmem_ld_wait();
        if ((cc + 1) * 32 <= Lk) {   // interior chunk: no key masking
#pragma unroll
          for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(r[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (cc * 32 + j < Lk) m = fmaxf(m, __uint_as_float(r[j]));
        }
      }
      // the CLS query row (qrow == Lq, fused mode) sees the CLS key only in frame 0: the T partials must count it once
      const int row_Lk = (qrow == Lq && c.f != 0) ? Lk - 1 : Lk;
      {
        uint32_t r[16];
        tmem_ld_32x16(ts + 192, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (192 + j < row_Lk) m = fmaxf(m, __uint_as_float(r[j]));
      }
      // ---- pass 2: p = exp2((s - m) * scale * log2e), row sum, bf16 P into the swizzled K-major tile
      float sum = 0.f;
      const float mb = m * sl2;
#pragma unroll 1
      for (int cc = 0; cc < 6; ++cc) {
        uint32_t r[32];
        tmem_ld_32x32(ts + cc * 32, r);
        tmem_ld_wait();
        float pv[32];
        if ((cc + 1) * 32 <= Lk) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            pv[j] = ex2_fast(fmaf(__uint_as_float(r[j]), sl2, -mb));
            sum += pv[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            pv[j] = (cc * 32 + j < Lk) ? ex2_fast(fmaf(__uint_as_float(r[j]), sl2, -mb)) : 0.f;
            sum += pv[j];
          }
        }
        const uint32_t atom = p_tile + (cc >> 1) * 16384;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int chunk = (cc & 1) * 4 + jj;
          st_shared_v4(atom + ((chunk ^ (row & 7)) << 4), pack_bf16x2(pv[jj * 8], pv[jj * 8 + 1]),
                       pack_bf16x2(pv[jj * 8 + 2], pv[jj * 8 + 3]), pack_bf16x2(pv[jj * 8 + 4], pv[jj * 8 + 5]),
                       pack_bf16x2(pv[jj * 8 + 6], pv[jj * 8 + 7]));
        }
      }
      {
        uint32_t r[16];
        tmem_ld_32x16(ts + 192, r);
        tmem_ld_wait();
        float pv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          pv[j] = (192 + j < row_Lk) ? ex2_fast(fmaf(__uint_as_float(r[j]), sl2, -mb)) : 0.f;
          sum += pv[j];
        }
        const uint32_t atom = p_tile + 3 * 16384;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
          st_shared_v4(atom + ((jj ^ (row & 7)) << 4), pack_bf16x2(pv[jj * 8], pv[jj * 8 + 1]),
                       pack_bf16x2(pv[jj * 8 + 2], pv[jj * 8 + 3]), pack_bf16x2(pv[jj * 8 + 4], pv[jj * 8 + 5]),
                       pack_bf16x2(pv[jj * 8 + 6], pv[jj * 8 + 7]));
      }
      fence_proxy_async_smem();   // P must be visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_p[t]);
      // ---- epilogue: O / sum -> bf16 -> global
      mbar_wait(&bar_o[t], ph);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(trow + TM_S + t * KROWS, o0);        // O_t aliases the first 64 columns of S_t
      tmem_ld_32x32(trow + TM_S + t * KROWS + 32, o1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_ofree[t]);
      if (qrow < Lq) {
        const float inv = 1.f / sum;
        const long long grow = c.base_row + qrow;
        __nv_bfloat16* dst = p.out + grow * p.ld_out + c.h * HD;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          *reinterpret_cast<uint4*>(dst + jj * 8) =
              make_uint4(pack_bf16x2(__uint_as_float(o0[jj * 8]) * inv, __uint_as_float(o0[jj * 8 + 1]) * inv),
                         pack_bf16x2(__uint_as_float(o0[jj * 8 + 2]) * inv, __uint_as_float(o0[jj * 8 + 3]) * inv),
                         pack_bf16x2(__uint_as_float(o0[jj * 8 + 4]) * inv, __uint_as_float(o0[jj * 8 + 5]) * inv),
                         pack_bf16x2(__uint_as_float(o0[jj * 8 + 6]) * inv, __uint_as_float(o0[jj * 8 + 7]) * inv));
          *reinterpret_cast<uint4*>(dst + 32 + jj * 8) =
              make_uint4(pack_bf16x2(__uint_as_float(o1[jj * 8]) * inv, __uint_as_float(o1[jj * 8 + 1]) * inv),
                         pack_bf16x2(__uint_as_float(o1[jj * 8 + 2]) * inv, __uint_as_float(o1[jj * 8 + 3]) * inv),
                         pack_bf16x2(__uint_as_float(o1[jj * 8 + 4]) * inv, __uint_as_float(o1[jj * 8 + 5]) * inv),
                         pack_bf16x2(__uint_as_float(o1[jj * 8 + 6]) * inv, __uint_as_float(o1[jj * 8 + 7]) * inv));
        }
        p.lse[grow * p.H + c.h] = m * p.scale + logf(sum);
      } else if (qrow == Lq && p.cls_part) {
        float* dst = p.cls_part + (((long long)c.b * p.H + c.h) * p.T + c.f) * 66;
        dst[0] = m;
        dst[1] = sum;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          dst[2 + j] = __uint_as_float(o0[j]);
          dst[34 + j] = __uint_as_float(o1[j]);
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// CLS output row of a clip / head from the T per-frame partials: o = sum_f w_f o_f / sum_f w_f l_f, w_f = 2^((m_f - M) s log2e)
__global__ void __launch_bounds__(64)
cls_combine_kernel(const float* __restrict__ part, __nv_bfloat16* __restrict__ out, long long ld_out, float* __restrict__ lse,
                   int H, int T, long long clip_rows, float scale) {
  const int b = blockIdx.x / H, h = blockIdx.x - b * H, d = threadIdx.x;
  const float* pp = part + ((long long)b * H + h) * T * 66;
  float M = -INFINITY;
  for (int f = 0; f < T; ++f) M = fmaxf(M, pp[f * 66]);
  const float sl2 = scale * LOG2E;
  float L = 0.f, o = 0.f;
  for (int f = 0; f < T; ++f) {
    const float w = ex2_fast((pp[f * 66] - M) * sl2);
    L += w * pp[f * 66 + 1];
    o += w * pp[f * 66 + 2 + d];
  }
  const long long row = (long long)b * clip_rows;
  out[row * ld_out + h * HD + d] = __float2bfloat16_rn(o / L);
  if (d == 0) lse[row * H + h] = M * scale + logf(L);
}

constexpr int FWD_SMEM = 1024 + SQ_BYTES + 2 * SK_BYTES + 2 * SP_BYTES + 128;
static_assert(FWD_SMEM <= 227 * 1024, "space attention forward: shared memory budget");


// ================================================================================================ backward
// Per group, with key tiles kt (128 keys = TMEM lanes) and query tiles qt (N = 128 or 80 columns):
//   (A)  S^T  = K_kt Q_qt^T ,  dP^T = V_kt dO_qt^T                        [128 x Nq] fp32 in TMEM
//   (E)  P^T  = exp2(S^T*scale*log2e - lse_q) ,  dS^T = P^T (dP^T - delta_q) * scale      (thread = key row)
//        written as bf16 to shared memory, K-major swizzled rows [key][q]
//   (B)  dV_kt += P^T dO_qt ,  dK_kt += dS^T Q_qt        (A operand K-major from the staged tiles)
//        dQ_qt += dS K_kt                                (A operand = the same dS^T tile read MN-major)
// dV/dK are complete after the two query tiles of a key tile, dQ after both key tiles.
// TMEM columns: S^T 128 | dP^T 128 | dV 64 | dK 64 | dQ0 64 | dQ1 64 = 512.
__device__ long long* g_dbg = nullptr;   // optional cycle-stamp buffer (lv_debug_set_buffer), CTA 0 only
#define LV_STAMP(slot) do { if (g_dbg && blockIdx.x == 0 && it < 4) g_dbg[it * 64 + (slot)] = clock64(); } while (0)

constexpr int B_ROWS = 208;                   // operand tile rows (n + 1 <= 208).  M = 128 operands of the second key tile
                                              // read 48 rows past the tile into the next buffer: finite bf16, masked keys
constexpr int B_TILE_BYTES = B_ROWS * 128;    // 26 KB, a multiple of the 1024-byte swizzle atom
constexpr int B_NBUF = 6;                     // Q, K, V, dO of the current group + K, V of the next one
constexpr int B_STAGE_BYTES = 2 * 128 * 128;  // P^T / dS^T: 128 key rows x 128 query columns (2 atoms of 64)
constexpr int B_VEC = 256;                    // lse / delta entries per stage
constexpr int TB_ST = 0, TB_DPT = 128, TB_DV = 256, TB_DK = 320, TB_DQ = 384;
constexpr int BWD_SMEM = 1024 + B_NBUF * B_TILE_BYTES + 2 * B_STAGE_BYTES + 2 * 2 * B_VEC * 4 + 128;
static_assert(B_TILE_BYTES % 1024 == 0, "tiles must keep the 1024-byte swizzle alignment");
static_assert(BWD_SMEM <= 227 * 1024, "space attention backward: shared memory budget");

struct BwdParams {
  const __nv_bfloat16* qkv;
  long long ld_qkv;
  const __nv_bfloat16* out;
  long long ld_out;
  const __nv_bfloat16* dout;
  long long ld_dout;
  const float* lse;
  __nv_bfloat16* dqkv;
  long long ld_dqkv;
  float* dcls_kv;
  // CLS-query fusion (optional): the clip's CLS query / its output gradient ride as row n of the Q / dO tiles (lse and delta of
  // that row come from the clip's CLS row), so dV / dK of every key receive the CLS query's contribution in the same MMAs; the
  // per-frame partial dQ_cls is accumulated into dcls_q fp32 [B][H][64] (zeroed by the caller).  Replaces lv_cls_attn_bwd.
  float* dcls_q;
  int accumulate_kv;
  int H, D, n, T;
  long long clip_rows;
  long long num_groups;
  float scale;
};

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void bar_sync_epi() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ float dot8(const uint4& a, const uint4& b) {
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 x = unpack_bf16x2(aw[j]), y = unpack_bf16x2(bw[j]);
    acc += x.x * y.x + x.y * y.y;
  }
  return acc;
}

// Write a 64-wide fp32 accumulator row (two x32 TMEM loads) as bf16 to global, optionally adding what is there.
__device__ __forceinline__ void store_row64(__nv_bfloat16* dst, const uint32_t (&a)[32], const uint32_t (&b)[32], bool accumulate) {
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const uint32_t* src = jj < 4 ? &a[jj * 8] : &b[(jj - 4) * 8];
    float v[8];
#pragma unroll
    for (int e2 = 0; e2 < 8; ++e2) v[e2] = __uint_as_float(src[e2]);
    if (accumulate) {
      const uint4 old = *reinterpret_cast<const uint4*>(dst + jj * 8);
      const float2 o0 = unpack_bf16x2(old.x), o1 = unpack_bf16x2(old.y), o2 = unpack_bf16x2(old.z), o3 = unpack_bf16x2(old.w);
      v[0] += o0.x; v[1] += o0.y; v[2] += o1.x; v[3] += o1.y; v[4] += o2.x; v[5] += o2.y; v[6] += o3.x; v[7] += o3.y;
    }
    *reinterpret_cast<uint4*>(dst + jj * 8) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

// Same row as bf16 into a 128B-swizzled [rows x 64] staging tile that a TMA store then writes out.
__device__ __forceinline__ void stage_row64(uint32_t tile, int row, const uint32_t (&a)[32], const uint32_t (&b)[32]) {
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const uint32_t* src = jj < 4 ? &a[jj * 8] : &b[(jj - 4) * 8];
    st_shared_v4(tile + row * 128 + ((jj ^ (row & 7)) << 4),
                 pack_bf16x2(__uint_as_float(src[0]), __uint_as_float(src[1])), pack_bf16x2(__uint_as_float(src[2]), __uint_as_float(src[3])),
                 pack_bf16x2(__uint_as_float(src[4]), __uint_as_float(src[5])), pack_bf16x2(__uint_as_float(src[6]), __uint_as_float(src[7])));
  }
}

// Pipeline across groups (one CTA walks groups it = 0, 1, ...):
//   * operand tiles live in 6 rotating buffers, role r of group `it` in buffer (4*it + r) % 6 with r = 0:K 1:V 2:Q 3:dO.
//     K and V of group it+1 therefore land in the two buffers group `it` does not use and are requested while it
//     computes; Q and dO reuse the K/V buffers of group it-1 and are requested when its last MMA has retired.
//   * lse (log2 units) and delta = rowsum(dO * O) of group it+1 are produced by the otherwise idle producer warp
//     during group `it` into the other half of a two-stage buffer.
//   * gradients leave through the (then free) P^T / dS^T staging tiles and TMA stores, not per-thread row stores.
__global__ void __launch_bounds__(NTHREADS, 1)
space_attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                         const __grid_constant__ CUtensorMap tm_st128, const __grid_constant__ CUtensorMap tm_sttail,
                         const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* tiles = smem;
  uint8_t* sPt = tiles + B_NBUF * B_TILE_BYTES;
  uint8_t* sdSt = sPt + B_STAGE_BYTES;
  float* s_lse = reinterpret_cast<float*>(sdSt + B_STAGE_BYTES);   // [2][B_VEC]
  float* s_delta = s_lse + 2 * B_VEC;                               // [2][B_VEC]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_delta + 2 * B_VEC);
  uint64_t* bar_loadA = bars;        // [2]  K, V landed (group parity selects the barrier)
  uint64_t* bar_loadB = bars + 2;    //      Q, dO landed
  uint64_t* bar_prep = bars + 3;     // [2]  lse / delta stage written
  uint64_t* bar_sdp = bars + 5;
  uint64_t* bar_pds = bars + 6;
  uint64_t* bar_mma3 = bars + 7;
  uint64_t* bar_free = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Lq = p.n, Lk = p.n + 1;
  const bool tma_out = p.accumulate_kv == 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_qkv);
    prefetch_tmap(&tm_do);
    prefetch_tmap(&tm_st128);
    prefetch_tmap(&tm_sttail);
    mbar_init(&bar_loadA[0], 1);
    mbar_init(&bar_loadA[1], 1);
    mbar_init(bar_loadB, 1);
    mbar_init(&bar_prep[0], 1);
    mbar_init(&bar_prep[1], 1);
    mbar_init(bar_sdp, 1);
    mbar_init(bar_pds, 8);
    mbar_init(bar_mma3, 1);
    mbar_init(bar_free, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  // zero once: every tile buffer (padding rows are never written again; over-read rows must be finite), both
  // staging tiles and the lse / delta stages (entries >= n are read as padding columns)
  for (int idx = threadIdx.x; idx < (B_NBUF * B_TILE_BYTES + 2 * B_STAGE_BYTES + 2 * 2 * B_VEC * 4) / 16; idx += NTHREADS)
    st_shared_v4(smem_u32(tiles) + idx * 16, 0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto tile_of = [&](int it, int r) -> uint8_t* { return tiles + ((4 * (it % 3) + r) % B_NBUF) * B_TILE_BYTES; };

  if (warp == 0) {
    // ------------------------------------------------------------------ producer (+ lse / delta of the next group)
    auto issue_kv = [&](int it2, const Coord& c) {
      uint8_t* bk = tile_of(it2, 0);
      uint8_t* bv = tile_of(it2, 1);
      if (lane < 16) {   // the CLS key / value row closes the tile
        const int part = lane >> 3, ch = lane & 7;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.qkv + c.cls_row * p.ld_qkv + (1 + part) * p.D + c.h * HD + ch * 8));
        st_shared_v4(smem_u32(part ? bv : bk) + Lq * 128 + ((ch ^ (Lq & 7)) << 4), v.x, v.y, v.z, v.w);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        uint64_t* bar = &bar_loadA[it2 & 1];
        mbar_arrive_expect_tx(bar, 2 * Lq * 128);
        tma_load_2d(bk, &tm_qkv, bar, p.D + c.h * HD, (int)c.base_row);
        tma_load_2d(bv, &tm_qkv, bar, 2 * p.D + c.h * HD, (int)c.base_row);
      }
    };
    auto issue_qdo = [&](int it2, const Coord& c) {
      if (p.dcls_q) {    // fused CLS query: q_cls -> row Lq of the Q tile, dO_cls -> row Lq of the dO tile
        if (lane < 16) {
          const int part = lane >> 3, ch = lane & 7;
          const __nv_bfloat16* src = part ? p.dout + c.cls_row * p.ld_dout + c.h * HD + ch * 8
                                          : p.qkv + c.cls_row * p.ld_qkv + c.h * HD + ch * 8;
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
          st_shared_v4(smem_u32(tile_of(it2, part ? 3 : 2)) + Lq * 128 + ((ch ^ (Lq & 7)) << 4), v.x, v.y, v.z, v.w);
        }
        fence_proxy_async_smem();
        __syncwarp();
      }
      if (lane == 0) {
        mbar_arrive_expect_tx(bar_loadB, 2 * Lq * 128);
        tma_load_2d(tile_of(it2, 2), &tm_qkv, bar_loadB, c.h * HD, (int)c.base_row);
        tma_load_2d(tile_of(it2, 3), &tm_do, bar_loadB, c.h * HD, (int)c.base_row);
      }
    };
    auto prep = [&](int it2, const Coord& c) {
      float* ls = s_lse + (it2 & 1) * B_VEC;
      float* dl = s_delta + (it2 & 1) * B_VEC;
      const int nrows = Lq + (p.dcls_q ? 1 : 0);      // fused mode: entry Lq = the clip's CLS row
#pragma unroll 1
      for (int r = lane; r < nrows; r += 32) {
        const long long grow = r < Lq ? c.base_row + r : c.cls_row;
        const __nv_bfloat16* orow = p.out + grow * p.ld_out + c.h * HD;
        const __nv_bfloat16* drow = p.dout + grow * p.ld_dout + c.h * HD;
        uint4 o8[8], d8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o8[j] = __ldg(reinterpret_cast<const uint4*>(orow + j * 8));
          d8[j] = __ldg(reinterpret_cast<const uint4*>(drow + j * 8));
        }
        const float l2 = __ldg(p.lse + grow * p.H + c.h) * LOG2E;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += dot8(o8[j], d8[j]);
        ls[r] = l2;
        dl[r] = acc;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_prep[it2 & 1]);
    };
    long long g = blockIdx.x;
    if (g < p.num_groups) {
      const Coord c = decode(p, g);
      issue_kv(0, c);
      issue_qdo(0, c);
      prep(0, c);
    }
    int it = 0;
    for (; g < p.num_groups; g += gridDim.x, ++it) {
      const long long gn = g + gridDim.x;
      const bool has_next = gn < p.num_groups;
      Coord cn{};
      if (has_next) {
        cn = decode(p, gn);
        issue_kv(it + 1, cn);   // buffers of group it-1's Q / dO: free since bar_free(it-1)
        if (lane == 0) {        // Q / dO of the next group can only land once this group's MMAs are done: pull them into L2 now
          tma_prefetch_2d(&tm_qkv, cn.h * HD, (int)cn.base_row);
          tma_prefetch_2d(&tm_do, cn.h * HD, (int)cn.base_row);
        }
        prep(it + 1, cn);
      }
      if (lane == 0) mbar_wait(bar_free, it & 1);   // every MMA of group `it` has read its operands
      __syncwarp();
      if (has_next) issue_qdo(it + 1, cn);          // buffers of group it's K / V
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      // descriptors: constant high word per layout, low word = (address >> 4) | (LBO >> 4) << 16
      constexpr uint32_t HI = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO 1024 B, version 1, SWIZZLE_128B
      auto dk = [](uint32_t addr) -> uint64_t { return ((uint64_t)HI << 32) | (uint32_t)(((addr & 0x3FFFFu) >> 4) | (1u << 16)); };           // K-major
      auto dmn = [](uint32_t addr, uint32_t lbo) -> uint64_t { return ((uint64_t)HI << 32) | (uint32_t)(((addr & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16)); };  // MN-major
      const uint32_t pt_base = smem_u32(sPt), ds_base = smem_u32(sdSt);
      constexpr uint32_t idesc_a128 = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_a80 = make_idesc_bf16(128, 80, 0, 0);
      constexpr uint32_t idesc_kv = make_idesc_bf16(128, HD, 0, 1);   // A K-major (staged tile), B MN-major
      constexpr uint32_t idesc_dq = make_idesc_bf16(128, HD, 1, 1);   // A MN-major (dS^T read transposed), B MN-major
      int it = 0;
      for (long long g = blockIdx.x; g < p.num_groups; g += gridDim.x, ++it) {
        const uint32_t k_base = smem_u32(tile_of(it, 0)), v_base = smem_u32(tile_of(it, 1));
        const uint32_t q_base = smem_u32(tile_of(it, 2)), do_base = smem_u32(tile_of(it, 3));
        LV_STAMP(0);
        mbar_wait(&bar_loadA[it & 1], (it >> 1) & 1);
        mbar_wait(bar_loadB, it & 1);
        LV_STAMP(1);
        tc_fence_after();
        auto issue_a = [&](int step) {   // S^T = K_kt Q_qt^T ,  dP^T = V_kt dO_qt^T
          const int kt = step >> 1, qt = step & 1;
          const uint32_t idesc_a = qt ? idesc_a80 : idesc_a128;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc_mma_bf16(tmem_base + TB_ST, dk(k_base + kt * 16384 + ks * 32), dk(q_base + qt * 16384 + ks * 32), idesc_a, ks > 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc_mma_bf16(tmem_base + TB_DPT, dk(v_base + kt * 16384 + ks * 32), dk(do_base + qt * 16384 + ks * 32), idesc_a, ks > 0);
          tc_commit(bar_sdp);
        };
        issue_a(0);
#pragma unroll 1
        for (int step = 0; step < 4; ++step) {
          const int kt = step >> 1, qt = step & 1;
          LV_STAMP(2 + step * 4);
          // (B) needs the staged P^T / dS^T (and E has finished reading S^T / dP^T of this step)
          mbar_wait(bar_pds, step & 1);
          LV_STAMP(3 + step * 4);
          tc_fence_after();
          // The MMA-issuing thread is on the critical path (one thread, ~110 tcgen05.mma per group): trip counts are compile-time
          // (5 or 8 k-steps of 16) and every descriptor is the first one of its tile plus an immediate (the 14-bit address field
          // counts 16-byte units and never carries here), so an MMA costs an add and the issue itself.
          const uint64_t d_pt = dk(pt_base), d_ds = dk(ds_base);
          const uint64_t d_do = dmn(do_base + qt * 128 * 128, 8192), d_q = dmn(q_base + qt * 128 * 128, 8192);
          const uint64_t d_dst = dmn(ds_base, 16384), d_k = dmn(k_base + kt * 128 * 128, 8192);
          const uint32_t acc0 = qt > 0 ? 1u : 0u, accq = kt > 0 ? 1u : 0u;
          auto issue_kv = [&](auto nsteps) {
#pragma unroll
            for (int s2 = 0; s2 < decltype(nsteps)::value; ++s2) {
              const uint64_t a_off = (uint64_t)(((s2 >> 2) * 16384 + (s2 & 3) * 32) >> 4);
              const uint64_t b_off = (uint64_t)((s2 * 16 * 128) >> 4);
              tc_mma_bf16(tmem_base + TB_DV, d_pt + a_off, d_do + b_off, idesc_kv, s2 > 0 ? 1u : acc0);
              tc_mma_bf16(tmem_base + TB_DK, d_ds + a_off, d_q + b_off, idesc_kv, s2 > 0 ? 1u : acc0);
            }
          };
          auto issue_dq = [&](auto nsteps) {
#pragma unroll
            for (int s2 = 0; s2 < decltype(nsteps)::value; ++s2)
              tc_mma_bf16(tmem_base + TB_DQ + qt * 64, d_dst + (uint64_t)((s2 * 2048) >> 4), d_k + (uint64_t)((s2 * 16 * 128) >> 4),
                          idesc_dq, s2 > 0 ? 1u : accq);
          };
          if (qt) issue_kv(std::integral_constant<int, 5>{}); else issue_kv(std::integral_constant<int, 8>{});   // 80 / 128 query columns
          if (kt) issue_dq(std::integral_constant<int, 5>{}); else issue_dq(std::integral_constant<int, 8>{});   // keys 128..207 / 0..127
          tc_commit(bar_mma3);
          // the next step's S^T / dP^T MMAs queue right behind: the elementwise warps find them ready as soon as the
          // staged tiles of this step have been consumed
          if (step < 3) issue_a(step + 1);
          LV_STAMP(4 + step * 4);
        }
        tc_commit(bar_free);
      }
    }
  } else {
    // ------------------------------------------------------------------ elementwise + epilogue warps
    const int e = warp - 2, hf = e >> 2, q = warp & 3;
    const int row = q * 32 + lane;   // TMEM lane = key row inside the key tile (or query row for the dQ epilogue)
    const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
    const float sl2 = p.scale * LOG2E;
    const bool elected = warp == 2 && lane == 0;
    const uint32_t pt_tile = smem_u32(sPt), ds_tile = smem_u32(sdSt);
    int it = 0;
    for (long long g = blockIdx.x; g < p.num_groups; g += gridDim.x, ++it) {
      const Coord c = decode(p, g);
      const float* lse_s = s_lse + (it & 1) * B_VEC;
      const float* delta_s = s_delta + (it & 1) * B_VEC;
      mbar_wait(&bar_prep[it & 1], (it >> 1) & 1);
      if (elected) LV_STAMP(20);
#pragma unroll 1
      for (int step = 0; step < 4; ++step) {
        const int kt = step >> 1, qt = step & 1;
        mbar_wait(bar_sdp, step & 1);
        if (elected) LV_STAMP(21 + step * 4);
        if (step > 0) mbar_wait(bar_mma3, (step - 1) & 1);   // previous (B) MMAs have consumed the staged tiles
        if (elected) LV_STAMP(22 + step * 4);
        tc_fence_after();
        if (step == 0 && it > 0 && tma_out) {
          // the stores of the previous group's last gradients must have read the staging tiles before they are rewritten
          if (elected) tma_store_wait_read();
          bar_sync_epi();
        }
        if (step == 2) {
          // ---- epilogue of key tile 0: dV (hf 0) / dK (hf 1)
          uint32_t a[32], b[32];
          tmem_ld_32x32(trow + (hf ? TB_DK : TB_DV), a);
          tmem_ld_32x32(trow + (hf ? TB_DK : TB_DV) + 32, b);
          tmem_ld_wait();
          if (tma_out) {
            stage_row64(pt_tile + hf * 16384, row, a, b);
            fence_proxy_async_smem();
            bar_sync_epi();
            if (elected) {
              tma_store_2d(&tm_st128, sPt, 2 * p.D + c.h * HD, (int)c.base_row);
              tma_store_2d(&tm_st128, sPt + 16384, p.D + c.h * HD, (int)c.base_row);
              tma_store_commit();
              tma_store_wait_read();
            }
            bar_sync_epi();
          } else {
            store_row64(p.dqkv + (c.base_row + row) * p.ld_dqkv + (hf ? 1 : 2) * p.D + c.h * HD, a, b, true);
          }
        }
        const int key = kt * 128 + row;
        const bool key_ok = key < Lk;
        const int ncols = qt ? 40 : 64;         // this warp's share of the query columns
        const int col_base = hf * ncols;
        const uint32_t pt_row = pt_tile + row * 128, ds_row = ds_tile + row * 128;
        // Key tile 1 holds keys 128 .. n: warps whose 32 key rows are all past the CLS key (warp-uniform) only clear their
        // rows of the staged tiles -- no TMEM traffic, no exponentials.
        const bool warp_all_masked = kt * 128 + q * 32 >= Lk;
        if (warp_all_masked) {
          for (int c8 = col_base; c8 < col_base + ncols; c8 += 8) {
            const uint32_t off = (c8 >> 6) * 16384 + ((((c8 & 63) >> 3) ^ (row & 7)) << 4);
            st_shared_v4(pt_row + off, 0, 0, 0, 0);
            st_shared_v4(ds_row + off, 0, 0, 0, 0);
          }
        }
        // 32 columns per TMEM round trip (one wait per 32x2 values instead of per 8)
        for (int cb = 0; cb < (warp_all_masked ? 0 : ncols); cb += 32) {
          const int col0 = col_base + cb;       // query column inside the tile
          uint32_t s32[32], d32[32];
          const bool wide = cb + 32 <= ncols;   // qt = 1 ends with an 8-column remainder (40 = 32 + 8)
          if (wide) {
            tmem_ld_32x32(trow + TB_ST + col0, s32);
            tmem_ld_32x32(trow + TB_DPT + col0, d32);
          } else {
            uint32_t s8[8], d8[8];
            tmem_ld_32x8(trow + TB_ST + col0, s8);
            tmem_ld_32x8(trow + TB_DPT + col0, d8);
#pragma unroll
            for (int j = 0; j < 8; ++j) { s32[j] = s8[j]; d32[j] = d8[j]; }
          }
          tmem_ld_wait();
          const int ngroups = wide ? 4 : 1;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            if (g8 < ngroups) {
              const int c8 = col0 + g8 * 8;
              const int qg = qt * 128 + c8;
              const float4 la = *reinterpret_cast<const float4*>(lse_s + qg), lb = *reinterpret_cast<const float4*>(lse_s + qg + 4);
              const float4 da = *reinterpret_cast<const float4*>(delta_s + qg), db = *reinterpret_cast<const float4*>(delta_s + qg + 4);
              const float l8[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
              const float dl8[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
              float pv[8], dv[8];
              if (!key_ok) {                    // key row past the CLS key: zeros (rows of a partly valid warp)
#pragma unroll
                for (int j = 0; j < 8; ++j) { pv[j] = 0.f; dv[j] = 0.f; }
              } else if (qg + 8 <= Lq) {        // interior: no masking
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float pj = ex2_fast(fmaf(__uint_as_float(s32[g8 * 8 + j]), sl2, -l8[j]));
                  pv[j] = pj;
                  dv[j] = pj * (__uint_as_float(d32[g8 * 8 + j]) - dl8[j]) * p.scale;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  // fused CLS query (column Lq): every key of the frame, the CLS key itself only in frame 0
                  const bool ok = key_ok && ((qg + j < Lq) || (p.dcls_q && qg + j == Lq && (key < Lk - 1 || c.f == 0)));
                  const float pj = ok ? ex2_fast(fmaf(__uint_as_float(s32[g8 * 8 + j]), sl2, -l8[j])) : 0.f;
                  pv[j] = pj;
                  dv[j] = ok ? pj * (__uint_as_float(d32[g8 * 8 + j]) - dl8[j]) * p.scale : 0.f;
                }
              }
              const uint32_t off = (c8 >> 6) * 16384 + ((((c8 & 63) >> 3) ^ (row & 7)) << 4);
              st_shared_v4(pt_row + off, pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]), pack_bf16x2(pv[4], pv[5]), pack_bf16x2(pv[6], pv[7]));
              st_shared_v4(ds_row + off, pack_bf16x2(dv[0], dv[1]), pack_bf16x2(dv[2], dv[3]), pack_bf16x2(dv[4], dv[5]), pack_bf16x2(dv[6], dv[7]));
            }
          }
        }
        if (elected) LV_STAMP(23 + step * 4);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pds);
        if (elected) LV_STAMP(24 + step * 4);
      }
      // ---- final epilogues: key tile 1 (dV / dK) and dQ0 / dQ1
      mbar_wait(bar_mma3, 1);
      if (elected) LV_STAMP(40);
      tc_fence_after();
      {
        uint32_t a[32], b[32];
        tmem_ld_32x32(trow + (hf ? TB_DK : TB_DV), a);
        tmem_ld_32x32(trow + (hf ? TB_DK : TB_DV) + 32, b);
        tmem_ld_wait();
        const int key = 128 + row;
        if (tma_out) {
          stage_row64(pt_tile + hf * 16384, row, a, b);   // rows >= n - 128 are staged but lie outside the store box
        } else if (key < Lq) {
          store_row64(p.dqkv + (c.base_row + key) * p.ld_dqkv + (hf ? 1 : 2) * p.D + c.h * HD, a, b, true);
        }
        if (key == Lq && p.dcls_kv) {  // CLS key / value: fp32 atomics into the per-(clip, head) accumulator
          float* base = p.dcls_kv + (((long long)c.b * p.H + c.h) * 2 + (hf ? 0 : 1)) * HD;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            atomicAdd(base + j, __uint_as_float(a[j]));
            atomicAdd(base + 32 + j, __uint_as_float(b[j]));
          }
        }
      }
      {
        uint32_t a[32], b[32];
        tmem_ld_32x32(trow + TB_DQ + hf * 64, a);
        tmem_ld_32x32(trow + TB_DQ + hf * 64 + 32, b);
        tmem_ld_wait();
        const int qrow = hf * 128 + row;
        if (tma_out) stage_row64(ds_tile + hf * 16384, row, a, b);
        else if (qrow < Lq) store_row64(p.dqkv + (c.base_row + qrow) * p.ld_dqkv + c.h * HD, a, b, false);
        if (qrow == Lq && p.dcls_q) {   // this frame's share of dQ of the CLS query
          float* base = p.dcls_q + ((long long)c.b * p.H + c.h) * HD;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            atomicAdd(base + j, __uint_as_float(a[j]));
            atomicAdd(base + 32 + j, __uint_as_float(b[j]));
          }
        }
      }
      tc_fence_before();
      if (tma_out) {
        fence_proxy_async_smem();
        bar_sync_epi();
        if (elected) {
          tma_store_2d(&tm_sttail, sPt, 2 * p.D + c.h * HD, (int)c.base_row + 128);           // dV keys 128..n-1
          tma_store_2d(&tm_sttail, sPt + 16384, p.D + c.h * HD, (int)c.base_row + 128);       // dK
          tma_store_2d(&tm_st128, sdSt, c.h * HD, (int)c.base_row);                           // dQ rows 0..127
          tma_store_2d(&tm_sttail, sdSt + 16384, c.h * HD, (int)c.base_row + 128);            // dQ rows 128..n-1
          tma_store_commit();
        }
      }
      if (elected) LV_STAMP(41);
    }
    if (elected && tma_out) tma_store_wait_all();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace attn_tc
}  // namespace lv

using namespace lv;

// Space attention forward on tcgen05 (patch-token rows; the CLS row is produced by lv_cls_attn_fwd).
// Requirements: 129 <= n <= 207 (TSF-B: 196), head_dim 64.  Other shapes use lv_group_attn_fwd.
extern "C" int lv_space_attn_fwd_tc(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* lse, int B, int H,
                                    int T, int n, void* stream) {
  LV_REQUIRE(qkv && out && lse && B > 0 && H > 0 && T > 0, "lv_space_attn_fwd_tc: bad arguments");
  LV_REQUIRE(n > 128 && n + 1 <= attn_tc::KROWS, "lv_space_attn_fwd_tc: n=%d unsupported (129..207)", n);
  LV_REQUIRE(ld_qkv % 8 == 0 && ld_out % 8 == 0, "lv_space_attn_fwd_tc: leading dimensions must be multiples of 8");
  attn_tc::Params p{};
  p.qkv = (const __nv_bfloat16*)qkv; p.ld_qkv = ld_qkv;
  p.out = (__nv_bfloat16*)out; p.ld_out = ld_out;
  p.lse = lse;
  p.H = H; p.D = H * attn_tc::HD; p.n = n; p.T = T;
  p.clip_rows = 1 + (long long)T * n;
  p.num_groups = (long long)B * H * T;
  p.scale = 0.125f;
  const long long rows = (long long)B * p.clip_rows;
  CUtensorMap tm;
  int rc = make_tmap_2d_bf16(&tm, qkv, (uint64_t)(3 * p.D), (uint64_t)rows, (uint64_t)ld_qkv, 64, (uint32_t)n);
  if (rc) return rc;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, []() {
    attr_err = cudaFuncSetAttribute(attn_tc::space_attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_tc::FWD_SMEM);
  });
  if (attr_err != cudaSuccess) return set_error((int)attr_err, "lv_space_attn_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  const long long grid = p.num_groups < sm_count() ? p.num_groups : sm_count();
  attn_tc::space_attn_fwd_tc_kernel<<<(unsigned)grid, attn_tc::NTHREADS, attn_tc::FWD_SMEM, (cudaStream_t)stream>>>(tm, p);
  return check_launch("lv_space_attn_fwd_tc");
}

// Space attention forward INCLUDING the CLS query row (VarAttention.forward, lavila/models/timesformer.py:116-134, both halves):
// the tcgen05 group kernel carries the clip's CLS query as an extra row of every frame's Q tile and writes one partial per
// (clip, head, frame) into cls_part (fp32 [B*H*T*66], caller-owned scratch); cls_combine_kernel merges them.  Replaces the
// pair lv_space_attn_fwd_tc + lv_cls_attn_fwd (which streams K and V of all tokens a second time).
extern "C" int lv_space_attn_fwd_tc_cls(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* lse, float* cls_part,
                                        int B, int H, int T, int n, void* stream) {
  LV_REQUIRE(qkv && out && lse && cls_part && B > 0 && H > 0 && T > 0, "lv_space_attn_fwd_tc_cls: bad arguments");
  LV_REQUIRE(n > 128 && n + 1 <= attn_tc::KROWS, "lv_space_attn_fwd_tc_cls: n=%d unsupported (129..207)", n);
  LV_REQUIRE(ld_qkv % 8 == 0 && ld_out % 8 == 0, "lv_space_attn_fwd_tc_cls: leading dimensions must be multiples of 8");
  attn_tc::Params p{};
  p.qkv = (const __nv_bfloat16*)qkv; p.ld_qkv = ld_qkv;
  p.out = (__nv_bfloat16*)out; p.ld_out = ld_out;
  p.lse = lse;
  p.cls_part = cls_part;
  p.H = H; p.D = H * attn_tc::HD; p.n = n; p.T = T;
  p.clip_rows = 1 + (long long)T * n;
  p.num_groups = (long long)B * H * T;
  p.scale = 0.125f;
  const long long rows = (long long)B * p.clip_rows;
  CUtensorMap tm;
  int rc = make_tmap_2d_bf16(&tm, qkv, (uint64_t)(3 * p.D), (uint64_t)rows, (uint64_t)ld_qkv, 64, (uint32_t)n);
  if (rc) return rc;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, []() {
    attr_err = cudaFuncSetAttribute(attn_tc::space_attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_tc::FWD_SMEM);
  });
  if (attr_err != cudaSuccess) return set_error((int)attr_err, "lv_space_attn_fwd_tc_cls: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  const long long grid = p.num_groups < sm_count() ? p.num_groups : sm_count();
  attn_tc::space_attn_fwd_tc_kernel<<<(unsigned)grid, attn_tc::NTHREADS, attn_tc::FWD_SMEM, (cudaStream_t)stream>>>(tm, p);
  rc = check_launch("lv_space_attn_fwd_tc_cls");
  if (rc) return rc;
  attn_tc::cls_combine_kernel<<<B * H, 64, 0, (cudaStream_t)stream>>>(cls_part, (__nv_bfloat16*)out, ld_out, lse, H, T, p.clip_rows, p.scale);
  return check_launch("lv_space_attn_fwd_tc_cls(combine)");
}

// Space attention backward on tcgen05.  Same contract / call order as lv_group_attn_bwd(mode 0):
// lv_cls_attn_bwd -> lv_space_attn_bwd_tc(accumulate_kv = 1) -> lv_cls_kv_finalize.
extern "C" int lv_space_attn_bwd_tc(const void* qkv, int64_t ld_qkv, const void* out, int64_t ld_out, const float* lse,
                                    const void* dout, int64_t ld_dout, void* dqkv, int64_t ld_dqkv, float* dcls_kv,
                                    int accumulate_kv, int B, int H, int T, int n, void* stream) {
  LV_REQUIRE(qkv && out && lse && dout && dqkv && dcls_kv && B > 0 && H > 0 && T > 0, "lv_space_attn_bwd_tc: bad arguments");
  LV_REQUIRE(n > 128 && n + 1 <= attn_tc::KROWS, "lv_space_attn_bwd_tc: n=%d unsupported (129..207)", n);
  LV_REQUIRE(ld_qkv % 8 == 0 && ld_out % 8 == 0 && ld_dout % 8 == 0 && ld_dqkv % 8 == 0, "lv_space_attn_bwd_tc: leading dimensions must be multiples of 8");
  attn_tc::BwdParams p{};
  p.qkv = (const __nv_bfloat16*)qkv; p.ld_qkv = ld_qkv;
  p.out = (const __nv_bfloat16*)out; p.ld_out = ld_out;
  p.dout = (const __nv_bfloat16*)dout; p.ld_dout = ld_dout;
  p.lse = lse;
  p.dqkv = (__nv_bfloat16*)dqkv; p.ld_dqkv = ld_dqkv;
  p.dcls_kv = dcls_kv; p.accumulate_kv = accumulate_kv;
  p.H = H; p.D = H * attn_tc::HD; p.n = n; p.T = T;
  p.clip_rows = 1 + (long long)T * n;
  p.num_groups = (long long)B * H * T;
  p.scale = 0.125f;
  const long long rows = (long long)B * p.clip_rows;
  CUtensorMap tm_qkv, tm_do;
  int rc = make_tmap_2d_bf16(&tm_qkv, qkv, (uint64_t)(3 * p.D), (uint64_t)rows, (uint64_t)ld_qkv, 64, (uint32_t)n);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tm_do, dout, (uint64_t)p.D, (uint64_t)rows, (uint64_t)ld_dout, 64, (uint32_t)n);
  if (rc) return rc;
  CUtensorMap tm_st128, tm_sttail;   // gradient stores: 128-row boxes and the (n - 128)-row tail
  rc = make_tmap_2d_bf16(&tm_st128, dqkv, (uint64_t)(3 * p.D), (uint64_t)rows, (uint64_t)ld_dqkv, 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tm_sttail, dqkv, (uint64_t)(3 * p.D), (uint64_t)rows, (uint64_t)ld_dqkv, 64, (uint32_t)(n - 128));
  if (rc) return rc;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, []() {
    attr_err = cudaFuncSetAttribute(attn_tc::space_attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_tc::BWD_SMEM);
  });
  if (attr_err != cudaSuccess) return set_error((int)attr_err, "lv_space_attn_bwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  const long long grid = p.num_groups < sm_count() ? p.num_groups : sm_count();
  attn_tc::space_attn_bwd_tc_kernel<<<(unsigned)grid, attn_tc::NTHREADS, attn_tc::BWD_SMEM, (cudaStream_t)stream>>>(tm_qkv, tm_do, tm_st128, tm_sttail, p);
  return check_launch("lv_space_attn_bwd_tc");
}

namespace lv {
namespace attn_tc {
// dqkv[cls row] = bf16([dq_cls | dk_cls | dv_cls]) from the fp32 accumulators
__global__ void cls_qkv_finalize_kernel(const float* __restrict__ dcls_kv, const float* __restrict__ dcls_q,
                                        __nv_bfloat16* __restrict__ dqkv, long long lddq, int H, int D, long long N) {
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int part = threadIdx.x >> 6, d = threadIdx.x & 63;       // 192 threads: [q 64 | k 64 | v 64]
  const float v = part == 0 ? dcls_q[((long long)b * H + h) * HD + d] : dcls_kv[(((long long)b * H + h) * 2 + (part - 1)) * HD + d];
  dqkv[(long long)b * N * lddq + part * D + h * HD + d] = __float2bfloat16_rn(v);
}
}  // namespace attn_tc
}  // namespace lv

// Space attention backward INCLUDING the CLS query (pairs with lv_space_attn_fwd_tc_cls): one tcgen05 kernel + a 3 x 64-value
// finalize per (clip, head).  dcls_kv fp32 [B][H][2][64] and dcls_q fp32 [B][H][64] are zeroed scratch of the caller; every row
// of dqkv (CLS rows included) is overwritten.  Replaces lv_space_attn_bwd_tc + lv_cls_attn_bwd + lv_cls_kv_finalize.
extern "C" int lv_space_attn_bwd_tc_cls(const void* qkv, int64_t ld_qkv, const void* out, int64_t ld_out, const float* lse,
                                        const void* dout, int64_t ld_dout, void* dqkv, int64_t ld_dqkv, float* dcls_kv,
                                        float* dcls_q, int B, int H, int T, int n, void* stream) {
  LV_REQUIRE(qkv && out && lse && dout && dqkv && dcls_kv && dcls_q && B > 0 && H > 0 && T > 0, "lv_space_attn_bwd_tc_cls: bad arguments");
  LV_REQUIRE(n > 128 && n + 1 <= attn_tc::KROWS, "lv_space_attn_bwd_tc_cls: n=%d unsupported (129..207)", n);
  LV_REQUIRE(ld_qkv % 8 == 0 && ld_out % 8 == 0 && ld_dout % 8 == 0 && ld_dqkv % 8 == 0, "lv_space_attn_bwd_tc_cls: leading dimensions must be multiples of 8");
  attn_tc::BwdParams p{};
  p.qkv = (const __nv_bfloat16*)qkv; p.ld_qkv = ld_qkv;
  p.out = (const __nv_bfloat16*)out; p.ld_out = ld_out;
  p.dout = (const __nv_bfloat16*)dout; p.ld_dout = ld_dout;
  p.lse = lse;
  p.dqkv = (__nv_bfloat16*)dqkv; p.ld_dqkv = ld_dqkv;
  p.dcls_kv = dcls_kv; p.dcls_q = dcls_q; p.accumulate_kv = 0;
  p.H = H; p.D = H * attn_tc::HD; p.n = n; p.T = T;
  p.clip_rows = 1 + (long long)T * n;
  p.num_groups = (long long)B * H * T;
  p.scale = 0.125f;
  const long long rows = (long long)B * p.clip_rows;
  CUtensorMap tm_qkv, tm_do, tm_st128, tm_sttail;
  int rc = make_tmap_2d_bf16(&tm_qkv, qkv, (uint64_t)(3 * p.D), (uint64_t)rows, (uint64_t)ld_qkv, 64, (uint32_t)n);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tm_do, dout, (uint64_t)p.D, (uint64_t)rows, (uint64_t)ld_dout, 64, (uint32_t)n);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tm_st128, dqkv, (uint64_t)(3 * p.D), (uint64_t)rows, (uint64_t)ld_dqkv, 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tm_sttail, dqkv, (uint64_t)(3 * p.D), (uint64_t)rows, (uint64_t)ld_dqkv, 64, (uint32_t)(n - 128));
  if (rc) return rc;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, []() {
    attr_err = cudaFuncSetAttribute(attn_tc::space_attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_tc::BWD_SMEM);
  });
  if (attr_err != cudaSuccess) return set_error((int)attr_err, "lv_space_attn_bwd_tc_cls: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  const long long grid = p.num_groups < sm_count() ? p.num_groups : sm_count();
  attn_tc::space_attn_bwd_tc_kernel<<<(unsigned)grid, attn_tc::NTHREADS, attn_tc::BWD_SMEM, (cudaStream_t)stream>>>(tm_qkv, tm_do, tm_st128, tm_sttail, p);
  rc = check_launch("lv_space_attn_bwd_tc_cls");
  if (rc) return rc;
  attn_tc::cls_qkv_finalize_kernel<<<B * H, 192, 0, (cudaStream_t)stream>>>(dcls_kv, dcls_q, (__nv_bfloat16*)dqkv, ld_dqkv, H, p.D, p.clip_rows);
  return check_launch("lv_space_attn_bwd_tc_cls(finalize)");
}

extern "C" int lv_debug_set_buffer(void* buf) {
  long long* p = (long long*)buf;
  cudaError_t e = cudaMemcpyToSymbol(lv::attn_tc::g_dbg, &p, sizeof(p));
  return e == cudaSuccess ? 0 : lv::set_error((int)e, "lv_debug_set_buffer: %s", cudaGetErrorString(e));
}
