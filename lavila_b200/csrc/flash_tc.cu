// Key-tiled ("flash") attention forward on tcgen05 tensor cores, head_dim 64, any number of keys.
//
// One work item = 128 query rows of one (batch, group, head).  Per 128-key tile:
//     S = Q K^T          4 x tcgen05.mma 128 x N x 16 into TMEM columns [0, 128)      (N = 128, or the 16-multiple tail)
//     softmax            4 warps, thread = query row (TMEM lane): running max / sum in registers, P (bf16) into a
//                        128B-swizzled K-major staging tile in shared memory
//     O_t = P V          8 x tcgen05.mma 128 x 64 x 16 into TMEM columns [128, 192), V consumed MN-major from its TMA tile
//     O = O * alpha + O_t in registers (64 fp32 per thread): the running rescale never touches TMEM
// Q / K / V tiles arrive by TMA straight from the caller's matrices (row pitch and column offset per head, so the packed
// qkv projection output, the narrator's [k | v] cross-attention buffer and a shared multi-query K/V head are all addressed
// in place); K / V are double buffered.  Two CTAs share an SM (112 KB shared memory, 256 TMEM columns each), so one CTA's
// softmax overlaps the other's MMAs.  Warp roles: warp 0 = TMA producer (+ tail fix-ups), warp 1 = MMA issuer, warps 2..5 =
// softmax / epilogue (warp w owns TMEM lanes 32 * (w % 4) ..).
//
// Replaces, for >= 64 query rows per group:
//   * the space attention of TimeSformer-L/14 (257 / 577 keys per group: lavila/models/timesformer.py:121-134 at the geometries
//     of lavila/models/models.py:374-491) -- the extra CLS key / value row of the group is appended to the last key tile;
//   * the narrator's attention forward: CoCa multi-query pooling (lavila/models/coca.py:100-125), gated cross-attention
//     prefill (lavila/models/gpt2_gated.py:320-360, _attn :206-238) and GPT-2 causal self-attention prefill (:464-477).
// Decoding steps (one query row per sequence) stay on the mma.sync kernels of flash_attn.cu: a 128-row MMA tile would be empty.
#include <cstdlib>
#include <mutex>

#include "../../include/lavila_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace lv {
namespace ftc {

constexpr int HD = 64;
constexpr int BQ = 128, BKEY = 128;
constexpr int Q_BYTES = BQ * 128;          // 16 KB
constexpr int KV_BYTES = BKEY * 128;       // 16 KB per K or V stage
constexpr int P_BYTES = 2 * 128 * 128;     // two 64-key atoms of [128 rows x 128 B]
constexpr int NTHREADS = 192;
constexpr int TM_S = 0, TM_O = 128, TMEM_COLS = 256;
constexpr int SMEM_BYTES = 1024 + Q_BYTES + 4 * KV_BYTES + P_BYTES + 128;
constexpr float LOG2E = 1.4426950408889634f;

struct Params {
  __nv_bfloat16* out;
  long long ld_out;
  float* lse;                 // optional [row][H]
  const __nv_bfloat16* kx;    // matrices the extra key / value row is read from (generic loads), with its pitch
  const __nv_bfloat16* vx;
  long long ld_kv;
  long long q_rows, kv_rows;            // rows per batch element
  long long q_grp_row0, q_grp_stride;   // first query row of group g inside a batch element: q_grp_row0 + g * q_grp_stride
  long long kv_grp_row0, kv_grp_stride;
  int G, H, Lq, Lk, has_extra, causal;
  int kv_head_stride;
  int QT;                               // query tiles per group
  long long num_items;
  float scale;
};

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

struct Item {
  int b, g, h, qt;
  long long q_row, kv_row, extra_row;
};
__device__ __forceinline__ Item decode(const Params& p, long long it) {
  Item c;
  c.qt = (int)(it % p.QT);
  long long r = it / p.QT;
  c.g = (int)(r % p.G);
  r /= p.G;
  c.h = (int)(r % p.H);
  c.b = (int)(r / p.H);
  c.q_row = (long long)c.b * p.q_rows + p.q_grp_row0 + (long long)c.g * p.q_grp_stride + (long long)c.qt * BQ;
  c.kv_row = (long long)c.b * p.kv_rows + p.kv_grp_row0 + (long long)c.g * p.kv_grp_stride;
  c.extra_row = (long long)c.b * p.kv_rows;
  return c;
}
// number of key tiles query tile qt has to visit (causal: keys j <= i + (Lk - Lq))
__device__ __forceinline__ int key_tiles(const Params& p, int qt) {
  const int Lkt = p.Lk + p.has_extra;
  int last = Lkt - 1;
  if (p.causal) {
    const int imax = min(p.Lq, (qt + 1) * BQ) - 1;
    last = min(last, imax + (p.Lk - p.Lq));
  }
  return last / BKEY + 1;
}

__global__ void __launch_bounds__(NTHREADS, 2)
flash_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;            // [2]
  uint8_t* sV = sK + 2 * KV_BYTES;       // [2]
  uint8_t* sP = sV + 2 * KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
  uint64_t* bar_q = bars;            // Q landed
  uint64_t* bar_qfree = bars + 1;    // all S MMAs of the item retired: the Q tile may be overwritten
  uint64_t* bar_tma = bars + 2;      // [2] K, V bytes of stage s landed (TMA)
  uint64_t* bar_kv = bars + 4;       // [2] stage s ready for the MMA warp (after the producer's fix-ups)
  uint64_t* bar_kvfree = bars + 6;   // [2] the MMAs reading stage s retired
  uint64_t* bar_s = bars + 8;        // S ready
  uint64_t* bar_p = bars + 9;        // P staged, S consumed (4 warps)
  uint64_t* bar_o = bars + 10;       // O_t ready
  uint64_t* bar_ofree = bars + 11;   // O_t consumed (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Lkt = p.Lk + p.has_extra;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(bar_q, 1);
    mbar_init(bar_qfree, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_tma[s], 1);
      mbar_init(&bar_kv[s], 1);
      mbar_init(&bar_kvfree[s], 1);
    }
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 4);
    mbar_init(bar_o, 1);
    mbar_init(bar_ofree, 4);
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

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    uint32_t n_q = 0, n_kv = 0;         // Q tiles / K-V stages issued so far (barrier parities)
    for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      const Item c = decode(p, it);
      const int KT = key_tiles(p, c.qt);
      if (n_q > 0) {
        if (lane == 0) mbar_wait(bar_qfree, (n_q - 1) & 1);
        __syncwarp();
      }
      if (lane == 0) {
        mbar_arrive_expect_tx(bar_q, Q_BYTES);
        tma_load_2d(sQ, &tmQ, bar_q, c.h * HD, (int)c.q_row);
      }
      ++n_q;
      for (int kt = 0; kt < KT; ++kt, ++n_kv) {
        const int s = n_kv & 1;
        const uint32_t use = n_kv >> 1;             // how many times stage s has been filled before
        if (use > 0) {
          if (lane == 0) mbar_wait(&bar_kvfree[s], (use - 1) & 1);
          __syncwarp();
        }
        uint8_t* k_dst = sK + s * KV_BYTES;
        uint8_t* v_dst = sV + s * KV_BYTES;
        if (lane == 0) {
          mbar_arrive_expect_tx(&bar_tma[s], 2 * KV_BYTES);
          tma_load_2d(k_dst, &tmK, &bar_tma[s], c.h * p.kv_head_stride, (int)(c.kv_row + (long long)kt * BKEY));
          tma_load_2d(v_dst, &tmV, &bar_tma[s], c.h * p.kv_head_stride, (int)(c.kv_row + (long long)kt * BKEY));
          mbar_wait(&bar_tma[s], use & 1);
        }
        __syncwarp();
        // tail tile: rows past the group's keys hold whatever follows in memory.  Append the extra (CLS) key / value row
        // and clear the V rows up to the next multiple of 16 (P is 0 there, but 0 x NaN must not reach the accumulator).
        const int valid = min(BKEY, Lkt - kt * BKEY);
        if (valid < BKEY) {
          const int n16 = (valid + 15) & ~15;
          if (p.has_extra && lane < 16) {
            const int part = lane >> 3, ch = lane & 7, r = valid - 1;
            const __nv_bfloat16* src = (part ? p.vx : p.kx) + c.extra_row * p.ld_kv + c.h * p.kv_head_stride + ch * 8;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
            st_shared_v4(smem_u32(part ? v_dst : k_dst) + r * 128 + ((ch ^ (r & 7)) << 4), v.x, v.y, v.z, v.w);
          }
          for (int idx = lane; idx < (n16 - valid) * 8; idx += 32) {
            const int r = valid + (idx >> 3), ch = idx & 7;
            st_shared_v4(smem_u32(v_dst) + r * 128 + ((ch ^ (r & 7)) << 4), 0, 0, 0, 0);
          }
          fence_proxy_async_smem();
          __syncwarp();
        }
        if (lane == 0) mbar_arrive(&bar_kv[s]);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_o = make_idesc_bf16(128, HD, 0, 1);
      const uint32_t q_base = smem_u32(sQ), p_base = smem_u32(sP);
      uint32_t n_q = 0, n_kv = 0, n_t = 0;       // items, K-V stages, key tiles processed so far
      for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x, ++n_q) {
        const Item c = decode(p, it);
        const int KT = key_tiles(p, c.qt);
        mbar_wait(bar_q, n_q & 1);
        auto issue_s = [&](int kt, uint32_t stage_idx) {
          const int s = stage_idx & 1;
          mbar_wait(&bar_kv[s], (stage_idx >> 1) & 1);
          tc_fence_after();
          const int valid = min(BKEY, Lkt - kt * BKEY);
          const uint32_t idesc_s = make_idesc_bf16(128, (valid + 15) & ~15, 0, 0);
          const uint32_t k_base = smem_u32(sK + s * KV_BYTES);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            tc_mma_bf16(tmem_base + TM_S, make_smem_desc_sw128(q_base + ks * 32, 16, 1024),
                        make_smem_desc_sw128(k_base + ks * 32, 16, 1024), idesc_s, ks > 0);
          tc_commit(bar_s);
        };
        issue_s(0, n_kv);
        for (int kt = 0; kt < KT; ++kt, ++n_kv, ++n_t) {
          const int s = n_kv & 1;
          mbar_wait(bar_p, n_t & 1);                           // P(kt) staged; S(kt) consumed
          if (n_t > 0) mbar_wait(bar_ofree, (n_t - 1) & 1);    // O_t of the previous tile has been read
          tc_fence_after();
          const int valid = min(BKEY, Lkt - kt * BKEY);
          const int nsteps = (valid + 15) >> 4;
          const uint32_t v_base = smem_u32(sV + s * KV_BYTES);
          for (int st = 0; st < nsteps; ++st)
            tc_mma_bf16(tmem_base + TM_O, make_smem_desc_sw128(p_base + (st >> 2) * 16384 + (st & 3) * 32, 16, 1024),
                        make_smem_desc_sw128(v_base + st * 2048, 8192, 1024), idesc_o, st > 0);
          tc_commit(bar_o);
          tc_commit(&bar_kvfree[s]);
          if (kt + 1 < KT) issue_s(kt + 1, n_kv + 1);
          else tc_commit(bar_qfree);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue warps
    const int q = warp & 3;
    const int row = q * 32 + lane;                  // TMEM lane == row of the query tile
    const uint32_t trow = tmem_base + (uint32_t(q * 32) << 16);
    const uint32_t p_row = smem_u32(sP) + row * 128;
    const float sl2 = p.scale * LOG2E;
    uint32_t n_t = 0;
    for (long long it = blockIdx.x; it < p.num_items; it += gridDim.x) {
      const Item c = decode(p, it);
      const int KT = key_tiles(p, c.qt);
      const int i_q = c.qt * BQ + row;              // query index inside the group
      const int jmax = p.causal ? i_q + (p.Lk - p.Lq) : (1 << 30);   // last visible key
      float o[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) o[j] = 0.f;
      float m = -INFINITY, l = 0.f;
      for (int kt = 0; kt < KT; ++kt, ++n_t) {
        const int valid = min(BKEY, Lkt - kt * BKEY);
        const int vis = min(valid, jmax - kt * BKEY + 1);          // keys of this tile visible to this row (may be <= 0)
        const int n16 = (valid + 15) & ~15;
        mbar_wait(bar_s, n_t & 1);
        tc_fence_after();
        // ---- pass 1: tile maximum over the visible keys
        float tmax = -INFINITY;
#pragma unroll 1
        for (int cc = 0; cc * 32 < n16; ++cc) {
          uint32_t r[32];
          tmem_ld_32x32(trow + TM_S + cc * 32, r);
          tmem_ld_wait();
          if ((cc + 1) * 32 <= vis) {
#pragma unroll
            for (int j = 0; j < 32; ++j) tmax = fmaxf(tmax, __uint_as_float(r[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (cc * 32 + j < vis) tmax = fmaxf(tmax, __uint_as_float(r[j]));
          }
        }
        const float m_new = fmaxf(m, tmax);
        const float mb = (m_new == -INFINITY) ? 0.f : m_new * sl2;
        const float alpha = ex2_fast(m * sl2 - mb);                // m = -inf on the first tile: 0
        // ---- pass 2: P = exp2(S * scale * log2e - mb), row sum, bf16 P into the swizzled K-major tile
        float sum = 0.f;
#pragma unroll 1
        for (int cc = 0; cc * 32 < n16; ++cc) {
          uint32_t r[32];
          tmem_ld_32x32(trow + TM_S + cc * 32, r);
          tmem_ld_wait();
          float pv[32];
          if ((cc + 1) * 32 <= vis) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              pv[j] = ex2_fast(fmaf(__uint_as_float(r[j]), sl2, -mb));
              sum += pv[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              pv[j] = (cc * 32 + j < vis) ? ex2_fast(fmaf(__uint_as_float(r[j]), sl2, -mb)) : 0.f;
              sum += pv[j];
            }
          }
          const uint32_t atom = p_row + (cc >> 1) * 16384;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int chunk = (cc & 1) * 4 + jj;
            st_shared_v4(atom + ((chunk ^ (row & 7)) << 4), pack_bf16x2(pv[jj * 8], pv[jj * 8 + 1]),
                         pack_bf16x2(pv[jj * 8 + 2], pv[jj * 8 + 3]), pack_bf16x2(pv[jj * 8 + 4], pv[jj * 8 + 5]),
                         pack_bf16x2(pv[jj * 8 + 6], pv[jj * 8 + 7]));
          }
        }
        m = m_new;
        l = l * alpha + sum;
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
        // ---- O = O * alpha + P V
        mbar_wait(bar_o, n_t & 1);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(trow + TM_O, o0);
        tmem_ld_32x32(trow + TM_O + 32, o1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ofree);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          o[j] = fmaf(o[j], alpha, __uint_as_float(o0[j]));
          o[32 + j] = fmaf(o[32 + j], alpha, __uint_as_float(o1[j]));
        }
      }
      // ---- epilogue: O / l -> bf16 -> global (each thread writes its row's 128 contiguous bytes)
      if (i_q < p.Lq) {
        const float inv = l > 0.f ? 1.f / l : 0.f;
        const long long grow = c.q_row + row;
        __nv_bfloat16* dst = p.out + grow * p.ld_out + c.h * HD;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          *reinterpret_cast<uint4*>(dst + jj * 8) =
              make_uint4(pack_bf16x2(o[jj * 8] * inv, o[jj * 8 + 1] * inv), pack_bf16x2(o[jj * 8 + 2] * inv, o[jj * 8 + 3] * inv),
                         pack_bf16x2(o[jj * 8 + 4] * inv, o[jj * 8 + 5] * inv), pack_bf16x2(o[jj * 8 + 6] * inv, o[jj * 8 + 7] * inv));
        if (p.lse) p.lse[grow * p.H + c.h] = m * p.scale + logf(l);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace ftc

bool flash_tc_enabled() {
  static const bool on = []() { const char* e = std::getenv("LAVILA_B200_FLASH_TC"); return !(e && e[0] == '0'); }();
  return on;
}

// Host launcher shared by the narrator entry point (lv_flash_attn_fwd) and the TimeSformer-L space attention
// (big_group_attn_fwd, mode 0).  q / k / v: first element of head 0 of batch element 0; *_cols: accessible columns from there.
int flash_attn_fwd_tc(const void* q, long long ld_q, long long q_rows, int q_cols, const void* k, const void* v, long long ld_kv,
                      long long kv_rows, int kv_cols, int kv_head_stride, void* out, long long ld_out, float* lse, int B, int G,
                      int H, int Lq, int Lk, long long q_grp_row0, long long q_grp_stride, long long kv_grp_row0,
                      long long kv_grp_stride, int has_extra, int causal, float scale, cudaStream_t st) {
  using namespace ftc;
  Params p{};
  p.out = (__nv_bfloat16*)out; p.ld_out = ld_out; p.lse = lse;
  p.kx = (const __nv_bfloat16*)k; p.vx = (const __nv_bfloat16*)v; p.ld_kv = ld_kv;
  p.q_rows = q_rows; p.kv_rows = kv_rows;
  p.q_grp_row0 = q_grp_row0; p.q_grp_stride = q_grp_stride; p.kv_grp_row0 = kv_grp_row0; p.kv_grp_stride = kv_grp_stride;
  p.G = G; p.H = H; p.Lq = Lq; p.Lk = Lk; p.has_extra = has_extra; p.causal = causal; p.kv_head_stride = kv_head_stride;
  p.QT = (Lq + BQ - 1) / BQ;
  p.num_items = (long long)B * G * H * p.QT;
  p.scale = scale;
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_tmap_2d_bf16(&tmQ, q, (uint64_t)q_cols, (uint64_t)((long long)B * q_rows), (uint64_t)ld_q, 64, BQ);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmK, k, (uint64_t)kv_cols, (uint64_t)((long long)B * kv_rows), (uint64_t)ld_kv, 64, BKEY);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmV, v, (uint64_t)kv_cols, (uint64_t)((long long)B * kv_rows), (uint64_t)ld_kv, 64, BKEY);
  if (rc) return rc;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, []() {
    attr_err = cudaFuncSetAttribute(flash_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return set_error((int)attr_err, "flash attention (tcgen05): cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  const long long cap = 2ll * sm_count();
  const long long grid = p.num_items < cap ? p.num_items : cap;
  flash_fwd_tc_kernel<<<(unsigned)grid, NTHREADS, SMEM_BYTES, st>>>(tmQ, tmK, tmV, p);
  return check_launch("flash attention forward (tcgen05)");
}

}  // namespace lv
