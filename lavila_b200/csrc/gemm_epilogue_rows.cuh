// Row-per-thread GEMM epilogue with TMA stores (2-CTA tcgen05 kernel, output-only and dQuickGELU epilogues).
//
// A warp owns 32 accumulator rows (its TMEM lane quarter) x 128 columns.  tcgen05.ld.32x32b hands thread t the 32
// consecutive fp32 columns of row t, so bias / activation run on registers with no shared-memory transposition; the
// bf16 (or fp32) results go into a swizzled [32 rows x 32 columns] staging tile (conflict-free st.shared.v4: the 8 lanes
// of a quarter-warp hit 8 different 16-byte slots of a 128-byte window) and ONE elected lane issues a TMA store of the box
// -- full 64/128-byte row segments per row, clipped by the tensor map at the M / N edges, completion tracked by bulk groups
// instead of per-thread st.global (the legacy epilogue spent 74 % of the LSU pipe there: profiles/ncu_r01_gemm2_b64_summary.txt).
// The next chunk's TMEM load is in flight while the current one is processed, and the accumulator stage is handed back to
// the MMA warp as soon as the last chunk is in registers.
//
// dQuickGELU (fc2 dgrad, lavila/models/timesformer.py:52-58 backward) also reads the saved bf16 pre-activation: the whole
// 32 x 128 slab of the NEXT tile is fetched by TMA into a per-warp 8 KB buffer while the current tile's mainloop runs.
#pragma once
#include "gemm_epilogue.cuh"

namespace lv {
namespace gemm {

constexpr int ROWS_STAGE_BYTES = 4096;   // per warp: 2 x [32 x 32 bf16] (64B swizzle) or 1 x [32 x 32 fp32] (128B swizzle)
constexpr int ROWS_AUX_BYTES = 8192;     // per warp: 4 x [32 x 32 bf16] (64B swizzle), dQuickGELU only

__device__ __forceinline__ uint32_t sw64_off(int r, int j) { return r * 64 + ((j ^ ((r >> 1) & 3)) << 4); }     // j: 16-byte unit 0..3
__device__ __forceinline__ uint32_t sw128_off(int r, int j) { return r * 128 + ((j ^ (r & 7)) << 4); }         // j: 16-byte unit 0..7

__device__ __forceinline__ void st_shared_v4u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read_n() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// Flag sets served by the rows epilogue (everything else keeps the transposing epilogue of gemm_epilogue.cuh).
__host__ __device__ constexpr bool rows_supported(int f) {
  return f == 0 || f == LV_EPI_BIAS || f == (LV_EPI_BIAS | LV_EPI_QUICKGELU) || f == LV_EPI_DQUICKGELU;
}
__host__ __device__ constexpr bool rows_needs_aux(int f) { return (f & LV_EPI_DQUICKGELU) != 0; }
// fp32 output + fp32 residual (proj / fc2 forward: y = x + [tanh(gate) *] (a W^T + b), timesformer.py:181-196): mode 3
__host__ __device__ constexpr bool rows_f32_resid(int f) {
  return f == (LV_EPI_BIAS | LV_EPI_RESID | LV_EPI_OUT_F32) ||
         f == (LV_EPI_BIAS | LV_EPI_SCALE | LV_EPI_SCALE_TANH | LV_EPI_RESID | LV_EPI_OUT_F32);
}
constexpr int ROWS_F32_BYTES = 8192;     // per warp: 2 x [32 rows x 32 fp32] (128B swizzle): residual box in, result box out, in place

// One tile.  stage: this warp's 4 KB staging tile (1024-byte aligned).  aux: this warp's 8 KB slab of the saved pre-activation
// (already requested; aux_bar completes when it has landed).  release(): hands the TMEM accumulator stage back.
template <int FLAGS, class Release>
__device__ __forceinline__ void epilogue_rows_tile(const Args& g, uint8_t* stage, const uint8_t* aux, uint64_t* aux_bar,
                                                   const uint32_t aux_phase, const CUtensorMap* tmO, const CUtensorMap* tmO2,
                                                   uint64_t* tfull, const uint32_t aphase, const uint32_t tmem_acc,
                                                   const int m_base, const int n0, const int half, const int q, const int lane,
                                                   Release release) {
  constexpr int CPW = BN / 64;      // 32-column chunks per warp (4)
  constexpr bool TWO_OUT = (FLAGS & LV_EPI_QUICKGELU) != 0;
  const uint32_t st = smem_u32(stage);
  const uint32_t ax = smem_u32(aux);
  const uint32_t trow = tmem_acc + (uint32_t(q * 32) << 16);
  mbar_wait(tfull, aphase);
  tc_fence_after();
  uint32_t r[2][32];
  const int nbase = n0 + half * CPW * 32;
  if (nbase < g.N) tmem_ld_32x32(trow + half * CPW * 32, r[0]);
  if (FLAGS & LV_EPI_DQUICKGELU) mbar_wait(aux_bar, aux_phase);
#pragma unroll
  for (int cc = 0; cc < CPW; ++cc) {
    const int n = nbase + cc * 32;
    const bool chunk_ok = n < g.N;               // warp-uniform
    const bool next_ok = (cc + 1 < CPW) && (n + 32 < g.N);
    if (chunk_ok) tmem_ld_wait();
    uint32_t (&cur)[32] = r[cc & 1];
    if (next_ok) tmem_ld_32x32(trow + (half * CPW + cc + 1) * 32, r[(cc + 1) & 1]);
    if (!next_ok && chunk_ok) {                  // the last chunk of this warp is in registers: free the accumulator stage
      tc_fence_before();
      __syncwarp();
      release();
    }
    if (!chunk_ok) {
      if (cc == 0) { tc_fence_before(); __syncwarp(); release(); }
      break;
    }
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(cur[j]);
    if (FLAGS & LV_EPI_BIAS) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (n + 4 * k < g.N) {                   // N is a multiple of 4
          const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + n) + k);   // same address in every lane: one broadcast
          v[4 * k] += b.x; v[4 * k + 1] += b.y; v[4 * k + 2] += b.z; v[4 * k + 3] += b.w;
        }
      }
    }
    uint32_t o[16], o2[16];
    if (FLAGS & LV_EPI_QUICKGELU) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint32_t hb = pack_bf16x2(v[2 * k], v[2 * k + 1]);     // the saved pre-activation (bf16), as the legacy epilogue
        o2[k] = hb;
        const float2 h = unpack_bf16x2(hb);
        o[k] = pack_bf16x2(h.x * sigmoidf_fast(1.702f * h.x), h.y * sigmoidf_fast(1.702f * h.y));
      }
    } else if (FLAGS & LV_EPI_DQUICKGELU) {
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const uint4 a4 = ld_shared_v4u(ax + cc * 2048 + sw64_off(lane, k4));
        const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 h = unpack_bf16x2(aw[e]);
          const float s0 = sigmoidf_fast(1.702f * h.x), s1 = sigmoidf_fast(1.702f * h.y);
          const int j = k4 * 8 + e * 2;
          o[k4 * 4 + e] = pack_bf16x2(v[j] * (s0 * (1.0f + 1.702f * h.x * (1.0f - s0))),
                                      v[j + 1] * (s1 * (1.0f + 1.702f * h.y * (1.0f - s1))));
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) o[k] = pack_bf16x2(v[2 * k], v[2 * k + 1]);
    }
    // ---- staging tile(s): a single output alternates the two 2 KB halves (the previous chunk's store may still be reading the
    //      other one); two outputs (fc1: activation + saved pre-activation) use one half each.  Measured alternative for the
    //      second output, st.global.v4 straight from registers (64 contiguous bytes per row, 32 rows per instruction): 846 vs
    //      1081 TF/s standalone -- the uncoalesced stores cost more L1 / LSU cycles than the staging tile costs shared memory.
    const int hsel = TWO_OUT ? 0 : (cc & 1);
    if (lane == 0) {
      if (TWO_OUT) tma_store_wait_read_n<0>();
      else tma_store_wait_read_n<1>();
    }
    __syncwarp();
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4)
      st_shared_v4u(st + hsel * 2048 + sw64_off(lane, k4), o[k4 * 4], o[k4 * 4 + 1], o[k4 * 4 + 2], o[k4 * 4 + 3]);
    if (TWO_OUT) {
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4)
        st_shared_v4u(st + 2048 + sw64_off(lane, k4), o2[k4 * 4], o2[k4 * 4 + 1], o2[k4 * 4 + 2], o2[k4 * 4 + 3]);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tmO, stage + hsel * 2048, n, m_base);
      if (TWO_OUT) tma_store_2d(tmO2, stage + 2048, n, m_base);
      tma_store_commit();
    }
  }
}

// fp32 output with an fp32 residual operand.  Per 32-column chunk the residual box [32 rows x 32 fp32] is fetched by TMA into one
// of two per-warp 4 KB buffers (the next chunk's box is requested before the current one is processed), every thread adds its
// accumulator row IN PLACE (8 conflict-free ld.shared.v4 / st.shared.v4 on the 128B-swizzled tile) and the same buffer is
// handed to a TMA store.  No per-thread global loads or stores: the transposing epilogue needed 8 + 8 of them per thread and
// chunk, each touching 8 rows x 64 bytes.  bars: two mbarriers of this warp (one per buffer); ph: their parities (kept by
// the caller across tiles).
template <int FLAGS, class Release>
__device__ __forceinline__ void epilogue_rows_f32_tile(const Args& g, const float scale, uint8_t* stage, uint64_t* bars, uint32_t (&ph)[2],
                                                       const CUtensorMap* tmO, const CUtensorMap* tmR, uint64_t* tfull,
                                                       const uint32_t aphase, const uint32_t tmem_acc, const int m_base, const int n0,
                                                       const int half, const int q, const int lane, Release release) {
  constexpr int CPW = BN / 64;
  const uint32_t st = smem_u32(stage);
  const uint32_t trow = tmem_acc + (uint32_t(q * 32) << 16);
  const int nbase = n0 + half * CPW * 32;
  auto request = [&](int cc) {          // lane 0: residual box of chunk cc -> buffer cc & 1
    mbar_arrive_expect_tx(&bars[cc & 1], 4096);
    tma_load_2d(stage + (cc & 1) * 4096, tmR, &bars[cc & 1], nbase + cc * 32, m_base);
  };
  if (nbase < g.N && lane == 0) {
    tma_store_wait_read_n<1>();         // buffer 0 was last read by the store of chunk 2 of the previous tile
    request(0);
  }
  mbar_wait(tfull, aphase);
  tc_fence_after();
  uint32_t r[2][32];
  if (nbase < g.N) tmem_ld_32x32(trow + half * CPW * 32, r[0]);
#pragma unroll
  for (int cc = 0; cc < CPW; ++cc) {
    const int n = nbase + cc * 32;
    const bool chunk_ok = n < g.N;
    const bool next_ok = (cc + 1 < CPW) && (n + 32 < g.N);
    if (chunk_ok) tmem_ld_wait();
    uint32_t (&cur)[32] = r[cc & 1];
    if (next_ok) tmem_ld_32x32(trow + (half * CPW + cc + 1) * 32, r[(cc + 1) & 1]);
    if (!next_ok && chunk_ok) {
      tc_fence_before();
      __syncwarp();
      release();
    }
    if (!chunk_ok) {
      if (cc == 0) { tc_fence_before(); __syncwarp(); release(); }
      break;
    }
    if (next_ok && lane == 0) {
      tma_store_wait_read_n<0>();       // the other buffer: its store (chunk cc - 1) must have been read before it is refilled
      request(cc + 1);
    }
    mbar_wait(&bars[cc & 1], ph[cc & 1]);
    ph[cc & 1] ^= 1u;
    const uint32_t buf = st + (cc & 1) * 4096;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t addr = buf + sw128_off(lane, k);
      const uint4 rr = ld_shared_v4u(addr);
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((FLAGS & LV_EPI_BIAS) && n + 4 * k < g.N) b = __ldg(reinterpret_cast<const float4*>(g.bias + n) + k);
      float v0 = __uint_as_float(cur[4 * k]) + b.x, v1 = __uint_as_float(cur[4 * k + 1]) + b.y;
      float v2 = __uint_as_float(cur[4 * k + 2]) + b.z, v3 = __uint_as_float(cur[4 * k + 3]) + b.w;
      if (FLAGS & LV_EPI_SCALE) { v0 *= scale; v1 *= scale; v2 *= scale; v3 *= scale; }
      v0 += __uint_as_float(rr.x); v1 += __uint_as_float(rr.y); v2 += __uint_as_float(rr.z); v3 += __uint_as_float(rr.w);
      st_shared_v4u(addr, __float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2), __float_as_uint(v3));
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tmO, stage + (cc & 1) * 4096, n, m_base);
      tma_store_commit();
    }
  }
}

}  // namespace gemm
}  // namespace lv
