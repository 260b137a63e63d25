// LayerNorm forward / backward over the last dimension -- HBM-bound, one warp per row, 128-bit accesses.
//
// forward : x fp32 [rows, D] (row stride ldx) -> y bf16 (GEMM A operand) and/or y fp32
// backward: dx = LN'(dy) (+ add1 + add2), dgamma += sum dy*xhat, dbeta += sum dy   (mean/rstd recomputed from x)
//
// Replaces nn.LayerNorm at lavila/models/timesformer.py:180,189,196,366,377 and openai_model.py:196-204 /
// models.py:156 (eps 1e-6 for norm1/2/3 + final norm, 1e-5 for ln_pre and the text tower).  Under CUDA autocast
// the reference keeps LayerNorm input/output in fp32 and then casts for the following bf16 GEMM; here the cast is
// fused into the store.
#include "../../include/lavila_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace lv {
namespace ln {

constexpr int MAX_V = 8;            // float4 per lane -> D <= 1024
constexpr int WARPS = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(WARPS * 32)
ln_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y_bf16, long long ldy,
              float* __restrict__ y_f32, long long ldyf, long long rows, int D) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = D >> 7;  // float4 per lane
  const float* xr = x + row * ldx;
  float4 v[MAX_V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_V; ++i)
    if (i < nv) {
      v[i] = __ldg(reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4));
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  const float mean = warp_sum(s) / D;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_V; ++i)
    if (i < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      ss += a * a + b * b + c * c + d * d;
    }
  const float rstd = rsqrtf(warp_sum(ss) / D + eps);
#pragma unroll
  for (int i = 0; i < MAX_V; ++i)
    if (i < nv) {
      const int col = (i * 32 + lane) * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (y_bf16)
        *reinterpret_cast<uint2*>(y_bf16 + row * ldy + col) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      if (y_f32) *reinterpret_cast<float4*>(y_f32 + row * ldyf + col) = o;
    }
}

// Any width with D % 4 == 0 (e.g. GPT-2 XL's 1600): warp per row, three passes over the (L1-resident) row.
__global__ void __launch_bounds__(WARPS * 32)
ln_fwd_generic_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y_bf16, long long ldy,
                      float* __restrict__ y_f32, long long ldyf, long long rows, int D) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * ldx;
  float s = 0.f;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    s += v.x + v.y + v.z + v.w;
  }
  const float mean = warp_sum(s) / D;
  float ss = 0.f;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float a = v.x - mean, b = v.y - mean, cc = v.z - mean, d = v.w - mean;
    ss += a * a + b * b + cc * cc + d * d;
  }
  const float rstd = rsqrtf(warp_sum(ss) / D + eps);
  for (int c = lane * 4; c < D; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (beta) b = __ldg(reinterpret_cast<const float4*>(beta + c));
    float4 o;
    o.x = (v.x - mean) * rstd * g.x + b.x;
    o.y = (v.y - mean) * rstd * g.y + b.y;
    o.z = (v.z - mean) * rstd * g.z + b.z;
    o.w = (v.w - mean) * rstd * g.w + b.w;
    if (y_bf16) *reinterpret_cast<uint2*>(y_bf16 + row * ldy + c) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    if (y_f32) *reinterpret_cast<float4*>(y_f32 + row * ldyf + c) = o;
  }
}

// Persistent: each warp walks rows with a grid stride, keeps dgamma/dbeta partials in registers, the CTA reduces
// them through shared memory and issues one red.add per column.
// HBM-latency-bound unless enough bytes are in flight: NV (float4 per lane) is a compile-time constant so that ALL
// global loads of R rows (x, dy, residual-path gradients) are issued back to back before the first use
// (R = 2 rows per warp iteration with <= 1 residual operand, 1 row with 2: ~12-15 KB in flight per warp).
template <int NV, int R, int NADD, bool DY_BF16>
__global__ void __launch_bounds__(WARPS * 32)
ln_bwd_kernel(const void* __restrict__ dy_, long long lddy, const float* __restrict__ x, long long ldx,
              const float* __restrict__ gamma, float eps, const float* __restrict__ add1, long long ld1,
              const float* __restrict__ add2, long long ld2, float* __restrict__ dx, long long lddx,
              __nv_bfloat16* __restrict__ dx_bf16, long long lddxb, float* __restrict__ dgamma,
              float* __restrict__ dbeta, long long rows) {
  extern __shared__ float red[];  // [WARPS][2][D]
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float4 dg[NV], db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const long long stride = (long long)gridDim.x * WARPS * R;
  for (long long row0 = ((long long)blockIdx.x * WARPS + warp) * R; row0 < rows; row0 += stride) {
    float4 xv[R][NV], gv[R][NV], a1v[R][NV], a2v[R][NV];
    // ---- issue every load of the R rows
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = row0 + r < rows ? row0 + r : rows - 1;   // tail: recompute the last row, store is masked
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = (i * 32 + lane) * 4;
        xv[r][i] = __ldg(reinterpret_cast<const float4*>(x + row * ldx + col));
        if (DY_BF16) {
          const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy_) + row * lddy + col));
          gv[r][i] = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u),
                                 __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
        } else {
          gv[r][i] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + row * lddy + col));
        }
        if (NADD >= 1) a1v[r][i] = __ldg(reinterpret_cast<const float4*>(add1 + row * ld1 + col));
        if (NADD >= 2) a2v[r][i] = __ldg(reinterpret_cast<const float4*>(add2 + row * ld2 + col));
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long row = row0 + r;
      const bool live = row < rows;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) s += xv[r][i].x + xv[r][i].y + xv[r][i].z + xv[r][i].w;
      const float mean = warp_sum(s) / D;
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        xv[r][i].x -= mean; xv[r][i].y -= mean; xv[r][i].z -= mean; xv[r][i].w -= mean;
        ss += xv[r][i].x * xv[r][i].x + xv[r][i].y * xv[r][i].y + xv[r][i].z * xv[r][i].z + xv[r][i].w * xv[r][i].w;
      }
      const float rstd = rsqrtf(warp_sum(ss) / D + eps);
      float c1 = 0.f, c2 = 0.f;
      const float lv_ = live ? 1.f : 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float4& xh = xv[r][i];
        float4& g = gv[r][i];
        xh.x *= rstd; xh.y *= rstd; xh.z *= rstd; xh.w *= rstd;  // xhat
        dg[i].x += lv_ * g.x * xh.x; dg[i].y += lv_ * g.y * xh.y; dg[i].z += lv_ * g.z * xh.z; dg[i].w += lv_ * g.w * xh.w;
        db[i].x += lv_ * g.x; db[i].y += lv_ * g.y; db[i].z += lv_ * g.z; db[i].w += lv_ * g.w;
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 4));  // L1-resident
        g.x *= gm.x; g.y *= gm.y; g.z *= gm.z; g.w *= gm.w;  // dy * gamma
        c1 += g.x + g.y + g.z + g.w;
        c2 += g.x * xh.x + g.y * xh.y + g.z * xh.z + g.w * xh.w;
      }
      c1 = warp_sum(c1) / D;
      c2 = warp_sum(c2) / D;
      if (live) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int col = (i * 32 + lane) * 4;
          float4 o;
          o.x = rstd * (gv[r][i].x - c1 - xv[r][i].x * c2);
          o.y = rstd * (gv[r][i].y - c1 - xv[r][i].y * c2);
          o.z = rstd * (gv[r][i].z - c1 - xv[r][i].z * c2);
          o.w = rstd * (gv[r][i].w - c1 - xv[r][i].w * c2);
          if (NADD >= 1) { o.x += a1v[r][i].x; o.y += a1v[r][i].y; o.z += a1v[r][i].z; o.w += a1v[r][i].w; }
          if (NADD >= 2) { o.x += a2v[r][i].x; o.y += a2v[r][i].y; o.z += a2v[r][i].z; o.w += a2v[r][i].w; }
          if (dx) *reinterpret_cast<float4*>(dx + row * lddx + col) = o;
          if (dx_bf16)
            *reinterpret_cast<uint2*>(dx_bf16 + row * lddxb + col) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        }
      }
    }
  }
  if (!dgamma) return;
  float* rg = red + (size_t)warp * 2 * D;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * 32 + lane) * 4;
    *reinterpret_cast<float4*>(rg + col) = dg[i];
    *reinterpret_cast<float4*>(rg + D + col) = db[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) t += red[(size_t)w * 2 * D + c];
    if (c < D) red_add_f32(dgamma + c, t);
    else red_add_f32(dbeta + (c - D), t);
  }
}

template <int NV, int R, int NADD, bool DY_BF16>
static int launch_ln_bwd(const void* dy, long long lddy, const float* x, long long ldx, const float* gamma, float eps,
                         const float* add1, long long ld1, const float* add2, long long ld2, float* dx, long long lddx,
                         void* dx_bf16, long long lddxb, float* dgamma, float* dbeta, long long rows, cudaStream_t st) {
  constexpr int D = NV * 128;
  const size_t smem = (size_t)WARPS * 2 * D * sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(ln_bwd_kernel<NV, R, NADD, DY_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    configured = true;
  }
  long long want = (rows + WARPS * R - 1) / (WARPS * R);
  long long cap = (long long)sm_count() * 4;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  ln_bwd_kernel<NV, R, NADD, DY_BF16><<<grid, WARPS * 32, smem, st>>>(dy, lddy, x, ldx, gamma, eps, add1, ld1, add2, ld2, dx,
                                                                      lddx, (__nv_bfloat16*)dx_bf16, lddxb, dgamma, dbeta, rows);
  return check_launch("lv_layernorm_bwd");
}

template <int NV, bool DY_BF16>
static int dispatch_ln_bwd(int nadd, const void* dy, long long lddy, const float* x, long long ldx, const float* gamma,
                           float eps, const float* add1, long long ld1, const float* add2, long long ld2, float* dx,
                           long long lddx, void* dx_bf16, long long lddxb, float* dgamma, float* dbeta, long long rows,
                           cudaStream_t st) {
  constexpr int R01 = NV <= 6 ? 2 : 1;   // two rows in flight while registers allow
  if (nadd == 0) return launch_ln_bwd<NV, R01, 0, DY_BF16>(dy, lddy, x, ldx, gamma, eps, add1, ld1, add2, ld2, dx, lddx, dx_bf16, lddxb, dgamma, dbeta, rows, st);
  if (nadd == 1) return launch_ln_bwd<NV, R01, 1, DY_BF16>(dy, lddy, x, ldx, gamma, eps, add1, ld1, add2, ld2, dx, lddx, dx_bf16, lddxb, dgamma, dbeta, rows, st);
  return launch_ln_bwd<NV, 1, 2, DY_BF16>(dy, lddy, x, ldx, gamma, eps, add1, ld1, add2, ld2, dx, lddx, dx_bf16, lddxb, dgamma, dbeta, rows, st);
}

}  // namespace ln
}  // namespace lv

extern "C" int lv_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps,
                                void* y_bf16, int64_t ldy, float* y_f32, int64_t ldyf, int64_t rows, int D,
                                void* stream) {
  using namespace lv;
  LV_REQUIRE(x && gamma && (y_bf16 || y_f32), "lv_layernorm_fwd: null pointer");
  LV_REQUIRE(D > 0 && D % 4 == 0, "lv_layernorm_fwd: D=%d must be a multiple of 4", D);
  LV_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0 && ldyf % 4 == 0, "lv_layernorm_fwd: leading dimensions must be multiples of 4");
  if (rows <= 0) return 0;
  const unsigned grid = (unsigned)((rows + ln::WARPS - 1) / ln::WARPS);
  if (beta && D % 128 == 0 && D <= 128 * ln::MAX_V)
    ln::ln_fwd_kernel<<<grid, ln::WARPS * 32, 0, (cudaStream_t)stream>>>(x, ldx, gamma, beta, eps, (__nv_bfloat16*)y_bf16, ldy,
                                                                        y_f32, ldyf, rows, D);
  else   // beta may be NULL (coca.LayerNorm keeps a zero buffer), any D % 4 == 0
    ln::ln_fwd_generic_kernel<<<grid, ln::WARPS * 32, 0, (cudaStream_t)stream>>>(x, ldx, gamma, beta, eps, (__nv_bfloat16*)y_bf16,
                                                                                ldy, y_f32, ldyf, rows, D);
  return check_launch("lv_layernorm_fwd");
}

extern "C" int lv_layernorm_bwd(const void* dy, int dy_is_bf16, int64_t lddy, const float* x, int64_t ldx,
                                const float* gamma, float eps, const float* add1, int64_t ld1, const float* add2,
                                int64_t ld2, float* dx, int64_t lddx, void* dx_bf16, int64_t lddxb, float* dgamma,
                                float* dbeta, int64_t rows, int D, void* stream) {
  using namespace lv;
  LV_REQUIRE(dy && x && gamma && (dx || dx_bf16), "lv_layernorm_bwd: null pointer");
  LV_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "lv_layernorm_bwd: dgamma/dbeta must both be given or both null");
  LV_REQUIRE(D > 0 && D % 128 == 0 && D <= 128 * ln::MAX_V, "lv_layernorm_bwd: D=%d must be a multiple of 128 and <= 1024", D);
  LV_REQUIRE(lddy % 4 == 0 && ldx % 4 == 0 && ld1 % 4 == 0 && ld2 % 4 == 0 && lddx % 4 == 0 && lddxb % 4 == 0,
             "lv_layernorm_bwd: leading dimensions must be multiples of 4");
  if (rows <= 0) return 0;
  LV_REQUIRE(!(add2 && !add1), "lv_layernorm_bwd: add2 without add1");
  const int nadd = add1 ? (add2 ? 2 : 1) : 0;
  cudaStream_t st = (cudaStream_t)stream;
#define LV_LN_BWD(NV)                                                                                                      \
  return dy_is_bf16 ? ln::dispatch_ln_bwd<NV, true>(nadd, dy, lddy, x, ldx, gamma, eps, add1, ld1, add2, ld2, dx, lddx,      \
                                                    dx_bf16, lddxb, dgamma, dbeta, rows, st)                                \
                    : ln::dispatch_ln_bwd<NV, false>(nadd, dy, lddy, x, ldx, gamma, eps, add1, ld1, add2, ld2, dx, lddx,     \
                                                     dx_bf16, lddxb, dgamma, dbeta, rows, st)
  switch (D / 128) {
    case 1: LV_LN_BWD(1);
    case 2: LV_LN_BWD(2);
    case 4: LV_LN_BWD(4);
    case 6: LV_LN_BWD(6);
    case 8: LV_LN_BWD(8);
    default: return set_error(-1, "lv_layernorm_bwd: D=%d not instantiated (128, 256, 512, 768, 1024)", D);
  }
#undef LV_LN_BWD
}
