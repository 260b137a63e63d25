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

// Persistent: each warp walks rows with a grid stride, keeps dgamma/dbeta partials in registers, the CTA reduces
// them through shared memory and issues one red.add per column.
template <bool DY_BF16>
__global__ void __launch_bounds__(WARPS * 32)
ln_bwd_kernel(const void* __restrict__ dy_, long long lddy, const float* __restrict__ x, long long ldx,
              const float* __restrict__ gamma, float eps, const float* __restrict__ add1, long long ld1,
              const float* __restrict__ add2, long long ld2, float* __restrict__ dx, long long lddx,
              __nv_bfloat16* __restrict__ dx_bf16, long long lddxb, float* __restrict__ dgamma,
              float* __restrict__ dbeta, long long rows, int D) {
  extern __shared__ float red[];  // [WARPS][2][D]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nv = D >> 7;
  float4 dg[MAX_V], db[MAX_V];
#pragma unroll
  for (int i = 0; i < MAX_V; ++i) {
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long row = (long long)blockIdx.x * WARPS + warp; row < rows; row += (long long)gridDim.x * WARPS) {
    const float* xr = x + row * ldx;
    float4 xv[MAX_V], gv[MAX_V], av[MAX_V];
    float s = 0.f;
    // all global loads of the row are issued up front (x, dy, residual-path gradients): one exposed latency per row
#pragma unroll
    for (int i = 0; i < MAX_V; ++i)
      if (i < nv) {
        const int col = (i * 32 + lane) * 4;
        av[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (add1) av[i] = __ldg(reinterpret_cast<const float4*>(add1 + row * ld1 + col));
        if (add2) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(add2 + row * ld2 + col));
          av[i].x += a.x; av[i].y += a.y; av[i].z += a.z; av[i].w += a.w;
        }
      }
#pragma unroll
    for (int i = 0; i < MAX_V; ++i)
      if (i < nv) {
        const int col = (i * 32 + lane) * 4;
        xv[i] = __ldg(reinterpret_cast<const float4*>(xr + col));
        s += xv[i].x + xv[i].y + xv[i].z + xv[i].w;
        if (DY_BF16) {
          const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy_) + row * lddy + col));
          const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
          gv[i] = make_float4(a.x, a.y, b.x, b.y);
        } else {
          gv[i] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + row * lddy + col));
        }
      }
    const float mean = warp_sum(s) / D;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_V; ++i)
      if (i < nv) {
        xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
        ss += xv[i].x * xv[i].x + xv[i].y * xv[i].y + xv[i].z * xv[i].z + xv[i].w * xv[i].w;
      }
    const float rstd = rsqrtf(warp_sum(ss) / D + eps);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_V; ++i)
      if (i < nv) {
        xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;  // xhat
        dg[i].x += gv[i].x * xv[i].x; dg[i].y += gv[i].y * xv[i].y; dg[i].z += gv[i].z * xv[i].z; dg[i].w += gv[i].w * xv[i].w;
        db[i].x += gv[i].x; db[i].y += gv[i].y; db[i].z += gv[i].z; db[i].w += gv[i].w;
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 4));  // L1-resident
        gv[i].x *= gm.x; gv[i].y *= gm.y; gv[i].z *= gm.z; gv[i].w *= gm.w;  // dy * gamma
        c1 += gv[i].x + gv[i].y + gv[i].z + gv[i].w;
        c2 += gv[i].x * xv[i].x + gv[i].y * xv[i].y + gv[i].z * xv[i].z + gv[i].w * xv[i].w;
      }
    c1 = warp_sum(c1) / D;
    c2 = warp_sum(c2) / D;
#pragma unroll
    for (int i = 0; i < MAX_V; ++i)
      if (i < nv) {
        const int col = (i * 32 + lane) * 4;
        float4 o;
        o.x = rstd * (gv[i].x - c1 - xv[i].x * c2);
        o.y = rstd * (gv[i].y - c1 - xv[i].y * c2);
        o.z = rstd * (gv[i].z - c1 - xv[i].z * c2);
        o.w = rstd * (gv[i].w - c1 - xv[i].w * c2);
        o.x += av[i].x; o.y += av[i].y; o.z += av[i].z; o.w += av[i].w;
        if (dx) *reinterpret_cast<float4*>(dx + row * lddx + col) = o;
        if (dx_bf16)
          *reinterpret_cast<uint2*>(dx_bf16 + row * lddxb + col) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      }
  }
  if (!dgamma) return;
  float* rg = red + (size_t)warp * 2 * D;
#pragma unroll
  for (int i = 0; i < MAX_V; ++i)
    if (i < nv) {
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

}  // namespace ln
}  // namespace lv

extern "C" int lv_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps,
                                void* y_bf16, int64_t ldy, float* y_f32, int64_t ldyf, int64_t rows, int D,
                                void* stream) {
  using namespace lv;
  LV_REQUIRE(x && gamma && beta && (y_bf16 || y_f32), "lv_layernorm_fwd: null pointer");
  LV_REQUIRE(D > 0 && D % 128 == 0 && D <= 128 * ln::MAX_V, "lv_layernorm_fwd: D=%d must be a multiple of 128 and <= 1024", D);
  LV_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0 && ldyf % 4 == 0, "lv_layernorm_fwd: leading dimensions must be multiples of 4");
  if (rows <= 0) return 0;
  const unsigned grid = (unsigned)((rows + ln::WARPS - 1) / ln::WARPS);
  ln::ln_fwd_kernel<<<grid, ln::WARPS * 32, 0, (cudaStream_t)stream>>>(x, ldx, gamma, beta, eps, (__nv_bfloat16*)y_bf16, ldy,
                                                                      y_f32, ldyf, rows, D);
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
  long long want = (rows + ln::WARPS - 1) / ln::WARPS;
  long long cap = (long long)sm_count() * 4;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  const size_t smem = (size_t)ln::WARPS * 2 * D * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  static bool configured = false;
  if (!configured) {  // D = 1024 needs 64 KB of dynamic shared memory for the dgamma/dbeta reduction
    cudaFuncSetAttribute(ln::ln_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ln::ln_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    configured = true;
  }
  if (dy_is_bf16)
    ln::ln_bwd_kernel<true><<<grid, ln::WARPS * 32, smem, st>>>(dy, lddy, x, ldx, gamma, eps, add1, ld1, add2, ld2, dx, lddx,
                                                                (__nv_bfloat16*)dx_bf16, lddxb, dgamma, dbeta, rows, D);
  else
    ln::ln_bwd_kernel<false><<<grid, ln::WARPS * 32, smem, st>>>(dy, lddy, x, ldx, gamma, eps, add1, ld1, add2, ld2, dx, lddx,
                                                                 (__nv_bfloat16*)dx_bf16, lddxb, dgamma, dbeta, rows, D);
  return check_launch("lv_layernorm_bwd");
}
