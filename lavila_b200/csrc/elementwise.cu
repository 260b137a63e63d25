// HBM-bound glue kernels of the dual-encoder step: casts, bias-gradient column sums, patch im2col, token/position
// embedding assembly and its gradients, row gather/scatter, argmax (EOT), L2 normalisation.
// All accesses are 128-bit where alignment allows; grids are sized in multiples of the SM count for the reductions.
#include "../../include/lavila_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace lv {
namespace ew {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------- cast fp32 -> bf16
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(in + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(in + i + 4));
    *reinterpret_cast<uint4*>(out + i) =
        make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
  } else {
    for (long long j = i; j < n; ++j) out[j] = __float2bfloat16_rn(in[j]);
  }
}

// ------------------------------------------------------------------------------------------- column sums (bias grads)
// out[n] += sum_m in[m, n];  in bf16 [M, N] (ld), N % 8 == 0.  CTA = 32 column-groups (8 cols) x 8 row lanes.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ in, long long ld, long long M, int N, float* __restrict__ out,
                   int row_chunks) {
  __shared__ float red[8][32][8 + 1];
  const int cg = threadIdx.x & 31;
  const int rl = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cg) * 8;
  const long long rows_per = (M + row_chunks - 1) / row_chunks;
  const long long r0 = (long long)blockIdx.y * rows_per;
  const long long r1 = (r0 + rows_per < M) ? r0 + rows_per : M;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < N) {
    for (long long r = r0 + rl; r < r1; r += 8) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(in + r * ld + col));
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
      acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][cg][j] = acc[j];
  __syncthreads();
  // 256 threads -> 256 columns of this CTA
  const int c_l = threadIdx.x;
  const int g2 = c_l >> 3, j2 = c_l & 7;
  const int ocol = blockIdx.x * 256 + c_l;
  if (ocol < N) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][g2][j2];
    red_add_f32(out + ocol, t);
  }
}

// ------------------------------------------------------------------------------------------- patch im2col
// frames fp32 [B, C, T, H, W]  ->  patches bf16 [B*T*gh*gw, ldp], column = (c*p + py)*p + px  (conv-weight order).
// Folds the reference's permute(0,2,1,3,4).contiguous() (timesformer.py:387) and the Conv2d input gather (:82-83).
// One thread per (b, t, c, y, gx): reads p contiguous floats of one image row, writes p bf16.
template <int P>
__global__ void im2col_kernel(const float* __restrict__ frames, __nv_bfloat16* __restrict__ patches, int B, int C,
                              int T, int H, int W, long long ldp) {
  const int gw = W / P, gh = H / P;
  const long long total = (long long)B * T * C * H * gw;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int gx = (int)(idx % gw);
  long long t1 = idx / gw;
  const int y = (int)(t1 % H);
  t1 /= H;
  const int c = (int)(t1 % C);
  t1 /= C;
  const int t = (int)(t1 % T);
  const int b = (int)(t1 / T);
  const int gy = y / P, py = y - gy * P;
  const float* src = frames + ((((long long)b * C + c) * T + t) * H + y) * W + gx * P;
  const long long row = (((long long)b * T + t) * gh + gy) * gw + gx;
  __nv_bfloat16* dst = patches + row * ldp + (c * P + py) * P;
  if (P % 8 == 0) {
#pragma unroll
    for (int j = 0; j < P; j += 8) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src + j));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(src + j + 4));
      *reinterpret_cast<uint4*>(dst + j) =
          make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b4.x, b4.y), pack_bf16x2(b4.z, b4.w));
    }
  } else {
#pragma unroll
    for (int j = 0; j < P; j += 2) {
      const float2 a = __ldg(reinterpret_cast<const float2*>(src + j));
      *reinterpret_cast<uint32_t*>(dst + j) = pack_bf16x2(a.x, a.y);
    }
  }
}

// ------------------------------------------------------------------------------------------- vision embedding assembly
// x0[b, 0]         = cls + pos[0]
// x0[b, 1+f*n+i]   = patch[(b*T+f)*n+i] + pos[1+i] + temporal[f]          (timesformer.py:353-364)
// One warp per row, fp32, D % 128 == 0.
__global__ void __launch_bounds__(256)
embed_assemble_kernel(const float* __restrict__ patch, const float* __restrict__ cls, const float* __restrict__ pos,
                      const float* __restrict__ temporal, float* __restrict__ x0, int B, int T, int n, int D) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int N = 1 + T * n;
  if (row >= (long long)B * N) return;
  const int b = (int)(row / N), tok = (int)(row - (long long)b * N);
  float* dst = x0 + row * D;
  if (tok == 0) {
    for (int c = lane * 4; c < D; c += 128) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(cls + c));
      const float4 p = __ldg(reinterpret_cast<const float4*>(pos + c));
      *reinterpret_cast<float4*>(dst + c) = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    }
  } else {
    const int f = (tok - 1) / n, i = (tok - 1) - f * n;
    const float* src = patch + (((long long)b * T + f) * n + i) * D;
    const float* pp = pos + (long long)(1 + i) * D;
    const float* tp = temporal + (long long)f * D;
    for (int c = lane * 4; c < D; c += 128) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src + c));
      const float4 p = __ldg(reinterpret_cast<const float4*>(pp + c));
      const float4 t = __ldg(reinterpret_cast<const float4*>(tp + c));
      *reinterpret_cast<float4*>(dst + c) = make_float4(a.x + p.x + t.x, a.y + p.y + t.y, a.z + p.z + t.z, a.w + p.w + t.w);
    }
  }
}

// Gradients of the assembly.  dx0 fp32 [B, N, D].
//  (a) dpos[1+i] += sum_{b,f} dx0[b,1+f*n+i];  dpos[0] += sum_b dx0[b,0];  dcls += sum_b dx0[b,0]
//      grid = (n+1, D/256); thread = column; coalesced across columns.
__global__ void __launch_bounds__(256)
embed_dpos_kernel(const float* __restrict__ dx0, float* __restrict__ dpos, float* __restrict__ dcls, int B, int T, int n,
                  int D) {
  const int i = blockIdx.x;  // 0 = CLS slot, 1.. = spatial position
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= D) return;
  const long long N = 1 + (long long)T * n;
  float acc = 0.f;
  if (i == 0) {
    for (int b = 0; b < B; ++b) acc += __ldg(dx0 + (long long)b * N * D + c);
    dcls[c] += acc;
    dpos[c] += acc;
  } else {
    for (int b = 0; b < B; ++b)
      for (int f = 0; f < T; ++f) acc += __ldg(dx0 + ((long long)b * N + 1 + (long long)f * n + (i - 1)) * D + c);
    dpos[(long long)i * D + c] += acc;
  }
}
//  (b) dtemporal[f] += sum_{b,i} dx0[b,1+f*n+i]  and the compact bf16 copy dpatch[(b*T+f)*n+i] = dx0[b,1+f*n+i]
//      (A operand of the patch-embedding weight gradient).  grid = (B*T, D/256).
__global__ void __launch_bounds__(256)
embed_dtemporal_kernel(const float* __restrict__ dx0, float* __restrict__ dtemporal, __nv_bfloat16* __restrict__ dpatch,
                       int B, int T, int n, int D) {
  const int bf = blockIdx.x;
  const int b = bf / T, f = bf - b * T;
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= D) return;
  const long long N = 1 + (long long)T * n;
  const float* src = dx0 + ((long long)b * N + 1 + (long long)f * n) * D + c;
  __nv_bfloat16* dst = dpatch + ((long long)bf * n) * D + c;
  float acc = 0.f;
  for (int i = 0; i < n; ++i) {
    const float v = __ldg(src + (long long)i * D);
    acc += v;
    dst[(long long)i * D] = __float2bfloat16_rn(v);
  }
  red_add_f32(dtemporal + (long long)f * D + c, acc);
}

// ------------------------------------------------------------------------------------------- text embedding
// x[b,l] = tok[text[b,l]] + pos[l]   (models.py:151-152); ids are int64, compared/gathered exactly.
__global__ void __launch_bounds__(256)
text_embed_kernel(const long long* __restrict__ text, const float* __restrict__ tok, const float* __restrict__ pos,
                  float* __restrict__ x, long long rows, int L, int W, int vocab) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  long long id = text[row];
  if (id < 0 || id >= vocab) id = 0;  // the reference would raise; never hit with valid tokenizer output
  const int l = (int)(row % L);
  for (int c = lane * 4; c < W; c += 128) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(tok + id * W + c));
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos + (long long)l * W + c));
    *reinterpret_cast<float4*>(x + row * W + c) = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
  }
}
__global__ void __launch_bounds__(256)
text_embed_bwd_kernel(const long long* __restrict__ text, const float* __restrict__ dx, float* __restrict__ dtok,
                      float* __restrict__ dpos, long long rows, int L, int W, int vocab) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  long long id = text[row];
  if (id < 0 || id >= vocab) id = 0;
  const int l = (int)(row % L);
  for (int c = lane; c < W; c += 32) {
    const float v = __ldg(dx + row * W + c);
    red_add_f32(dtok + id * W + c, v);
    red_add_f32(dpos + (long long)l * W + c, v);
  }
}

// argmax over the last dim of int64 [B, L] -> int32 [B]; first occurrence of the maximum (torch.argmax semantics,
// models.py:160: the EOT token has the highest id).
__global__ void argmax_i64_kernel(const long long* __restrict__ text, int* __restrict__ out, int B, int L) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  long long best = text[(long long)b * L];
  int bi = 0;
  for (int l = 1; l < L; ++l) {
    const long long v = text[(long long)b * L + l];
    if (v > best) { best = v; bi = l; }
  }
  out[b] = bi;
}

// dst[r, :] = src[(r*rows_per + idx[r]), :]   /   dst[(r*rows_per + idx[r]), :] = src[r, :]   (fp32, W % 4 == 0)
__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx, float* __restrict__ dst,
                                   int R, int rows_per, int W, int scatter) {
  const int r = blockIdx.x;
  const long long big = ((long long)r * rows_per + idx[r]) * W;
  const long long small = (long long)r * W;
  for (int c = threadIdx.x * 4; c < W; c += blockDim.x * 4) {
    if (scatter) *reinterpret_cast<float4*>(dst + big + c) = __ldg(reinterpret_cast<const float4*>(src + small + c));
    else *reinterpret_cast<float4*>(dst + small + c) = __ldg(reinterpret_cast<const float4*>(src + big + c));
  }
}

// ------------------------------------------------------------------------------------------- L2 normalise (F.normalize)
// y = x / max(||x||, 1e-12)  (models.py:169-170);  backward: dx = (dy - y * <dy, y>) / max(||x||, eps)
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ norm, int R,
                                  int E) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  float s = 0.f;
  for (int c = lane; c < E; c += 32) { const float v = x[(long long)r * E + c]; s += v * v; }
  const float nrm = fmaxf(sqrtf(warp_sum(s)), 1e-12f);
  for (int c = lane; c < E; c += 32) y[(long long)r * E + c] = x[(long long)r * E + c] / nrm;
  if (lane == 0) norm[r] = nrm;
}
__global__ void l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                  const float* __restrict__ norm, float* __restrict__ dx, int R, int E) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  float s = 0.f;
  for (int c = lane; c < E; c += 32) s += dy[(long long)r * E + c] * y[(long long)r * E + c];
  s = warp_sum(s);
  const float inv = 1.0f / norm[r];
  for (int c = lane; c < E; c += 32)
    dx[(long long)r * E + c] = (dy[(long long)r * E + c] - y[(long long)r * E + c] * s) * inv;
}

// dst[r*stride + c] += src[r*W + c]  (dst bf16 or fp32; R rows picked with a row stride, e.g. the CLS row of every clip)
template <bool BF16>
__global__ void add_rows_kernel(void* __restrict__ dst, long long stride, const float* __restrict__ src, int R, int W) {
  const int r = blockIdx.x;
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    const float v = src[(long long)r * W + c];
    if (BF16) {
      __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dst) + (long long)r * stride + c;
      *d = __float2bfloat16_rn(__bfloat162float(*d) + v);
    } else {
      float* d = reinterpret_cast<float*>(dst) + (long long)r * stride + c;
      *d += v;
    }
  }
}

}  // namespace ew
}  // namespace lv

using namespace lv;

extern "C" int lv_add_rows(void* dst, int dst_is_bf16, int64_t stride, const float* src, int R, int W, void* stream) {
  LV_REQUIRE(dst && src && R > 0 && W > 0, "lv_add_rows: bad arguments");
  if (dst_is_bf16) ew::add_rows_kernel<true><<<R, 256, 0, (cudaStream_t)stream>>>(dst, stride, src, R, W);
  else ew::add_rows_kernel<false><<<R, 256, 0, (cudaStream_t)stream>>>(dst, stride, src, R, W);
  return check_launch("lv_add_rows");
}

extern "C" int lv_cast_f32_bf16(const float* in, void* out, int64_t n, void* stream) {
  LV_REQUIRE(in && out && n >= 0, "lv_cast_f32_bf16: bad arguments");
  LV_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "lv_cast_f32_bf16: pointers must be 16-byte aligned");
  if (n == 0) return 0;
  const long long threads = (n + 7) / 8;
  ew::cast_f32_bf16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out, n);
  return check_launch("lv_cast_f32_bf16");
}

extern "C" int lv_colsum_bf16(const void* in, int64_t ld, int64_t M, int N, float* out, void* stream) {
  LV_REQUIRE(in && out && M > 0 && N > 0 && N % 8 == 0 && ld % 8 == 0, "lv_colsum_bf16: bad arguments (N, ld must be multiples of 8)");
  const int col_blocks = (N + 255) / 256;
  int row_chunks = (4 * sm_count() + col_blocks - 1) / col_blocks;
  if (row_chunks > (M + 63) / 64) row_chunks = (int)((M + 63) / 64);
  if (row_chunks < 1) row_chunks = 1;
  ew::colsum_bf16_kernel<<<dim3(col_blocks, row_chunks), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, ld, M, N, out, row_chunks);
  return check_launch("lv_colsum_bf16");
}

extern "C" int lv_patch_im2col(const float* frames, void* patches, int B, int C, int T, int H, int W, int P, int64_t ldp,
                               void* stream) {
  LV_REQUIRE(frames && patches && B > 0 && C > 0 && T > 0 && H % P == 0 && W % P == 0, "lv_patch_im2col: bad shape");
  LV_REQUIRE(ldp >= (int64_t)C * P * P && ldp % 8 == 0, "lv_patch_im2col: ldp must be >= C*P*P and a multiple of 8");
  const long long total = (long long)B * T * C * H * (W / P);
  const unsigned grid = (unsigned)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (P == 16) ew::im2col_kernel<16><<<grid, 256, 0, st>>>(frames, (__nv_bfloat16*)patches, B, C, T, H, W, ldp);
  else if (P == 14) ew::im2col_kernel<14><<<grid, 256, 0, st>>>(frames, (__nv_bfloat16*)patches, B, C, T, H, W, ldp);
  else if (P == 32) ew::im2col_kernel<32><<<grid, 256, 0, st>>>(frames, (__nv_bfloat16*)patches, B, C, T, H, W, ldp);
  else return set_error(-1, "lv_patch_im2col: unsupported patch size %d (14, 16, 32)", P);
  return check_launch("lv_patch_im2col");
}

extern "C" int lv_embed_assemble(const float* patch, const float* cls, const float* pos, const float* temporal, float* x0,
                                 int B, int T, int n, int D, void* stream) {
  LV_REQUIRE(patch && cls && pos && temporal && x0 && D % 128 == 0, "lv_embed_assemble: bad arguments");
  const long long rows = (long long)B * (1 + (long long)T * n);
  ew::embed_assemble_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(patch, cls, pos, temporal, x0, B, T, n, D);
  return check_launch("lv_embed_assemble");
}

extern "C" int lv_embed_assemble_bwd(const float* dx0, float* dpos, float* dcls, float* dtemporal, void* dpatch_bf16,
                                     int B, int T, int n, int D, void* stream) {
  LV_REQUIRE(dx0 && dpos && dcls && dtemporal && dpatch_bf16 && D % 128 == 0, "lv_embed_assemble_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ew::embed_dpos_kernel<<<dim3(n + 1, (D + 255) / 256), 256, 0, st>>>(dx0, dpos, dcls, B, T, n, D);
  int rc = check_launch("lv_embed_assemble_bwd(dpos)");
  if (rc) return rc;
  ew::embed_dtemporal_kernel<<<dim3(B * T, (D + 255) / 256), 256, 0, st>>>(dx0, dtemporal, (__nv_bfloat16*)dpatch_bf16, B, T, n, D);
  return check_launch("lv_embed_assemble_bwd(dtemporal)");
}

extern "C" int lv_text_embed(const int64_t* text, const float* tok, const float* pos, float* x, int64_t rows, int L, int W,
                             int vocab, void* stream) {
  LV_REQUIRE(text && tok && pos && x && W % 4 == 0, "lv_text_embed: bad arguments (W must be a multiple of 4)");
  ew::text_embed_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>((const long long*)text, tok, pos, x, rows, L, W, vocab);
  return check_launch("lv_text_embed");
}

extern "C" int lv_text_embed_bwd(const int64_t* text, const float* dx, float* dtok, float* dpos, int64_t rows, int L, int W,
                                 int vocab, void* stream) {
  LV_REQUIRE(text && dx && dtok && dpos, "lv_text_embed_bwd: bad arguments");
  ew::text_embed_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>((const long long*)text, dx, dtok, dpos, rows, L, W, vocab);
  return check_launch("lv_text_embed_bwd");
}

extern "C" int lv_argmax_i64(const int64_t* text, int32_t* out, int B, int L, void* stream) {
  LV_REQUIRE(text && out && B > 0 && L > 0, "lv_argmax_i64: bad arguments");
  ew::argmax_i64_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const long long*)text, out, B, L);
  return check_launch("lv_argmax_i64");
}

extern "C" int lv_gather_rows_f32(const float* src, const int32_t* idx, float* dst, int R, int rows_per, int W, int scatter,
                                  void* stream) {
  LV_REQUIRE(src && idx && dst && R > 0 && W % 4 == 0, "lv_gather_rows_f32: bad arguments");
  ew::gather_rows_kernel<<<R, 128, 0, (cudaStream_t)stream>>>(src, idx, dst, R, rows_per, W, scatter);
  return check_launch("lv_gather_rows_f32");
}

extern "C" int lv_l2norm_fwd(const float* x, float* y, float* norm, int R, int E, void* stream) {
  LV_REQUIRE(x && y && norm && R > 0 && E > 0, "lv_l2norm_fwd: bad arguments");
  ew::l2norm_fwd_kernel<<<(R + 3) / 4, 128, 0, (cudaStream_t)stream>>>(x, y, norm, R, E);
  return check_launch("lv_l2norm_fwd");
}

extern "C" int lv_l2norm_bwd(const float* dy, const float* y, const float* norm, float* dx, int R, int E, void* stream) {
  LV_REQUIRE(dy && y && norm && dx && R > 0 && E > 0, "lv_l2norm_bwd: bad arguments");
  ew::l2norm_bwd_kernel<<<(R + 3) / 4, 128, 0, (cudaStream_t)stream>>>(dy, y, norm, dx, R, E);
  return check_launch("lv_l2norm_bwd");
}
