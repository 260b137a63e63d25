// Clip pre-processing on the GPU: decoded frames (T x H x W x 3, uint8 or fp32, as the decoder hands them over) -> the
// normalised C x T x S x S fp32 clip the video encoder takes, for a whole batch in one launch.
//
// Replaces the per-sample CPU transform chain the reference's DataLoader workers run (main_pretrain.py:263-272 train,
// :274-281 val; lavila/data/video_transforms.py:15-32 `Permute`; torchvision 0.11.2 `RandomResizedCrop` / `Resize` +
// `CenterCrop` on float tensors = crop + `interpolate(mode="bilinear", align_corners=False)`, NO antialiasing in that
// version; `NormalizeVideo`): 64 clips x (permute copy + crop + resize + normalise) per step on 10 workers per GPU.
//
// Geometry (one descriptor per clip, so a batch may mix source resolutions):
//   source box (box_i, box_j, box_h, box_w) inside the H x W frame  --bilinear-->  virtual resized image RH x RW
//   output pixel (y, x) = resized pixel (y + off_y, x + off_x), y < OH, x < OW
// train: box = the sampled crop, RH = RW = OH = OW = S, off = 0.   val: box = the frame, (RH, RW) = short side scaled to S,
// off = the centre-crop offset.  Taps are clamped to the BOX (the reference resizes the cropped tensor), not to the frame.
//
// Byte work, HBM-bound: per clip 3*T*OH*OW*4 B written (9.6 MB at 16 x 224^2) + the box read once from DRAM (taps hit L1/L2).
// One thread per 4 consecutive output pixels of a row, the 3 channels of the interleaved source pixel together: one 16-byte
// store per channel plane, the row taps shared by the 4 pixels; the clip's descriptor is staged in shared memory once per CTA.
#include "../../include/lavila_b200.h"
#include "host_common.h"

namespace lv {
namespace inp {

struct ClipDesc {          // 12 x int64, the layout of the `desc` table (include/lavila_b200.h)
  long long src, H, W, box_i, box_j, box_h, box_w, RH, RW, off_y, off_x, frame_stride;
};

template <typename T>
__device__ __forceinline__ void load3(const T* p, float& r, float& g, float& b) {
  r = (float)p[0];
  g = (float)p[1];
  b = (float)p[2];
}

// ATen's source index for bilinear, align_corners = False (aten/src/ATen/native/UpSample.h area_pixel_compute_source_index)
__device__ __forceinline__ void lin_taps(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
}

// antialiased (triangle filter stretched by the scale when down-sampling): ATen's _compute_weights_span / _compute_weights
__device__ __forceinline__ void aa_span(int dst, float scale, int in_size, int& xmin, int& xsize, float& c0, float& invscale) {
  const float support = scale >= 1.f ? scale : 1.f;
  const float center = scale * ((float)dst + 0.5f);
  xmin = max((int)(center - support + 0.5f), 0);
  xsize = min((int)(center + support + 0.5f), in_size) - xmin;
  if (xsize < 0) xsize = 0;
  c0 = (float)xmin - center + 0.5f;
  invscale = scale >= 1.f ? 1.f / scale : 1.f;
}
__device__ __forceinline__ float tri(float x) {
  x = fabsf(x);
  return x < 1.f ? 1.f - x : 0.f;
}

// a / b with b's reciprocal rb = RN(1 / b) precomputed: quotient, exact remainder by FMA, one correction step.  Correctly rounded for
// the value range here (checked against true division over all 0..255 inputs of both statistics sets and 1.2 M random values), so the
// result equals `clip.sub_(mean).div_(std)`; 3 instructions instead of the ~10 + slow-path call of an IEEE division.
__device__ __forceinline__ float div_by(float a, float b, float rb) {
  const float q = a * rb;
  return fmaf(fmaf(-b, q, a), rb, q);
}

constexpr int PX = 4;   // output pixels per thread along x (one 16-byte store per channel plane)

template <typename T, bool AA>
__global__ void __launch_bounds__(256)
clip_transform_kernel(const ClipDesc* __restrict__ desc, float* __restrict__ out, int frames, int OH, int OW, float m0, float m1,
                      float m2, float s0, float s1, float s2, float r0, float r1, float r2) {
  __shared__ long long sd[12];
  const int b = blockIdx.z, t = blockIdx.y;
  if (threadIdx.x < 12) sd[threadIdx.x] = reinterpret_cast<const long long*>(desc + b)[threadIdx.x];
  __syncthreads();
  const int groups = (OW + PX - 1) / PX;
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= OH * groups) return;
  const int y = gi / groups, xg = (gi - y * groups) * PX;
  const int W = (int)sd[2], bh = (int)sd[5], bw = (int)sd[6];
  const T* __restrict__ frame = reinterpret_cast<const T*>(sd[0]) + (long long)t * sd[11] + (sd[3] * W + sd[4]) * 3;
  const float sy = (float)bh / (float)(int)sd[7], sx = (float)bw / (float)(int)sd[8];
  const int ry = y + (int)sd[9], rx0 = xg + (int)sd[10];
  float res[3][PX];
  if (!AA) {
    int y0, y1;
    float ly;
    lin_taps(ry, sy, bh, y0, y1, ly);
    const float hy = 1.f - ly;
    const T* __restrict__ row0 = frame + (long long)y0 * W * 3;
    const T* __restrict__ row1 = frame + (long long)y1 * W * 3;
#pragma unroll
    for (int k = 0; k < PX; ++k) {
      int x0, x1;
      float lx;
      lin_taps(rx0 + k, sx, bw, x0, x1, lx);
      x0 *= 3;
      x1 *= 3;
      const float hx = 1.f - lx;
      float a0, a1, a2, b0, b1, b2, c0, c1, c2, e0, e1, e2;
      load3(row0 + x0, a0, a1, a2);
      load3(row0 + x1, b0, b1, b2);
      load3(row1 + x0, c0, c1, c2);
      load3(row1 + x1, e0, e1, e2);
      res[0][k] = hy * (hx * a0 + lx * b0) + ly * (hx * c0 + lx * e0);
      res[1][k] = hy * (hx * a1 + lx * b1) + ly * (hx * c1 + lx * e1);
      res[2][k] = hy * (hx * a2 + lx * b2) + ly * (hx * c2 + lx * e2);
    }
  } else {
    int ymin, ysz;
    float cy, iy;
    aa_span(ry, sy, bh, ymin, ysz, cy, iy);
    float ty = 0.f;
    for (int i = 0; i < ysz; ++i) ty += tri(((float)i + cy) * iy);
    const float ny = ty != 0.f ? 1.f / ty : 0.f;
#pragma unroll 1
    for (int k = 0; k < PX; ++k) {
      int xmin, xsz;
      float cx, ix;
      aa_span(rx0 + k, sx, bw, xmin, xsz, cx, ix);
      float tx = 0.f;
      for (int j = 0; j < xsz; ++j) tx += tri(((float)j + cx) * ix);
      const float nx = tx != 0.f ? 1.f / tx : 0.f;
      float r = 0.f, g = 0.f, bl = 0.f;
      for (int i = 0; i < ysz; ++i) {
        const float wy = tri(((float)i + cy) * iy) * ny;
        const T* __restrict__ row = frame + ((long long)(ymin + i) * W + xmin) * 3;
        float rr = 0.f, rg = 0.f, rb = 0.f;
        for (int j = 0; j < xsz; ++j) {
          const float wx = tri(((float)j + cx) * ix) * nx;
          float p0, p1, p2;
          load3(row + j * 3, p0, p1, p2);
          rr += wx * p0;
          rg += wx * p1;
          rb += wx * p2;
        }
        r += wy * rr;
        g += wy * rg;
        bl += wy * rb;
      }
      res[0][k] = r;
      res[1][k] = g;
      res[2][k] = bl;
    }
  }
  const long long plane = (long long)OH * OW;
  float* o = out + (((long long)b * 3) * frames + t) * plane + (long long)y * OW + xg;
  const float mean[3] = {m0, m1, m2}, sd3[3] = {s0, s1, s2}, rc[3] = {r0, r1, r2};
  const bool vec = (OW % PX) == 0 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float* oc = o + (long long)c * frames * plane;
    float v[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) v[k] = div_by(res[c][k] - mean[c], sd3[c], rc[c]);
    if (vec) {
      *reinterpret_cast<float4*>(oc) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int k = 0; k < PX; ++k)
        if (xg + k < OW) oc[k] = v[k];
    }
  }
}

}  // namespace inp
}  // namespace lv

extern "C" int lv_clip_transform(const int64_t* desc, int clips, int frames, int src_dtype, int antialias, const float* mean,
                                 const float* std, float* out, int OH, int OW, void* stream) {
  using namespace lv;
  using namespace lv::inp;
  LV_REQUIRE(desc && out && mean && std, "lv_clip_transform: null argument");
  LV_REQUIRE(clips >= 0 && clips <= 65535 && frames >= 1 && frames <= 65535 && OH >= 1 && OW >= 1,
             "lv_clip_transform: clips=%d frames=%d OH=%d OW=%d out of range", clips, frames, OH, OW);
  LV_REQUIRE(src_dtype == 0 || src_dtype == 1, "lv_clip_transform: src_dtype %d (0 = uint8, 1 = fp32)", src_dtype);
  LV_REQUIRE(std[0] != 0.f && std[1] != 0.f && std[2] != 0.f, "lv_clip_transform: zero std");
  if (clips == 0) return 0;
  const dim3 grid((unsigned)((OH * ((OW + inp::PX - 1) / inp::PX) + 255) / 256), (unsigned)frames, (unsigned)clips);
  const float r0 = (float)(1.0 / (double)std[0]), r1 = (float)(1.0 / (double)std[1]), r2 = (float)(1.0 / (double)std[2]);
  const ClipDesc* d = reinterpret_cast<const ClipDesc*>(desc);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define LV_LAUNCH(T, AA) \
  clip_transform_kernel<T, AA><<<grid, 256, 0, st>>>(d, out, frames, OH, OW, mean[0], mean[1], mean[2], std[0], std[1], std[2], r0, r1, r2)
  if (src_dtype == 0) {
    if (antialias) LV_LAUNCH(uint8_t, true); else LV_LAUNCH(uint8_t, false);
  } else {
    if (antialias) LV_LAUNCH(float, true); else LV_LAUNCH(float, false);
  }
#undef LV_LAUNCH
  return check_launch("lv_clip_transform");
}
