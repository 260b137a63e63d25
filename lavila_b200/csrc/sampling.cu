// Temperature + nucleus (top-p) filtering of next-token logits in one kernel, one CTA per sequence.
//
// Replaces TemperatureLogitsWarper + TopPLogitsWarper of the narrator's sampling loop (lavila/models/narrator.py:131,368-389;
// transformers `generation/logits_process.py`): scores /= T; sort ascending; softmax; cumsum; remove every token whose
// cumulative probability is <= 1 - top_p, never the largest one; removed scores become -inf.  The library path sorts
// [sequences x 50 257] key-value pairs per decoding step (4.7 + 1.2 ms of the 23 ms step at 320 sequences); the token SET that
// survives only depends on a threshold, which a radix select over the row finds without sorting:
//   * the row (V fp32 logits, 201 KB at V = 50 257) lives in shared memory, scaled by 1/T;
//   * keys = order-preserving bit patterns of the scaled logits, mass(i) = exp(v_i - max);
//   * 4 passes of 8 bits: per pass a 256-bin histogram of MASS over the keys matching the prefix found so far, scanned in
//     ascending order until the running mass exceeds (1 - top_p) * Z -- this yields k* = the smallest key whose ascending
//     cumulative mass is > (1 - top_p) * Z and the mass strictly below it;
//   * keys below k* are removed; among keys EQUAL to k* (ties) the first floor((T - below) / mass(k*)) in index order are
//     removed (the reference removes the same NUMBER of them, which ones is up to its unstable sort).
// Index op in spirit: the surviving set equals the library's whenever no cumulative sum lands within fp32 round-off of the
// threshold (tests/test_gpu_narrator.py compares the masks on real decoder logits).
#include "../../include/lavila_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace lv {
namespace sampling {

constexpr int THREADS = 1024;

__device__ __forceinline__ uint32_t key_of(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(~0u, v, o);
    v = is_max ? fmaxf(v, w) : v + w;
  }
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(~0u, r, o);
    r = is_max ? fmaxf(r, w) : r + w;
  }
  return r;
}

__global__ void __launch_bounds__(THREADS, 1)
top_p_filter_kernel(float* __restrict__ logits, long long ld, int V, float inv_temperature, float temperature, float top_p) {
  extern __shared__ float row[];          // [V] scaled logits
  __shared__ float red[32];
  __shared__ float hist[256];
  __shared__ unsigned int s_prefix, s_count;
  __shared__ float s_below;
  float* g = logits + (long long)blockIdx.x * ld;
  const int tid = threadIdx.x;
  float m = -INFINITY;
  for (int i = tid; i < V; i += THREADS) {
    // `scores / temperature` with a Python-scalar divisor: ATen's CUDA kernel multiplies by the fp32 reciprocal
    // (aten/src/ATen/native/cuda/BinaryDivTrueKernel.cu, the CPU-scalar fast path) -- done the same way so the kept logits are
    // bit-identical to the library's
    const float v = (temperature != 1.0f) ? g[i] * inv_temperature : g[i];
    row[i] = v;
    m = fmaxf(m, v);
  }
  m = block_reduce(m, red, true);
  float z = 0.f;
  for (int i = tid; i < V; i += THREADS) z += expf(row[i] - m);
  z = block_reduce(z, red, false);
  const float target = (1.0f - top_p) * z;       // ascending cumulative mass <= target  ->  removed
  if (tid == 0) { s_prefix = 0u; s_below = 0.f; }
  __syncthreads();
#pragma unroll 1
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    const uint32_t pmask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    if (tid < 256) hist[tid] = 0.f;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    for (int i = tid; i < V; i += THREADS) {
      const float v = row[i];
      const uint32_t k = key_of(v);
      if ((k & pmask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], expf(v - m));
    }
    __syncthreads();
    if (tid == 0) {
      float acc = s_below;
      int d = 0;
      for (; d < 255; ++d) {
        if (acc + hist[d] > target) break;
        acc += hist[d];
      }
      s_below = acc;
      s_prefix = prefix | ((uint32_t)d << shift);
    }
    __syncthreads();
  }
  const uint32_t kstar = s_prefix;
  const float below = s_below;
  // ties at k*: how many of them still fall at or under the threshold
  if (tid == 0) s_count = 0u;
  __syncthreads();
  unsigned int mine = 0;
  float e_star = 0.f;
  for (int i = tid; i < V; i += THREADS)
    if (key_of(row[i]) == kstar) { ++mine; e_star = expf(row[i] - m); }
  if (mine) atomicAdd(&s_count, mine);
  __syncthreads();
  const unsigned int ties = s_count;
  e_star = block_reduce(e_star, red, true);
  unsigned int n_rm = 0;
  if (ties > 1 && e_star > 0.f) {
    const float room = (target - below) / e_star;
    n_rm = room <= 0.f ? 0u : (unsigned int)fminf(floorf(room), (float)(ties - 1));
  }
  if (n_rm == 0) {
    for (int i = tid; i < V; i += THREADS) {
      const float v = row[i];
      g[i] = (key_of(v) < kstar) ? -INFINITY : v;
    }
  } else {
    // rare: remove the first n_rm tied tokens in index order (serial scan by one thread keeps it deterministic)
    for (int i = tid; i < V; i += THREADS) {
      const float v = row[i];
      const uint32_t k = key_of(v);
      if (k != kstar) g[i] = (k < kstar) ? -INFINITY : v;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int seen = 0;
      for (int i = 0; i < V; ++i)
        if (key_of(row[i]) == kstar) { g[i] = (seen < n_rm) ? -INFINITY : row[i]; ++seen; }
    }
  }
}

}  // namespace sampling
}  // namespace lv

using namespace lv;

extern "C" int lv_top_p_filter(float* logits, int64_t ld, int rows, int V, float temperature, float top_p, void* stream) {
  LV_REQUIRE(logits && rows > 0 && V > 0 && ld >= V, "lv_top_p_filter: bad arguments");
  LV_REQUIRE(temperature > 0.f && top_p > 0.f && top_p < 1.f, "lv_top_p_filter: needs temperature > 0 and 0 < top_p < 1");
  const size_t smem = (size_t)V * sizeof(float);
  LV_REQUIRE(smem <= 220 * 1024, "lv_top_p_filter: a row of %d logits does not fit shared memory", V);
  static cudaError_t attr_err = cudaFuncSetAttribute(sampling::top_p_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (attr_err != cudaSuccess) return set_error((int)attr_err, "lv_top_p_filter: cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
  sampling::top_p_filter_kernel<<<rows, sampling::THREADS, smem, (cudaStream_t)stream>>>(logits, ld, V, 1.0f / temperature, temperature, top_p);
  return check_launch("lv_top_p_filter");
}
