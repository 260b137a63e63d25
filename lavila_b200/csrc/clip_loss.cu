// CLIPLoss forward/backward on the (gathered) global batch -- lavila/models/loss.py:76-79,107-116.
//
//   logits[i,j] = s * <I_i, T_j>,  loss = (CE(logits, arange) + CE(logits^T, arange)) / 2,  acc = 100 * mean(argmax_j == i)
//
// Ng <= 8 * per-GPU batch (512 at 8 GPUs), E = 256: 134 MFLOP -- latency-bound, so ONE kernel does logits, both
// log-sum-exps, the label pick, the arg-max and the mean (last-CTA-done reduction), and one kernel produces the
// gradients of the LOCAL rows only.  Every rank evaluates the same global loss, so the reference's backward
// all_reduce(SUM) of identical per-rank gradients (distributed_utils.py:64-67) is just a factor W: it is applied as
// `grad_scale` and the collective disappears from the backward pass.
// fp32 throughout (the reference's autocast runs the matmul in bf16 and CE in fp32; fp32 is >= that precision).
#include "../../include/lavila_b200.h"
#include "host_common.h"
#include "ptx.cuh"

namespace lv {
namespace loss {

constexpr int THREADS = 256;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dot(a_row (smem), B[j,:]) for all j handled by this thread
__device__ __forceinline__ float dot_row(const float* __restrict__ a_s, const float* __restrict__ b, int E) {
  float acc = 0.f;
  for (int c = 0; c < E; c += 4) {
    const float4 x = *reinterpret_cast<const float4*>(a_s + c);
    const float4 y = __ldg(reinterpret_cast<const float4*>(b + c));
    acc += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
  }
  return acc;
}

// block-wide (max, argmax-first, sum-exp) over values v[j], j = tid + k*THREADS
struct RowStat { float lse; int argmax; float label_logit; };

__device__ RowStat row_stats(const float* vals, int Ng, int label, float* sred, int* ired) {
  // vals in shared memory [Ng]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float m = -INFINITY;
  int am = 0x7fffffff;
  for (int j = tid; j < Ng; j += THREADS) {
    const float v = vals[j];
    if (v > m) { m = v; am = j; }
  }
  // warp arg-max with first-index tie-break
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > m || (om == m && oa < am)) { m = om; am = oa; }
  }
  if (lane == 0) { sred[warp] = m; ired[warp] = am; }
  __syncthreads();
  float bm = sred[0];
  int ba = ired[0];
  for (int w = 1; w < THREADS / 32; ++w) {
    const float om = sred[w];
    const int oa = ired[w];
    if (om > bm || (om == bm && oa < ba)) { bm = om; ba = oa; }
  }
  __syncthreads();
  float s = 0.f;
  for (int j = tid; j < Ng; j += THREADS) s += __expf(vals[j] - bm);
  s = warp_sum(s);
  if (lane == 0) sred[warp] = s;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < THREADS / 32; ++w) tot += sred[w];
  __syncthreads();
  RowStat r;
  r.lse = bm + logf(tot);
  r.argmax = ba;
  r.label_logit = vals[label];
  return r;
}

// grid = Ng CTAs; CTA i handles image row i and text row i.
// out: lse_img[Ng], lse_txt[Ng], result[3] = {loss, acc(%), scratch}, counter for the last-CTA reduction.
__global__ void __launch_bounds__(THREADS)
clip_loss_fwd_kernel(const float* __restrict__ img, const float* __restrict__ txt, const float* __restrict__ scale_ptr,
                     int Ng, int E, float* __restrict__ lse_img, float* __restrict__ lse_txt,
                     float* __restrict__ partial /*[Ng][2]*/, unsigned int* __restrict__ counter,
                     float* __restrict__ result) {
  extern __shared__ float sm[];
  float* a_img = sm;            // [E]
  float* a_txt = sm + E;        // [E]
  float* vals = sm + 2 * E;     // [Ng]
  __shared__ float sred[THREADS / 32];
  __shared__ int ired[THREADS / 32];
  __shared__ bool is_last;
  const int i = blockIdx.x, tid = threadIdx.x;
  const float s = __ldg(scale_ptr);
  for (int c = tid; c < E; c += THREADS) {
    a_img[c] = s * img[(long long)i * E + c];   // (logit_scale * I) @ T^T : scale applied to the image row first
    a_txt[c] = txt[(long long)i * E + c];
  }
  __syncthreads();
  // image -> text logits, row i
  for (int j = tid; j < Ng; j += THREADS) vals[j] = dot_row(a_img, txt + (long long)j * E, E);
  __syncthreads();
  const RowStat ri = row_stats(vals, Ng, i, sred, ired);
  __syncthreads();
  // text -> image logits, row i of logits^T : s * <I_j, T_i>
  for (int j = tid; j < Ng; j += THREADS) vals[j] = s * dot_row(a_txt, img + (long long)j * E, E);
  __syncthreads();
  const RowStat rt = row_stats(vals, Ng, i, sred, ired);
  if (tid == 0) {
    lse_img[i] = ri.lse;
    lse_txt[i] = rt.lse;
    partial[2 * i] = 0.5f * ((ri.lse - ri.label_logit) + (rt.lse - rt.label_logit));
    partial[2 * i + 1] = (ri.argmax == i) ? 1.f : 0.f;
    __threadfence();
    const unsigned int done = atomicAdd(counter, 1u);
    is_last = (done == (unsigned)Ng - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float l = 0.f, a = 0.f;
    for (int j = tid; j < Ng; j += THREADS) {
      l += __ldcg(partial + 2 * j);
      a += __ldcg(partial + 2 * j + 1);
    }
    l = warp_sum(l);
    a = warp_sum(a);
    __shared__ float lr[THREADS / 32], ar[THREADS / 32];
    if ((tid & 31) == 0) { lr[tid >> 5] = l; ar[tid >> 5] = a; }
    __syncthreads();
    if (tid == 0) {
      float L = 0.f, A = 0.f;
      for (int w = 0; w < THREADS / 32; ++w) { L += lr[w]; A += ar[w]; }
      result[0] = L / Ng;
      result[1] = 100.f * A / Ng;
      *counter = 0;  // self-reset for the next call
    }
  }
}

// grid = 2 * Nl CTAs (Nl local rows starting at global row r0).  CTA k < Nl: d loss / d I_{r0+k};  k >= Nl: d loss / d T_{r0+k-Nl}.
//   dlogits[i,j] = ( softmax_img[i,j] + softmax_txt[j,i] - 2*delta_ij ) / (2 Ng)
//   dI_i = g * s * sum_j dlogits[i,j] T_j        dT_j = g * s * sum_i dlogits[i,j] I_i
//   dscale += g0 * sum_{i local, j} dlogits[i,j] <I_i,T_j>   (each rank contributes its local rows; callers sum / DDP averages)
// g = upstream grad * grad_scale (W for the vissl path).
__global__ void __launch_bounds__(THREADS)
clip_loss_bwd_kernel(const float* __restrict__ img, const float* __restrict__ txt, const float* __restrict__ scale_ptr,
                     const float* __restrict__ lse_img, const float* __restrict__ lse_txt,
                     const float* __restrict__ gout_ptr, float grad_scale, float scale_grad_scale, int Ng, int E, int r0,
                     int Nl, float* __restrict__ d_img, float* __restrict__ d_txt, float* __restrict__ d_scale) {
  extern __shared__ float sm[];
  float* a_row = sm;          // [E] the row this CTA differentiates
  float* w = sm + E;          // [Ng] dlogits coefficients
  __shared__ float sred[THREADS / 32];
  const int tid = threadIdx.x;
  const bool is_img = blockIdx.x < (unsigned)Nl;
  const int i = r0 + (is_img ? blockIdx.x : blockIdx.x - Nl);
  const float s = __ldg(scale_ptr);
  const float gout = __ldg(gout_ptr);
  const float* self = is_img ? img : txt;
  const float* other = is_img ? txt : img;
  for (int c = tid; c < E; c += THREADS) a_row[c] = self[(long long)i * E + c];
  __syncthreads();
  const float my_lse_img = is_img ? lse_img[i] : 0.f;
  const float my_lse_txt = is_img ? 0.f : lse_txt[i];
  float ds = 0.f;
  for (int j = tid; j < Ng; j += THREADS) {
    const float l = s * dot_row(a_row, other + (long long)j * E, E);   // logits[i,j] (img CTA) or logits[j,i] (txt CTA)
    float p_img, p_txt;
    if (is_img) { p_img = __expf(l - my_lse_img); p_txt = __expf(l - lse_txt[j]); }
    else        { p_img = __expf(l - lse_img[j]); p_txt = __expf(l - my_lse_txt); }
    const float d = (p_img + p_txt - ((j == i) ? 2.f : 0.f)) / (2.f * Ng);
    w[j] = d;
    ds += d * l;
  }
  __syncthreads();
  // d(row) = g * s * sum_j w[j] * other[j,:]
  float* dst = (is_img ? d_img : d_txt) + (long long)(i - r0) * E;
  for (int c = tid; c < E; c += THREADS) {
    float acc = 0.f;
    for (int j = 0; j < Ng; ++j) acc += w[j] * __ldg(other + (long long)j * E + c);
    dst[c] = gout * grad_scale * s * acc;
  }
  if (is_img && d_scale) {
    ds = warp_sum(ds);
    if ((tid & 31) == 0) sred[tid >> 5] = ds;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int k = 0; k < THREADS / 32; ++k) t += sred[k];
      // ds accumulated d * l = d * s * <I,T>;  d loss / d s = sum d * <I,T> = t / s
      atomicAdd(d_scale, gout * scale_grad_scale * t / s);
    }
  }
}

// ================================================================================================ SSLCLIPLoss
// lavila/models/loss.py:148-213: the LaViLa recipe's loss over human (gt = 1) and pseudo-narrated (gt = 0) pairs.
//   logits[i,j] = c(i,j) * <I_i, T_j>,   c = s_p (both pseudo) | sqrt(s_p * s) (mixed) | s (both human)   (:160-164)
// c is symmetric, so logits_per_text = logits^T (:165 / :178).  Same structure as the CLIPLoss kernels above.
__device__ __forceinline__ float pair_scale(int gi, int gj, float s, float sp, float smix) {
  const int m = gi + gj;
  return m == 2 ? s : (m == 1 ? smix : sp);
}

// result[6] = {loss, acc, acc_gt, acc_pseudo, num_gt, num_pseudo};  partial [Ng][2] = {loss term, correct}.
__global__ void __launch_bounds__(THREADS)
ssl_clip_loss_fwd_kernel(const float* __restrict__ img, const float* __restrict__ txt, const float* __restrict__ scale_ptr,
                         const float* __restrict__ scale_pseudo_ptr, const int* __restrict__ gt, int Ng, int E,
                         float* __restrict__ lse_img, float* __restrict__ lse_txt, float* __restrict__ partial,
                         unsigned int* __restrict__ counter, float* __restrict__ result) {
  extern __shared__ float sm[];
  float* a_img = sm;            // [E]
  float* a_txt = sm + E;        // [E]
  float* vals = sm + 2 * E;     // [Ng]
  __shared__ float sred[THREADS / 32];
  __shared__ int ired[THREADS / 32];
  __shared__ bool is_last;
  const int i = blockIdx.x, tid = threadIdx.x;
  const float s = __ldg(scale_ptr), sp = __ldg(scale_pseudo_ptr), smix = sqrtf(sp * s);
  const int gi = gt[i];
  for (int c = tid; c < E; c += THREADS) {
    a_img[c] = img[(long long)i * E + c];
    a_txt[c] = txt[(long long)i * E + c];
  }
  __syncthreads();
  for (int j = tid; j < Ng; j += THREADS) vals[j] = pair_scale(gi, gt[j], s, sp, smix) * dot_row(a_img, txt + (long long)j * E, E);
  __syncthreads();
  const RowStat ri = row_stats(vals, Ng, i, sred, ired);
  __syncthreads();
  for (int j = tid; j < Ng; j += THREADS) vals[j] = pair_scale(gi, gt[j], s, sp, smix) * dot_row(a_txt, img + (long long)j * E, E);
  __syncthreads();
  const RowStat rt = row_stats(vals, Ng, i, sred, ired);
  if (tid == 0) {
    lse_img[i] = ri.lse;
    lse_txt[i] = rt.lse;
    partial[2 * i] = 0.5f * ((ri.lse - ri.label_logit) + (rt.lse - rt.label_logit));
    partial[2 * i + 1] = (ri.argmax == i) ? 1.f : 0.f;
    __threadfence();
    const unsigned int done = atomicAdd(counter, 1u);
    is_last = (done == (unsigned)Ng - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float v[4] = {0.f, 0.f, 0.f, 0.f};   // loss, correct, correct among gt, number of gt
    for (int j = tid; j < Ng; j += THREADS) {
      const float c = __ldcg(partial + 2 * j + 1);
      const float isg = gt[j] == 1 ? 1.f : 0.f;
      v[0] += __ldcg(partial + 2 * j);
      v[1] += c;
      v[2] += c * isg;
      v[3] += isg;
    }
    __shared__ float red[4][THREADS / 32];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = warp_sum(v[k]);
      if ((tid & 31) == 0) red[k][tid >> 5] = v[k];
    }
    __syncthreads();
    if (tid == 0) {
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < 4; ++k)
        for (int w = 0; w < THREADS / 32; ++w) t[k] += red[k][w];
      const float n_gt = t[3], n_ps = (float)Ng - t[3];
      result[0] = t[0] / Ng;
      result[1] = 100.f * t[1] / Ng;
      result[2] = 100.f * t[2] / n_gt;              // 0/0 -> NaN, like the reference's empty-selection mean (:203-204)
      result[3] = 100.f * (t[1] - t[2]) / n_ps;
      result[4] = n_gt;
      result[5] = n_ps;
      *counter = 0;
    }
  }
}

// Same CTA mapping as clip_loss_bwd_kernel.  d_scales[0] += d loss / d s, d_scales[1] += d loss / d s_p (local image rows).
__global__ void __launch_bounds__(THREADS)
ssl_clip_loss_bwd_kernel(const float* __restrict__ img, const float* __restrict__ txt, const float* __restrict__ scale_ptr,
                         const float* __restrict__ scale_pseudo_ptr, const int* __restrict__ gt,
                         const float* __restrict__ lse_img, const float* __restrict__ lse_txt,
                         const float* __restrict__ gout_ptr, float grad_scale, float scale_grad_scale, int Ng, int E, int r0,
                         int Nl, float* __restrict__ d_img, float* __restrict__ d_txt, float* __restrict__ d_scales) {
  extern __shared__ float sm[];
  float* a_row = sm;          // [E]
  float* w = sm + E;          // [Ng] dlogits * pair scale
  __shared__ float sred[2][THREADS / 32];
  const int tid = threadIdx.x;
  const bool is_img = blockIdx.x < (unsigned)Nl;
  const int i = r0 + (is_img ? blockIdx.x : blockIdx.x - Nl);
  const float s = __ldg(scale_ptr), sp = __ldg(scale_pseudo_ptr), smix = sqrtf(sp * s);
  const float gout = __ldg(gout_ptr);
  const int gi = gt[i];
  const float* self = is_img ? img : txt;
  const float* other = is_img ? txt : img;
  for (int c = tid; c < E; c += THREADS) a_row[c] = self[(long long)i * E + c];
  __syncthreads();
  const float my_lse_img = is_img ? lse_img[i] : 0.f;
  const float my_lse_txt = is_img ? 0.f : lse_txt[i];
  float ds = 0.f, dsp = 0.f;
  for (int j = tid; j < Ng; j += THREADS) {
    const int m = gi + gt[j];
    const float c = m == 2 ? s : (m == 1 ? smix : sp);
    const float dot = dot_row(a_row, other + (long long)j * E, E);
    const float l = c * dot;
    float p_img, p_txt;
    if (is_img) { p_img = __expf(l - my_lse_img); p_txt = __expf(l - lse_txt[j]); }
    else        { p_img = __expf(l - lse_img[j]); p_txt = __expf(l - my_lse_txt); }
    const float d = (p_img + p_txt - ((j == i) ? 2.f : 0.f)) / (2.f * Ng);
    w[j] = d * c;
    // d c / d s and d c / d s_p
    ds += d * dot * (m == 2 ? 1.f : (m == 1 ? 0.5f * smix / s : 0.f));
    dsp += d * dot * (m == 0 ? 1.f : (m == 1 ? 0.5f * smix / sp : 0.f));
  }
  __syncthreads();
  float* dst = (is_img ? d_img : d_txt) + (long long)(i - r0) * E;
  for (int c = tid; c < E; c += THREADS) {
    float acc = 0.f;
    for (int j = 0; j < Ng; ++j) acc += w[j] * __ldg(other + (long long)j * E + c);
    dst[c] = gout * grad_scale * acc;
  }
  if (is_img && d_scales) {
    ds = warp_sum(ds);
    dsp = warp_sum(dsp);
    if ((tid & 31) == 0) { sred[0][tid >> 5] = ds; sred[1][tid >> 5] = dsp; }
    __syncthreads();
    if (tid == 0) {
      float t0 = 0.f, t1 = 0.f;
      for (int k = 0; k < THREADS / 32; ++k) { t0 += sred[0][k]; t1 += sred[1][k]; }
      atomicAdd(d_scales, gout * scale_grad_scale * t0);
      atomicAdd(d_scales + 1, gout * scale_grad_scale * t1);
    }
  }
}

// ================================================================================================ fused gather + loss
// Multi-GPU CLIPLoss forward with the embedding all-gather INSIDE the kernel (loss.py:76-79 + distributed_utils.py:51-62):
// every rank owns a symmetric block (NVLink peer-mapped on all ranks) of two slots (step parity), each
//   [Bl rows x 2E fp32 = image | text][32-word pad, word 0 = ready flag].
// One kernel per rank, launched cooperatively with Ng = W*Bl CTAs:
//   phase 0  CTAs of this rank's own rows publish their row into the rank's symmetric slot; the last one releases the
//            slot flag (= step) at system scope;
//   phase 1  CTA i acquires the flag of the rank that owns global row i (spin on the peer's memory over NVLink), pulls
//            that row (2E floats, peer loads) into the local gathered buffers all_img / all_txt;
//   phase 2  grid barrier, then exactly the single-GPU loss of clip_loss_fwd_kernel on the gathered rows.
// NVLink traffic is the minimum (every remote row crosses once); no NCCL call, no gathered staging copy, no host sync.
// Slot reuse is safe by the flag protocol alone: a rank rewrites slot s in its step s+2 kernel, which starts after its step s+1
// kernel completed, which needed every peer's step s+1 flag, which a peer only raises after its own step s kernel (the one that
// reads slot s) has finished.  A peer that never publishes (crashed rank) makes the spin time out: result = NaN, no hang.
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

constexpr int P2P_PAD = 32;   // floats after the data of a slot; word 0 = flag

// ctrl[0] = last-CTA counter (as clip_loss_fwd_kernel), ctrl[1] = own-rows-published counter, ctrl[2] = grid-barrier
// arrivals, ctrl[3] = grid-barrier generation, ctrl[4] = sticky error word (the step at which a peer's rows did not arrive
// within timeout_ns; the host raises on it).  All zero before the first call; words 0..3 are self-resetting.
__global__ void __launch_bounds__(THREADS)
clip_loss_fwd_gather_kernel(const float* __restrict__ img_local, const float* __restrict__ txt_local,
                            float* const* __restrict__ peers, int rank, int W, int Bl, unsigned int step,
                            float* __restrict__ all_img, float* __restrict__ all_txt, const float* __restrict__ scale_ptr,
                            int E, float* __restrict__ lse_img, float* __restrict__ lse_txt, float* __restrict__ partial,
                            unsigned int* __restrict__ ctrl, float* __restrict__ result, unsigned long long timeout_ns) {
  extern __shared__ float sm[];
  float* a_img = sm;
  float* a_txt = sm + E;
  float* vals = sm + 2 * E;
  __shared__ float sred[THREADS / 32];
  __shared__ int ired[THREADS / 32];
  __shared__ bool is_last;
  __shared__ int timed_out;
  const int Ng = W * Bl;
  const int i = blockIdx.x, tid = threadIdx.x;
  const int owner = i / Bl, lrow = i - owner * Bl;
  const long long slot_floats = (long long)Bl * 2 * E + P2P_PAD;
  const long long slot_off = (long long)(step & 1u) * slot_floats;
  if (tid == 0) timed_out = 0;
  // ---- phase 0: publish own rows
  if (owner == rank) {
    float* mine = peers[rank] + slot_off + (long long)lrow * 2 * E;
    for (int c = tid; c < E; c += THREADS) {
      const float a = img_local[(long long)lrow * E + c], b = txt_local[(long long)lrow * E + c];
      mine[c] = a;
      mine[E + c] = b;
      all_img[(long long)i * E + c] = a;
      all_txt[(long long)i * E + c] = b;
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      const unsigned int done = atomicAdd(ctrl + 1, 1u);
      if (done == (unsigned)Bl - 1) {
        ctrl[1] = 0;
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned int*>(peers[rank] + slot_off + (long long)Bl * 2 * E), step);
      }
    }
  } else {
    // ---- phase 1: wait for the owner's slot, pull the row over NVLink
    const float* theirs = peers[owner] + slot_off;
    if (tid == 0) {
      const unsigned int* flag = reinterpret_cast<const unsigned int*>(theirs + (long long)Bl * 2 * E);
      const unsigned long long t0 = globaltimer_ns();
      unsigned int spins = 0;
      while (ld_acquire_sys(flag) != step) {
        __nanosleep(spins < 64 ? 100 : 2000);
        // the wall-clock bound matches the process group's collective timeout (host passes it): a peer that is merely
        // slow (checkpoint write, data stall) is waited for like NCCL would; one that is gone fails loudly -- NaN loss
        // AND the sticky error word the host turns into LavilaB200Error
        if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > timeout_ns) {
          timed_out = 1;
          atomicExch(ctrl + 4, step);
          break;
        }
      }
    }
    __syncthreads();
    const float* row = theirs + (long long)lrow * 2 * E;
    for (int c = tid; c < E; c += THREADS) {
      all_img[(long long)i * E + c] = ld_relaxed_sys_f32(row + c);
      all_txt[(long long)i * E + c] = ld_relaxed_sys_f32(row + E + c);
    }
  }
  // ---- grid barrier (all Ng CTAs are co-resident: cooperative launch)
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int gen = ld_volatile_u32(ctrl + 3);
    if (atomicAdd(ctrl + 2, 1u) == (unsigned)Ng - 1) {
      ctrl[2] = 0;
      __threadfence();
      atomicAdd(ctrl + 3, 1u);
    } else {
      while (ld_volatile_u32(ctrl + 3) == gen) __nanosleep(100);
    }
    __threadfence();
  }
  __syncthreads();
  // ---- phase 2: the single-GPU loss on the gathered rows (plain loads: the rows were written by other CTAs of this kernel)
  const float s = __ldg(scale_ptr);
  for (int c = tid; c < E; c += THREADS) {
    a_img[c] = s * __ldcg(all_img + (long long)i * E + c);
    a_txt[c] = __ldcg(all_txt + (long long)i * E + c);
  }
  __syncthreads();
  for (int j = tid; j < Ng; j += THREADS) {
    const float* b = all_txt + (long long)j * E;
    float acc = 0.f;
    for (int c = 0; c < E; c += 4) {
      const float4 x = *reinterpret_cast<const float4*>(a_img + c);
      const float4 y = __ldcg(reinterpret_cast<const float4*>(b + c));
      acc += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
    }
    vals[j] = acc;
  }
  __syncthreads();
  const RowStat ri = row_stats(vals, Ng, i, sred, ired);
  __syncthreads();
  for (int j = tid; j < Ng; j += THREADS) {
    const float* b = all_img + (long long)j * E;
    float acc = 0.f;
    for (int c = 0; c < E; c += 4) {
      const float4 x = *reinterpret_cast<const float4*>(a_txt + c);
      const float4 y = __ldcg(reinterpret_cast<const float4*>(b + c));
      acc += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
    }
    vals[j] = s * acc;
  }
  __syncthreads();
  const RowStat rt = row_stats(vals, Ng, i, sred, ired);
  if (tid == 0) {
    lse_img[i] = ri.lse;
    lse_txt[i] = rt.lse;
    partial[2 * i] = timed_out ? NAN : 0.5f * ((ri.lse - ri.label_logit) + (rt.lse - rt.label_logit));
    partial[2 * i + 1] = (ri.argmax == i) ? 1.f : 0.f;
    __threadfence();
    const unsigned int done = atomicAdd(ctrl, 1u);
    is_last = (done == (unsigned)Ng - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float l = 0.f, a = 0.f;
    for (int j = tid; j < Ng; j += THREADS) {
      l += __ldcg(partial + 2 * j);
      a += __ldcg(partial + 2 * j + 1);
    }
    l = warp_sum(l);
    a = warp_sum(a);
    __shared__ float lr[THREADS / 32], ar[THREADS / 32];
    if ((tid & 31) == 0) { lr[tid >> 5] = l; ar[tid >> 5] = a; }
    __syncthreads();
    if (tid == 0) {
      float Lsum = 0.f, A = 0.f;
      for (int w = 0; w < THREADS / 32; ++w) { Lsum += lr[w]; A += ar[w]; }
      result[0] = Lsum / Ng;
      result[1] = 100.f * A / Ng;
      ctrl[0] = 0;
    }
  }
}

}  // namespace loss
}  // namespace lv

using namespace lv;

extern "C" int lv_clip_loss_fwd(const float* img, const float* txt, const float* scale_ptr, int Ng, int E, float* lse_img,
                                float* lse_txt, float* partial, uint32_t* counter, float* result, void* stream) {
  LV_REQUIRE(img && txt && scale_ptr && lse_img && lse_txt && partial && counter && result, "lv_clip_loss_fwd: null pointer");
  LV_REQUIRE(Ng > 0 && E > 0 && E % 4 == 0, "lv_clip_loss_fwd: bad shape Ng=%d E=%d", Ng, E);
  const size_t smem = (size_t)(2 * E + Ng) * sizeof(float);
  LV_REQUIRE(smem <= 48 * 1024, "lv_clip_loss_fwd: Ng=%d too large for one CTA row buffer", Ng);
  loss::clip_loss_fwd_kernel<<<Ng, loss::THREADS, smem, (cudaStream_t)stream>>>(img, txt, scale_ptr, Ng, E, lse_img, lse_txt, partial, counter, result);
  return check_launch("lv_clip_loss_fwd");
}

extern "C" int lv_clip_loss_bwd(const float* img, const float* txt, const float* scale_ptr, const float* lse_img,
                                const float* lse_txt, const float* gout, float grad_scale, float scale_grad_scale, int Ng,
                                int E, int r0, int Nl, float* d_img, float* d_txt, float* d_scale, void* stream) {
  LV_REQUIRE(img && txt && scale_ptr && lse_img && lse_txt && gout && d_img && d_txt, "lv_clip_loss_bwd: null pointer");
  LV_REQUIRE(Ng > 0 && E % 4 == 0 && r0 >= 0 && Nl > 0 && r0 + Nl <= Ng, "lv_clip_loss_bwd: bad shape");
  const size_t smem = (size_t)(E + Ng) * sizeof(float);
  LV_REQUIRE(smem <= 48 * 1024, "lv_clip_loss_bwd: Ng too large");
  loss::clip_loss_bwd_kernel<<<2 * Nl, loss::THREADS, smem, (cudaStream_t)stream>>>(img, txt, scale_ptr, lse_img, lse_txt, gout, grad_scale, scale_grad_scale, Ng, E, r0, Nl, d_img, d_txt, d_scale);
  return check_launch("lv_clip_loss_bwd");
}

extern "C" int lv_ssl_clip_loss_fwd(const float* img, const float* txt, const float* scale_ptr, const float* scale_pseudo_ptr,
                                    const int32_t* gt, int Ng, int E, float* lse_img, float* lse_txt, float* partial,
                                    uint32_t* counter, float* result, void* stream) {
  LV_REQUIRE(img && txt && scale_ptr && scale_pseudo_ptr && gt && lse_img && lse_txt && partial && counter && result,
             "lv_ssl_clip_loss_fwd: null pointer");
  LV_REQUIRE(Ng > 0 && E > 0 && E % 4 == 0, "lv_ssl_clip_loss_fwd: bad shape Ng=%d E=%d", Ng, E);
  const size_t smem = (size_t)(2 * E + Ng) * sizeof(float);
  LV_REQUIRE(smem <= 48 * 1024, "lv_ssl_clip_loss_fwd: Ng=%d too large for one CTA row buffer", Ng);
  loss::ssl_clip_loss_fwd_kernel<<<Ng, loss::THREADS, smem, (cudaStream_t)stream>>>(img, txt, scale_ptr, scale_pseudo_ptr, gt, Ng, E, lse_img, lse_txt, partial, counter, result);
  return check_launch("lv_ssl_clip_loss_fwd");
}

extern "C" int lv_ssl_clip_loss_bwd(const float* img, const float* txt, const float* scale_ptr, const float* scale_pseudo_ptr,
                                    const int32_t* gt, const float* lse_img, const float* lse_txt, const float* gout,
                                    float grad_scale, float scale_grad_scale, int Ng, int E, int r0, int Nl, float* d_img,
                                    float* d_txt, float* d_scales, void* stream) {
  LV_REQUIRE(img && txt && scale_ptr && scale_pseudo_ptr && gt && lse_img && lse_txt && gout && d_img && d_txt,
             "lv_ssl_clip_loss_bwd: null pointer");
  LV_REQUIRE(Ng > 0 && E % 4 == 0 && r0 >= 0 && Nl > 0 && r0 + Nl <= Ng, "lv_ssl_clip_loss_bwd: bad shape");
  const size_t smem = (size_t)(E + Ng) * sizeof(float);
  LV_REQUIRE(smem <= 48 * 1024, "lv_ssl_clip_loss_bwd: Ng too large");
  loss::ssl_clip_loss_bwd_kernel<<<2 * Nl, loss::THREADS, smem, (cudaStream_t)stream>>>(img, txt, scale_ptr, scale_pseudo_ptr, gt, lse_img, lse_txt, gout, grad_scale, scale_grad_scale, Ng, E, r0, Nl, d_img, d_txt, d_scales);
  return check_launch("lv_ssl_clip_loss_bwd");
}

static int gather_max_rows(int E) {
  // Two bounds on Ng = W * Bl: the [2E + Ng] fp32 row buffer (dynamic) plus the kernel's static shared memory must fit the
  // 48 KB default limit, and all Ng CTAs of the cooperative launch must be co-resident (occupancy falls as the buffer grows).
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, loss::clip_loss_fwd_gather_kernel) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  const long long by_smem = ((long long)48 * 1024 - (long long)fa.sharedSizeBytes) / (long long)sizeof(float) - 2ll * E;
  if (by_smem <= 0) return 0;
  const int sms = sm_count();
  for (long long ng = by_smem; ng >= 1; ng -= (ng > 64 ? 32 : 1)) {
    int per_sm = 0;
    const size_t smem = (size_t)(2 * E + ng) * sizeof(float);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, loss::clip_loss_fwd_gather_kernel, loss::THREADS, smem) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    if ((long long)per_sm * sms >= ng) return (int)ng;
  }
  return 0;
}

extern "C" int lv_clip_loss_gather_max_rows(int E) {
  if (E <= 0 || E % 4 != 0) return 0;
  return gather_max_rows(E);
}

extern "C" int lv_clip_loss_fwd_gather(const float* img_local, const float* txt_local, void* const* peers, int rank, int W,
                                       int Bl, uint32_t step, float* all_img, float* all_txt, const float* scale_ptr, int E,
                                       float* lse_img, float* lse_txt, float* partial, uint32_t* ctrl, float* result,
                                       int64_t timeout_ms, void* stream) {
  LV_REQUIRE(img_local && txt_local && peers && all_img && all_txt && scale_ptr && lse_img && lse_txt && partial && ctrl && result,
             "lv_clip_loss_fwd_gather: null pointer");
  LV_REQUIRE(W > 0 && rank >= 0 && rank < W && Bl > 0 && E > 0 && E % 4 == 0 && step > 0, "lv_clip_loss_fwd_gather: bad arguments");
  LV_REQUIRE(timeout_ms > 0, "lv_clip_loss_fwd_gather: timeout_ms must be positive");
  int Ng = W * Bl;
  const size_t smem = (size_t)(2 * E + Ng) * sizeof(float);
  LV_REQUIRE(smem <= 48 * 1024, "lv_clip_loss_fwd_gather: Ng=%d too large for one CTA row buffer (see lv_clip_loss_gather_max_rows)", Ng);
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, loss::clip_loss_fwd_gather_kernel, loss::THREADS, smem);
  if (e != cudaSuccess) return set_error((int)e, "lv_clip_loss_fwd_gather: occupancy query: %s", cudaGetErrorString(e));
  LV_REQUIRE((long long)per_sm * sm_count() >= Ng, "lv_clip_loss_fwd_gather: %d CTAs cannot be co-resident (%d per SM; see lv_clip_loss_gather_max_rows)", Ng, per_sm);
  float* const* peers_f = reinterpret_cast<float* const*>(peers);
  unsigned long long timeout_ns = (unsigned long long)timeout_ms * 1000000ull;
  void* args[] = {(void*)&img_local, (void*)&txt_local, (void*)&peers_f, (void*)&rank, (void*)&W, (void*)&Bl, (void*)&step,
                  (void*)&all_img, (void*)&all_txt, (void*)&scale_ptr, (void*)&E, (void*)&lse_img, (void*)&lse_txt,
                  (void*)&partial, (void*)&ctrl, (void*)&result, (void*)&timeout_ns};
  e = cudaLaunchCooperativeKernel((const void*)loss::clip_loss_fwd_gather_kernel, dim3(Ng), dim3(loss::THREADS), args, smem,
                                  (cudaStream_t)stream);
  if (e != cudaSuccess) return set_error((int)e, "lv_clip_loss_fwd_gather: cooperative launch: %s", cudaGetErrorString(e));
  return check_launch("lv_clip_loss_fwd_gather");
}
