/*
 * lavila_b200 -- C ABI of the B200-native (sm_100a) kernels behind LaViLa's dual-encoder training step.
 *
 * The reference (facebookresearch/LaViLa) is pure PyTorch and has no FFI of its own; every entry point
 * below replaces a chain of torch library calls, cited as `reference file:line` (paths relative to the
 * reference root).  The reference-side binding is the ctypes stub in lavila_b200/_lib.py.
 *
 * Conventions:
 *   - plain pointers + sizes; every pointer is a DEVICE pointer unless stated; no torch types;
 *   - the caller owns and allocates every buffer and keeps it alive until the stream work completes;
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises the device or touches stream 0
 *     unless stream 0 is what the caller passed;
 *   - return value: 0 = ok, <0 = invalid argument / unsupported shape, >0 = cudaError_t;
 *     lv_last_error() returns a thread-local message for the last non-zero return;
 *   - re-entrant; safe to call from the Python main thread (forward) and the autograd thread (backward).
 *   - bf16 = __nv_bfloat16 bits (uint16_t), row-major matrices with explicit leading dimensions (elements).
 */
#ifndef LAVILA_B200_H
#define LAVILA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LV_ABI_VERSION 1

int lv_version(void);
const char* lv_last_error(void);
/* Number of kernels this library has launched in this process (all threads). bench.py reports it. */
int64_t lv_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM (tcgen05 tensor cores, TMA-staged operands, TMEM accumulators)
 *   C[M,N] = epilogue( A_op[M,K] * B_op[N,K]^T )
 * Replaces nn.Linear / F.linear calls on the hot path:
 *   lavila/models/timesformer.py:110 (qkv), :142 (proj), :53,:56 (Mlp fc1/fc2), :82-83 (patch-embed conv as
 *   im2col GEMM); lavila/models/openai_model.py:186,198 (in_proj/out_proj), :188-192 (c_fc/c_proj);
 *   lavila/models/models.py:146,160 (projections) and all of their autograd backward GEMMs.
 *
 *   a_mn = 0: A is [M][lda], K contiguous ("K-major").   a_mn = 1: A is stored [K][lda], M contiguous.
 *   b_mn = 0: B is [N][ldb], K contiguous.               b_mn = 1: B is stored [K][ldb], N contiguous.
 *     forward  y = x W^T          : A = x  (a_mn 0), B = W  [out,in]      (b_mn 0)
 *     dgrad    dx = dy W          : A = dy (a_mn 0), B = W  [out,in] as [K=out][N=in] (b_mn 1)
 *     wgrad    dW = dy^T x        : A = dy [tok][out] (a_mn 1), B = x [tok][in] (b_mn 1), K = tokens
 *   k_splits > 1 splits the reduction across CTAs; requires LV_EPI_ATOMIC (fp32 red.add into `out`).
 * ---------------------------------------------------------------------------------------------- */
enum {
  LV_EPI_BIAS = 1,        /* v += bias[n]                         (fp32 bias)                       */
  LV_EPI_QUICKGELU = 2,   /* out2 <- bf16(v) (pre-activation); v = v * sigmoid(1.702 v)              */
  LV_EPI_DQUICKGELU = 4,  /* v *= d/dh quickgelu(h), h = aux[m,n] (bf16)                             */
  LV_EPI_SCALE = 8,       /* v *= s, s = *scale_ptr (or tanh(*scale_ptr) with LV_EPI_SCALE_TANH)     */
  LV_EPI_SCALE_TANH = 16,
  LV_EPI_RESID = 32,      /* v += resid[m,n] (fp32)                                                  */
  LV_EPI_OUT_F32 = 64,    /* `out` is fp32 (default bf16)                                            */
  LV_EPI_COPY_BF16 = 128, /* out2 <- bf16(v) (final value), e.g. bf16 shadow of an fp32 stream       */
  LV_EPI_ATOMIC = 256,    /* out[m,n] += v with fp32 atomics (implies LV_EPI_OUT_F32)                */
  LV_EPI_ROWBIAS = 512    /* v += bias[m] instead of bias[n] (unused by torch layouts; tests only)   */
};

typedef struct LvGemmEpilogue {
  int32_t flags;
  int32_t _pad;
  void* out;          int64_t ldo;
  void* out2;         int64_t ldo2;   /* bf16 */
  const float* bias;
  const float* resid; int64_t ldr;
  const void* aux;    int64_t ldaux;  /* bf16 */
  const float* scale_ptr;
} LvGemmEpilogue;

int lv_gemm_bf16(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, int64_t M, int64_t N,
                 int64_t K, int k_splits, const LvGemmEpilogue* epi, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAVILA_B200_H */
