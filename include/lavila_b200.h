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
/* Bytes of caller-allocated scratch an op needs (the library allocates nothing); -1 = unknown op / bad shape.
 *   LV_WS_GEMM_SKINNY (M, N, K)    workspace of lv_gemm_skinny_bf16
 *   LV_WS_ATTN_BWD_DCLS (B, H, 0)  dcls_kv of the attention backward entry points (must be zeroed)
 *   LV_WS_CLIP_LOSS (Ng, 0, 0)     lse_img + lse_txt + partial + control words + result of the loss entry points
 *   LV_WS_P2P_BLOCK (Bl, E, 0)     one rank's symmetric block for lv_clip_loss_fwd_gather */
enum { LV_WS_GEMM_SKINNY = 1, LV_WS_ATTN_BWD_DCLS = 2, LV_WS_CLIP_LOSS = 3, LV_WS_P2P_BLOCK = 4 };
int64_t lv_workspace_bytes(int op, int64_t a, int64_t b, int64_t c);

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
  LV_EPI_ROWBIAS = 512,   /* v += bias[m] instead of bias[n] (unused by torch layouts; tests only)   */
  LV_EPI_GELU_TANH = 1024,/* v = 0.5 v (1 + tanh(sqrt(2/pi)(v + 0.044715 v^3)))  (HF 'gelu_new', inference) */
  LV_EPI_SQRELU = 2048    /* v = relu(v)^2  (gpt2_gated.SqReLU, inference)                           */
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
/* Same contract; 256 x 256 tiles computed by CTA pairs (tcgen05 cta_group::2): each SM stages its own 128 rows of A and
 * half of B, which cuts shared-memory traffic per FLOP by a third.  Preferred when M >= 256. */
int lv_gemm_bf16_2cta(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, int64_t M, int64_t N,
                      int64_t K, int k_splits, const LvGemmEpilogue* epi, void* stream);

/* Skinny GEMM for KV-cached decoding (M = sequences being decoded, up to a few hundred rows): C = epilogue(A[M,K] x W) with
 * W stored [K][ldb], N contiguous (HF Conv1D, gpt2_gated.py:47; the b_mn = 1 layout above).  Weight streaming is the
 * roofline: the grid is (N/64) x splits x ceil(M/64) CTAs so that every SM streams W; partial tiles go to `workspace`
 * (fp32 [splits][M][N], caller-allocated) with plain stores and are summed in order by a second kernel that applies the
 * epilogue (deterministic: no atomics).  Supported flags: BIAS, GELU_TANH, SQRELU, SCALE[_TANH], RESID, OUT_F32.
 * N % 64 == 0, K % 64 == 0.  lv_gemm_skinny_splits returns the split count to use (0 = shape not supported). */
int lv_gemm_skinny_splits(int64_t M, int64_t N, int64_t K);
int lv_gemm_skinny_bf16(const void* A, int64_t lda, const void* W, int64_t ldb, int64_t M, int64_t N, int64_t K,
                        float* workspace, int splits, const LvGemmEpilogue* epi, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm over the last dimension (one warp per row, fp32 statistics).  D % 128 == 0, D <= 1024.
 * Replaces nn.LayerNorm at lavila/models/timesformer.py:180,189,196 (eps 1e-6), :366 ln_pre (1e-5), :377 norm;
 * lavila/models/openai_model.py:196,200 ln_1/ln_2 and models.py:156 ln_final (1e-5).
 *   fwd: y = LN(x) written as bf16 (GEMM operand) and/or fp32.  Any of y_bf16 / y_f32 may be NULL (not both).
 *        fwd also accepts any D % 4 == 0 (GPT-2 XL: 1600) and beta == NULL (coca.LayerNorm, coca.py:28-35).
 *   bwd: dx = dLN(dy) [+ add1] [+ add2] written as fp32 and/or bf16; dgamma/dbeta accumulated with fp32 atomics
 *        (may be NULL together).  mean/rstd are recomputed from x.
 * ---------------------------------------------------------------------------------------------- */
int lv_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, void* y_bf16,
                     int64_t ldy, float* y_f32, int64_t ldyf, int64_t rows, int D, void* stream);
int lv_layernorm_bwd(const void* dy, int dy_is_bf16, int64_t lddy, const float* x, int64_t ldx, const float* gamma,
                     float eps, const float* add1, int64_t ld1, const float* add2, int64_t ld2, float* dx, int64_t lddx,
                     void* dx_bf16, int64_t lddxb, float* dgamma, float* dbeta, int64_t rows, int D, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Divided space-time attention + causal text attention (head_dim 64), in place on the packed projection output.
 * Replaces VarAttention.forward's regroup/concat/bmm/softmax/bmm chain, lavila/models/timesformer.py:111-140 and
 * attn() :35-39; nn.MultiheadAttention's core at lavila/models/openai_model.py:196-198 (mode 2).
 *   qkv  bf16 [rows, 3*D] = [q | k | v] (D = H*64), out bf16 [rows, D], lse fp32 [rows, H].
 *   mode 0 = space  (group = (clip, head, frame):      n queries, n + CLS keys)
 *   mode 1 = time   (group = (clip, head, position):   T queries, T + CLS keys, token stride n)
 *   mode 2 = causal (group = (caption, head):          L queries/keys, additive -inf upper triangle)
 *   rows = B * (1 + T*n) for modes 0/1 (token 0 of each clip is CLS), B * L for mode 2.
 * The group kernels write the patch-token rows; lv_cls_attn_* handle token 0 (the CLS query attends to all N tokens,
 * timesformer.py:119).  Backward call order for modes 0/1 (dcls_kv: fp32 [B, H, 2, 64] scratch, ZEROED by the caller):
 * lv_group_attn_bwd / lv_space_attn_bwd_tc (accumulate_kv=0: plain stores, CLS-key partials -> dcls_kv atomics)
 * -> lv_cls_attn_bwd(accumulate=1: streams over all keys adding its contribution) -> lv_cls_kv_finalize.
 * (The opposite order, cls first then group with accumulate_kv=1, is also supported but exposes the read-modify-write
 * latency inside the tensor-core kernel: measured 3.8 ms vs 2.0 ms per launch.)
 * ---------------------------------------------------------------------------------------------- */
int lv_group_attn_fwd(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* lse, int mode, int B, int H,
                      int T, int n, int L, void* stream);
int lv_group_attn_bwd(const void* qkv, int64_t ld_qkv, const void* out, int64_t ld_out, const float* lse,
                      const void* dout, int64_t ld_dout, void* dqkv, int64_t ld_dqkv, float* dcls_kv, int accumulate_kv,
                      int mode, int B, int H, int T, int n, int L, void* stream);
/* Space attention forward on tcgen05 tensor cores (S, O accumulators in TMEM; TMA-loaded Q/K/V tiles; P staged in
 * swizzled shared memory).  Same contract as lv_group_attn_fwd(mode 0); requires 129 <= n <= 207 (TSF-B: 196). */
int lv_space_attn_fwd_tc(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* lse, int B, int H, int T,
                         int n, void* stream);
/* Space attention backward on tcgen05 (S^T/dP^T, dV, dK, dQ accumulators in TMEM; P^T/dS^T staged in shared memory and
 * consumed both K-major and MN-major).  Same contract and call order as lv_group_attn_bwd(mode 0). */
int lv_space_attn_bwd_tc(const void* qkv, int64_t ld_qkv, const void* out, int64_t ld_out, const float* lse,
                         const void* dout, int64_t ld_dout, void* dqkv, int64_t ld_dqkv, float* dcls_kv, int accumulate_kv,
                         int B, int H, int T, int n, void* stream);
/* General attention forward (inference): flash-style, head_dim 64, bf16 in/out, optional causal mask, multi-query
 * K/V when kv_head_stride = 0.  Q element (b,h,i,d) at q[(b*q_rows+i)*ld_q + h*64 + d]; K/V element (b,h,j,d) at
 * k[(b*kv_rows+j)*ld_kv + h*kv_head_stride + d].  Replaces gpt2_gated.GPT2Attention._attn (:206-238) and
 * coca.CrossAttention's einsum/softmax/einsum (coca.py:114-121). */
int lv_flash_attn_fwd(const void* q, int64_t ld_q, int64_t q_rows, const void* k, const void* v, int64_t ld_kv,
                      int64_t kv_rows, int kv_head_stride, void* out, int64_t ld_out, int B, int H, int Lq, int Lk,
                      int causal, float scale, void* stream);
/* Same, but the key count Lk is read from device memory (*lk_dev, 1 <= *lk_dev <= kv_rows) at run time: one captured
 * CUDA graph of a KV-cached decoding step (gpt2_gated.py:331-345 `layer_past`) can be replayed for every position. */
int lv_flash_attn_fwd_dyn(const void* q, int64_t ld_q, int64_t q_rows, const void* k, const void* v, int64_t ld_kv,
                          int64_t kv_rows, int kv_head_stride, void* out, int64_t ld_out, int B, int H, int Lq,
                          const int32_t* lk_dev, int causal, float scale, void* stream);
/* Debug: device buffer (>= 256 int64) receiving clock64() phase stamps of CTA 0 of lv_space_attn_bwd_tc; NULL disables. */
int lv_debug_set_buffer(void* buf);
/* Space attention INCLUDING the CLS query row in one pass (VarAttention.forward of the space half, timesformer.py:116-134): the
 * clip's CLS query rides as an extra row of every frame's Q tile.  fwd: cls_part = fp32 scratch [B*H*T*66]; bwd: dcls_kv fp32
 * [B][H][2][64] and dcls_q fp32 [B][H][64] zeroed by the caller; every row of out / lse / dqkv (CLS rows included) is written.
 * 129 <= n <= 207.  Replace lv_space_attn_fwd_tc + lv_cls_attn_fwd and lv_space_attn_bwd_tc + lv_cls_attn_bwd +
 * lv_cls_kv_finalize (which stream K, V, dK, dV of all tokens a second time). */
int lv_space_attn_fwd_tc_cls(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* lse, float* cls_part, int B,
                             int H, int T, int n, void* stream);
int lv_space_attn_bwd_tc_cls(const void* qkv, int64_t ld_qkv, const void* out, int64_t ld_out, const float* lse,
                             const void* dout, int64_t ld_dout, void* dqkv, int64_t ld_dqkv, float* dcls_kv, float* dcls_q,
                             int B, int H, int T, int n, void* stream);
/* Time attention (<= 16 frames per group) INCLUDING the CLS query row: as lv_space_attn_*_tc_cls with one partial per
 * (clip, head, spatial position): cls_part = fp32 scratch [B*H*n*66]. */
int lv_time_attn_fwd_cls(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* lse, float* cls_part, int B, int H,
                         int T, int n, void* stream);
int lv_time_attn_bwd_cls(const void* qkv, int64_t ld_qkv, const void* out, int64_t ld_out, const float* lse, const void* dout,
                         int64_t ld_dout, void* dqkv, int64_t ld_dqkv, float* dcls_kv, float* dcls_q, int B, int H, int T, int n,
                         void* stream);
int lv_cls_attn_fwd(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, float* lse, int B, int H, int N,
                    void* stream);
int lv_cls_attn_bwd(const void* qkv, int64_t ld_qkv, const void* out, int64_t ld_out, const void* dout, int64_t ld_dout,
                    const float* lse, void* dqkv, int64_t ld_dqkv, float* dcls_kv, int accumulate, int B, int H, int N,
                    void* stream);
int lv_cls_kv_finalize(const float* dcls_kv, void* dqkv, int64_t ld_dqkv, int B, int H, int N, void* stream);
/* CLS query attention with separate operands: q bf16 [B, D], kv bf16 [B*N, 2D] = [k | v], out bf16 [B, D], lse [B, H].
 * Used by the last SpaceTimeBlock, of which only the CLS row is consumed (timesformer.py:376-378). */
int lv_cls_query_attn_fwd(const void* q, const void* kv, void* out, float* lse, int B, int H, int N, void* stream);
int lv_cls_query_attn_bwd(const void* q, const void* kv, const void* out, const void* dout, const float* lse, void* dq,
                          void* dkv, int B, int H, int N, void* stream);

/* ------------------------------------------------------------------------------------------------
 * HBM-bound glue (elementwise.cu)
 * ---------------------------------------------------------------------------------------------- */
/* fp32 -> bf16 (weights each step; autocast's implicit casts in the reference). */
int lv_cast_f32_bf16(const float* in, void* out, int64_t n, void* stream);
/* out[n] += sum_m in[m,n] -- bias gradients (autograd of nn.Linear bias).  in bf16 [M,N]. */
int lv_colsum_bf16(const void* in, int64_t ld, int64_t M, int N, float* out, void* stream);
/* frames fp32 [B,C,T,H,W] -> patch matrix bf16 [B*T*(H/P)*(W/P), ldp], column (c*P+py)*P+px.
 * Folds timesformer.py:387 permute+contiguous and the Conv2d(k=P,s=P) input gather (:77,:82-83). */
int lv_patch_im2col(const float* frames, void* patches, int B, int C, int T, int H, int W, int P, int64_t ldp,
                    void* stream);
/* x0[b,0] = cls + pos[0]; x0[b,1+f*n+i] = patch[(b*T+f)*n+i] + pos[1+i] + temporal[f]  (timesformer.py:353-364). */
int lv_embed_assemble(const float* patch, const float* cls, const float* pos, const float* temporal, float* x0, int B,
                      int T, int n, int D, void* stream);
/* Gradients of the above: dpos/dcls/dtemporal accumulated (+=), dpatch written as compact bf16 [B*T*n, D]. */
int lv_embed_assemble_bwd(const float* dx0, float* dpos, float* dcls, float* dtemporal, void* dpatch_bf16, int B, int T,
                          int n, int D, void* stream);
/* x[r] = tok[text[r]] + pos[r % L]  (models.py:151-152) and its gradient (fp32 atomics). text is int64. */
int lv_text_embed(const int64_t* text, const float* tok, const float* pos, float* x, int64_t rows, int L, int W, int vocab,
                  void* stream);
int lv_text_embed_bwd(const int64_t* text, const float* dx, float* dtok, float* dpos, int64_t rows, int L, int W,
                      int vocab, void* stream);
/* argmax over the last dim of int64 [B,L], first maximum (models.py:160 EOT pick) -- bit exact. */
int lv_argmax_i64(const int64_t* text, int32_t* out, int B, int L, void* stream);
/* scatter=0: dst[r] = src[r*rows_per + idx[r]];  scatter=1: dst[r*rows_per + idx[r]] = src[r].  fp32 rows of W. */
int lv_gather_rows_f32(const float* src, const int32_t* idx, float* dst, int R, int rows_per, int W, int scatter,
                       void* stream);
/* dst[r*stride + c] += src[r*W + c] for R strided rows (dst bf16 or fp32). */
int lv_add_rows(void* dst, int dst_is_bf16, int64_t stride, const float* src, int R, int W, void* stream);
/* F.normalize(dim=-1) (models.py:169-170) and its gradient. */
int lv_l2norm_fwd(const float* x, float* y, float* norm, int R, int E, void* stream);
int lv_l2norm_bwd(const float* dy, const float* y, const float* norm, float* dx, int R, int E, void* stream);

/* ------------------------------------------------------------------------------------------------
 * CLIPLoss on the gathered global batch (lavila/models/loss.py:76-79,107-116).
 *   fwd: result[0] = loss, result[1] = clip_acc (%).  lse_img/lse_txt [Ng], partial [2*Ng], counter (zeroed once).
 *   bwd: gradients of the LOCAL rows [r0, r0+Nl) only, multiplied by grad_scale (= world size for the
 *        --contrastive-use-vissl path, distributed_utils.py:64-67, where the reference all_reduce-SUMs identical
 *        per-rank gradients); d_scale (may be NULL) accumulates scale_grad_scale * sum over local image rows.
 * ---------------------------------------------------------------------------------------------- */
int lv_clip_loss_fwd(const float* img, const float* txt, const float* scale_ptr, int Ng, int E, float* lse_img,
                     float* lse_txt, float* partial, uint32_t* counter, float* result, void* stream);
int lv_clip_loss_bwd(const float* img, const float* txt, const float* scale_ptr, const float* lse_img,
                     const float* lse_txt, const float* gout, float grad_scale, float scale_grad_scale, int Ng, int E,
                     int r0, int Nl, float* d_img, float* d_txt, float* d_scale, void* stream);

/* Multi-GPU CLIPLoss forward with the embedding all-gather fused into the kernel (replaces gather_from_all x2 +
 * the logits/CE chain, lavila/models/loss.py:74-79,107-116 and distributed_utils.py:51-62).  `peers` is a DEVICE array of W
 * pointers to the ranks' symmetric (NVLink peer-mapped) blocks, each 2 slots of [Bl x 2E fp32 = image | text][32 words,
 * word 0 = ready flag]; `step` (> 0, +1 per call on every rank) selects slot step & 1 and is the flag value.  The kernel
 * publishes this rank's rows, pulls every other rank's rows over NVLink into all_img / all_txt [W*Bl, E] (kept for the
 * backward, which is lv_clip_loss_bwd on those buffers) and evaluates the global loss: result[0] = loss, result[1] = acc.
 * ctrl: uint32[8], zeroed once.  Launched cooperatively: W*Bl CTAs must be co-resident and the [2E + W*Bl] fp32 row buffer
 * must fit 48 KB -- lv_clip_loss_gather_max_rows(E) returns the largest supported W*Bl on the current device (0 = none), so
 * the caller can take the all_gather + lv_clip_loss_fwd route for larger global batches instead of failing.
 * timeout_ms (> 0): how long a CTA waits for a peer's rows (the reference's NCCL gather waits for the process-group
 * timeout and then raises, distributed_utils.py:56); on expiry the loss is NaN AND ctrl[4] is set to `step` (sticky), which
 * the caller must check after the stream has drained and turn into an error. */
int lv_clip_loss_gather_max_rows(int E);
int lv_clip_loss_fwd_gather(const float* img_local, const float* txt_local, void* const* peers, int rank, int W, int Bl,
                            uint32_t step, float* all_img, float* all_txt, const float* scale_ptr, int E, float* lse_img,
                            float* lse_txt, float* partial, uint32_t* ctrl, float* result, int64_t timeout_ms, void* stream);

/* SSLCLIPLoss (lavila/models/loss.py:148-213): gt[i] = 1 human narration / 0 pseudo narration; pair scale
 * c(i,j) = *scale_pseudo_ptr (0 + 0) | sqrt(*scale_pseudo_ptr * *scale_ptr) (0 + 1) | *scale_ptr (1 + 1), both pointers hold the
 * already exponentiated scales.  result[6] = {loss, clip_acc, clip_acc_gt, clip_acc_pseudo, num_gt, num_pseudo} (an empty
 * class gives NaN accuracy, as in the reference).  bwd: as lv_clip_loss_bwd; d_scales[2] (may be NULL) accumulates
 * {d loss / d scale, d loss / d scale_pseudo} * scale_grad_scale over the local image rows. */
int lv_ssl_clip_loss_fwd(const float* img, const float* txt, const float* scale_ptr, const float* scale_pseudo_ptr,
                         const int32_t* gt, int Ng, int E, float* lse_img, float* lse_txt, float* partial, uint32_t* counter,
                         float* result, void* stream);
int lv_ssl_clip_loss_bwd(const float* img, const float* txt, const float* scale_ptr, const float* scale_pseudo_ptr,
                         const int32_t* gt, const float* lse_img, const float* lse_txt, const float* gout, float grad_scale,
                         float scale_grad_scale, int Ng, int E, int r0, int Nl, float* d_img, float* d_txt, float* d_scales,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sampling: temperature + nucleus (top-p) filter of next-token logits, in place (lavila/models/narrator.py:131,368-389:
 * TemperatureLogitsWarper then TopPLogitsWarper(min_tokens_to_keep = 1) of transformers).  logits fp32 [rows][ld], V <= 56 320
 * (a row is staged in shared memory); every entry becomes logits / temperature, or -inf if the library would remove it
 * (ascending cumulative probability <= 1 - top_p; ties at the threshold: the same number removed, lowest indices first).
 * ---------------------------------------------------------------------------------------------- */
int lv_top_p_filter(float* logits, int64_t ld, int rows, int V, float temperature, float top_p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Input pipeline: decoded frames -> normalised clip, a batch per launch.  Replaces the CPU transform chain of the DataLoader
 * workers: main_pretrain.py:263-272 (train: Permute, RandomResizedCrop, NormalizeVideo), :274-281 (val: Permute, Resize,
 * CenterCrop, NormalizeVideo), lavila/data/video_transforms.py:15-32.  The random crop box is drawn on the host
 * (lavila_b200/data/video_transforms.py restates torchvision's get_params with the same RNG calls).
 *   desc      device table, 12 x int64 per clip: { src pointer (frames x H x W x 3, channel-interleaved, as the decoder returns
 *             them: lavila/data/datasets.py:25-75), H, W, box_i, box_j, box_h, box_w (source rectangle, inside the frame),
 *             RH, RW (size the rectangle is resized to), off_y, off_x (top-left of the output window in the resized image:
 *             off + OH <= RH, off + OW <= RW), frame_stride (elements between frames, >= H*W*3) }
 *   src_dtype 0 = uint8, 1 = fp32 (the reference converts to fp32 before transforming; values are the same 0..255)
 *   antialias 0 = plain bilinear, align_corners = False (torchvision 0.11.2 on tensors, the reference's pinned version);
 *             1 = antialiased bilinear (torchvision >= 0.17 default)
 *   mean, std HOST pointers to 3 floats (per channel, in 0..255 units as in main_pretrain.py:268-270)
 *   out       fp32 [clips][3][frames][OH][OW] = (resized - mean) / std, the layout SpaceTimeTransformer.forward takes
 * ---------------------------------------------------------------------------------------------- */
int lv_clip_transform(const int64_t* desc, int clips, int frames, int src_dtype, int antialias, const float* mean,
                      const float* std, float* out, int OH, int OW, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAVILA_B200_H */
