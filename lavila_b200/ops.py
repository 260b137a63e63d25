"""Thin tensor-level wrappers over the C ABI (include/lavila_b200.h).  PyTorch supplies device memory and the
stream; every FLOP/byte of the hot path is moved by liblavila_b200.so.  No fallbacks: a missing library or a
failed launch raises LavilaB200Error."""
import ctypes

import torch

from . import _lib as L

BF16, F32 = torch.bfloat16, torch.float32


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.LavilaB200Error("lavila_b200 kernels need CUDA tensors (there is no CPU path)")


def gemm(A, B, M, N, K, out, *, a_mn=0, b_mn=0, flags=0, out2=None, bias=None, resid=None, aux=None, scale=None,
         k_splits=1, lda=None, ldb=None):
    """out[M,N] = epilogue(A_op[M,K] @ B_op[N,K]^T); see lv_gemm_bf16 for the operand layouts."""
    _check_cuda(A, B, out)
    e = L.LvGemmEpilogue()
    if out.dtype == F32:
        flags |= L.EPI_OUT_F32
    e.flags = flags
    e.out, e.ldo = out.data_ptr(), out.stride(0)
    if out2 is not None:
        e.out2, e.ldo2 = out2.data_ptr(), out2.stride(0)
    if bias is not None:
        e.bias = bias.data_ptr()
    if resid is not None:
        e.resid, e.ldr = resid.data_ptr(), resid.stride(0)
    if aux is not None:
        e.aux, e.ldaux = aux.data_ptr(), aux.stride(0)
    if scale is not None:
        e.scale_ptr = scale.data_ptr()
    fn = L.lib().lv_gemm_bf16_2cta if (USE_2CTA_GEMM and M >= 256) else L.lib().lv_gemm_bf16
    rc = fn(A.data_ptr(), lda if lda is not None else A.stride(0), a_mn, B.data_ptr(),
            ldb if ldb is not None else B.stride(0), b_mn, M, N, K, k_splits, ctypes.byref(e), _stream())
    L.check(rc, "lv_gemm_bf16")
    return out


SKINNY_MAX_M = 512      # decode-sized GEMMs (rows = sequences) stream W from every SM instead of N/256 CTAs


def gemm_skinny(A, W, M, N, K, out, *, flags=0, bias=None, resid=None, scale=None):
    """out[M,N] = epilogue(A[M,K] @ W[K,N]) for small M (lv_gemm_skinny_bf16); returns False if the shape is not supported
    (the caller then uses the tiled GEMM).  W is the [K, N] matrix itself (HF Conv1D weight), bf16."""
    if M > SKINNY_MAX_M:
        return False
    splits = L.lib().lv_gemm_skinny_splits(M, N, K)
    if splits <= 0:
        return False
    _check_cuda(A, W, out)
    e = L.LvGemmEpilogue()
    if out.dtype == F32:
        flags |= L.EPI_OUT_F32
    e.flags = flags
    e.out, e.ldo = out.data_ptr(), out.stride(0)
    if bias is not None:
        e.bias = bias.data_ptr()
    if resid is not None:
        e.resid, e.ldr = resid.data_ptr(), resid.stride(0)
    if scale is not None:
        e.scale_ptr = scale.data_ptr()
    ws = torch.empty(splits * M * N, device=A.device, dtype=F32)
    rc = L.lib().lv_gemm_skinny_bf16(A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), M, N, K, ws.data_ptr(), splits,
                                     ctypes.byref(e), _stream())
    L.check(rc, "lv_gemm_skinny_bf16")
    return True


USE_2CTA_GEMM = True    # 256x256 CTA-pair tiles (lv_gemm_bf16_2cta) whenever M >= 256


def wgrad_splits(m_out, n_in, k_tokens, sms=148):
    """Split count of a weight-gradient GEMM (dW[out,in] = dY^T X, K = tokens).  Items (tile, split) are spread round-robin
    over the CTA pairs, all equally long, so the makespan is ceil(items / pairs) item times: pick the split count in
    [2 waves, 4.5 waves] whose items fill whole waves best (11 splits of the 27-tile qkv gradient = 297 items on 74 pairs ran
    5 waves for 4.01 waves of work; 8 splits = 216 items = 2.92 waves run 3).  Items of one split are consecutive, so the
    pairs of a wave share operand panels in L2."""
    bm = 256 if USE_2CTA_GEMM and m_out >= 256 else 128
    units = sms // 2 if bm == 256 else sms
    tiles = ((m_out + bm - 1) // bm) * ((n_in + 255) // 256)
    kb = (k_tokens + 63) // 64
    lo = max(1, (2 * units) // tiles)
    hi = max(lo, min(kb, (9 * units) // (2 * tiles)))
    best, best_eff = lo, -1.0
    for s in range(lo, hi + 1):
        items = tiles * s
        eff = items / (units * ((items + units - 1) // units))
        if eff > best_eff + 1e-9:
            best, best_eff = s, eff
    return max(1, min(kb, best))


def layernorm_fwd(x, gamma, beta, eps, rows, D, *, ldx=None, y_bf16=None, y_f32=None):
    rc = L.lib().lv_layernorm_fwd(x.data_ptr(), ldx if ldx is not None else D, gamma.data_ptr(), _p(beta), eps,
                                  _p(y_bf16), D, _p(y_f32), D, rows, D, _stream())
    L.check(rc, "lv_layernorm_fwd")


def layernorm_bwd(dy, x, gamma, eps, rows, D, *, ldx=None, lddy=None, add1=None, add2=None, dx=None, lddx=None,
                  dx_bf16=None, dgamma=None, dbeta=None):
    rc = L.lib().lv_layernorm_bwd(dy.data_ptr(), 1 if dy.dtype == BF16 else 0, lddy if lddy is not None else D,
                                  x.data_ptr(), ldx if ldx is not None else D, gamma.data_ptr(), eps, _p(add1), D,
                                  _p(add2), D, _p(dx), lddx if lddx is not None else D, _p(dx_bf16), D, _p(dgamma),
                                  _p(dbeta), rows, D, _stream())
    L.check(rc, "lv_layernorm_bwd")


USE_TC_ATTN_FWD = True   # tcgen05 space attention (TSF-B geometry); the mma.sync kernels cover every other shape
USE_TC_ATTN_BWD = True


def space_attn_tc_supported(n):
    return 128 < n <= 207


def group_attn_fwd(qkv, out, lse, mode, B, H, T=0, n=0, Lctx=0):
    if mode == 0 and space_attn_tc_supported(n) and USE_TC_ATTN_FWD:   # tcgen05 path for TSF-B geometry
        rc = L.lib().lv_space_attn_fwd_tc(qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0), lse.data_ptr(), B,
                                          H, T, n, _stream())
        L.check(rc, "lv_space_attn_fwd_tc")
        return
    rc = L.lib().lv_group_attn_fwd(qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0), lse.data_ptr(), mode,
                                   B, H, T, n, Lctx, _stream())
    L.check(rc, "lv_group_attn_fwd")


def group_attn_bwd(qkv, out, lse, dout, dqkv, dcls_kv, accumulate_kv, mode, B, H, T=0, n=0, Lctx=0):
    if mode == 0 and space_attn_tc_supported(n) and USE_TC_ATTN_BWD:
        rc = L.lib().lv_space_attn_bwd_tc(qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0), lse.data_ptr(),
                                          dout.data_ptr(), dout.stride(0), dqkv.data_ptr(), dqkv.stride(0),
                                          dcls_kv.data_ptr(), accumulate_kv, B, H, T, n, _stream())
        L.check(rc, "lv_space_attn_bwd_tc")
        return
    rc = L.lib().lv_group_attn_bwd(qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0), lse.data_ptr(),
                                   dout.data_ptr(), dout.stride(0), dqkv.data_ptr(), dqkv.stride(0), _p(dcls_kv),
                                   accumulate_kv, mode, B, H, T, n, Lctx, _stream())
    L.check(rc, "lv_group_attn_bwd")


import os as _os
USE_TC_CLS_FUSION = _os.environ.get("LAVILA_B200_CLS_FUSION", "1") == "1"   # space attention: the CLS query rides inside the tcgen05 group kernels


def space_attn_cls_fused_supported(n):
    return USE_TC_CLS_FUSION and USE_TC_ATTN_FWD and USE_TC_ATTN_BWD and space_attn_tc_supported(n)


def space_attn_fwd_cls(qkv, out, lse, B, H, T, n):
    """Space attention forward for every row, CLS rows included (lv_space_attn_fwd_tc_cls)."""
    part = torch.empty(B * H * T * 66, device=qkv.device, dtype=F32)
    rc = L.lib().lv_space_attn_fwd_tc_cls(qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0), lse.data_ptr(),
                                          part.data_ptr(), B, H, T, n, _stream())
    L.check(rc, "lv_space_attn_fwd_tc_cls")


def space_attn_bwd_cls(qkv, out, lse, dout, dqkv, B, H, T, n):
    """Space attention backward for every row, CLS rows included (lv_space_attn_bwd_tc_cls)."""
    scratch = torch.zeros(B * H * 3 * 64, device=qkv.device, dtype=F32)
    dkv, dq = scratch[:B * H * 128], scratch[B * H * 128:]
    rc = L.lib().lv_space_attn_bwd_tc_cls(qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0), lse.data_ptr(),
                                          dout.data_ptr(), dout.stride(0), dqkv.data_ptr(), dqkv.stride(0), dkv.data_ptr(),
                                          dq.data_ptr(), B, H, T, n, _stream())
    L.check(rc, "lv_space_attn_bwd_tc_cls")


def time_attn_cls_fused_supported(T):
    return USE_TC_CLS_FUSION and 0 < T <= 16


def time_attn_fwd_cls(qkv, out, lse, B, H, T, n):
    """Time attention forward for every row, CLS rows included (lv_time_attn_fwd_cls)."""
    part = torch.empty(B * H * n * 66, device=qkv.device, dtype=F32)
    rc = L.lib().lv_time_attn_fwd_cls(qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0), lse.data_ptr(),
                                      part.data_ptr(), B, H, T, n, _stream())
    L.check(rc, "lv_time_attn_fwd_cls")


def time_attn_bwd_cls(qkv, out, lse, dout, dqkv, B, H, T, n):
    """Time attention backward for every row, CLS rows included (lv_time_attn_bwd_cls)."""
    scratch = torch.zeros(B * H * 3 * 64, device=qkv.device, dtype=F32)
    dkv, dq = scratch[:B * H * 128], scratch[B * H * 128:]
    rc = L.lib().lv_time_attn_bwd_cls(qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0), lse.data_ptr(),
                                      dout.data_ptr(), dout.stride(0), dqkv.data_ptr(), dqkv.stride(0), dkv.data_ptr(),
                                      dq.data_ptr(), B, H, T, n, _stream())
    L.check(rc, "lv_time_attn_bwd_cls")


def flash_attn_fwd(q, k, v, out, B, H, Lq, Lk, *, q_rows, kv_rows, ld_q, ld_kv, ld_out, kv_head_stride=64, causal=False,
                   scale=0.125):
    rc = L.lib().lv_flash_attn_fwd(q.data_ptr(), ld_q, q_rows, k.data_ptr(), v.data_ptr(), ld_kv, kv_rows, kv_head_stride,
                                   out.data_ptr(), ld_out, B, H, Lq, Lk, int(causal), float(scale), _stream())
    L.check(rc, "lv_flash_attn_fwd")


def flash_attn_fwd_dyn(q, k, v, out, B, H, Lq, lk_dev, *, q_rows, kv_rows, ld_q, ld_kv, ld_out, kv_head_stride=64, causal=False,
                       scale=0.125):
    """lv_flash_attn_fwd with the key count in device memory (int32 tensor lk_dev) -- CUDA-graph friendly."""
    rc = L.lib().lv_flash_attn_fwd_dyn(q.data_ptr(), ld_q, q_rows, k.data_ptr(), v.data_ptr(), ld_kv, kv_rows, kv_head_stride,
                                       out.data_ptr(), ld_out, B, H, Lq, lk_dev.data_ptr(), int(causal), float(scale), _stream())
    L.check(rc, "lv_flash_attn_fwd_dyn")


def cls_attn_fwd(qkv, out, lse, B, H, N):
    rc = L.lib().lv_cls_attn_fwd(qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0), lse.data_ptr(), B, H, N,
                                 _stream())
    L.check(rc, "lv_cls_attn_fwd")


def cls_attn_bwd(qkv, out, dout, lse, dqkv, dcls_kv, B, H, N, accumulate=False):
    rc = L.lib().lv_cls_attn_bwd(qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0), dout.data_ptr(),
                                 dout.stride(0), lse.data_ptr(), dqkv.data_ptr(), dqkv.stride(0), dcls_kv.data_ptr(),
                                 int(accumulate), B, H, N, _stream())
    L.check(rc, "lv_cls_attn_bwd")


def cls_query_attn_fwd(q, kv, out, lse, B, H, N):
    rc = L.lib().lv_cls_query_attn_fwd(q.data_ptr(), kv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, H, N, _stream())
    L.check(rc, "lv_cls_query_attn_fwd")


def cls_query_attn_bwd(q, kv, out, dout, lse, dq, dkv, B, H, N):
    rc = L.lib().lv_cls_query_attn_bwd(q.data_ptr(), kv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(),
                                       dq.data_ptr(), dkv.data_ptr(), B, H, N, _stream())
    L.check(rc, "lv_cls_query_attn_bwd")


def add_rows(dst, stride, src, R, W):
    rc = L.lib().lv_add_rows(dst.data_ptr(), 1 if dst.dtype == BF16 else 0, stride, src.data_ptr(), R, W, _stream())
    L.check(rc, "lv_add_rows")


def cls_kv_finalize(dcls_kv, dqkv, B, H, N):
    rc = L.lib().lv_cls_kv_finalize(dcls_kv.data_ptr(), dqkv.data_ptr(), dqkv.stride(0), B, H, N, _stream())
    L.check(rc, "lv_cls_kv_finalize")


def cast_bf16(x, out=None):
    x = x.contiguous()
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=BF16)
    rc = L.lib().lv_cast_f32_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream())
    L.check(rc, "lv_cast_f32_bf16")
    return out


def colsum_bf16(x, M, N, out):
    rc = L.lib().lv_colsum_bf16(x.data_ptr(), x.stride(0), M, N, out.data_ptr(), _stream())
    L.check(rc, "lv_colsum_bf16")
    return out


def patch_im2col(frames, patches, B, C, T, H, W, P):
    rc = L.lib().lv_patch_im2col(frames.data_ptr(), patches.data_ptr(), B, C, T, H, W, P, patches.stride(0), _stream())
    L.check(rc, "lv_patch_im2col")


def embed_assemble(patch, cls, pos, temporal, x0, B, T, n, D):
    rc = L.lib().lv_embed_assemble(patch.data_ptr(), cls.data_ptr(), pos.data_ptr(), temporal.data_ptr(),
                                   x0.data_ptr(), B, T, n, D, _stream())
    L.check(rc, "lv_embed_assemble")


def embed_assemble_bwd(dx0, dpos, dcls, dtemporal, dpatch, B, T, n, D):
    rc = L.lib().lv_embed_assemble_bwd(dx0.data_ptr(), dpos.data_ptr(), dcls.data_ptr(), dtemporal.data_ptr(),
                                       dpatch.data_ptr(), B, T, n, D, _stream())
    L.check(rc, "lv_embed_assemble_bwd")


def text_embed(text, tok, pos, x, rows, Lctx, W, vocab):
    rc = L.lib().lv_text_embed(text.data_ptr(), tok.data_ptr(), pos.data_ptr(), x.data_ptr(), rows, Lctx, W, vocab,
                               _stream())
    L.check(rc, "lv_text_embed")


def text_embed_bwd(text, dx, dtok, dpos, rows, Lctx, W, vocab):
    rc = L.lib().lv_text_embed_bwd(text.data_ptr(), dx.data_ptr(), dtok.data_ptr(), dpos.data_ptr(), rows, Lctx, W,
                                   vocab, _stream())
    L.check(rc, "lv_text_embed_bwd")


def argmax_i64(text, out, B, Lctx):
    rc = L.lib().lv_argmax_i64(text.data_ptr(), out.data_ptr(), B, Lctx, _stream())
    L.check(rc, "lv_argmax_i64")


def gather_rows(src, idx, dst, R, rows_per, W, scatter=False):
    rc = L.lib().lv_gather_rows_f32(src.data_ptr(), idx.data_ptr(), dst.data_ptr(), R, rows_per, W, int(scatter),
                                    _stream())
    L.check(rc, "lv_gather_rows_f32")


def l2norm_fwd(x, y, norm, R, E):
    rc = L.lib().lv_l2norm_fwd(x.data_ptr(), y.data_ptr(), norm.data_ptr(), R, E, _stream())
    L.check(rc, "lv_l2norm_fwd")


def l2norm_bwd(dy, y, norm, dx, R, E):
    rc = L.lib().lv_l2norm_bwd(dy.data_ptr(), y.data_ptr(), norm.data_ptr(), dx.data_ptr(), R, E, _stream())
    L.check(rc, "lv_l2norm_bwd")


def clip_loss_fwd(img, txt, scale, Ng, E, lse_img, lse_txt, partial, counter, result):
    rc = L.lib().lv_clip_loss_fwd(img.data_ptr(), txt.data_ptr(), scale.data_ptr(), Ng, E, lse_img.data_ptr(),
                                  lse_txt.data_ptr(), partial.data_ptr(), counter.data_ptr(), result.data_ptr(),
                                  _stream())
    L.check(rc, "lv_clip_loss_fwd")


def clip_loss_bwd(img, txt, scale, lse_img, lse_txt, gout, grad_scale, scale_grad_scale, Ng, E, r0, Nl, d_img, d_txt,
                  d_scale):
    rc = L.lib().lv_clip_loss_bwd(img.data_ptr(), txt.data_ptr(), scale.data_ptr(), lse_img.data_ptr(),
                                  lse_txt.data_ptr(), gout.data_ptr(), float(grad_scale), float(scale_grad_scale), Ng, E,
                                  r0, Nl, d_img.data_ptr(), d_txt.data_ptr(), _p(d_scale), _stream())
    L.check(rc, "lv_clip_loss_bwd")


def ssl_clip_loss_fwd(img, txt, scale, scale_pseudo, gt, Ng, E, lse_img, lse_txt, partial, counter, result):
    rc = L.lib().lv_ssl_clip_loss_fwd(img.data_ptr(), txt.data_ptr(), scale.data_ptr(), scale_pseudo.data_ptr(),
                                      gt.data_ptr(), Ng, E, lse_img.data_ptr(), lse_txt.data_ptr(), partial.data_ptr(),
                                      counter.data_ptr(), result.data_ptr(), _stream())
    L.check(rc, "lv_ssl_clip_loss_fwd")


def ssl_clip_loss_bwd(img, txt, scale, scale_pseudo, gt, lse_img, lse_txt, gout, grad_scale, scale_grad_scale, Ng, E, r0,
                      Nl, d_img, d_txt, d_scales):
    rc = L.lib().lv_ssl_clip_loss_bwd(img.data_ptr(), txt.data_ptr(), scale.data_ptr(), scale_pseudo.data_ptr(),
                                      gt.data_ptr(), lse_img.data_ptr(), lse_txt.data_ptr(), gout.data_ptr(),
                                      float(grad_scale), float(scale_grad_scale), Ng, E, r0, Nl, d_img.data_ptr(),
                                      d_txt.data_ptr(), _p(d_scales), _stream())
    L.check(rc, "lv_ssl_clip_loss_bwd")


def clip_loss_fwd_gather(img_local, txt_local, peers_dev_ptr, rank, W, Bl, step, all_img, all_txt, scale, E, lse_img, lse_txt,
                         partial, ctrl, result, timeout_ms=600000):
    """Fused NVLink gather + loss forward (lv_clip_loss_fwd_gather).  peers_dev_ptr: int, device address of the W-pointer array."""
    rc = L.lib().lv_clip_loss_fwd_gather(img_local.data_ptr(), txt_local.data_ptr(), int(peers_dev_ptr), rank, W, Bl, int(step),
                                         all_img.data_ptr(), all_txt.data_ptr(), scale.data_ptr(), E, lse_img.data_ptr(),
                                         lse_txt.data_ptr(), partial.data_ptr(), ctrl.data_ptr(), result.data_ptr(),
                                         int(timeout_ms), _stream())
    L.check(rc, "lv_clip_loss_fwd_gather")


def clip_loss_gather_max_rows(E):
    """Largest global batch W*B the fused gather + loss kernel supports on the current device (0: not supported)."""
    return int(L.lib().lv_clip_loss_gather_max_rows(int(E)))


def top_p_filter_(logits, temperature, top_p):
    """In place: logits <- logits / temperature, -inf where nucleus filtering (top_p, min_tokens_to_keep = 1) removes the token."""
    _check_cuda(logits)
    rows, V = logits.shape
    rc = L.lib().lv_top_p_filter(logits.data_ptr(), logits.stride(0), rows, V, float(temperature), float(top_p), _stream())
    L.check(rc, "lv_top_p_filter")
    return logits


def clip_transform(desc, sources, frames, antialias, mean, std, out):
    """out[B,3,T,OH,OW] = normalise(resize(crop(frames))) for a batch of clips (lv_clip_transform); `desc` is the device table of
    12 int64 per clip, `sources` the frame tensors it points at (kept alive by the caller; only their dtype is read here)."""
    _check_cuda(desc, out, *sources)
    if out.dtype != F32 or not out.is_contiguous() or desc.dtype != torch.int64 or not desc.is_contiguous():
        raise L.LavilaB200Error("clip_transform: out must be contiguous fp32, desc contiguous int64")
    B, _, T, OH, OW = out.shape
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    rc = L.lib().lv_clip_transform(desc.data_ptr(), B, T, 0 if sources[0].dtype == torch.uint8 else 1, int(bool(antialias)), m, s,
                                   out.data_ptr(), OH, OW, _stream())
    L.check(rc, "lv_clip_transform")
    return out
