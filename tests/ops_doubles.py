"""TEST DOUBLES of the kernel wrappers in lavila_b200/ops.py, in plain fp32 torch on the CPU.

Test infrastructure only: `install()` swaps them in for `lavila_b200.ops.*` so that the HOST logic of the product -- the
autograd Functions of lavila_b200/engine.py (what is saved, how gradients are routed and accumulated, strides, the CLS-only
tail, recompute), the `lavila_b200.models` mirrors and the loss modules -- runs on a machine without a GPU and can be
compared with golden vectors of the unmodified reference.  Each double implements the contract written in
include/lavila_b200.h for its entry point (same arguments, same in-place outputs, same accumulate-vs-store behaviour, bf16
rounding where the kernel writes bf16); none of them is a fallback: nothing in the product imports this file.
"""
import torch

BF16, F32 = torch.bfloat16, torch.float32
SCALE = 0.125      # head_dim 64


def _mat(t, rows, cols, ld=None):
    """Logical row-major [rows, cols] window starting at t's first element, leading dimension ld (elements)."""
    if ld is None:
        ld = t.stride(0) if t.dim() >= 2 else cols
    return t.as_strided((rows, cols), (ld, 1), t.storage_offset())


# ----------------------------------------------------------------------------------------------------------------- GEMM
EPI = dict(BIAS=1, QUICKGELU=2, DQUICKGELU=4, SCALE=8, SCALE_TANH=16, RESID=32, OUT_F32=64, COPY_BF16=128, ATOMIC=256,
           ROWBIAS=512, GELU_TANH=1024, SQRELU=2048)


def gemm(A, B, M, N, K, out, *, a_mn=0, b_mn=0, flags=0, out2=None, bias=None, resid=None, aux=None, scale=None, k_splits=1,
         lda=None, ldb=None):
    Aop = _mat(A, M, K, lda) if not a_mn else _mat(A, K, M, lda).t()
    Bop = _mat(B, N, K, ldb) if not b_mn else _mat(B, K, N, ldb).t()
    v = Aop.float() @ Bop.float().t()
    if flags & EPI["ROWBIAS"]:
        v = v + bias.float()[:M, None]
    elif flags & EPI["BIAS"]:
        v = v + bias.float()[None, :N]
    if flags & EPI["QUICKGELU"]:
        h = v.to(BF16)
        _mat(out2, M, N).copy_(h)
        h = h.float()
        v = h * torch.sigmoid(1.702 * h)
    if flags & EPI["GELU_TANH"]:
        v = torch.nn.functional.gelu(v, approximate="tanh")
    if flags & EPI["SQRELU"]:
        v = torch.relu(v) ** 2
    if flags & EPI["DQUICKGELU"]:
        h = _mat(aux, M, N).float()
        s = torch.sigmoid(1.702 * h)
        v = v * (s * (1.0 + 1.702 * h * (1.0 - s)))
    if flags & EPI["SCALE"]:
        s = scale.detach().float().reshape(())
        v = v * (torch.tanh(s) if flags & EPI["SCALE_TANH"] else s)
    if flags & EPI["RESID"]:
        v = v + _mat(resid, M, N).float()
    o = _mat(out, M, N)
    if flags & EPI["ATOMIC"]:
        o += v
    else:
        o.copy_(v.to(out.dtype))
    if flags & EPI["COPY_BF16"]:
        _mat(out2, M, N).copy_(v.to(BF16))
    return out


def wgrad_splits(m_out, n_in, k_tokens, sms=148):
    return 1


# ------------------------------------------------------------------------------------------------------------ LayerNorm
def layernorm_fwd(x, gamma, beta, eps, rows, D, *, ldx=None, y_bf16=None, y_f32=None):
    xr = _mat(x, rows, D, ldx).float()
    y = torch.nn.functional.layer_norm(xr, (D,), gamma.float(), None if beta is None else beta.float(), eps)
    if y_bf16 is not None:
        _mat(y_bf16, rows, D, D).copy_(y.to(BF16))
    if y_f32 is not None:
        _mat(y_f32, rows, D, D).copy_(y)


def layernorm_bwd(dy, x, gamma, eps, rows, D, *, ldx=None, lddy=None, add1=None, add2=None, dx=None, lddx=None, dx_bf16=None,
                  dgamma=None, dbeta=None):
    xr = _mat(x, rows, D, ldx).float()
    g = _mat(dy, rows, D, lddy).float()
    mu = xr.mean(-1, keepdim=True)
    rstd = torch.rsqrt(xr.var(-1, unbiased=False, keepdim=True) + eps)
    xh = (xr - mu) * rstd
    gw = g * gamma.float()
    d = rstd * (gw - gw.mean(-1, keepdim=True) - xh * (gw * xh).mean(-1, keepdim=True))
    if add1 is not None:
        d = d + _mat(add1, rows, D, D).float()
    if add2 is not None:
        d = d + _mat(add2, rows, D, D).float()
    if dx is not None:
        _mat(dx, rows, D, lddx).copy_(d)
    if dx_bf16 is not None:
        _mat(dx_bf16, rows, D, D).copy_(d.to(BF16))
    if dgamma is not None:
        dgamma += (g * xh).sum(0)
        dbeta += g.sum(0)


# ------------------------------------------------------------------------------------------------------------ attention
def _group_rows(mode, B, T, n, Lctx):
    """Row indices [B, G, L] of the groups' own tokens and, per clip, the CLS row (or None)."""
    if mode == 2:
        rows = torch.arange(B * Lctx).view(B, 1, Lctx)
        return rows, None
    N = 1 + T * n
    base = (torch.arange(B) * N).view(B, 1, 1)
    f = torch.arange(T).view(1, T, 1)
    i = torch.arange(n).view(1, 1, n)
    rows = base + 1 + f * n + i                       # [B, T, n]: space groups = frames
    if mode == 1:
        rows = rows.transpose(1, 2).contiguous()      # [B, n, T]: time groups = spatial positions
    return rows, (torch.arange(B) * N)


def _heads(t, H):          # [..., H*64] -> [..., H, 64]
    return t.reshape(*t.shape[:-1], H, 64)


def _group_attention(q, k, v, kc, vc, causal):
    """q, k, v: [B, G, L, H, 64]; kc, vc: [B, H, 64] or None.  Returns (O [B, G, L, H, 64], lse [B, G, L, H])."""
    qh, kh, vh = (t.permute(0, 3, 1, 2, 4) for t in (q, k, v))           # [B, H, G, L, 64]
    if kc is not None:
        G = qh.shape[2]
        kh = torch.cat((kh, kc[:, :, None, None, :].expand(-1, -1, G, 1, -1)), 3)
        vh = torch.cat((vh, vc[:, :, None, None, :].expand(-1, -1, G, 1, -1)), 3)
    s = torch.einsum("bhgqd,bhgkd->bhgqk", qh, kh) * SCALE
    if causal:
        L = s.shape[-1]
        s = s.masked_fill(torch.ones(L, L, dtype=torch.bool).triu(1), float("-inf"))
    lse = torch.logsumexp(s, -1)
    o = torch.einsum("bhgqk,bhgkd->bhgqd", torch.softmax(s, -1), vh)
    return o.permute(0, 2, 3, 1, 4), lse.permute(0, 2, 3, 1)


def _split_qkv(qkv, rows, H):
    D = H * 64
    x = _mat(qkv, qkv.shape[0], 3 * D).float()[rows]                    # [B, G, L, 3D]
    return _heads(x[..., :D], H), _heads(x[..., D:2 * D], H), _heads(x[..., 2 * D:], H)


def group_attn_fwd(qkv, out, lse, mode, B, H, T=0, n=0, Lctx=0):
    D = H * 64
    rows, cls = _group_rows(mode, B, T, n, Lctx)
    q, k, v = _split_qkv(qkv, rows, H)
    kc = vc = None
    if cls is not None:
        c = _mat(qkv, qkv.shape[0], 3 * D).float()[cls]
        kc, vc = _heads(c[:, D:2 * D], H), _heads(c[:, 2 * D:], H)
    o, l = _group_attention(q, k, v, kc, vc, causal=(mode == 2))
    _mat(out, out.shape[0], D)[rows] = o.reshape(*rows.shape, D).to(out.dtype)
    lse[rows] = l


def group_attn_bwd(qkv, out, lse, dout, dqkv, dcls_kv, accumulate_kv, mode, B, H, T=0, n=0, Lctx=0):
    D = H * 64
    rows, cls = _group_rows(mode, B, T, n, Lctx)
    with torch.enable_grad():
        q, k, v = (t.detach().requires_grad_(True) for t in _split_qkv(qkv, rows, H))
        kc = vc = None
        if cls is not None:
            c = _mat(qkv, qkv.shape[0], 3 * D).float()[cls]
            kc = _heads(c[:, D:2 * D], H).detach().requires_grad_(True)
            vc = _heads(c[:, 2 * D:], H).detach().requires_grad_(True)
        o, _ = _group_attention(q, k, v, kc, vc, causal=(mode == 2))
        do = _heads(_mat(dout, dout.shape[0], D).float()[rows], H)
        inputs = (q, k, v) + ((kc, vc) if cls is not None else ())
        grads = torch.autograd.grad(o, inputs, do)
    dq, dk, dv = (g.reshape(*rows.shape, D) for g in grads[:3])
    d = _mat(dqkv, dqkv.shape[0], 3 * D)
    d[rows, :D] = dq.to(dqkv.dtype)
    if accumulate_kv:
        d[rows, D:2 * D] = (d[rows, D:2 * D].float() + dk).to(dqkv.dtype)
        d[rows, 2 * D:] = (d[rows, 2 * D:].float() + dv).to(dqkv.dtype)
    else:
        d[rows, D:2 * D] = dk.to(dqkv.dtype)
        d[rows, 2 * D:] = dv.to(dqkv.dtype)
    if cls is not None:
        dcls_kv[:, :, 0] += grads[3]
        dcls_kv[:, :, 1] += grads[4]


def _cls_attention(q, k, v):
    """q [B, H, 64]; k, v [B, N, H, 64] -> (o [B, H, 64], lse [B, H])."""
    s = torch.einsum("bhd,bjhd->bhj", q, k) * SCALE
    return torch.einsum("bhj,bjhd->bhd", torch.softmax(s, -1), v), torch.logsumexp(s, -1)


def cls_attn_fwd(qkv, out, lse, B, H, N):
    D = H * 64
    x = _mat(qkv, B * N, 3 * D).float().view(B, N, 3 * D)
    o, l = _cls_attention(_heads(x[:, 0, :D], H), _heads(x[:, :, D:2 * D], H), _heads(x[:, :, 2 * D:], H))
    _mat(out, out.shape[0], D)[torch.arange(B) * N] = o.reshape(B, D).to(out.dtype)
    lse[torch.arange(B) * N] = l


def cls_attn_bwd(qkv, out, dout, lse, dqkv, dcls_kv, B, H, N, accumulate=False):
    D = H * 64
    x = _mat(qkv, B * N, 3 * D).float().view(B, N, 3 * D)
    cls = torch.arange(B) * N
    with torch.enable_grad():
        q = _heads(x[:, 0, :D], H).detach().requires_grad_(True)
        k = _heads(x[:, :, D:2 * D], H).detach().requires_grad_(True)
        v = _heads(x[:, :, 2 * D:], H).detach().requires_grad_(True)
        o, _ = _cls_attention(q, k, v)
        do = _heads(_mat(dout, dout.shape[0], D).float()[cls], H)
        dq, dk, dv = torch.autograd.grad(o, (q, k, v), do)
    d = _mat(dqkv, B * N, 3 * D).view(B, N, 3 * D)
    d[:, 0, :D] = dq.reshape(B, D).to(dqkv.dtype)
    dk, dv = dk.reshape(B, N, D), dv.reshape(B, N, D)
    if accumulate:     # the group backward ran first: add onto its k/v gradients (patch rows) and onto dcls_kv (CLS key)
        d[:, 1:, D:2 * D] = (d[:, 1:, D:2 * D].float() + dk[:, 1:]).to(dqkv.dtype)
        d[:, 1:, 2 * D:] = (d[:, 1:, 2 * D:].float() + dv[:, 1:]).to(dqkv.dtype)
        dcls_kv[:, :, 0] += _heads(dk[:, 0], H)
        dcls_kv[:, :, 1] += _heads(dv[:, 0], H)
    else:
        d[:, 1:, D:2 * D] = dk[:, 1:].to(dqkv.dtype)
        d[:, 1:, 2 * D:] = dv[:, 1:].to(dqkv.dtype)
        dcls_kv[:, :, 0] = _heads(dk[:, 0], H)
        dcls_kv[:, :, 1] = _heads(dv[:, 0], H)


def cls_kv_finalize(dcls_kv, dqkv, B, H, N):
    D = H * 64
    d = _mat(dqkv, B * N, 3 * D).view(B, N, 3 * D)
    d[:, 0, D:2 * D] = dcls_kv[:, :, 0].reshape(B, D).to(dqkv.dtype)
    d[:, 0, 2 * D:] = dcls_kv[:, :, 1].reshape(B, D).to(dqkv.dtype)


def cls_query_attn_fwd(q, kv, out, lse, B, H, N):
    D = H * 64
    x = kv.float().view(B, N, 2 * D)
    o, l = _cls_attention(_heads(q.float(), H), _heads(x[..., :D], H), _heads(x[..., D:], H))
    out.copy_(o.reshape(B, D).to(out.dtype))
    lse.copy_(l)


def cls_query_attn_bwd(q, kv, out, dout, lse, dq, dkv, B, H, N):
    D = H * 64
    x = kv.float().view(B, N, 2 * D)
    with torch.enable_grad():
        qq = _heads(q.float(), H).detach().requires_grad_(True)
        k = _heads(x[..., :D], H).detach().requires_grad_(True)
        v = _heads(x[..., D:], H).detach().requires_grad_(True)
        o, _ = _cls_attention(qq, k, v)
        gq, gk, gv = torch.autograd.grad(o, (qq, k, v), _heads(dout.float(), H))
    dq.copy_(gq.reshape(B, D).to(dq.dtype))
    dkv.copy_(torch.cat((gk.reshape(B * N, D), gv.reshape(B * N, D)), 1).to(dkv.dtype))


def flash_attn_fwd(q, k, v, out, B, H, Lq, Lk, *, q_rows, kv_rows, ld_q, ld_kv, ld_out, kv_head_stride=64, causal=False,
                   scale=0.125):
    """Q element (b,h,i,d) at q[(b*q_rows+i)*ld_q + h*64 + d]; K/V (b,h,j,d) at k[(b*kv_rows+j)*ld_kv + h*kv_head_stride + d];
    causal: key j visible to query i iff j <= i + (Lk - Lq)."""
    qq = _mat(q, B * q_rows, H * 64, ld_q).float().view(B, q_rows, H, 64)[:, :Lq]
    width = 64 + (H - 1) * kv_head_stride
    kk = _mat(k, B * kv_rows, width, ld_kv).float().view(B, kv_rows, width)[:, :Lk]
    vv = _mat(v, B * kv_rows, width, ld_kv).float().view(B, kv_rows, width)[:, :Lk]
    heads = [slice(h * kv_head_stride, h * kv_head_stride + 64) for h in range(H)]
    kh = torch.stack([kk[..., sl] for sl in heads], 2)                    # [B, Lk, H, 64]
    vh = torch.stack([vv[..., sl] for sl in heads], 2)
    s = torch.einsum("bihd,bjhd->bhij", qq, kh) * scale
    if causal:
        i = torch.arange(Lq).view(Lq, 1)
        j = torch.arange(Lk).view(1, Lk)
        s = s.masked_fill(j > i + (Lk - Lq), float("-inf"))
    o = torch.einsum("bhij,bjhd->bihd", torch.softmax(s, -1), vh)
    _mat(out, B * q_rows, H * 64, ld_out).view(B, q_rows, H * 64)[:, :Lq] = o.reshape(B, Lq, H * 64).to(out.dtype)


def flash_attn_fwd_dyn(q, k, v, out, B, H, Lq, lk_dev, **kw):
    flash_attn_fwd(q, k, v, out, B, H, Lq, int(lk_dev.item()), **kw)


SKINNY_MAX_M = 512


def gemm_skinny(A, W, M, N, K, out, *, flags=0, bias=None, resid=None, scale=None):
    if M > SKINNY_MAX_M or N % 64 or K % 64:
        return False
    gemm(A, W, M, N, K, out, b_mn=1, flags=flags, bias=bias, resid=resid, scale=scale)
    return True


# ------------------------------------------------------------------------------------------------------------------ glue
def add_rows(dst, stride, src, R, W):
    d = _mat(dst, R, W, stride)
    d.copy_((d.float() + src.float().view(R, W)).to(dst.dtype))


def cast_bf16(x, out=None):
    y = x.contiguous().to(BF16)
    if out is None:
        return y
    out.copy_(y.view(out.shape))
    return out


def colsum_bf16(x, M, N, out):
    out += _mat(x, M, N).float().sum(0)
    return out


def patch_im2col(frames, patches, B, C, T, H, W, P):
    f = frames.float().view(B, C, T, H // P, P, W // P, P).permute(0, 2, 3, 5, 1, 4, 6)      # b t hy wx c py px
    _mat(patches, B * T * (H // P) * (W // P), C * P * P).copy_(f.reshape(-1, C * P * P).to(patches.dtype))


def embed_assemble(patch, cls, pos, temporal, x0, B, T, n, D):
    x = x0.view(B, 1 + T * n, D)
    x[:, 0] = cls.view(1, D) + pos.view(-1, D)[0]
    p = patch.float().view(B, T, n, D) + pos.view(-1, D)[1:1 + n].view(1, 1, n, D) + temporal.view(-1, D)[:T].view(1, T, 1, D)
    x[:, 1:] = p.view(B, T * n, D)


def embed_assemble_bwd(dx0, dpos, dcls, dtemporal, dpatch, B, T, n, D):
    g = dx0.float().view(B, 1 + T * n, D)
    dcls.view(-1)[:D] += g[:, 0].sum(0)
    dpos.view(-1, D)[0] += g[:, 0].sum(0)
    gp = g[:, 1:].view(B, T, n, D)
    dpos.view(-1, D)[1:1 + n] += gp.sum((0, 1))
    dtemporal.view(-1, D)[:T] += gp.sum((0, 2))
    dpatch.copy_(gp.reshape(B * T * n, D).to(dpatch.dtype))


def text_embed(text, tok, pos, x, rows, Lctx, W, vocab):
    ids = text.reshape(-1)[:rows]
    x.view(rows, W).copy_(tok[ids] + pos[:Lctx].repeat(rows // Lctx, 1))


def text_embed_bwd(text, dx, dtok, dpos, rows, Lctx, W, vocab):
    ids = text.reshape(-1)[:rows]
    g = dx.float().view(rows, W)
    dtok.index_add_(0, ids, g)
    dpos[:Lctx] += g.view(rows // Lctx, Lctx, W).sum(0)


def argmax_i64(text, out, B, Lctx):
    out.copy_(text.view(B, Lctx).argmax(-1).to(out.dtype))       # torch.argmax returns the first maximum, like the kernel


def gather_rows(src, idx, dst, R, rows_per, W, scatter=False):
    rows = torch.arange(R) * rows_per + idx.long()
    if scatter:
        dst.view(-1, W)[rows] = src.view(R, W)
    else:
        dst.view(R, W).copy_(src.view(-1, W)[rows])


def l2norm_fwd(x, y, norm, R, E):
    nrm = x.view(R, E).norm(dim=-1).clamp_min(1e-12)
    norm.copy_(nrm)
    y.view(R, E).copy_(x.view(R, E) / nrm[:, None])


def l2norm_bwd(dy, y, norm, dx, R, E):
    g, yy = dy.view(R, E), y.view(R, E)
    dx.view(R, E).copy_((g - yy * (g * yy).sum(-1, keepdim=True)) / norm[:, None])


def clip_loss_fwd(img, txt, scale, Ng, E, lse_i, lse_t, partial, counter, result):
    logits = scale * img @ txt.t()
    lab = torch.arange(Ng)
    lse_i.copy_(torch.logsumexp(logits, 1))
    lse_t.copy_(torch.logsumexp(logits.t(), 1))
    result[0] = (torch.nn.functional.cross_entropy(logits, lab) + torch.nn.functional.cross_entropy(logits.t(), lab)) / 2
    result[1] = 100.0 * (logits.argmax(-1) == lab).float().mean()


def clip_loss_bwd(img, txt, scale, lse_i, lse_t, gout, grad_scale, scale_grad_scale, Ng, E, r0, Nl, d_img, d_txt, d_scale):
    with torch.enable_grad():
        i, t, s = img.clone().requires_grad_(True), txt.clone().requires_grad_(True), scale.clone().requires_grad_(True)
        raw = i @ t.t()
        logits = s * raw
        lab = torch.arange(Ng)
        loss = (torch.nn.functional.cross_entropy(logits, lab) + torch.nn.functional.cross_entropy(logits.t(), lab)) / 2
        gi, gt, gl = torch.autograd.grad(loss, (i, t, logits))
    d_img.copy_(gout * grad_scale * gi[r0:r0 + Nl])
    d_txt.copy_(gout * grad_scale * gt[r0:r0 + Nl])
    if d_scale is not None:      # header contract: the double sum over the LOCAL image rows x all texts
        d_scale += gout * scale_grad_scale * (gl[r0:r0 + Nl] * raw.detach()[r0:r0 + Nl]).sum()


def _ssl_logits(img, txt, s, sp, gt):
    m = gt[:, None] + gt[None, :]
    c = torch.where(m == 2, s, torch.where(m == 1, torch.sqrt(sp * s), sp))
    return c * (img @ txt.t())


def ssl_clip_loss_fwd(img, txt, scale, scale_p, gt, Ng, E, lse_i, lse_t, partial, counter, result):
    logits = _ssl_logits(img, txt, scale, scale_p, gt)
    lab = torch.arange(Ng)
    lse_i.copy_(torch.logsumexp(logits, 1))
    lse_t.copy_(torch.logsumexp(logits.t(), 1))
    ok = (logits.argmax(-1) == lab).float()
    g = (gt == 1).float()
    result[0] = (torch.nn.functional.cross_entropy(logits, lab) + torch.nn.functional.cross_entropy(logits.t(), lab)) / 2
    result[1] = 100 * ok.mean()
    result[2] = 100 * (ok * g).sum() / g.sum()
    result[3] = 100 * (ok * (1 - g)).sum() / (1 - g).sum()
    result[4], result[5] = g.sum(), (1 - g).sum()


def ssl_clip_loss_bwd(img, txt, scale, scale_p, gt, lse_i, lse_t, gout, grad_scale, scale_grad_scale, Ng, E, r0, Nl, d_img, d_txt,
                      d_scales):
    with torch.enable_grad():
        i, t = img.clone().requires_grad_(True), txt.clone().requires_grad_(True)
        s, sp = scale.clone().requires_grad_(True), scale_p.clone().requires_grad_(True)
        logits = _ssl_logits(i, t, s, sp, gt)
        lab = torch.arange(Ng)
        loss = (torch.nn.functional.cross_entropy(logits, lab) + torch.nn.functional.cross_entropy(logits.t(), lab)) / 2
        gi, gtx, gs, gp = torch.autograd.grad(loss, (i, t, s, sp), allow_unused=True)
    d_img.copy_(gout * grad_scale * gi[r0:r0 + Nl])
    d_txt.copy_(gout * grad_scale * gtx[r0:r0 + Nl])
    if gs is not None:
        d_scales[0] += (gout * scale_grad_scale * gs).reshape(())
    if gp is not None:
        d_scales[1] += (gout * scale_grad_scale * gp).reshape(())


DOUBLES = ("gemm", "wgrad_splits", "layernorm_fwd", "layernorm_bwd", "group_attn_fwd", "group_attn_bwd", "cls_attn_fwd",
           "cls_attn_bwd", "cls_kv_finalize", "cls_query_attn_fwd", "cls_query_attn_bwd", "add_rows", "cast_bf16", "colsum_bf16",
           "patch_im2col", "embed_assemble", "embed_assemble_bwd", "text_embed", "text_embed_bwd", "argmax_i64", "gather_rows",
           "l2norm_fwd", "l2norm_bwd", "clip_loss_fwd", "clip_loss_bwd", "flash_attn_fwd", "flash_attn_fwd_dyn", "gemm_skinny",
           "ssl_clip_loss_fwd", "ssl_clip_loss_bwd", "top_p_filter_", "clip_transform",
           "space_attn_cls_fused_supported", "space_attn_fwd_cls", "space_attn_bwd_cls",
           "time_attn_cls_fused_supported", "time_attn_fwd_cls", "time_attn_bwd_cls")


def top_p_filter_(logits, temperature, top_p):
    """TemperatureLogitsWarper + TopPLogitsWarper(min_tokens_to_keep=1), the transformers algorithm."""
    x = logits / temperature if temperature != 1.0 else logits.clone()
    srt, idx = torch.sort(x, descending=False, dim=-1)
    rm = srt.softmax(dim=-1).cumsum(dim=-1) <= (1 - top_p)
    rm[..., -1:] = False
    logits.copy_(x.masked_fill(rm.scatter(1, idx, rm), float("-inf")))
    return logits


def clip_transform(desc, sources, frames, antialias, mean, std, out):
    """lv_clip_transform's contract (include/lavila_b200.h) row by row of the descriptor table: source rectangle -> bilinear resize
    to RH x RW (taps clamped to the rectangle) -> OH x OW window at (off_y, off_x) -> (v - mean) / std, channels first."""
    OH, OW = out.shape[-2:]
    m = torch.tensor(mean, dtype=torch.float32).view(3, 1, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(3, 1, 1, 1)
    for b, src in enumerate(sources):
        _, H, W, i, j, h, w, RH, RW, oy, ox, _ = (int(v) for v in desc[b].tolist())
        box = src[:frames, i:i + h, j:j + w, :].permute(0, 3, 1, 2).float()                     # T x 3 x h x w
        r = torch.nn.functional.interpolate(box, size=(RH, RW), mode="bilinear", align_corners=False, antialias=bool(antialias))
        out[b].copy_(((r[:, :, oy:oy + OH, ox:ox + OW].permute(1, 0, 2, 3) - m) / s).to(out.device))
    return out


def space_attn_cls_fused_supported(n):
    return 128 < n <= 207


def space_attn_fwd_cls(qkv, out, lse, B, H, T, n):
    """Contract of lv_space_attn_fwd_tc_cls: the group pass and the CLS-query pass of the space attention, every row written."""
    group_attn_fwd(qkv, out, lse, 0, B, H, T=T, n=n)
    cls_attn_fwd(qkv, out, lse, B, H, 1 + T * n)


def space_attn_bwd_cls(qkv, out, lse, dout, dqkv, B, H, T, n):
    dcls = torch.zeros(B, H, 2, 64)
    group_attn_bwd(qkv, out, lse, dout, dqkv, dcls, 0, 0, B, H, T=T, n=n)
    cls_attn_bwd(qkv, out, dout, lse, dqkv, dcls, B, H, 1 + T * n, accumulate=True)
    cls_kv_finalize(dcls, dqkv, B, H, 1 + T * n)


def time_attn_cls_fused_supported(T):
    return 0 < T <= 16


def time_attn_fwd_cls(qkv, out, lse, B, H, T, n):
    group_attn_fwd(qkv, out, lse, 1, B, H, T=T, n=n)
    cls_attn_fwd(qkv, out, lse, B, H, 1 + T * n)


def time_attn_bwd_cls(qkv, out, lse, dout, dqkv, B, H, T, n):
    dcls = torch.zeros(B, H, 2, 64)
    group_attn_bwd(qkv, out, lse, dout, dqkv, dcls, 0, 1, B, H, T=T, n=n)
    cls_attn_bwd(qkv, out, dout, lse, dqkv, dcls, B, H, 1 + T * n, accumulate=True)
    cls_kv_finalize(dcls, dqkv, B, H, 1 + T * n)


def install(monkeypatch):
    """Swap the doubles in for lavila_b200.ops (pytest's monkeypatch restores the real wrappers afterwards)."""
    from lavila_b200 import engine, ops
    for name in DOUBLES:
        monkeypatch.setattr(ops, name, globals()[name])
    engine.SHADOW.clear()
    engine._GRAD_BF16.clear()
