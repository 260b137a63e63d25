"""Forward/backward schedules of the dual-encoder hot path, expressed as sequences of C-ABI kernel launches.

Numerics contract: fp32 residual stream and LayerNorm statistics, bf16 GEMM operands with fp32
accumulation (TMEM), fp32 softmax, bf16 saved activations.  The reference under CUDA autocast makes the same choices
op by op; here they are fused: LN writes the bf16 GEMM operand, GEMM epilogues add bias / QuickGELU / gate / the fp32
residual, attention reads the packed qkv in place.

Two reusable sub-blocks cover both towers:
    attn_sub : y = resid + gate * Proj(Attention(QKV(LN(x_ln))))      (time / space / causal)
    mlp_sub  : y = r + FC2(QuickGELU(FC1(LN(r))))
"""
import contextlib
import os

import torch

from . import _lib as L
from . import ops

BF16, F32 = torch.bfloat16, torch.float32
MODE_SPACE, MODE_TIME, MODE_CAUSAL = 0, 1, 2


# ----------------------------------------------------------------------------------------------- bf16 weight shadows
# Derived copies of parameters (bf16 GEMM operands, padded LM head, captured decode graphs) are keyed on
# (parameter identity, `_version`, PARAM_GENERATION).  `_version` alone is not enough: writes through `param.data`
# (ZeroRedundancyOptimizer broadcasting the non-owned shards, `main_pretrain.py`'s `logit_scale.data.clamp_`, manual
# `p.data.copy_()`) do not bump it.  Every optimizer step bumps the generation through a global post-step hook -- the hook
# of an outer optimizer such as ZeRO fires after its parameter broadcast -- and `invalidate_param_caches()` is the manual
# switch for any other out-of-band write.  Cost: one cast kernel per weight per step (0.6 ms at TSF-B), which the
# `_version` change after a normal optimizer step incurs anyway.
_GEN = [0]


def param_generation():
    return _GEN[0]


def invalidate_param_caches(*_a, **_k):
    """Force every derived parameter copy to be rebuilt at its next use (call after writing through `param.data`)."""
    _GEN[0] += 1


def _install_generation_hook():
    try:
        from torch.optim.optimizer import register_optimizer_step_post_hook
        register_optimizer_step_post_hook(invalidate_param_caches)
    except Exception:          # pragma: no cover  (very old torch: fall back to re-casting on every forward)
        _GEN.append("always")


_install_generation_hook()


class _Shadow:
    """bf16 copies of fp32 parameters, refreshed when the parameter may have been modified (see above)."""

    def __init__(self):
        self._c = {}

    def get(self, p):
        key = id(p)
        ent = self._c.get(key)
        ver = (p._version, _GEN[0])
        if ent is None or ent[0] is not p or ent[1] != ver or ent[2].device != p.device or len(_GEN) > 1:
            with torch.no_grad():
                src = p.detach()
                w = ops.cast_bf16(src.reshape(-1)).view(src.shape)
            ent = (p, ver, w)
            self._c[key] = ent
        return ent[2]

    def clear(self):
        self._c.clear()


SHADOW = _Shadow()

# bf16 copy of an fp32 gradient handed from one backward stage to the next.  The fp32 tensor is kept alive while
# stashed so its storage cannot be recycled for an unrelated gradient; a miss simply re-casts.
_GRAD_BF16 = {}


def stash_bf16(t_f32, t_bf16):
    _GRAD_BF16.clear()
    if t_bf16 is not None:
        _GRAD_BF16["g"] = (t_f32, t_f32.data_ptr(), t_f32.numel(), t_bf16)


def take_bf16(t_f32):
    ent = _GRAD_BF16.pop("g", None)
    if ent is not None and ent[1] == t_f32.data_ptr() and ent[2] == t_f32.numel():
        return ent[3].view(t_f32.shape)
    return ops.cast_bf16(t_f32.reshape(-1)).view(t_f32.shape)


# ----------------------------------------------------------------------------------------------- side stream for dW / db
# Weight and bias gradients are off the backward's critical chain (nothing downstream reads them before the autograd
# node returns).  With WGRAD_STREAM on they are enqueued on a second CUDA stream, so the tensor-core-bound split-K wgrad
# GEMMs overlap the chain's HBM-bound kernels (LayerNorm backward, CLS / time attention) and the HBM-bound bias column
# sums overlap the chain's GEMMs.  Every autograd node joins the side stream before it returns its gradients.
WGRAD_STREAM = os.environ.get("LAVILA_B200_WGRAD_STREAM", "0") == "1"
_SIDE = {}


def _side_stream(dev):
    st = _SIDE.get(dev)
    if st is None:
        st = {"stream": torch.cuda.Stream(device=dev), "dirty": False}
        _SIDE[dev] = st
    return st


@contextlib.contextmanager
def side_work(*tensors):
    """Run the enclosed launches on the side stream, ordered after everything already enqueued on the current stream.
    `tensors` (operands and outputs) are recorded on the side stream so the caching allocator cannot recycle them early."""
    if not WGRAD_STREAM:
        yield
        return
    dev = tensors[0].device
    st = _side_stream(dev)
    main = torch.cuda.current_stream(dev)
    ev = torch.cuda.Event()
    ev.record(main)
    st["stream"].wait_event(ev)
    with torch.cuda.stream(st["stream"]):
        yield
    for t in tensors:
        if t is not None:
            t.record_stream(st["stream"])
    st["dirty"] = True


def side_join(dev):
    """Make the current stream wait for all side-stream work issued so far (no-op if there was none)."""
    st = _SIDE.get(dev)
    if st is None or not st["dirty"]:
        return
    ev = torch.cuda.Event()
    ev.record(st["stream"])
    torch.cuda.current_stream(dev).wait_event(ev)
    st["dirty"] = False


# fp32 gradient buffers of one autograd node come out of ONE zero-filled allocation (a fill launch per node instead of one per
# parameter: 18 per SpaceTimeBlock); each is a contiguous, 256-byte aligned view, so the split-K atomics and DDP see ordinary tensors.
_GRAD_POOL = []


@contextlib.contextmanager
def grad_pool(params):
    params = [p for p in params if p is not None]
    total = sum((p.numel() + 63) // 64 * 64 for p in params)
    pool = {"buf": torch.zeros(total, device=params[0].device, dtype=F32) if params else None, "off": 0}
    _GRAD_POOL.append(pool)
    try:
        yield
    finally:
        _GRAD_POOL.pop()


def _zeros_like_param(p):
    if _GRAD_POOL:
        pool = _GRAD_POOL[-1]
        n = p.numel()
        if pool["buf"] is not None and pool["buf"].device == p.device and pool["off"] + n <= pool["buf"].numel():
            v = pool["buf"][pool["off"]:pool["off"] + n].view(p.shape)
            pool["off"] += (n + 63) // 64 * 64
            return v
    return torch.zeros(p.shape, device=p.device, dtype=F32)


def _wgrad(dy_b, x_b, out_features, in_features, tokens, dW):
    """dW[out,in] += dy^T x  (both operands read MN-major straight from the activation tensors)."""
    ops.gemm(dy_b, x_b, out_features, in_features, tokens, dW.view(out_features, in_features), a_mn=1, b_mn=1,
             flags=L.EPI_ATOMIC, k_splits=ops.wgrad_splits(out_features, in_features, tokens))


# ----------------------------------------------------------------------------------------------- attention sub-block
def attn_sub_fwd(x_ln, resid, P, dims, mode, eps, gate=None):
    """x_ln, resid: fp32 [M, D].  P: dict(ln_w, ln_b, qkv_w, qkv_b, proj_w, proj_b) fp32 params.
    dims: dict(B, H, T, n, L, N).  Returns (y fp32 [M, D], saved)."""
    M, D = x_ln.shape
    dev = x_ln.device
    B, H = dims["B"], dims["H"]
    ln = torch.empty(M, D, device=dev, dtype=BF16)
    ops.layernorm_fwd(x_ln, P["ln_w"], P["ln_b"], eps, M, D, y_bf16=ln)
    qkv = torch.empty(M, 3 * D, device=dev, dtype=BF16)
    ops.gemm(ln, SHADOW.get(P["qkv_w"]), M, 3 * D, D, qkv, flags=L.EPI_BIAS, bias=P["qkv_b"])
    att = torch.empty(M, D, device=dev, dtype=BF16)
    lse = torch.empty(M, H, device=dev, dtype=F32)
    if mode == MODE_CAUSAL:
        ops.group_attn_fwd(qkv, att, lse, mode, B, H, Lctx=dims["L"])
    elif mode == MODE_SPACE and ops.space_attn_cls_fused_supported(dims["n"]):
        ops.space_attn_fwd_cls(qkv, att, lse, B, H, dims["T"], dims["n"])        # CLS query inside the tcgen05 group kernel
    elif mode == MODE_TIME and ops.time_attn_cls_fused_supported(dims["T"]):
        ops.time_attn_fwd_cls(qkv, att, lse, B, H, dims["T"], dims["n"])         # CLS query next to every (clip, position) unit
    else:
        ops.group_attn_fwd(qkv, att, lse, mode, B, H, T=dims["T"], n=dims["n"])
        ops.cls_attn_fwd(qkv, att, lse, B, H, dims["N"])
    y = torch.empty(M, D, device=dev, dtype=F32)
    flags = L.EPI_BIAS | L.EPI_RESID
    if gate is not None:
        flags |= L.EPI_SCALE | L.EPI_SCALE_TANH
    ops.gemm(att, SHADOW.get(P["proj_w"]), M, D, D, y, flags=flags, bias=P["proj_b"], resid=resid, scale=gate)
    saved = dict(x_ln=x_ln, ln=ln, qkv=qkv, att=att, lse=lse, mode=mode, eps=eps, gate=gate, dims=dims)
    return y, saved


def attn_sub_bwd(dy, dy_b, P, S, adds=(), want_bf16=True):
    """dy: fp32 grad of the sub-block output (only its bf16 copy dy_b is consumed by the GEMMs).
    Returns (dx_ln fp32 = LN'(...) + sum(adds), dx_ln bf16 or None, grads dict).  The caller owns the residual path."""
    M, D = S["ln"].shape
    dev = dy_b.device
    dims, mode, gate = S["dims"], S["mode"], S["gate"]
    B, H = dims["B"], dims["H"]
    g = {}
    # ---- proj:  y = resid + gate * (att W^T + b)
    g["proj_w"] = _zeros_like_param(P["proj_w"])
    g["proj_b"] = _zeros_like_param(P["proj_b"])
    with side_work(dy_b, S["att"], g["proj_w"], g["proj_b"]):
        _wgrad(dy_b, S["att"], D, D, M, g["proj_w"])
        ops.colsum_bf16(dy_b, M, D, g["proj_b"])
    datt = torch.empty(M, D, device=dev, dtype=BF16)
    if gate is not None:
        side_join(dev)      # the gate gradient below reads proj_w / proj_b gradients
    if gate is None:
        ops.gemm(dy_b, SHADOW.get(P["proj_w"]), M, D, D, datt, b_mn=1)
    else:
        ops.gemm(dy_b, SHADOW.get(P["proj_w"]), M, D, D, datt, b_mn=1, flags=L.EPI_SCALE | L.EPI_SCALE_TANH, scale=gate)
        # d tanh(alpha): sum(dy * y_pre) = <W, dy^T att> + <b, colsum(dy)>  (tiny [D,D] reductions)
        tg = torch.tanh(gate.detach())
        g["gate"] = (1 - tg * tg) * ((P["proj_w"].detach() * g["proj_w"]).sum() + (P["proj_b"].detach() * g["proj_b"]).sum())
        g["proj_w"].mul_(tg)
        g["proj_b"].mul_(tg)
    # ---- attention
    dqkv = torch.empty(M, 3 * D, device=dev, dtype=BF16)
    if mode == MODE_CAUSAL:
        ops.group_attn_bwd(S["qkv"], S["att"], S["lse"], datt, dqkv, None, 0, mode, B, H, Lctx=dims["L"])
    elif mode == MODE_SPACE and ops.space_attn_cls_fused_supported(dims["n"]):
        ops.space_attn_bwd_cls(S["qkv"], S["att"], S["lse"], datt, dqkv, B, H, dims["T"], dims["n"])
    elif mode == MODE_TIME and ops.time_attn_cls_fused_supported(dims["T"]):
        ops.time_attn_bwd_cls(S["qkv"], S["att"], S["lse"], datt, dqkv, B, H, dims["T"], dims["n"])
    else:
        # group kernel first (plain stores), then the streaming CLS-query kernel accumulates on top: keeps the
        # read-modify-write latency out of the tensor-core kernel's critical path
        dcls = torch.zeros(B, H, 2, 64, device=dev, dtype=F32)
        ops.group_attn_bwd(S["qkv"], S["att"], S["lse"], datt, dqkv, dcls, 0, mode, B, H, T=dims["T"], n=dims["n"])
        ops.cls_attn_bwd(S["qkv"], S["att"], datt, S["lse"], dqkv, dcls, B, H, dims["N"], accumulate=True)
        ops.cls_kv_finalize(dcls, dqkv, B, H, dims["N"])
    del datt
    # ---- qkv
    g["qkv_w"] = _zeros_like_param(P["qkv_w"])
    g["qkv_b"] = _zeros_like_param(P["qkv_b"])
    with side_work(dqkv, S["ln"], g["qkv_w"], g["qkv_b"]):
        _wgrad(dqkv, S["ln"], 3 * D, D, M, g["qkv_w"])
        ops.colsum_bf16(dqkv, M, 3 * D, g["qkv_b"])
    dln = torch.empty(M, D, device=dev, dtype=BF16)
    ops.gemm(dqkv, SHADOW.get(P["qkv_w"]), M, D, 3 * D, dln, b_mn=1)
    del dqkv
    # ---- LayerNorm backward (+ residual-path gradients folded in)
    g["ln_w"] = _zeros_like_param(P["ln_w"])
    g["ln_b"] = _zeros_like_param(P["ln_b"])
    dx = torch.empty(M, D, device=dev, dtype=F32)
    dx_b = torch.empty(M, D, device=dev, dtype=BF16) if want_bf16 else None
    a1 = adds[0] if len(adds) > 0 else None
    a2 = adds[1] if len(adds) > 1 else None
    ops.layernorm_bwd(dln, S["x_ln"], P["ln_w"], S["eps"], M, D, add1=a1, add2=a2, dx=dx, dx_bf16=dx_b,
                      dgamma=g["ln_w"], dbeta=g["ln_b"])
    return dx, dx_b, g


# ----------------------------------------------------------------------------------------------- MLP sub-block
def mlp_sub_fwd(r, P, eps):
    """r fp32 [M, D] -> r + fc2(quickgelu(fc1(LN(r)))).  P: ln_w, ln_b, fc1_w, fc1_b, fc2_w, fc2_b."""
    M, D = r.shape
    dev = r.device
    Hd = P["fc1_w"].shape[0]
    ln = torch.empty(M, D, device=dev, dtype=BF16)
    ops.layernorm_fwd(r, P["ln_w"], P["ln_b"], eps, M, D, y_bf16=ln)
    act = torch.empty(M, Hd, device=dev, dtype=BF16)
    pre = torch.empty(M, Hd, device=dev, dtype=BF16)
    ops.gemm(ln, SHADOW.get(P["fc1_w"]), M, Hd, D, act, flags=L.EPI_BIAS | L.EPI_QUICKGELU, bias=P["fc1_b"], out2=pre)
    y = torch.empty(M, D, device=dev, dtype=F32)
    ops.gemm(act, SHADOW.get(P["fc2_w"]), M, D, Hd, y, flags=L.EPI_BIAS | L.EPI_RESID, bias=P["fc2_b"], resid=r)
    return y, dict(r=r, ln=ln, act=act, pre=pre, eps=eps)


def mlp_sub_bwd(dy, dy_b, P, S, want_bf16=True):
    """Returns (dr fp32 = dy + LN'(...), dr bf16, grads)."""
    M, D = S["ln"].shape
    Hd = S["act"].shape[1]
    dev = dy.device
    g = {}
    g["fc2_w"] = _zeros_like_param(P["fc2_w"])
    g["fc2_b"] = _zeros_like_param(P["fc2_b"])
    with side_work(dy_b, S["act"], g["fc2_w"], g["fc2_b"]):
        _wgrad(dy_b, S["act"], D, Hd, M, g["fc2_w"])
        ops.colsum_bf16(dy_b, M, D, g["fc2_b"])
    dh = torch.empty(M, Hd, device=dev, dtype=BF16)
    ops.gemm(dy_b, SHADOW.get(P["fc2_w"]), M, Hd, D, dh, b_mn=1, flags=L.EPI_DQUICKGELU, aux=S["pre"])
    g["fc1_w"] = _zeros_like_param(P["fc1_w"])
    g["fc1_b"] = _zeros_like_param(P["fc1_b"])
    with side_work(dh, S["ln"], g["fc1_w"], g["fc1_b"]):
        _wgrad(dh, S["ln"], Hd, D, M, g["fc1_w"])
        ops.colsum_bf16(dh, M, Hd, g["fc1_b"])
    dln = torch.empty(M, D, device=dev, dtype=BF16)
    ops.gemm(dh, SHADOW.get(P["fc1_w"]), M, D, Hd, dln, b_mn=1)
    del dh
    g["ln_w"] = _zeros_like_param(P["ln_w"])
    g["ln_b"] = _zeros_like_param(P["ln_b"])
    dr = torch.empty(M, D, device=dev, dtype=F32)
    dr_b = torch.empty(M, D, device=dev, dtype=BF16) if want_bf16 else None
    ops.layernorm_bwd(dln, S["r"], P["ln_w"], S["eps"], M, D, add1=dy, dx=dr, dx_bf16=dr_b, dgamma=g["ln_w"],
                      dbeta=g["ln_b"])
    return dr, dr_b, g


# ----------------------------------------------------------------------------------------------- SpaceTimeBlock
BLOCK_PARAM_ORDER = (
    "norm3.weight", "norm3.bias", "timeattn.qkv.weight", "timeattn.qkv.bias", "timeattn.proj.weight", "timeattn.proj.bias",
    "norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias",
    "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias",
)


def _attn_params(ps, ln, at):
    return dict(ln_w=ps[ln + ".weight"], ln_b=ps[ln + ".bias"], qkv_w=ps[at + ".qkv.weight"], qkv_b=ps[at + ".qkv.bias"],
                proj_w=ps[at + ".proj.weight"], proj_b=ps[at + ".proj.bias"])


def _mlp_params(ps, ln="norm2", fc1="mlp.fc1", fc2="mlp.fc2"):
    return dict(ln_w=ps[ln + ".weight"], ln_b=ps[ln + ".bias"], fc1_w=ps[fc1 + ".weight"], fc1_b=ps[fc1 + ".bias"],
                fc2_w=ps[fc2 + ".weight"], fc2_b=ps[fc2 + ".bias"])


class SpaceTimeBlockFn(torch.autograd.Function):
    """lavila/models/timesformer.py:173-198 as one autograd node: time attention -> space attention (residual from
    the block input, 'frozen-in-time') -> QuickGELU MLP.  x: fp32 [B, N, D]."""

    @staticmethod
    def _run(x2, ps, dims, eps, gate):
        xt, s_t = attn_sub_fwd(x2, x2, _attn_params(ps, "norm3", "timeattn"), dims, MODE_TIME, eps, gate=gate)
        r, s_s = attn_sub_fwd(xt, x2, _attn_params(ps, "norm1", "attn"), dims, MODE_SPACE, eps)
        y, s_m = mlp_sub_fwd(r, _mlp_params(ps), eps)
        return y, s_t, s_s, s_m

    @staticmethod
    def forward(ctx, x, heads, frames, patches, eps, gate, checkpoint, *params):
        ps = dict(zip(BLOCK_PARAM_ORDER, params))
        B, N, D = x.shape
        x2 = x.contiguous().view(B * N, D)
        if x2.dtype != F32:
            x2 = x2.float()
        dims = dict(B=B, H=heads, T=frames, n=patches, N=N, L=0)
        y, s_t, s_s, s_m = SpaceTimeBlockFn._run(x2, ps, dims, eps, gate)
        if checkpoint:
            # use_checkpoint=True (timesformer.py:175-190 checkpoints the two attentions; here the whole block): keep only
            # the block input (4 bytes/element instead of ~50) and re-run the forward kernels in backward
            ctx.saved = (None, None, None, ps, gate)
            ctx.recompute = (x2, dims, eps)
        else:
            ctx.saved = (s_t, s_s, s_m, ps, gate)
            ctx.recompute = None
        ctx.shape = (B, N, D)
        return y.view(B, N, D)

    @staticmethod
    def backward(ctx, dy):
        s_t, s_s, s_m, ps, gate = ctx.saved
        ctx.saved = None
        if ctx.recompute is not None:
            x2, dims, eps = ctx.recompute
            ctx.recompute = None
            y, s_t, s_s, s_m = SpaceTimeBlockFn._run(x2, ps, dims, eps, gate)
            del y
        B, N, D = ctx.shape
        dy = dy.contiguous().view(B * N, D)
        dy_b = take_bf16(dy)
        with grad_pool([ps[k] for k in BLOCK_PARAM_ORDER]):
            dr, dr_b, g_m = mlp_sub_bwd(dy, dy_b, _mlp_params(ps), s_m)
            del dy_b
            dxt, dxt_b, g_s = attn_sub_bwd(dr, dr_b, _attn_params(ps, "norm1", "attn"), s_s)
            del dr_b
            # x feeds norm3 (LN), the time residual (grad dxt) and the space residual (grad dr)
            dx, dx_b, g_t = attn_sub_bwd(dxt, dxt_b, _attn_params(ps, "norm3", "timeattn"), s_t, adds=(dr, dxt))
        stash_bf16(dx, dx_b)
        grads = {
            "norm3.weight": g_t["ln_w"], "norm3.bias": g_t["ln_b"], "timeattn.qkv.weight": g_t["qkv_w"],
            "timeattn.qkv.bias": g_t["qkv_b"], "timeattn.proj.weight": g_t["proj_w"], "timeattn.proj.bias": g_t["proj_b"],
            "norm1.weight": g_s["ln_w"], "norm1.bias": g_s["ln_b"], "attn.qkv.weight": g_s["qkv_w"],
            "attn.qkv.bias": g_s["qkv_b"], "attn.proj.weight": g_s["proj_w"], "attn.proj.bias": g_s["proj_b"],
            "norm2.weight": g_m["ln_w"], "norm2.bias": g_m["ln_b"], "mlp.fc1.weight": g_m["fc1_w"],
            "mlp.fc1.bias": g_m["fc1_b"], "mlp.fc2.weight": g_m["fc2_w"], "mlp.fc2.bias": g_m["fc2_b"],
        }
        dgate = g_t.get("gate") if gate is not None else None
        side_join(dx.device)
        return (dx.view(B, N, D), None, None, None, None, dgate, None) + tuple(grads[k] for k in BLOCK_PARAM_ORDER)


class LastBlockClsFn(torch.autograd.Function):
    """The LAST SpaceTimeBlock when only its CLS row is consumed (`norm(x)[:, 0]`, timesformer.py:376-378).

    Same arithmetic as SpaceTimeBlockFn for the rows that matter: the time sub-block runs in full (every token's
    key/value of the space attention depends on it), but the space attention is evaluated for the CLS query only (it
    attends to all N tokens, timesformer.py:119), and the output projection and the MLP run on B rows instead of B*N.
    Skips 40/64 of the block's GEMM FLOPs, the whole space group attention and two LayerNorm passes, fwd and bwd.
    x: fp32 [B, N, D] -> fp32 [B, D] (CLS rows of the block output)."""

    @staticmethod
    def forward(ctx, x, heads, frames, patches, eps, gate, *params):
        ps = dict(zip(BLOCK_PARAM_ORDER, params))
        B, N, D = x.shape
        M = B * N
        dev = x.device
        x2 = x.contiguous().view(M, D)
        if x2.dtype != F32:
            x2 = x2.float()
        dims = dict(B=B, H=heads, T=frames, n=patches, N=N, L=0)
        xt, s_t = attn_sub_fwd(x2, x2, _attn_params(ps, "norm3", "timeattn"), dims, MODE_TIME, eps, gate=gate)
        # ---- space attention, CLS query only
        ln1 = torch.empty(M, D, device=dev, dtype=BF16)
        ops.layernorm_fwd(xt, ps["norm1.weight"], ps["norm1.bias"], eps, M, D, y_bf16=ln1)
        wqkv, bqkv = SHADOW.get(ps["attn.qkv.weight"]), ps["attn.qkv.bias"]
        kv = torch.empty(M, 2 * D, device=dev, dtype=BF16)
        ops.gemm(ln1, wqkv[D:], M, 2 * D, D, kv, flags=L.EPI_BIAS, bias=bqkv[D:])
        q = torch.empty(B, D, device=dev, dtype=BF16)
        ops.gemm(ln1, wqkv[:D], B, D, D, q, flags=L.EPI_BIAS, bias=bqkv[:D], lda=N * D)       # CLS rows: row stride N*D
        att = torch.empty(B, D, device=dev, dtype=BF16)
        lse = torch.empty(B, heads, device=dev, dtype=F32)
        ops.cls_query_attn_fwd(q, kv, att, lse, B, heads, N)
        r = torch.empty(B, D, device=dev, dtype=F32)                                              # r = x_cls + proj(att)
        ops.gemm(att, SHADOW.get(ps["attn.proj.weight"]), B, D, D, r, flags=L.EPI_BIAS | L.EPI_RESID,
                 bias=ps["attn.proj.bias"], resid=x2.view(B, N * D))       # residual = CLS rows of x (row stride N*D)
        y, s_m = mlp_sub_fwd(r, _mlp_params(ps), eps)
        ctx.saved = (s_t, s_m, ps, gate, xt, ln1, kv, q, att, lse)
        ctx.shape = (B, N, D, heads, eps)
        return y

    @staticmethod
    def backward(ctx, dy):
        s_t, s_m, ps, gate, xt, ln1, kv, q, att, lse = ctx.saved
        ctx.saved = None
        B, N, D, heads, eps = ctx.shape
        M = B * N
        dev = dy.device
        _GRAD_BF16.clear()
        dy = dy.contiguous().float()
        dy_b = ops.cast_bf16(dy)
        dr, dr_b, g_m = mlp_sub_bwd(dy, dy_b, _mlp_params(ps), s_m)            # [B, D]
        # ---- proj (B rows)
        g_pw, g_pb = _zeros_like_param(ps["attn.proj.weight"]), _zeros_like_param(ps["attn.proj.bias"])
        _wgrad(dr_b, att, D, D, B, g_pw)
        ops.colsum_bf16(dr_b, B, D, g_pb)
        datt = torch.empty(B, D, device=dev, dtype=BF16)
        ops.gemm(dr_b, SHADOW.get(ps["attn.proj.weight"]), B, D, D, datt, b_mn=1)
        # ---- CLS attention backward: dq [B, D], dk/dv for every token [M, 2D]
        dq = torch.empty(B, D, device=dev, dtype=BF16)
        dkv = torch.empty(M, 2 * D, device=dev, dtype=BF16)
        ops.cls_query_attn_bwd(q, kv, att, datt, lse, dq, dkv, B, heads, N)
        # ---- qkv projection: k/v part over all rows, q part over the CLS rows
        g_qw, g_qb = _zeros_like_param(ps["attn.qkv.weight"]), _zeros_like_param(ps["attn.qkv.bias"])
        wqkv = SHADOW.get(ps["attn.qkv.weight"])
        _wgrad(dkv, ln1, 2 * D, D, M, g_qw[D:])
        ops.colsum_bf16(dkv, M, 2 * D, g_qb[D:])
        ops.gemm(dq, ln1, D, D, B, g_qw[:D], a_mn=1, b_mn=1, flags=L.EPI_ATOMIC, ldb=N * D)       # dWq += dq^T ln1[cls rows]
        ops.colsum_bf16(dq, B, D, g_qb[:D])
        dln1 = torch.empty(M, D, device=dev, dtype=BF16)
        ops.gemm(dkv, wqkv[D:], M, D, 2 * D, dln1, b_mn=1)
        dln1_cls = torch.empty(B, D, device=dev, dtype=F32)
        ops.gemm(dq, wqkv[:D], B, D, D, dln1_cls, b_mn=1)
        ops.add_rows(dln1, N * D, dln1_cls, B, D)                                                  # CLS rows also feed q
        del dkv
        # ---- norm1 backward (all rows), then the time sub-block
        g_n1w, g_n1b = _zeros_like_param(ps["norm1.weight"]), _zeros_like_param(ps["norm1.bias"])
        dxt = torch.empty(M, D, device=dev, dtype=F32)
        dxt_b = torch.empty(M, D, device=dev, dtype=BF16)
        ops.layernorm_bwd(dln1, xt, ps["norm1.weight"], eps, M, D, dx=dxt, dx_bf16=dxt_b, dgamma=g_n1w, dbeta=g_n1b)
        del dln1
        dx, dx_b, g_t = attn_sub_bwd(dxt, dxt_b, _attn_params(ps, "norm3", "timeattn"), s_t, adds=(dxt,), want_bf16=False)
        ops.add_rows(dx, N * D, dr, B, D)                       # space residual r = x_cls + ... reaches x at the CLS rows only
        grads = {
            "norm3.weight": g_t["ln_w"], "norm3.bias": g_t["ln_b"], "timeattn.qkv.weight": g_t["qkv_w"],
            "timeattn.qkv.bias": g_t["qkv_b"], "timeattn.proj.weight": g_t["proj_w"], "timeattn.proj.bias": g_t["proj_b"],
            "norm1.weight": g_n1w, "norm1.bias": g_n1b, "attn.qkv.weight": g_qw, "attn.qkv.bias": g_qb,
            "attn.proj.weight": g_pw, "attn.proj.bias": g_pb,
            "norm2.weight": g_m["ln_w"], "norm2.bias": g_m["ln_b"], "mlp.fc1.weight": g_m["fc1_w"],
            "mlp.fc1.bias": g_m["fc1_b"], "mlp.fc2.weight": g_m["fc2_w"], "mlp.fc2.bias": g_m["fc2_b"],
        }
        dgate = g_t.get("gate") if gate is not None else None
        side_join(dx.device)
        return (dx.view(B, N, D), None, None, None, None, dgate) + tuple(grads[k] for k in BLOCK_PARAM_ORDER)


# ----------------------------------------------------------------------------------------------- text block
TEXT_PARAM_ORDER = (
    "ln_1.weight", "ln_1.bias", "attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight", "attn.out_proj.bias",
    "ln_2.weight", "ln_2.bias", "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias",
)


def _text_attn_params(ps):
    return dict(ln_w=ps["ln_1.weight"], ln_b=ps["ln_1.bias"], qkv_w=ps["attn.in_proj_weight"], qkv_b=ps["attn.in_proj_bias"],
                proj_w=ps["attn.out_proj.weight"], proj_b=ps["attn.out_proj.bias"])


class TextBlockFn(torch.autograd.Function):
    """lavila/models/openai_model.py:182-216 (pre-LN causal MHA + QuickGELU MLP).  x: fp32 [B, L, W] (batch-major; the
    reference's LND layout is a pure permutation)."""

    @staticmethod
    def forward(ctx, x, heads, *params):
        ps = dict(zip(TEXT_PARAM_ORDER, params))
        B, Lc, W = x.shape
        x2 = x.contiguous().view(B * Lc, W)
        dims = dict(B=B, H=heads, T=0, n=0, N=Lc, L=Lc)
        x1, s_a = attn_sub_fwd(x2, x2, _text_attn_params(ps), dims, MODE_CAUSAL, 1e-5)
        y, s_m = mlp_sub_fwd(x1, _mlp_params(ps, "ln_2", "mlp.c_fc", "mlp.c_proj"), 1e-5)
        ctx.saved = (s_a, s_m, ps)
        ctx.shape = (B, Lc, W)
        return y.view(B, Lc, W)

    @staticmethod
    def backward(ctx, dy):
        s_a, s_m, ps = ctx.saved
        ctx.saved = None
        B, Lc, W = ctx.shape
        dy = dy.contiguous().view(B * Lc, W)
        dy_b = take_bf16(dy)
        with grad_pool([ps[k] for k in TEXT_PARAM_ORDER]):
            dx1, dx1_b, g_m = mlp_sub_bwd(dy, dy_b, _mlp_params(ps, "ln_2", "mlp.c_fc", "mlp.c_proj"), s_m)
            dx, dx_b, g_a = attn_sub_bwd(dx1, dx1_b, _text_attn_params(ps), s_a, adds=(dx1,))
        stash_bf16(dx, dx_b)
        grads = {
            "ln_1.weight": g_a["ln_w"], "ln_1.bias": g_a["ln_b"], "attn.in_proj_weight": g_a["qkv_w"],
            "attn.in_proj_bias": g_a["qkv_b"], "attn.out_proj.weight": g_a["proj_w"], "attn.out_proj.bias": g_a["proj_b"],
            "ln_2.weight": g_m["ln_w"], "ln_2.bias": g_m["ln_b"], "mlp.c_fc.weight": g_m["fc1_w"],
            "mlp.c_fc.bias": g_m["fc1_b"], "mlp.c_proj.weight": g_m["fc2_w"], "mlp.c_proj.bias": g_m["fc2_b"],
        }
        side_join(dx.device)
        return (dx.view(B, Lc, W), None) + tuple(grads[k] for k in TEXT_PARAM_ORDER)


# ----------------------------------------------------------------------------------------------- vision stem / head
class PatchEmbedStemFn(torch.autograd.Function):
    """Frames [B,C,T,H,W] fp32 -> token stream [B, 1+T*n, D] fp32:
    permute (timesformer.py:387) + Conv2d patch embedding (:77-84) + CLS concat + positional/temporal embedding
    (:353-364) + ln_pre (:366), as im2col -> tcgen05 GEMM -> assemble -> LayerNorm."""

    @staticmethod
    def forward(ctx, frames, conv_w, conv_b, cls_token, pos_embed, temporal_embed, lnpre_w, lnpre_b, patch):
        B, C, T, H, W = frames.shape
        D = conv_w.shape[0]
        n = (H // patch) * (W // patch)
        N = 1 + T * n
        dev = frames.device
        K = C * patch * patch
        Kp = (K + 7) // 8 * 8
        frames = frames.contiguous()
        if frames.dtype != F32:
            frames = frames.float()
        if Kp != K:
            patches = torch.zeros(B * T * n, Kp, device=dev, dtype=BF16)
        else:
            patches = torch.empty(B * T * n, Kp, device=dev, dtype=BF16)
        ops.patch_im2col(frames, patches, B, C, T, H, W, patch)
        wb = SHADOW.get(conv_w).view(D, K)
        if Kp != K:
            wpad = torch.zeros(D, Kp, device=dev, dtype=BF16)
            wpad[:, :K] = wb
            wb = wpad
        pe = torch.empty(B * T * n, D, device=dev, dtype=F32)
        ops.gemm(patches, wb, B * T * n, D, Kp, pe, flags=(L.EPI_BIAS if conv_b is not None else 0), bias=conv_b)
        x0 = torch.empty(B * N, D, device=dev, dtype=F32)
        ops.embed_assemble(pe, cls_token.reshape(-1), pos_embed.reshape(-1, D), temporal_embed.reshape(-1, D), x0, B, T, n, D)
        del pe
        if lnpre_w is not None:
            x = torch.empty(B * N, D, device=dev, dtype=F32)
            ops.layernorm_fwd(x0, lnpre_w, lnpre_b, 1e-5, B * N, D, y_f32=x)
        else:
            x = x0
        ctx.saved = (patches, x0, conv_w, conv_b, cls_token, pos_embed, temporal_embed, lnpre_w, lnpre_b)
        ctx.meta = (B, C, T, H, W, D, n, N, K, Kp, patch)
        return x.view(B, N, D)

    @staticmethod
    def backward(ctx, dx):
        patches, x0, conv_w, conv_b, cls_token, pos_embed, temporal_embed, lnpre_w, lnpre_b = ctx.saved
        ctx.saved = None
        B, C, T, H, W, D, n, N, K, Kp, patch = ctx.meta
        dev = dx.device
        dx = dx.contiguous().view(B * N, D)
        _GRAD_BF16.clear()
        d_lw = d_lb = None
        if lnpre_w is not None:
            d_lw, d_lb = _zeros_like_param(lnpre_w), _zeros_like_param(lnpre_b)
            dx0 = torch.empty(B * N, D, device=dev, dtype=F32)
            ops.layernorm_bwd(dx, x0, lnpre_w, 1e-5, B * N, D, dx=dx0, dgamma=d_lw, dbeta=d_lb)
        else:
            dx0 = dx
        d_pos, d_cls, d_tmp = _zeros_like_param(pos_embed), _zeros_like_param(cls_token), _zeros_like_param(temporal_embed)
        dpatch = torch.empty(B * T * n, D, device=dev, dtype=BF16)
        ops.embed_assemble_bwd(dx0, d_pos, d_cls, d_tmp, dpatch, B, T, n, D)
        d_w = torch.zeros(D, Kp, device=dev, dtype=F32)
        _wgrad(dpatch, patches, D, Kp, B * T * n, d_w)
        d_b = None
        if conv_b is not None:
            d_b = _zeros_like_param(conv_b)
            ops.colsum_bf16(dpatch, B * T * n, D, d_b)
        d_w = d_w[:, :K].reshape(conv_w.shape)
        return None, d_w, d_b, d_cls, d_pos, d_tmp, d_lw, d_lb, None


class StridedLayerNormFn(torch.autograd.Function):
    """LayerNorm of selected rows (row stride `step` tokens): the final `norm(x)[:, 0]` (timesformer.py:377) touches
    only the CLS row of every clip, so only those B rows are normalised / differentiated."""

    @staticmethod
    def forward(ctx, x, w, b, eps, step, rows):
        D = x.shape[-1]
        x = x.contiguous()
        y = torch.empty(rows, D, device=x.device, dtype=F32)
        ops.layernorm_fwd(x, w, b, eps, rows, D, ldx=step * D, y_f32=y)
        ctx.saved = (x, w, b)
        ctx.meta = (eps, step, rows, D)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, b = ctx.saved
        ctx.saved = None
        eps, step, rows, D = ctx.meta
        dy = dy.contiguous().float()
        dx = torch.zeros_like(x)
        dw, db = _zeros_like_param(w), _zeros_like_param(b)
        ops.layernorm_bwd(dy, x, w, eps, rows, D, ldx=step * D, dx=dx, lddx=step * D, dgamma=dw, dbeta=db)
        _GRAD_BF16.clear()
        return dx, dw, db, None, None, None


class LayerNormFn(torch.autograd.Function):
    """Plain fp32 -> fp32 LayerNorm over all rows (forward_features(cls_at_last=False), ln_final on gathered rows)."""

    @staticmethod
    def forward(ctx, x, w, b, eps):
        shp = x.shape
        D = shp[-1]
        x2 = x.contiguous().view(-1, D)
        y = torch.empty_like(x2)
        ops.layernorm_fwd(x2, w, b, eps, x2.shape[0], D, y_f32=y)
        ctx.saved = (x2, w, b)
        ctx.meta = (eps, shp)
        return y.view(shp)

    @staticmethod
    def backward(ctx, dy):
        x2, w, b = ctx.saved
        ctx.saved = None
        eps, shp = ctx.meta
        D = shp[-1]
        dy = dy.contiguous().float().view(-1, D)
        dx = torch.empty_like(x2)
        dw, db = _zeros_like_param(w), _zeros_like_param(b)
        ops.layernorm_bwd(dy, x2, w, eps, x2.shape[0], D, dx=dx, dgamma=dw, dbeta=db)
        _GRAD_BF16.clear()
        return dx.view(shp), dw, db, None


class ProjectFn(torch.autograd.Function):
    """y = x @ P  with P [in, out] (models.py:146 image_projection, :160 text_projection), bf16 operands, fp32 out."""

    @staticmethod
    def forward(ctx, x, proj):
        R, Din = x.shape
        E = proj.shape[1]
        xb = ops.cast_bf16(x.contiguous().float())
        y = torch.empty(R, E, device=x.device, dtype=F32)
        ops.gemm(xb, SHADOW.get(proj), R, E, Din, y, b_mn=1)
        ctx.saved = (xb, proj)
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, proj = ctx.saved
        ctx.saved = None
        R, Din = xb.shape
        E = proj.shape[1]
        dyb = ops.cast_bf16(dy.contiguous().float())
        dx = torch.empty(R, Din, device=dy.device, dtype=F32)
        ops.gemm(dyb, SHADOW.get(proj), R, Din, E, dx)            # dx = dy P^T : B_op[N=Din][K=E] = P row-major
        dproj = torch.zeros(Din, E, device=dy.device, dtype=F32)
        ops.gemm(xb, dyb, Din, E, R, dproj, a_mn=1, b_mn=1, flags=L.EPI_ATOMIC)
        return dx, dproj


class TextEmbedFn(torch.autograd.Function):
    """token_embedding(text) + positional_embedding (models.py:151-152); ids are int64 and used exactly."""

    @staticmethod
    def forward(ctx, text, tok_w, pos):
        B, Lc = text.shape
        W = tok_w.shape[1]
        text = text.contiguous()
        x = torch.empty(B * Lc, W, device=tok_w.device, dtype=F32)
        ops.text_embed(text, tok_w, pos, x, B * Lc, Lc, W, tok_w.shape[0])
        ctx.saved = (text, tok_w, pos)
        return x.view(B, Lc, W)

    @staticmethod
    def backward(ctx, dx):
        text, tok_w, pos = ctx.saved
        ctx.saved = None
        B, Lc = text.shape
        W = tok_w.shape[1]
        dx = dx.contiguous().float()
        dtok, dpos = _zeros_like_param(tok_w), _zeros_like_param(pos)
        ops.text_embed_bwd(text, dx, dtok, dpos, B * Lc, Lc, W, tok_w.shape[0])
        _GRAD_BF16.clear()
        return None, dtok, dpos


class GatherEotFn(torch.autograd.Function):
    """x[arange(B), text.argmax(-1)]  (models.py:160): exact first-max index, fp32 row gather / scatter."""

    @staticmethod
    def forward(ctx, x, text):
        B, Lc, W = x.shape
        idx = torch.empty(B, device=x.device, dtype=torch.int32)
        ops.argmax_i64(text.contiguous(), idx, B, Lc)
        y = torch.empty(B, W, device=x.device, dtype=F32)
        ops.gather_rows(x.contiguous(), idx, y, B, Lc, W, scatter=False)
        ctx.saved = (idx,)
        ctx.shape = (B, Lc, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved
        B, Lc, W = ctx.shape
        dx = torch.zeros(B, Lc, W, device=dy.device, dtype=F32)
        ops.gather_rows(dy.contiguous().float(), idx, dx, B, Lc, W, scatter=True)
        _GRAD_BF16.clear()
        return dx, None


class L2NormalizeFn(torch.autograd.Function):
    """F.normalize(x, dim=-1) (models.py:169-170)."""

    @staticmethod
    def forward(ctx, x):
        R, E = x.shape
        x = x.contiguous().float()
        y = torch.empty_like(x)
        nrm = torch.empty(R, device=x.device, dtype=F32)
        ops.l2norm_fwd(x, y, nrm, R, E)
        ctx.saved = (y, nrm)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, nrm = ctx.saved
        ctx.saved = None
        R, E = y.shape
        dx = torch.empty_like(y)
        ops.l2norm_bwd(dy.contiguous().float(), y, nrm, dx, R, E)
        return dx


class VarAttentionFn(torch.autograd.Function):
    """Stand-alone VarAttention.forward (timesformer.py:107-144) for API parity / unit tests: no LayerNorm, no residual.
    x: [B, N, D] (any float dtype) -> [B, N, D] fp32."""

    @staticmethod
    def forward(ctx, x, qkv_w, qkv_b, proj_w, proj_b, heads, mode, frames, patches):
        B, N, D = x.shape
        M = B * N
        dev = x.device
        xb = ops.cast_bf16(x.contiguous().float().view(M, D))
        qkv = torch.empty(M, 3 * D, device=dev, dtype=BF16)
        ops.gemm(xb, SHADOW.get(qkv_w), M, 3 * D, D, qkv, flags=L.EPI_BIAS, bias=qkv_b)
        att = torch.empty(M, D, device=dev, dtype=BF16)
        lse = torch.empty(M, heads, device=dev, dtype=F32)
        if mode == MODE_SPACE and ops.space_attn_cls_fused_supported(patches):
            ops.space_attn_fwd_cls(qkv, att, lse, B, heads, frames, patches)
        elif mode == MODE_TIME and ops.time_attn_cls_fused_supported(frames):
            ops.time_attn_fwd_cls(qkv, att, lse, B, heads, frames, patches)
        else:
            ops.group_attn_fwd(qkv, att, lse, mode, B, heads, T=frames, n=patches)
            ops.cls_attn_fwd(qkv, att, lse, B, heads, N)
        y = torch.empty(M, D, device=dev, dtype=F32)
        ops.gemm(att, SHADOW.get(proj_w), M, D, D, y, flags=L.EPI_BIAS, bias=proj_b)
        ctx.saved = (xb, qkv, att, lse, qkv_w, qkv_b, proj_w, proj_b)
        ctx.meta = (B, N, D, heads, mode, frames, patches)
        return y.view(B, N, D)

    @staticmethod
    def backward(ctx, dy):
        xb, qkv, att, lse, qkv_w, qkv_b, proj_w, proj_b = ctx.saved
        ctx.saved = None
        B, N, D, heads, mode, frames, patches = ctx.meta
        M = B * N
        dev = dy.device
        dyb = ops.cast_bf16(dy.contiguous().float().view(M, D))
        d_pw, d_pb = _zeros_like_param(proj_w), _zeros_like_param(proj_b)
        _wgrad(dyb, att, D, D, M, d_pw)
        ops.colsum_bf16(dyb, M, D, d_pb)
        datt = torch.empty(M, D, device=dev, dtype=BF16)
        ops.gemm(dyb, SHADOW.get(proj_w), M, D, D, datt, b_mn=1)
        dqkv = torch.empty(M, 3 * D, device=dev, dtype=BF16)
        if mode == MODE_SPACE and ops.space_attn_cls_fused_supported(patches):
            ops.space_attn_bwd_cls(qkv, att, lse, datt, dqkv, B, heads, frames, patches)
        elif mode == MODE_TIME and ops.time_attn_cls_fused_supported(frames):
            ops.time_attn_bwd_cls(qkv, att, lse, datt, dqkv, B, heads, frames, patches)
        else:
            dcls = torch.zeros(B, heads, 2, 64, device=dev, dtype=F32)
            ops.group_attn_bwd(qkv, att, lse, datt, dqkv, dcls, 0, mode, B, heads, T=frames, n=patches)
            ops.cls_attn_bwd(qkv, att, datt, lse, dqkv, dcls, B, heads, N, accumulate=True)
            ops.cls_kv_finalize(dcls, dqkv, B, heads, N)
        d_qw, d_qb = _zeros_like_param(qkv_w), _zeros_like_param(qkv_b)
        _wgrad(dqkv, xb, 3 * D, D, M, d_qw)
        ops.colsum_bf16(dqkv, M, 3 * D, d_qb)
        dx = torch.empty(M, D, device=dev, dtype=F32)
        ops.gemm(dqkv, SHADOW.get(qkv_w), M, D, 3 * D, dx, b_mn=1)
        _GRAD_BF16.clear()
        return dx.view(B, N, D), d_qw, d_qb, d_pw, d_pb, None, None, None, None


class MlpFn(torch.autograd.Function):
    """Stand-alone Mlp.forward (timesformer.py:52-58) with QuickGELU: fc2(quickgelu(fc1(x))), no LN / residual."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        shp = x.shape
        D = shp[-1]
        xb = ops.cast_bf16(x.contiguous().float().view(-1, D))
        M, Hd = xb.shape[0], w1.shape[0]
        act = torch.empty(M, Hd, device=x.device, dtype=BF16)
        pre = torch.empty(M, Hd, device=x.device, dtype=BF16)
        ops.gemm(xb, SHADOW.get(w1), M, Hd, D, act, flags=L.EPI_BIAS | L.EPI_QUICKGELU, bias=b1, out2=pre)
        y = torch.empty(M, w2.shape[0], device=x.device, dtype=F32)
        ops.gemm(act, SHADOW.get(w2), M, w2.shape[0], Hd, y, flags=L.EPI_BIAS, bias=b2)
        ctx.saved = (xb, act, pre, w1, b1, w2, b2)
        ctx.shp = shp
        return y.view(shp[:-1] + (w2.shape[0],))

    @staticmethod
    def backward(ctx, dy):
        xb, act, pre, w1, b1, w2, b2 = ctx.saved
        ctx.saved = None
        M, D = xb.shape
        Hd, Do = w1.shape[0], w2.shape[0]
        dyb = ops.cast_bf16(dy.contiguous().float().view(M, Do))
        d_w2, d_b2 = _zeros_like_param(w2), _zeros_like_param(b2)
        _wgrad(dyb, act, Do, Hd, M, d_w2)
        ops.colsum_bf16(dyb, M, Do, d_b2)
        dh = torch.empty(M, Hd, device=dy.device, dtype=BF16)
        ops.gemm(dyb, SHADOW.get(w2), M, Hd, Do, dh, b_mn=1, flags=L.EPI_DQUICKGELU, aux=pre)
        d_w1, d_b1 = _zeros_like_param(w1), _zeros_like_param(b1)
        _wgrad(dh, xb, Hd, D, M, d_w1)
        ops.colsum_bf16(dh, M, Hd, d_b1)
        dx = torch.empty(M, D, device=dy.device, dtype=F32)
        ops.gemm(dh, SHADOW.get(w1), M, D, Hd, dx, b_mn=1)
        _GRAD_BF16.clear()
        return dx.view(ctx.shp), d_w1, d_b1, d_w2, d_b2
