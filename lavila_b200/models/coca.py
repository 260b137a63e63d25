"""CoCa-style attention pooling -- mirror of lavila/models/coca.py (LayerNorm :28-35, CrossAttention :55-131).
Inference forward on the B200 kernels: LN -> bf16, tcgen05 GEMMs for to_q / to_kv / to_out, flash attention with ONE
shared 64-d key/value head (multi-query).  `parallel_ff` is never enabled by the reference's VCLM_HF (narrator.py:49-53)."""
import torch
import torch.nn as nn

from .. import ops
from ..engine import SHADOW

BF16, F32 = torch.bfloat16, torch.float32


class LayerNorm(nn.Module):
    """coca.py:28-35: learnable gamma, constant zero beta (a buffer, kept for state_dict compatibility)."""

    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))
        self.register_buffer("beta", torch.zeros(dim))

    def forward(self, x):
        shp = x.shape
        D = shp[-1]
        x2 = x.contiguous().float().view(-1, D)
        y = torch.empty_like(x2)
        ops.layernorm_fwd(x2, self.gamma, None, 1e-5, x2.shape[0], D, y_f32=y)
        return y.view(shp)


class CrossAttention(nn.Module):
    """coca.py:55-131."""

    def __init__(self, dim, *, context_dim=None, dim_head=64, heads=8, parallel_ff=False, ff_mult=4, norm_context=False):
        super().__init__()
        if dim_head != 64:
            raise NotImplementedError("attention kernels are specialised for head_dim 64")
        if parallel_ff:
            raise NotImplementedError("parallel_ff is never used by VCLM_HF")
        self.heads = heads
        self.scale = dim_head ** -0.5
        inner_dim = heads * dim_head
        context_dim = context_dim if context_dim is not None else dim
        self.norm = LayerNorm(dim)
        self.context_norm = LayerNorm(context_dim) if norm_context else nn.Identity()
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_kv = nn.Linear(context_dim, dim_head * 2, bias=False)
        self.to_out = nn.Linear(inner_dim, dim, bias=False)
        self.ff = None

    @torch.no_grad()
    def forward(self, x, context):
        """x: [B, Q, dim] queries, context: [B, N, context_dim] -> [B, Q, dim] fp32 (no residual)."""
        B, Q, dim = x.shape
        N, Dc = context.shape[1], context.shape[2]
        dev = x.device
        inner = self.heads * 64
        xq = torch.empty(B * Q, dim, device=dev, dtype=BF16)
        ops.layernorm_fwd(x.contiguous().float().view(B * Q, dim), self.norm.gamma, None, 1e-5, B * Q, dim, y_bf16=xq)
        ctx = context.contiguous().float().view(B * N, Dc)
        cb = torch.empty(B * N, Dc, device=dev, dtype=BF16)
        if isinstance(self.context_norm, LayerNorm):
            ops.layernorm_fwd(ctx, self.context_norm.gamma, None, 1e-5, B * N, Dc, y_bf16=cb)
        else:
            ops.cast_bf16(ctx, out=cb)
        q = torch.empty(B * Q, inner, device=dev, dtype=BF16)
        ops.gemm(xq, SHADOW.get(self.to_q.weight), B * Q, inner, dim, q)
        kv = torch.empty(B * N, 128, device=dev, dtype=BF16)
        ops.gemm(cb, SHADOW.get(self.to_kv.weight), B * N, 128, Dc, kv)
        att = torch.empty(B * Q, inner, device=dev, dtype=BF16)
        ops.flash_attn_fwd(q, kv, kv[:, 64:], att, B, self.heads, Q, N, q_rows=Q, kv_rows=N, ld_q=inner, ld_kv=128,
                           ld_out=inner, kv_head_stride=0, causal=False, scale=self.scale)
        out = torch.empty(B * Q, dim, device=dev, dtype=F32)
        ops.gemm(att, SHADOW.get(self.to_out.weight), B * Q, dim, inner, out)
        return out.view(B, Q, dim)
