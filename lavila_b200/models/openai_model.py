"""CLIP text-tower blocks -- mirror of lavila/models/openai_model.py:177-232 (QuickGELU, ResidualAttentionBlock,
Transformer).  Same parameter names (attn.in_proj_weight, attn.out_proj.*, ln_1, mlp.c_fc, mlp.c_proj, ln_2)."""
from collections import OrderedDict

import torch
from torch import nn

from .. import engine as E
from .timesformer import QuickGELU  # noqa: F401  (re-exported, openai_model.py:177)


class ResidualAttentionBlock(nn.Module):
    """openai_model.py:182-216.  Input layout is the reference's [L, N, D] (LND)."""

    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None):
        super().__init__()
        if d_model // n_head != 64:
            raise NotImplementedError("attention kernels are specialised for head_dim 64")
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
        ]))
        self.ln_2 = nn.LayerNorm(d_model)
        self.attn_mask = attn_mask   # always the causal mask built by CLIP.build_attention_mask (models.py:131-137)
        self.n_head = n_head

    def _params(self):
        return (self.ln_1.weight, self.ln_1.bias, self.attn.in_proj_weight, self.attn.in_proj_bias,
                self.attn.out_proj.weight, self.attn.out_proj.bias, self.ln_2.weight, self.ln_2.bias,
                self.mlp.c_fc.weight, self.mlp.c_fc.bias, self.mlp.c_proj.weight, self.mlp.c_proj.bias)

    def forward_bld(self, x):
        """Batch-major fast path used by Transformer.forward: x [B, L, D] fp32."""
        return E.TextBlockFn.apply(x, self.n_head, *self._params())

    def forward(self, x: torch.Tensor, use_checkpoint=False):
        return self.forward_bld(x.permute(1, 0, 2)).permute(1, 0, 2)


class Transformer(nn.Module):
    """openai_model.py:219-232."""

    def __init__(self, width: int, layers: int, heads: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.width = width
        self.layers = layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])

    def forward(self, x: torch.Tensor, use_checkpoint=False):
        """x: [L, N, D] (reference contract).  One permute in, one out; blocks run batch-major."""
        y = x.permute(1, 0, 2)
        for blk in self.resblocks:
            y = blk.forward_bld(y)
        return y.permute(1, 0, 2)

    def forward_bld(self, x):
        for blk in self.resblocks:
            x = blk.forward_bld(x)
        return x
