"""B200-native TimeSformer (divided space-time attention) -- mirror of lavila/models/timesformer.py.

Same classes, constructor arguments, parameter names/shapes (=> identical state_dict) and forward signatures as the
reference; the arithmetic is done by liblavila_b200.so through lavila_b200.engine.  Sub-modules such as `qkv`/`proj`
(nn.Linear) and `norm1` (nn.LayerNorm) exist to own the parameters under the reference's names; their own forward()
is never on the hot path.
"""
from collections import OrderedDict
from functools import partial

import torch
from torch import nn

from .. import engine as E


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class DropPath(nn.Module):
    """Stochastic depth (timm.models.layers.DropPath, used at timesformer.py:165).  p = 0 in every LaViLa recipe."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        raise NotImplementedError("drop_path > 0 is not supported by the fused SpaceTimeBlock (reference recipes use 0)")


class QuickGELU(nn.Module):
    """lavila/models/openai_model.py:177-179 (only used as the `act_layer` marker; fused into the fc1 GEMM epilogue)."""

    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class Mlp(nn.Module):
    """timesformer.py:42-58."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)
        if not isinstance(self.act, QuickGELU) and type(self.act).__name__ != "QuickGELU":
            raise NotImplementedError("only act_layer=QuickGELU (the CLIP_OPENAI_TIMESFORMER_* recipes) has a fused kernel")

    def forward(self, x):
        return E.MlpFn.apply(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)


class VideoPatchEmbed(nn.Module):
    """timesformer.py:61-84.  forward() keeps the reference contract ([B,F,C,H,W] -> [B*F, D, H/p, W/p])."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, num_frames=8, ln_pre=False):
        super().__init__()
        img_size = to_2tuple(img_size)
        patch_size = to_2tuple(patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0]) * num_frames
        self.num_frames = num_frames
        self.embed_dim = embed_dim
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=not ln_pre)

    def forward(self, x):
        B, F, C, H, W = x.shape
        assert F <= self.num_frames
        # [B,F,C,H,W] -> kernel layout [B,C,F,H,W]; stem without CLS/pos/LN = plain patch projection
        raise NotImplementedError("VideoPatchEmbed is fused into SpaceTimeTransformer.forward_features (PatchEmbedStemFn)")


class VarAttention(nn.Module):
    """timesformer.py:87-144."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0., initialize='random'):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        if head_dim != 64:
            raise NotImplementedError("attention kernels are specialised for head_dim 64 (got %d)" % head_dim)
        if qk_scale is not None and abs(qk_scale - head_dim ** -0.5) > 1e-12:
            raise NotImplementedError("qk_scale override is not supported")
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        if not qkv_bias:
            raise NotImplementedError("qkv_bias=False is not supported (every LaViLa factory uses qkv_bias=True)")
        if initialize == 'zeros':
            self.qkv.weight.data.fill_(0)
            self.qkv.bias.data.fill_(0)
            self.proj.weight.data.fill_(1)
            self.proj.bias.data.fill_(0)
        self.attn_drop = nn.Dropout(attn_drop)   # defined but never applied by the reference either
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x, einops_from, einops_to, einops_dims):
        B, N, D = x.shape
        if "f" in einops_dims:      # space: 'b (f n) d -> (b f) n d'
            frames = int(einops_dims["f"])
            patches = (N - 1) // frames
            mode = E.MODE_SPACE
        else:                       # time: 'b (f n) d -> (b n) f d'
            patches = int(einops_dims["n"])
            frames = (N - 1) // patches
            mode = E.MODE_TIME
        return E.VarAttentionFn.apply(x, self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias,
                                      self.num_heads, mode, frames, patches)


class SpaceTimeBlock(nn.Module):
    """timesformer.py:147-198."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm, time_init='zeros',
                 attention_style='frozen-in-time', is_tanh_gating=False):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = VarAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                                 proj_drop=drop)
        self.timeattn = VarAttention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                                     proj_drop=drop, initialize=time_init)
        if is_tanh_gating:
            self.alpha_timeattn = nn.Parameter(torch.zeros([]))
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.norm3 = norm_layer(dim)
        self.attention_style = attention_style
        self.num_heads = num_heads
        if drop > 0 or drop_path > 0:
            raise NotImplementedError("dropout / drop_path > 0 are not supported by the fused block")

    def _params(self):
        return (self.norm3.weight, self.norm3.bias, self.timeattn.qkv.weight, self.timeattn.qkv.bias,
                self.timeattn.proj.weight, self.timeattn.proj.bias, self.norm1.weight, self.norm1.bias,
                self.attn.qkv.weight, self.attn.qkv.bias, self.attn.proj.weight, self.attn.proj.bias,
                self.norm2.weight, self.norm2.bias, self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight,
                self.mlp.fc2.bias)

    def forward_cls_only(self, x, time_n, space_f):
        """Last block of a `cls_at_last` forward: returns only the CLS rows [B, D] of the block output."""
        gate = getattr(self, "alpha_timeattn", None)
        return E.LastBlockClsFn.apply(x, self.num_heads, int(space_f), int(time_n), float(self.norm1.eps), gate,
                                      *self._params())

    def forward(self, x, einops_from_space, einops_to_space, einops_from_time, einops_to_time, time_n, space_f,
                use_checkpoint=False):
        # use_checkpoint: the block keeps only its input and re-runs its forward kernels in backward (engine.py)
        if self.attention_style != 'frozen-in-time':
            raise NotImplementedError
        gate = getattr(self, "alpha_timeattn", None)
        return E.SpaceTimeBlockFn.apply(x, self.num_heads, int(space_f), int(time_n), float(self.norm1.eps), gate,
                                        bool(use_checkpoint) and torch.is_grad_enabled(), *self._params())


class SpaceTimeTransformer(nn.Module):
    """timesformer.py:201-390."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=True, qk_scale=None, representation_size=None,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0., hybrid_backbone=None, norm_layer=None,
                 num_frames=8, time_init='rand', attention_style='frozen-in-time', ln_pre=False,
                 act_layer=nn.GELU, is_tanh_gating=False):
        super().__init__()
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.num_frames = num_frames
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        if hybrid_backbone is not None:
            raise NotImplementedError('hybrid backbone not implemented')
        self.patch_embed = VideoPatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans,
                                           embed_dim=embed_dim, num_frames=num_frames, ln_pre=ln_pre)
        num_patches = self.patch_embed.num_patches
        self.patches_per_frame = num_patches // num_frames
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patches_per_frame + 1, embed_dim))
        self.temporal_embed = nn.Parameter(torch.zeros(1, num_frames, embed_dim))
        self.ln_pre = nn.LayerNorm(embed_dim) if ln_pre else None
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            SpaceTimeBlock(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                           drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer,
                           time_init=time_init, attention_style=attention_style, act_layer=act_layer,
                           is_tanh_gating=is_tanh_gating)
            for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        if representation_size:
            self.num_features = representation_size
            self.pre_logits = nn.Sequential(OrderedDict([('fc', nn.Linear(embed_dim, representation_size)),
                                                         ('act', nn.Tanh())]))
        else:
            self.pre_logits = nn.Identity()
        self.head = nn.Linear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.trunc_normal_(self.cls_token, std=.02)
        if num_frames == 1:
            self.apply(self._init_weights)
        self.cls_only_tail = True   # see _features_bcthw
        self.einops_from_space = 'b (f n) d'
        self.einops_to_space = '(b f) n d'
        self.einops_from_time = 'b (f n) d'
        self.einops_to_time = '(b n) f d'

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    def get_classifier(self):
        return self.head

    def reset_classifier(self, num_classes, global_pool=''):
        self.num_classes = num_classes
        self.head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()

    def freeze_spatial_weights(self):
        freeze_list = []
        for n, p in self.named_parameters():
            if not ('temporal_embed' in n or 'timeattn' in n or 'norm3' in n):
                p.requires_grad = False
                freeze_list.append(n)
        print("Freeze the pretrained parts in vision model: {}".format(freeze_list))

    def freeze_temporal_weights(self):
        freeze_list = []
        for n, p in self.named_parameters():
            if 'temporal_embed' in n or 'timeattn' in n or 'norm3' in n:
                p.requires_grad = False
                freeze_list.append(n)
        print("Freeze the pretrained parts in vision model: {}".format(freeze_list))

    def _features_bcthw(self, x_bcthw, use_checkpoint=False, cls_at_last=True):
        B, C, T, H, W = x_bcthw.shape
        assert T <= self.num_frames
        pe = self.patch_embed.proj
        lw = self.ln_pre.weight if self.ln_pre is not None else None
        lb = self.ln_pre.bias if self.ln_pre is not None else None
        x = E.PatchEmbedStemFn.apply(x_bcthw, pe.weight, pe.bias, self.cls_token, self.pos_embed, self.temporal_embed,
                                     lw, lb, self.patch_embed.patch_size[0])
        n, f = self.patches_per_frame, T
        last = len(self.blocks) - 1
        for i, blk in enumerate(self.blocks):
            if cls_at_last and i == last and self.cls_only_tail:
                # only norm(x)[:, 0] is consumed: the last block produces its CLS rows directly (same numbers, 40/64 of
                # its GEMM work and the whole space group attention skipped)
                x = blk.forward_cls_only(x, time_n=n, space_f=f)
                x = E.LayerNormFn.apply(x, self.norm.weight, self.norm.bias, float(self.norm.eps))
                return self.pre_logits(x)
            x = blk(x, self.einops_from_space, self.einops_to_space, self.einops_from_time, self.einops_to_time,
                    time_n=n, space_f=f, use_checkpoint=use_checkpoint)
        if cls_at_last:
            N = x.shape[1]
            x = E.StridedLayerNormFn.apply(x, self.norm.weight, self.norm.bias, float(self.norm.eps), N, B)
            return self.pre_logits(x)
        return E.LayerNormFn.apply(x, self.norm.weight, self.norm.bias, float(self.norm.eps))

    def forward_features(self, x, use_checkpoint=False, cls_at_last=True):
        """x: [B, T, C, H, W] (reference contract, timesformer.py:345).  The kernels read [B, C, T, H, W] directly, so
        a caller that permuted a BCTHW tensor (as VCLM_HF.encode_image does) pays no copy: the permute is a view."""
        return self._features_bcthw(x.permute(0, 2, 1, 3, 4), use_checkpoint=use_checkpoint, cls_at_last=cls_at_last)

    def forward(self, x, use_checkpoint=False):
        """x: [B, C, T, H, W] (timesformer.py:384-390); the reference's permute+contiguous copy is folded into im2col."""
        x = self._features_bcthw(x, use_checkpoint=use_checkpoint)
        return self.head(x)
