"""GPT-2 with tanh-gated cross-attention -- mirror of lavila/models/gpt2_gated.py (hot subset: augment_gpt2_config :84-89,
GPT2Attention :149-360, SqReLU :363-376, GPT2MLP :379-396, GPT2Block :399-495, GPT2Model.forward :802-994,
GPT2LMHeadModel :1004-1161).  Same parameter / buffer names, so reference checkpoints load with strict=True.

Inference forward on the B200 kernels.  HF `Conv1D` stores weights [in, out]; the tcgen05 GEMM reads that layout
directly as an MN-major B operand.  Training of the narrator is outside this hot path (AMP is disabled
for it upstream) and raises.
"""
import copy
from types import SimpleNamespace

import torch
import torch.nn as nn

from .. import _lib as L
from .. import ops
from ..engine import SHADOW

BF16, F32 = torch.bfloat16, torch.float32


def augment_gpt2_config(config, cross_attn_freq=1, gated_xattn=True):
    """gpt2_gated.py:84-89."""
    new_config = copy.deepcopy(config)
    new_config.add_cross_attention = True
    new_config.add_cross_attention_freq = cross_attn_freq
    new_config.is_tanh_gating = gated_xattn
    return new_config


def _cfg(config, name, default=None):
    alias = {"hidden_size": "n_embd", "num_attention_heads": "n_head", "num_hidden_layers": "n_layer",
             "max_position_embeddings": "n_positions"}
    for n in (name, alias.get(name)):
        if n is not None and hasattr(config, n) and getattr(config, n) is not None:
            return getattr(config, n)
    return default


class Conv1D(nn.Module):
    """transformers.pytorch_utils.Conv1D parameter holder: weight [in, out], bias [out]."""

    def __init__(self, nf, nx):
        super().__init__()
        self.nf = nf
        self.weight = nn.Parameter(torch.empty(nx, nf))
        self.bias = nn.Parameter(torch.zeros(nf))
        nn.init.normal_(self.weight, std=0.02)


def _linear(x_bf16, conv, M, out, flags=0, **kw):
    """out[M, nf] = epilogue(x @ W + b) with W stored [in, out] (B operand MN-major)."""
    K, N = conv.weight.shape
    w = SHADOW.get(conv.weight)
    if M <= ops.SKINNY_MAX_M and ops.gemm_skinny(x_bf16, w, M, N, K, out, flags=flags | L.EPI_BIAS, bias=conv.bias, **kw):
        return out     # decoding-sized M: weight-streaming kernel (csrc/gemm_skinny.cu)
    return ops.gemm(x_bf16, w, M, N, K, out, b_mn=1, flags=flags | L.EPI_BIAS, bias=conv.bias, **kw)


class GPT2Attention(nn.Module):
    """gpt2_gated.py:149-360."""

    def __init__(self, config, is_cross_attention=False, layer_idx=None):
        super().__init__()
        max_positions = _cfg(config, "max_position_embeddings", 1024)
        self.register_buffer("bias", torch.tril(torch.ones((max_positions, max_positions), dtype=torch.uint8)).view(
            1, 1, max_positions, max_positions))
        self.register_buffer("masked_bias", torch.tensor(-1e4))
        self.embed_dim = _cfg(config, "hidden_size")
        self.num_heads = _cfg(config, "num_attention_heads")
        self.head_dim = self.embed_dim // self.num_heads
        if self.head_dim != 64:
            raise NotImplementedError("attention kernels are specialised for head_dim 64")
        if getattr(config, "scale_attn_by_inverse_layer_idx", False) or getattr(config, "reorder_and_upcast_attn", False):
            raise NotImplementedError("scale_attn_by_inverse_layer_idx / reorder_and_upcast_attn are not used by LaViLa")
        self.is_cross_attention = is_cross_attention
        self.layer_idx = layer_idx
        if is_cross_attention:
            self.c_attn = Conv1D(2 * self.embed_dim, self.embed_dim)
            self.q_attn = Conv1D(self.embed_dim, self.embed_dim)
        else:
            self.c_attn = Conv1D(3 * self.embed_dim, self.embed_dim)
        self.c_proj = Conv1D(self.embed_dim, self.embed_dim)


class GPT2MLP(nn.Module):
    """gpt2_gated.py:379-396."""

    def __init__(self, intermediate_size, config, squared_relu=False):
        super().__init__()
        embed_dim = _cfg(config, "hidden_size")
        self.c_fc = Conv1D(intermediate_size, embed_dim)
        self.c_proj = Conv1D(embed_dim, intermediate_size)
        self.squared_relu = squared_relu
        act = getattr(config, "activation_function", "gelu_new")
        if not squared_relu and act != "gelu_new":
            raise NotImplementedError("only activation_function='gelu_new' (GPT-2) has a fused epilogue")


class GPT2Block(nn.Module):
    """gpt2_gated.py:399-495."""

    def __init__(self, config, layer_idx=None):
        super().__init__()
        hidden = _cfg(config, "hidden_size")
        inner = getattr(config, "n_inner", None) or 4 * hidden
        eps = getattr(config, "layer_norm_epsilon", 1e-5)
        self.ln_1 = nn.LayerNorm(hidden, eps=eps)
        self.attn = GPT2Attention(config, layer_idx=layer_idx)
        self.ln_2 = nn.LayerNorm(hidden, eps=eps)
        self.add_cross_attention_freq = getattr(config, "add_cross_attention_freq", 1)
        if getattr(config, "add_cross_attention", False) and layer_idx % self.add_cross_attention_freq == 0:
            self.crossattention = GPT2Attention(config, is_cross_attention=True, layer_idx=layer_idx)
            self.ln_cross_attn = nn.LayerNorm(hidden, eps=eps)
            self.mlp_crossattention = GPT2MLP(inner, config, squared_relu=True)
            self.ln_2_crossattention = nn.LayerNorm(hidden, eps=eps)
            if getattr(config, "is_tanh_gating", False):
                self.alpha_cattn = nn.Parameter(torch.zeros([]))
                self.alpha_dense = nn.Parameter(torch.zeros([]))
        self.mlp = GPT2MLP(inner, config)
        self.eps = eps
        self.layer_idx = layer_idx

    def _ln(self, h, ln, M, H):
        y = torch.empty(M, H, device=h.device, dtype=BF16)
        ops.layernorm_fwd(h, ln.weight, ln.bias, self.eps, M, H, y_bf16=y)
        return y

    def _ffn(self, h, ln, mlp, M, H, act_flag, gate):
        y = self._ln(h, ln, M, H)
        inner = mlp.c_fc.weight.shape[1]
        a = torch.empty(M, inner, device=h.device, dtype=BF16)
        _linear(y, mlp.c_fc, M, a, flags=act_flag)
        out = torch.empty(M, H, device=h.device, dtype=F32)
        fl = L.EPI_RESID | ((L.EPI_SCALE | L.EPI_SCALE_TANH) if gate is not None else 0)
        _linear(a, mlp.c_proj, M, out, flags=fl, resid=h, scale=gate)
        return out

    def forward_rows(self, h, B, Lq, ctx_kv_cache, ctx_bf16, ctx_rows, self_kv_cache=None, past_len=0, dyn=None):
        """h: fp32 [B*Lq, H] residual stream.  ctx_bf16: bf16 [B*ctx_rows, H] (already cast) or None.
        self_kv_cache (dict with "max_len", filled per layer) + past_len: incremental decoding -- the rows are positions
        past_len .. past_len+Lq-1, their keys/values are appended to the layer's cache and attention runs over the
        past_len+Lq cached keys (same numbers as re-running the whole prefix, gpt2_gated.py:331-345 `layer_past`).
        dyn (dict: "pos_idx" int64[1], "lk_dev" int32[1] on the device): the position lives in device memory, so that ONE
        captured CUDA graph of this step can be replayed for every position (Lq must be 1; every launch has static arguments)."""
        M, H = h.shape
        heads = self.attn.num_heads
        dev = h.device
        if ctx_rows > 0 and hasattr(self, "crossattention"):
            ca = self.crossattention
            y = self._ln(h, self.ln_cross_attn, M, H)
            q = torch.empty(M, H, device=dev, dtype=BF16)
            _linear(y, ca.q_attn, M, q)
            kv = ctx_kv_cache.get(self.layer_idx)
            stale = ctx_kv_cache.get("_stale")    # layers whose (persistent) K/V buffer belongs to a previous batch of clips
            # `generate(num_return_sequences=R)` repeats every clip's video tokens R times (narrator.py:108): the R sequences of
            # a clip share ONE set of cross-attention keys / values.  They are projected once per CLIP and the R x Lq query rows
            # of a clip (consecutive rows of q) attend to them as one attention problem -- 1/R of the K/V traffic of a decoding
            # step (12.6 GB -> 1.3 GB per step for GPT-2 XL at 32 clips x 10 sequences), same numbers.
            rep = int(ctx_kv_cache.get("_repeat", 1))
            if rep < 1 or B % rep != 0:
                rep = 1
            clips = B // rep
            if kv is None or (stale is not None and self.layer_idx in stale):
                # K/V of the video tokens are projected once per clip, not once per decoding step
                if kv is None:
                    kv = torch.empty(clips * ctx_rows, 2 * H, device=dev, dtype=BF16)
                    ctx_kv_cache[self.layer_idx] = kv
                src = ctx_bf16 if rep == 1 else ctx_bf16.view(clips, rep, ctx_rows, H)[:, 0].contiguous().view(clips * ctx_rows, H)
                _linear(src, ca.c_attn, clips * ctx_rows, kv)
                if stale is not None:
                    stale.discard(self.layer_idx)
            att = torch.empty(M, H, device=dev, dtype=BF16)
            ops.flash_attn_fwd(q, kv, kv[:, H:], att, clips, heads, rep * Lq, ctx_rows, q_rows=rep * Lq, kv_rows=ctx_rows, ld_q=H,
                               ld_kv=2 * H, ld_out=H, causal=False)
            h2 = torch.empty(M, H, device=dev, dtype=F32)
            gate = getattr(self, "alpha_cattn", None)
            fl = L.EPI_RESID | ((L.EPI_SCALE | L.EPI_SCALE_TANH) if gate is not None else 0)
            _linear(att, ca.c_proj, M, h2, flags=fl, resid=h, scale=gate)
            h = self._ffn(h2, self.ln_2_crossattention, self.mlp_crossattention, M, H, L.EPI_SQRELU,
                          getattr(self, "alpha_dense", None))
        # causal self-attention (where(tril, w, -1e4) == -inf after softmax in fp32)
        y = self._ln(h, self.ln_1, M, H)
        qkv = torch.empty(M, 3 * H, device=dev, dtype=BF16)
        _linear(y, self.attn.c_attn, M, qkv)
        att = torch.empty(M, H, device=dev, dtype=BF16)
        if self_kv_cache is None:
            ops.flash_attn_fwd(qkv, qkv[:, H:], qkv[:, 2 * H:], att, B, heads, Lq, Lq, q_rows=Lq, kv_rows=Lq, ld_q=3 * H,
                               ld_kv=3 * H, ld_out=H, causal=True)
        else:
            Lmax = self_kv_cache["max_len"]
            kv = self_kv_cache.get(self.layer_idx)
            if kv is None:
                kv = torch.zeros(B, Lmax, 2 * H, device=dev, dtype=BF16)
                self_kv_cache[self.layer_idx] = kv
            kv2 = kv.view(B * Lmax, 2 * H)
            if dyn is None:
                kv[:, past_len:past_len + Lq].copy_(qkv.view(B, Lq, 3 * H)[:, :, H:])
                ops.flash_attn_fwd(qkv, kv2, kv2[:, H:], att, B, heads, Lq, past_len + Lq, q_rows=Lq, kv_rows=Lmax,
                                   ld_q=3 * H, ld_kv=2 * H, ld_out=H, causal=True)
            else:
                kv.index_copy_(1, dyn["pos_idx"], qkv.view(B, 1, 3 * H)[:, :, H:])
                ops.flash_attn_fwd_dyn(qkv, kv2, kv2[:, H:], att, B, heads, 1, dyn["lk_dev"], q_rows=1, kv_rows=Lmax,
                                       ld_q=3 * H, ld_kv=2 * H, ld_out=H, causal=True)
        h2 = torch.empty(M, H, device=dev, dtype=F32)
        _linear(att, self.attn.c_proj, M, h2, flags=L.EPI_RESID, resid=h)
        return self._ffn(h2, self.ln_2, self.mlp, M, H, L.EPI_GELU_TANH, None)


class GPT2Model(nn.Module):
    """gpt2_gated.py:783-994 (forward only)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embed_dim = _cfg(config, "hidden_size")
        self.wte = nn.Embedding(_cfg(config, "vocab_size"), self.embed_dim)
        self.wpe = nn.Embedding(_cfg(config, "max_position_embeddings", 1024), self.embed_dim)
        self.h = nn.ModuleList([GPT2Block(config, layer_idx=i) for i in range(_cfg(config, "num_hidden_layers"))])
        self.ln_f = nn.LayerNorm(self.embed_dim, eps=getattr(config, "layer_norm_epsilon", 1e-5))

    @torch.no_grad()
    def forward_rows(self, input_ids, encoder_hidden_states=None, ctx_kv_cache=None, self_kv_cache=None, past_len=0, dyn=None):
        B, Lq = input_ids.shape
        H = self.embed_dim
        dev = self.wte.weight.device
        cache = ctx_kv_cache if ctx_kv_cache is not None else {}
        ctx_b, ctx_rows = None, 0
        if dyn is None:
            h = torch.empty(B * Lq, H, device=dev, dtype=F32)
            # positions past_len .. past_len+Lq-1 (gpt2_gated.py:855-858 position_ids with past_length)
            ops.text_embed(input_ids.contiguous(), self.wte.weight, self.wpe.weight[past_len:], h, B * Lq, Lq, H,
                           self.wte.weight.shape[0])
            if encoder_hidden_states is not None:
                ctx_rows = encoder_hidden_states.shape[1]
                ctx_b = ops.cast_bf16(encoder_hidden_states.contiguous().float().view(-1, H))
        else:
            # graph-replayable step: the position is a device scalar, the cross-attention K/V are already cached
            assert Lq == 1 and self_kv_cache is not None
            h = self.wte.weight[input_ids.reshape(-1)] + self.wpe.weight[dyn["pos_idx"]]
            if encoder_hidden_states is not None:
                ctx_rows = encoder_hidden_states.shape[1]
        for blk in self.h:
            h = blk.forward_rows(h, B, Lq, cache, ctx_b, ctx_rows, self_kv_cache, past_len, dyn)
        y = torch.empty(B * Lq, H, device=dev, dtype=BF16)
        ops.layernorm_fwd(h, self.ln_f.weight, self.ln_f.bias, self.ln_f.eps, B * Lq, H, y_bf16=y)
        return y


class GPT2LMHeadModel(nn.Module):
    """gpt2_gated.py:1004-1161 (forward / freeze_lm_weights).  lm_head is tied to wte (as transformers 4.27 does)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.transformer = GPT2Model(config)
        self.lm_head = nn.Linear(self.transformer.embed_dim, _cfg(config, "vocab_size"), bias=False)
        self.lm_head.weight = self.transformer.wte.weight
        self._gc = False

    def freeze_lm_weights(self):
        """gpt2_gated.py:1019-1029."""
        freeze_list, unfreeze_list = [], []
        for n, p in self.named_parameters():
            if 'crossattention' in n or 'cross_attn' in n or 'alpha_cattn' in n or 'alpha_dense' in n:
                p.requires_grad = True
                unfreeze_list.append(n)
            else:
                p.requires_grad = False
                freeze_list.append(n)
        print("Freeze the pretrained parts in LM: {}".format(freeze_list))
        print(" Learn the rest parts in LM: {}".format(unfreeze_list))

    def gradient_checkpointing_enable(self):
        self._gc = True

    def gradient_checkpointing_disable(self):
        self._gc = False

    def _padded_head(self):
        w = self.lm_head.weight
        V, H = w.shape
        Vp = (V + 3) // 4 * 4
        from ..engine import param_generation
        key = (w.data_ptr(), w._version, param_generation(), w.device)
        if getattr(self, "_head_key", None) != key:
            wb = torch.zeros(Vp, H, device=w.device, dtype=BF16)
            wb[:V] = SHADOW.get(w)
            self._head_bf16, self._head_key = wb, key
        return self._head_bf16, V, Vp

    @torch.no_grad()
    def forward(self, input_ids=None, encoder_hidden_states=None, last_only=False, ctx_kv_cache=None, self_kv_cache=None,
                past_len=0, dyn=None, **kwargs):
        """Returns an object with `.logits` [B, L, vocab] fp32 (or [B, 1, vocab] with last_only=True).
        With self_kv_cache, input_ids holds only the NEW positions past_len .. past_len+L-1."""
        B, Lq = input_ids.shape
        H = self.transformer.embed_dim
        y = self.transformer.forward_rows(input_ids, encoder_hidden_states, ctx_kv_cache, self_kv_cache, past_len, dyn)
        wb, V, Vp = self._padded_head()
        if last_only:
            y = y.view(B, Lq, H)[:, -1].contiguous()
            rows, Lout = B, 1
        else:
            rows, Lout = B * Lq, Lq
        logits = torch.empty(rows, Vp, device=y.device, dtype=F32)
        ops.gemm(y, wb, rows, Vp, H, logits)
        return SimpleNamespace(logits=logits.view(B, Lout, Vp)[:, :, :V])
