"""ORACLE -- CPU restatement (plain PyTorch fp32) of LaViLa's narrator: TimeSformer features -> CoCa attention pooling
-> gated cross-attention GPT-2 decoder -> logits / sampling.   TEST INFRASTRUCTURE ONLY (see oracle/dual_encoder.py).

Pinned by tests/test_oracle_narrator.py against golden vectors produced by the unmodified reference
(tests/golden/make_golden_narrator.py).  Parameters use the reference's state_dict names (VCLM_HF).
"""
import math

import torch
import torch.nn.functional as F

from .dual_encoder import layer_norm, timesformer_features


def gelu_new(x):
    """transformers ACT2FN['gelu_new'] (tanh approximation), used by the stock GPT-2 FFN (gpt2_gated.py:388)."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x.pow(3))))


def conv1d(x, p, prefix):
    """HF Conv1D: weight [in, out], y = x @ W + b (gpt2_gated.py:179-184)."""
    return x @ p[prefix + "weight"] + p[prefix + "bias"]


def attn_pool(queries, context, p, prefix, heads, dim_head=64):
    """coca.CrossAttention.forward (coca.py:90-131): pre-LN on queries and context, multi-query K/V (one 64-d head
    shared by all query heads), softmax(sim - amax), no residual.  queries [B, Q, dim], context [B, N, ctx_dim]."""
    B, Q, dim = queries.shape
    x = F.layer_norm(queries, (dim,), p[prefix + "norm.gamma"], torch.zeros(dim, device=queries.device), 1e-5)            # :100
    c = F.layer_norm(context, (context.shape[-1],), p[prefix + "context_norm.gamma"], torch.zeros(context.shape[-1], device=context.device), 1e-5)
    q = (x @ p[prefix + "to_q.weight"].t()).reshape(B, Q, heads, dim_head).permute(0, 2, 1, 3)      # :104-105
    q = q * dim_head ** -0.5                                                                        # :108
    k, v = (c @ p[prefix + "to_kv.weight"].t()).chunk(2, dim=-1)                                    # :111  [B, N, 64]
    sim = q @ k.unsqueeze(1).transpose(-1, -2)                                                      # :114  [B, h, Q, N]
    sim = sim - sim.amax(dim=-1, keepdim=True)                                                      # :117
    out = torch.softmax(sim, dim=-1) @ v.unsqueeze(1)                                               # :118-121
    out = out.permute(0, 2, 1, 3).reshape(B, Q, heads * dim_head)                                   # :124
    return out @ p[prefix + "to_out.weight"].t()                                                    # :125


def gpt2_attention(h, p, prefix, heads, ctx=None):
    """GPT2Attention.forward/_attn (gpt2_gated.py:309-360,206-238).  Self: causal where(tril, w, -1e4); cross: no mask."""
    B, L, H = h.shape
    dh = H // heads
    if ctx is not None:
        q = conv1d(h, p, prefix + "q_attn.")                                                       # :327
        k, v = conv1d(ctx, p, prefix + "c_attn.").split(H, dim=2)                                  # :328
    else:
        q, k, v = conv1d(h, p, prefix + "c_attn.").split(H, dim=2)                                 # :331
    sp = lambda t: t.reshape(t.shape[0], t.shape[1], heads, dh).permute(0, 2, 1, 3)
    q, k, v = sp(q), sp(k), sp(v)
    w = (q @ k.transpose(-1, -2)) / (dh ** 0.5)                                                    # :207-210
    if ctx is None:
        Lq, Lk = w.shape[-2:]
        causal = torch.tril(torch.ones(Lk, Lk, dtype=torch.bool, device=h.device))[Lk - Lq:Lk, :Lk]
        w = torch.where(causal, w, torch.tensor(-1e4))                                             # :216-220
    a = torch.softmax(w, dim=-1) @ v                                                               # :226-238
    a = a.permute(0, 2, 1, 3).reshape(B, L, H)
    return conv1d(a, p, prefix + "c_proj.")                                                        # :352


def gpt2_block(h, ctx, p, prefix, heads, has_xattn, eps=1e-5):
    """GPT2Block.forward (gpt2_gated.py:421-495)."""
    H = h.shape[-1]
    ln = lambda t, n: layer_norm(t, p[prefix + n + ".weight"], p[prefix + n + ".bias"], eps)
    if has_xattn and ctx is not None:
        a = gpt2_attention(ln(h, "ln_cross_attn"), p, prefix + "crossattention.", heads, ctx=ctx)  # :440-449
        if prefix + "alpha_cattn" in p:
            a = torch.tanh(p[prefix + "alpha_cattn"]) * a                                          # :450-451
        h = h + a                                                                                  # :453
        f = conv1d(ln(h, "ln_2_crossattention"), p, prefix + "mlp_crossattention.c_fc.")
        f = conv1d(torch.relu(f) ** 2, p, prefix + "mlp_crossattention.c_proj.")                   # SqReLU :363-376
        if prefix + "alpha_dense" in p:
            f = torch.tanh(p[prefix + "alpha_dense"]) * f                                          # :458-459
        h = h + f                                                                                  # :461
    h = gpt2_attention(ln(h, "ln_1"), p, prefix + "attn.", heads) + h                              # :464-477
    f = conv1d(gelu_new(conv1d(ln(h, "ln_2"), p, prefix + "mlp.c_fc.")), p, prefix + "mlp.c_proj.")
    return h + f                                                                                   # :483-488


def gpt2_lm_logits(ids, ctx, p, cfg, prefix="text_decoder."):
    """GPT2Model.forward + lm_head (gpt2_gated.py:802-994,1092-1161).  ids int64 [B, L] -> logits [B, L, vocab]."""
    t = prefix + "transformer."
    L = ids.shape[1]
    h = p[t + "wte.weight"][ids] + p[t + "wpe.weight"][:L]                                         # :890-893
    for i in range(cfg["n_layer"]):
        h = gpt2_block(h, ctx, p, "%sh.%d." % (t, i), cfg["n_head"], i % cfg["cross_attn_freq"] == 0)
    h = layer_norm(h, p[t + "ln_f.weight"], p[t + "ln_f.bias"], 1e-5)                              # :974
    return h @ p[prefix + "lm_head.weight"].t()                                                    # :1139


def vclm_encode_image(image_bcthw, p, cfg):
    """VCLM_HF.encode_image (narrator.py:63-87), SpaceTimeTransformer branch."""
    x = timesformer_features(image_bcthw.permute(0, 2, 1, 3, 4), p, cfg["visual"], prefix="visual.", cls_at_last=False)
    B = x.shape[0]
    q = p["img_queries"].unsqueeze(0).expand(B, -1, -1)                                            # :84
    q = attn_pool(q, x, p, "img_attn_pool.", cfg["pool_heads"])                                    # :85
    W = q.shape[-1]
    return F.layer_norm(q, (W,), p["img_attn_pool_norm.gamma"], torch.zeros(W, device=q.device), 1e-5)              # :86


def vclm_forward(image, text, p, cfg):
    """VCLM_HF.forward (narrator.py:89-104): teacher forcing; logits rearranged 'b n c -> b c n'."""
    tok = vclm_encode_image(image, p, cfg)
    logits = gpt2_lm_logits(text[:, :-1], tok, p, cfg)
    return {"text_tokens_logits": logits.permute(0, 2, 1), "labels": text[:, 1:]}


def warp_logits(logits, temperature=1.0, top_p=None):
    """TemperatureLogitsWarper then TopPLogitsWarper(min_tokens_to_keep=1) as configured by narrator.py:368-389
    (transformers: ascending sort -> softmax -> cumsum -> drop where cum <= 1 - p, always keep the last one)."""
    if temperature is not None and temperature != 1.0:
        logits = logits / temperature
    if top_p is not None and top_p < 1.0:
        sl, si = torch.sort(logits, descending=False, dim=-1)
        cum = sl.softmax(dim=-1).cumsum(dim=-1)
        remove = cum <= (1 - top_p)
        remove[..., -1:] = False
        mask = remove.scatter(1, si, remove)
        logits = logits.masked_fill(mask, float("-inf"))
    return logits


def init_narrator_params(cfg, seed=0):
    """Random parameters with the reference's VCLM_HF state_dict names (small configs for tests)."""
    from .dual_encoder import init_params
    g = torch.Generator().manual_seed(seed + 17)
    vcfg = dict(cfg["visual"], text_width=64, text_heads=1, text_layers=0, context_length=8, vocab_size=8, project_dim=8)
    full = init_params(vcfg, seed=seed)
    p = {k: v for k, v in full.items() if k.startswith("visual.")}
    H, Dv = cfg["n_embd"], cfg["visual"]["embed_dim"]
    rn = lambda *s, std=0.02: torch.randn(*s, generator=g) * std
    p["img_queries"] = rn(cfg["num_img_queries"], H, std=H ** -0.5)
    inner = cfg["pool_heads"] * 64
    p["img_attn_pool.norm.gamma"] = 1 + rn(H, std=0.1)
    p["img_attn_pool.context_norm.gamma"] = 1 + rn(Dv, std=0.1)
    p["img_attn_pool.to_q.weight"] = rn(inner, H, std=H ** -0.5)
    p["img_attn_pool.to_kv.weight"] = rn(128, Dv, std=Dv ** -0.5)
    p["img_attn_pool.to_out.weight"] = rn(H, inner, std=inner ** -0.5)
    p["img_attn_pool_norm.gamma"] = 1 + rn(H, std=0.1)
    t = "text_decoder.transformer."
    p[t + "wte.weight"] = rn(cfg["vocab_size"], H)
    p[t + "wpe.weight"] = rn(cfg["n_positions"], H, std=0.01)
    for i in range(cfg["n_layer"]):
        b = "%sh.%d." % (t, i)
        names = ["ln_1", "ln_2"]
        lin = [("attn.c_attn", H, 3 * H), ("attn.c_proj", H, H), ("mlp.c_fc", H, 4 * H), ("mlp.c_proj", 4 * H, H)]
        if i % cfg["cross_attn_freq"] == 0:
            names += ["ln_cross_attn", "ln_2_crossattention"]
            lin += [("crossattention.c_attn", H, 2 * H), ("crossattention.q_attn", H, H), ("crossattention.c_proj", H, H),
                    ("mlp_crossattention.c_fc", H, 4 * H), ("mlp_crossattention.c_proj", 4 * H, H)]
            p[b + "alpha_cattn"] = torch.tensor(0.5)
            p[b + "alpha_dense"] = torch.tensor(-0.4)
        for n in names:
            p[b + n + ".weight"] = 1 + rn(H, std=0.1)
            p[b + n + ".bias"] = rn(H, std=0.1)
        for n, i_, o_ in lin:
            p[b + n + ".weight"] = rn(i_, o_, std=i_ ** -0.5)
            p[b + n + ".bias"] = rn(o_, std=0.02)
    p[t + "ln_f.weight"] = 1 + rn(H, std=0.1)
    p[t + "ln_f.bias"] = rn(H, std=0.1)
    p["text_decoder.lm_head.weight"] = p[t + "wte.weight"]   # tied (transformers 4.27 behaviour the reference relies on)
    return p


SMALL_NARRATOR = dict(
    visual=dict(img_size=32, patch_size=16, embed_dim=128, depth=2, num_heads=2, num_frames=4, ln_pre=True),
    n_embd=128, n_head=2, n_layer=2, cross_attn_freq=2, vocab_size=512, n_positions=32, num_img_queries=8, pool_heads=2)
