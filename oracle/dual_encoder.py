"""ORACLE -- CPU restatement (plain PyTorch, fp32) of LaViLa's dual-encoder contrastive training path.

TEST INFRASTRUCTURE ONLY.  Importers allowed: tests/, __graft_entry__.smoke(), bench.py (`cpu_baseline` /
`--impl reference`).  The product package `lavila_b200` never imports this module.

Pinned: tests/test_oracle.py checks every function here against golden vectors produced by running the
UNMODIFIED reference modules in the build container (tests/golden/make_golden.py, fixtures in tests/golden/).
The reference itself ships no tests or golden vectors.

Each function restates one piece of the reference, cited as file:line relative to the reference root.
Parameters are passed as a flat dict using the reference's `state_dict` names, so the same dict drives the
reference modules, this oracle and the CUDA product.  Everything is differentiable through autograd.
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------- helpers
def quick_gelu(x):
    """lavila/models/openai_model.py:177-179"""
    return x * torch.sigmoid(1.702 * x)


def layer_norm(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def softmax_attend(q, k, v):
    """lavila/models/timesformer.py:35-39 -- plain softmax(q k^T) v, batched over leading dims (q pre-scaled)."""
    s = q @ k.transpose(-1, -2)
    return torch.softmax(s, dim=-1) @ v


# ----------------------------------------------------------------------------------------------- TimeSformer
def var_attention(x, p, prefix, heads, mode, frames, patches):
    """Divided attention with CLS broadcast -- lavila/models/timesformer.py:107-144.

    x: [B, 1 + frames*patches, D] (token order: CLS, then frame-major).  mode: 'time' | 'space'.
    The CLS query attends to all tokens (:119); patch queries attend within their group (same spatial
    position across frames for 'time', same frame for 'space') plus the CLS key/value (:124-128).
    """
    B, N, D = x.shape
    dh = D // heads
    qkv = F.linear(x, p[prefix + "qkv.weight"], p[prefix + "qkv.bias"])          # :110
    q, k, v = qkv.split(D, dim=-1)
    # [B, N, h, dh] -> [B, h, N, dh]                                              # :111
    q, k, v = (t.reshape(B, N, heads, dh).permute(0, 2, 1, 3) for t in (q, k, v))
    q = q * (dh ** -0.5)                                                           # :113
    cls_out = softmax_attend(q[:, :, :1], k, v)                                    # :116-119  [B,h,1,dh]

    def group(t):  # patch tokens -> [B, h, G, L, dh]                              # :121
        t = t[:, :, 1:].reshape(B, heads, frames, patches, dh)
        return t.transpose(2, 3) if mode == "time" else t

    qg, kg, vg = group(q), group(k), group(v)
    G = qg.shape[2]
    cls_k = k[:, :, :1].unsqueeze(2).expand(B, heads, G, 1, dh)                    # :124-125
    cls_v = v[:, :, :1].unsqueeze(2).expand(B, heads, G, 1, dh)
    kg = torch.cat((cls_k, kg), dim=3)                                             # :127-128
    vg = torch.cat((cls_v, vg), dim=3)
    og = softmax_attend(qg, kg, vg)                                                # :131
    if mode == "time":                                                             # :134
        og = og.transpose(2, 3)
    og = og.reshape(B, heads, frames * patches, dh)
    out = torch.cat((cls_out, og), dim=2)                                          # :137
    out = out.permute(0, 2, 1, 3).reshape(B, N, D)                                 # :140
    return F.linear(out, p[prefix + "proj.weight"], p[prefix + "proj.bias"])      # :142


def space_time_block(x, p, prefix, heads, frames, patches, eps=1e-6):
    """lavila/models/timesformer.py:173-198 ('frozen-in-time' residual; DropPath p=0; QuickGELU MLP)."""
    t = var_attention(layer_norm(x, p[prefix + "norm3.weight"], p[prefix + "norm3.bias"], eps), p,
                      prefix + "timeattn.", heads, "time", frames, patches)       # :180
    if prefix + "alpha_timeattn" in p:                                             # :181-182
        t = torch.tanh(p[prefix + "alpha_timeattn"]) * t
    xt = x + t                                                                     # :183
    s = var_attention(layer_norm(xt, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], eps), p,
                      prefix + "attn.", heads, "space", frames, patches)          # :189
    r = x + s                                                                      # :192 (residual from x, not x+t)
    h = F.linear(layer_norm(r, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], eps),
                 p[prefix + "mlp.fc1.weight"], p[prefix + "mlp.fc1.bias"])        # :53
    h = F.linear(quick_gelu(h), p[prefix + "mlp.fc2.weight"], p[prefix + "mlp.fc2.bias"])  # :54-56
    return r + h                                                                   # :196


def timesformer_features(frames_btchw, p, cfg, prefix="visual.", cls_at_last=True):
    """lavila/models/timesformer.py:345-382.  frames: [B, T, C, H, W]."""
    B, T, C, H, W = frames_btchw.shape
    ps, D = cfg["patch_size"], cfg["embed_dim"]
    w = p[prefix + "patch_embed.proj.weight"]
    bias = p.get(prefix + "patch_embed.proj.bias")
    x = F.conv2d(frames_btchw.reshape(B * T, C, H, W), w, bias, stride=ps)         # :82-83
    n = x.shape[2] * x.shape[3]
    x = x.flatten(2).transpose(1, 2).reshape(B, T * n, D)                          # :349-350
    x = torch.cat((p[prefix + "cls_token"].expand(B, -1, -1), x), dim=1)           # :353-354
    pos = p[prefix + "pos_embed"]
    tile_pos = pos[:, 1:].repeat(1, cfg["num_frames"], 1)                          # :357
    tile_tmp = p[prefix + "temporal_embed"].repeat_interleave(n, dim=1)            # :359
    total = torch.cat((pos[:, :1], tile_pos + tile_tmp), dim=1)                    # :360-361
    x = x + total[:, : x.shape[1]]                                                 # :364
    if cfg.get("ln_pre", True):
        x = layer_norm(x, p[prefix + "ln_pre.weight"], p[prefix + "ln_pre.bias"], 1e-5)  # :366
    for i in range(cfg["depth"]):                                                  # :371-374
        x = space_time_block(x, p, "%sblocks.%d." % (prefix, i), cfg["num_heads"], T, n)
    x = layer_norm(x, p[prefix + "norm.weight"], p[prefix + "norm.bias"], 1e-6)
    return x[:, 0] if cls_at_last else x                                           # :376-382


def timesformer_forward(frames_bcthw, p, cfg, prefix="visual."):
    """lavila/models/timesformer.py:384-390 (head is Identity in the CLIP factories, models.py:347-349)."""
    return timesformer_features(frames_bcthw.permute(0, 2, 1, 3, 4), p, cfg, prefix)


# ----------------------------------------------------------------------------------------------- CLIP text tower
def text_block(x, p, prefix, heads):
    """lavila/models/openai_model.py:182-216: pre-LN nn.MultiheadAttention with additive causal mask + QuickGELU MLP.
    x: [B, L, W] (the reference runs LND; the math is batch-layout independent)."""
    B, L, W = x.shape
    dh = W // heads
    y = layer_norm(x, p[prefix + "ln_1.weight"], p[prefix + "ln_1.bias"], 1e-5)
    qkv = F.linear(y, p[prefix + "attn.in_proj_weight"], p[prefix + "attn.in_proj_bias"])
    q, k, v = (t.reshape(B, L, heads, dh).permute(0, 2, 1, 3) for t in qkv.split(W, dim=-1))
    s = (q * dh ** -0.5) @ k.transpose(-1, -2)
    mask = torch.full((L, L), float("-inf"), device=x.device, dtype=x.dtype).triu(1)   # models.py:131-137
    a = torch.softmax(s + mask, dim=-1) @ v
    a = a.permute(0, 2, 1, 3).reshape(B, L, W)
    x = x + F.linear(a, p[prefix + "attn.out_proj.weight"], p[prefix + "attn.out_proj.bias"])
    y = layer_norm(x, p[prefix + "ln_2.weight"], p[prefix + "ln_2.bias"], 1e-5)
    h = F.linear(y, p[prefix + "mlp.c_fc.weight"], p[prefix + "mlp.c_fc.bias"])
    return x + F.linear(quick_gelu(h), p[prefix + "mlp.c_proj.weight"], p[prefix + "mlp.c_proj.bias"])


def encode_text(text, p, cfg):
    """lavila/models/models.py:150-162."""
    x = p["token_embedding.weight"][text] + p["positional_embedding"]              # :151-152
    for i in range(cfg["text_layers"]):                                            # :154
        x = text_block(x, p, "transformer.resblocks.%d." % i, cfg["text_heads"])
    x = layer_norm(x, p["ln_final.weight"], p["ln_final.bias"], 1e-5)              # :156
    eot = text.argmax(dim=-1)                                                      # :160 (EOT is the largest id)
    return x[torch.arange(x.shape[0], device=x.device), eot] @ p["text_projection"]


def encode_image(image_bcthw, p, cfg):
    """lavila/models/models.py:139-148."""
    return timesformer_forward(image_bcthw, p, cfg) @ p["image_projection"]


def clip_forward(image, text, p, cfg, norm_embed=False):
    """lavila/models/models.py:164-173."""
    ie, te = encode_image(image, p, cfg), encode_text(text, p, cfg)
    if norm_embed:
        ie, te = F.normalize(ie, dim=-1), F.normalize(te, dim=-1)
    return {"image_embed": ie, "text_embed": te, "logit_scale": p["logit_scale"].exp()}


# ----------------------------------------------------------------------------------------------- loss
def clip_loss(all_image, all_text, logit_scale):
    """lavila/models/loss.py:76-79,107-116 on the (already gathered) global batch.
    Returns dict(loss, clip_loss, clip_acc) exactly like CLIPLoss.forward."""
    logits_i = (logit_scale * all_image) @ all_text.t()                            # :78 precedence: (s*I)@T^T
    logits_t = logits_i.t()
    labels = torch.arange(logits_i.shape[0], device=logits_i.device)
    loss = (F.cross_entropy(logits_i, labels) + F.cross_entropy(logits_t, labels)) / 2
    with torch.no_grad():
        acc = 100 * (logits_i.argmax(dim=-1) == labels).sum() / logits_i.shape[0]
    return {"loss": loss, "clip_loss": loss, "clip_acc": acc}


def clip_loss_multi_rank(image_per_rank, text_per_rank, logit_scale):
    """World-size W > 1 with --contrastive-use-vissl (loss.py:74-79 + distributed_utils.py:51-67), evaluated in one
    process: every rank sees the same concatenated batch, so loss/acc are identical on all ranks; the gradient that
    reaches rank r's local embeddings is W x d(loss)/d(embeddings_r) (all_reduce-SUM of identical per-rank grads,
    SURVEY.md 8(a) a13) -- DDP's 1/W parameter averaging restores the global-loss gradient."""
    out = clip_loss(torch.cat(image_per_rank), torch.cat(text_per_rank), logit_scale)
    out["embed_grad_scale"] = float(len(image_per_rank))
    return out


def ssl_clip_loss(all_image, all_text, logit_scale, logit_scale_pseudo, gt_indicators):
    """lavila/models/loss.py:148-213 (SSLCLIPLoss.forward) on the (already gathered) global batch.
    `logit_scale` is the model output (already exp'ed, models.py:173); `logit_scale_pseudo` is the loss module's
    log-parameter (loss.py:140, exp'ed at :152); gt_indicators[i] = 1 for a human narration, 0 for a pseudo one.
    Pair scale: both pseudo -> s_p, mixed -> sqrt(s_p * s), both human -> s  (loss.py:160-164 / :172-176)."""
    sp = logit_scale_pseudo.exp()
    num = gt_indicators.shape[0]
    mask = gt_indicators.repeat(num, 1) + gt_indicators.repeat(num, 1).t()
    mat = torch.ones((num, num), device=all_image.device) * sp
    mat = torch.where(mask == 1, torch.sqrt(sp * logit_scale), mat)
    mat = torch.where(mask == 2, logit_scale * torch.ones_like(mat), mat)
    logits_i = mat * (all_image @ all_text.t())
    logits_t = logits_i.t()                                    # mat is symmetric: equals mat * (T @ I^T) of :178
    labels = torch.arange(num, device=logits_i.device)
    loss = (F.cross_entropy(logits_i, labels) + F.cross_entropy(logits_t, labels)) / 2
    with torch.no_grad():
        pred = logits_i.argmax(dim=-1)
        acc = 100 * pred.eq(labels).sum() / num
        is_gt = gt_indicators == 1
        is_ps = gt_indicators == 0
        num_gt, num_pseudo = int(is_gt.sum()), int(is_ps.sum())
        acc_gt = 100 * pred[is_gt].eq(labels[is_gt]).sum() / num_gt
        acc_pseudo = 100 * pred[is_ps].eq(labels[is_ps]).sum() / num_pseudo
    return {"loss": loss, "clip_loss": loss, "num_gt": torch.tensor([num_gt]), "num_pseudo": torch.tensor([num_pseudo]),
            "clip_acc": acc, "clip_acc_gt": acc_gt, "clip_acc_pseudo": acc_pseudo}


# ----------------------------------------------------------------------------------------------- configs / inputs
def tsf_base_config(num_frames=16, img_size=224):
    """CLIP_OPENAI_TIMESFORMER_BASE -- lavila/models/models.py:316-361."""
    return dict(img_size=img_size, patch_size=16, embed_dim=768, depth=12, num_heads=12, num_frames=num_frames,
                ln_pre=True, text_width=512, text_heads=8, text_layers=12, context_length=77, vocab_size=49408,
                project_dim=256)


def init_params(cfg, seed=0, gated=False, dtype=torch.float32):
    """Random parameters with the reference's names/shapes (SURVEY.md 8(b) checkpoint contract).  Statistics follow
    the reference initialisers loosely (models.py:115-129, timesformer.py:97-103,257-293) but the zero-initialised
    time-attention / temporal embedding are randomised on purpose (SURVEY.md 7.2 'zero-init trap')."""
    g = torch.Generator().manual_seed(seed)
    D, W, E = cfg["embed_dim"], cfg["text_width"], cfg["project_dim"]
    ps, n = cfg["patch_size"], (cfg["img_size"] // cfg["patch_size"]) ** 2

    def rn(*shape, std=0.02):
        return (torch.randn(*shape, generator=g) * std).to(dtype)

    p = {}
    v = "visual."
    p[v + "cls_token"] = rn(1, 1, D)
    p[v + "pos_embed"] = rn(1, n + 1, D)
    p[v + "temporal_embed"] = rn(1, cfg["num_frames"], D)
    p[v + "patch_embed.proj.weight"] = rn(D, 3, ps, ps, std=(3 * ps * ps) ** -0.5)
    p[v + "ln_pre.weight"] = 1 + rn(D, std=0.1)
    p[v + "ln_pre.bias"] = rn(D, std=0.1)
    for i in range(cfg["depth"]):
        b = "%sblocks.%d." % (v, i)
        for nm in ("norm1", "norm2", "norm3"):
            p[b + nm + ".weight"] = 1 + rn(D, std=0.1)
            p[b + nm + ".bias"] = rn(D, std=0.1)
        for a in ("attn.", "timeattn."):
            p[b + a + "qkv.weight"] = rn(3 * D, D, std=D ** -0.5)
            p[b + a + "qkv.bias"] = rn(3 * D, std=0.02)
            p[b + a + "proj.weight"] = rn(D, D, std=D ** -0.5)
            p[b + a + "proj.bias"] = rn(D, std=0.02)
        p[b + "mlp.fc1.weight"] = rn(4 * D, D, std=D ** -0.5)
        p[b + "mlp.fc1.bias"] = rn(4 * D, std=0.02)
        p[b + "mlp.fc2.weight"] = rn(D, 4 * D, std=(4 * D) ** -0.5)
        p[b + "mlp.fc2.bias"] = rn(D, std=0.02)
        if gated:
            p[b + "alpha_timeattn"] = torch.tensor(0.5, dtype=dtype)
    p[v + "norm.weight"] = 1 + rn(D, std=0.1)
    p[v + "norm.bias"] = rn(D, std=0.1)
    L = cfg["text_layers"]
    for i in range(L):
        b = "transformer.resblocks.%d." % i
        p[b + "attn.in_proj_weight"] = rn(3 * W, W, std=W ** -0.5)
        p[b + "attn.in_proj_bias"] = rn(3 * W, std=0.02)
        p[b + "attn.out_proj.weight"] = rn(W, W, std=W ** -0.5 * (2 * L) ** -0.5)
        p[b + "attn.out_proj.bias"] = rn(W, std=0.02)
        for nm in ("ln_1", "ln_2"):
            p[b + nm + ".weight"] = 1 + rn(W, std=0.1)
            p[b + nm + ".bias"] = rn(W, std=0.1)
        p[b + "mlp.c_fc.weight"] = rn(4 * W, W, std=(2 * W) ** -0.5)
        p[b + "mlp.c_fc.bias"] = rn(4 * W, std=0.02)
        p[b + "mlp.c_proj.weight"] = rn(W, 4 * W, std=W ** -0.5 * (2 * L) ** -0.5)
        p[b + "mlp.c_proj.bias"] = rn(W, std=0.02)
    p["token_embedding.weight"] = rn(cfg["vocab_size"], W, std=0.02)
    p["positional_embedding"] = rn(cfg["context_length"], W, std=0.01)
    p["ln_final.weight"] = 1 + rn(W, std=0.1)
    p["ln_final.bias"] = rn(W, std=0.1)
    p["image_projection"] = rn(D, E, std=D ** -0.5)
    p["text_projection"] = rn(W, E, std=W ** -0.5)
    p["logit_scale"] = torch.tensor(math.log(1 / 0.07), dtype=dtype)
    return p


def synthetic_batch(cfg, batch, seed=1234, frames=None):
    """Synthetic inputs: frames fp32 [B,3,T,H,W] ~ N(0,1); tokens int64 [B,ctx] =
    [SOT, U{1..EOT-2} x len, EOT, 0-pad], len ~ U{4..20} capped by the context, EOT = vocab-1 (largest id)."""
    g = torch.Generator().manual_seed(seed)
    T = frames or cfg["num_frames"]
    x = torch.randn(batch, 3, T, cfg["img_size"], cfg["img_size"], generator=g)
    ctx, vocab = cfg["context_length"], cfg["vocab_size"]
    sot, eot = vocab - 2, vocab - 1
    text = torch.zeros(batch, ctx, dtype=torch.int64)
    for b in range(batch):
        ln = int(torch.randint(4, min(20, ctx - 2) + 1, (1,), generator=g))
        text[b, 0] = sot
        text[b, 1:1 + ln] = torch.randint(1, sot, (ln,), generator=g)
        text[b, 1 + ln] = eot
    return x, text
