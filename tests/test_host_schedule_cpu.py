"""The product's HOST logic on a machine without a GPU: lavila_b200.models + lavila_b200/engine.py (autograd Functions: what is
saved, gradient routing and accumulation, strides, CLS-only tail, use_checkpoint recompute) run with the kernel wrappers replaced
by the fp32 torch test doubles of tests/ops_doubles.py, and are compared with golden vectors of the UNMODIFIED reference.  What
this does not test is the kernels themselves: that is tests/test_gpu_*.py on a B200."""
import os

import pytest
import torch

from oracle import dual_encoder as O
from tests import ops_doubles
from tests.util import cosine, rel_l2

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "dual_encoder_small.pt"), weights_only=False)
EXTRA = torch.load(os.path.join(os.path.dirname(__file__), "golden", "dual_encoder_extra.pt"), weights_only=False)
CASES = {"plain": GOLD["plain"], "norm": GOLD["norm"], "gated_norm": GOLD["gated_norm"], "p14": EXTRA["p14"],
         "wide_text": EXTRA["wide_text"], "fewframes": EXTRA["fewframes"]}


def _build(cfg, params, gated):
    from lavila_b200.models.models import CLIP
    from lavila_b200.models.timesformer import QuickGELU, SpaceTimeTransformer
    vis = SpaceTimeTransformer(img_size=cfg["img_size"], patch_size=cfg["patch_size"], embed_dim=cfg["embed_dim"],
                               depth=cfg["depth"], num_heads=cfg["num_heads"], num_frames=cfg["num_frames"], time_init="zeros",
                               ln_pre=True, act_layer=QuickGELU, is_tanh_gating=gated)
    vis.head = torch.nn.Identity()
    vis.pre_logits = torch.nn.Identity()
    m = CLIP(embed_dim=cfg["project_dim"], vision_width=cfg["embed_dim"], vision_model=vis, context_length=cfg["context_length"],
             vocab_size=cfg["vocab_size"], transformer_width=cfg["text_width"], transformer_heads=cfg["text_heads"],
             transformer_layers=cfg["text_layers"])
    res = m.load_state_dict(params, strict=False)
    assert not res.unexpected_keys and not res.missing_keys, res
    return m


def _run(c, monkeypatch, cls_only_tail=True, use_checkpoint=False):
    from lavila_b200.models.loss import CLIPLoss
    ops_doubles.install(monkeypatch)
    cfg = c["cfg"]
    params = O.init_params(cfg, seed=c["param_seed"], gated=c["gated"])
    model = _build(cfg, params, c["gated"])
    model.visual.cls_only_tail = cls_only_tail
    frames, text = O.synthetic_batch(cfg, c["batch"], seed=c["input_seed"], frames=c.get("frames"))
    out = model(frames, text, use_checkpoint=use_checkpoint, norm_embed=c["norm_embed"])
    ld = CLIPLoss()(out)
    ld["loss"].backward()
    return model, out, ld


@pytest.mark.parametrize("case", list(CASES))
def test_host_schedule_matches_reference_golden(case, monkeypatch):
    c = CASES[case]
    model, out, ld = _run(c, monkeypatch)
    assert rel_l2(out["image_embed"], c["image_embed"]) < 2e-2 and cosine(out["image_embed"], c["image_embed"]) > 0.999
    assert rel_l2(out["text_embed"], c["text_embed"]) < 2e-2 and cosine(out["text_embed"], c["text_embed"]) > 0.999
    assert abs(float(out["logit_scale"]) - float(c["logit_scale"])) < 1e-4
    if not c["norm_embed"]:
        return
    assert abs(float(ld["loss"]) - float(c["loss"])) < 3e-2
    named = dict(model.named_parameters())
    for name, ref in c["grads"].items():
        g = named[name].grad
        assert g is not None, name
        got, want = (g.flatten(), ref["full"].flatten()) if "full" in ref else (g.flatten()[ref["idx"]], ref["sample"])
        if float(want.norm()) < 1e-7 or want.numel() == 1:
            continue
        assert cosine(got, want) > 0.99 and rel_l2(got, want) < 6e-2, "%s: rel_l2 %.3e" % (name, rel_l2(got, want))


def test_full_last_block_and_recompute_paths(monkeypatch):
    """The same numbers through the full last block (cls_only_tail off) and with use_checkpoint=True (block recompute)."""
    c = CASES["gated_norm"]
    base_model, base_out, _ = _run(c, monkeypatch)
    for kw in (dict(cls_only_tail=False), dict(cls_only_tail=False, use_checkpoint=True), dict(use_checkpoint=True)):
        model, out, _ = _run(c, monkeypatch, **kw)
        assert rel_l2(out["image_embed"], base_out["image_embed"]) < 1e-2
        for (n1, p1), (n2, p2) in zip(base_model.named_parameters(), model.named_parameters()):
            if p1.grad is None or float(p1.grad.norm()) < 1e-7 or p1.numel() == 1:
                continue
            assert rel_l2(p2.grad, p1.grad) < 4e-2, (kw, n1, rel_l2(p2.grad, p1.grad))


SSL = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ssl_loss_small.pt"), weights_only=False)


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_sslcliploss_module_host_logic(idx, monkeypatch):
    """The SSLCLIPLoss module (exp of the pseudo scale, gt handling, result dict, autograd plumbing of both scales) on the doubles,
    against the unmodified reference: incl. the all-human / all-pseudo cases whose empty class gives a NaN accuracy."""
    from lavila_b200.models.loss import SSLCLIPLoss
    ops_doubles.install(monkeypatch)
    c = SSL["world1"][idx]
    crit = SSLCLIPLoss(scale_init=SSL["scale_init"])
    img, txt = c["image"].clone().requires_grad_(True), c["text"].clone().requires_grad_(True)
    s = torch.tensor(SSL["scale"], requires_grad=True)
    out = crit({"image_embed": img, "text_embed": txt, "logit_scale": s}, c["gt"])
    gi, gt, gs, gp = torch.autograd.grad(out["loss"], (img, txt, s, crit.logit_scale_pseudo))
    torch.testing.assert_close(out["loss"], c["loss"], rtol=1e-5, atol=1e-6)
    for k in ("clip_acc", "clip_acc_gt", "clip_acc_pseudo"):
        a, b = float(out[k]), float(c[k])
        assert (a != a and b != b) or abs(a - b) < 1e-3, (k, a, b)
    assert torch.equal(out["num_gt"], c["num_gt"]) and torch.equal(out["num_pseudo"], c["num_pseudo"])
    torch.testing.assert_close(gi, c["grad_image"], rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(gt, c["grad_text"], rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(gs, c["grad_scale"], rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(gp, c["grad_scale_pseudo"], rtol=1e-4, atol=1e-7)


def test_clip_hf_host_logic_matches_reference_golden(monkeypatch):
    """CLIP_HF (video tower / projections / normalise / loss through the engine on the doubles, DistilBERT through the HF module)
    against the unmodified reference CLIP_HF with the same DistilBERT weights (tests/golden/make_golden_clip_hf.py)."""
    from transformers import DistilBertConfig, DistilBertModel
    from lavila_b200.models.loss import CLIPLoss
    from lavila_b200.models.models import CLIP_HF
    from lavila_b200.models.timesformer import QuickGELU, SpaceTimeTransformer
    from tests.golden.make_golden_clip_hf import text_inputs
    ops_doubles.install(monkeypatch)
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "clip_hf_small.pt"), weights_only=False)
    cfg = g["cfg"]
    params = O.init_params(cfg, seed=g["param_seed"])
    bert = DistilBertModel(DistilBertConfig(**g["bert"])).eval()
    bert.load_state_dict(g["bert_state"])
    vis = SpaceTimeTransformer(img_size=cfg["img_size"], patch_size=cfg["patch_size"], embed_dim=cfg["embed_dim"], depth=cfg["depth"],
                               num_heads=cfg["num_heads"], num_frames=cfg["num_frames"], time_init="zeros", ln_pre=True,
                               act_layer=QuickGELU)
    vis.head = torch.nn.Identity()
    vis.pre_logits = torch.nn.Identity()
    m = CLIP_HF(embed_dim=cfg["project_dim"], vision_width=cfg["embed_dim"], vision_model=vis, text_width=g["bert"]["dim"],
                text_model=bert, text_use_cls_token=True, text_is_regressive=False)
    assert not m.visual.load_state_dict({k[len("visual."):]: v for k, v in params.items() if k.startswith("visual.")},
                                        strict=False).unexpected_keys
    with torch.no_grad():
        m.image_projection.copy_(params["image_projection"])
        m.text_projection.copy_(g["text_projection"])
    frames, _ = O.synthetic_batch(cfg, 3, seed=1234)
    ids, mask = text_inputs()
    out = m(frames, ids, mask=mask, norm_embed=True)
    ld = CLIPLoss()(out)
    ld["loss"].backward()
    assert rel_l2(out["image_embed"], g["image_embed"]) < 2e-2 and rel_l2(out["text_embed"], g["text_embed"]) < 2e-2
    assert abs(float(ld["loss"]) - float(g["loss"])) < 3e-2 and abs(float(ld["clip_acc"]) - float(g["clip_acc"])) < 1e-3
    named = dict(m.named_parameters())
    for name, want in g["grads"].items():
        got = named[name].grad
        if want.numel() == 1:
            assert abs(float(got) - float(want)) < 0.1 * abs(float(want)) + 1e-3, name
        else:
            assert cosine(got, want) > 0.99 and rel_l2(got, want) < 6e-2, "%s: rel_l2 %.3e" % (name, rel_l2(got, want))


def test_training_loop_body_reduces_the_loss(monkeypatch):
    """The body of main_pretrain.py's train() (:486-530: zero_grad -> model(frames, tokens, use_checkpoint, norm_embed) -> criterion
    -> backward -> AdamW.step over the driver's weight-decay groups -> logit_scale clamp) with this package's modules: a few steps
    on one fixed toy batch must drive the contrastive loss down (gradient signs and routing are right end to end)."""
    import bench
    from lavila_b200.models.loss import CLIPLoss
    ops_doubles.install(monkeypatch)
    cfg = GOLD["norm"]["cfg"]
    model = _build(cfg, O.init_params(cfg, seed=2), gated=False)
    crit = CLIPLoss(use_vissl=False, cache_labels=True, rank=0, world_size=1)
    opt = torch.optim.AdamW(bench.param_groups(model), lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    frames, text = O.synthetic_batch(cfg, 4, seed=3)
    losses = []
    for it in range(8):
        opt.zero_grad(set_to_none=True)
        out = model(frames, text, use_checkpoint=(it % 2 == 1), norm_embed=True)
        ld = crit(out)
        ld["loss"].backward()
        opt.step()
        model.logit_scale.data.clamp_(0, 4.6052)
        losses.append(float(ld["loss"].detach()))
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < 0.6 * losses[0], losses


EDGE_CONFIGS = {
    "depth1_only_cls_tail": dict(img_size=32, patch_size=16, embed_dim=64, depth=1, num_heads=1, num_frames=2, text_width=64, text_heads=1,
                                 text_layers=1, context_length=7, vocab_size=64, project_dim=32),
    "one_frame": dict(img_size=32, patch_size=16, embed_dim=128, depth=2, num_heads=2, num_frames=1, text_width=64, text_heads=1,
                      text_layers=2, context_length=9, vocab_size=128, project_dim=16),
    "one_patch_per_frame": dict(img_size=16, patch_size=16, embed_dim=64, depth=2, num_heads=1, num_frames=5, text_width=128, text_heads=2,
                                text_layers=1, context_length=12, vocab_size=100, project_dim=24),
    "nine_patches_three_heads": dict(img_size=48, patch_size=16, embed_dim=192, depth=3, num_heads=3, num_frames=3, text_width=192,
                                     text_heads=3, text_layers=2, context_length=10, vocab_size=90, project_dim=48),
}


@pytest.mark.parametrize("name", list(EDGE_CONFIGS))
@pytest.mark.parametrize("gated", [False, True])
def test_host_schedule_edge_geometries_vs_oracle(name, gated, monkeypatch):
    """Corner geometries (a single block = only the CLS tail; one frame; one patch per frame; 3 heads / 9 patches) against the
    oracle (itself pinned to the reference): outputs, loss and all gradients."""
    from lavila_b200.models.loss import CLIPLoss
    ops_doubles.install(monkeypatch)
    cfg = dict(EDGE_CONFIGS[name], ln_pre=True)
    params = O.init_params(cfg, seed=17, gated=gated)
    model = _build(cfg, params, gated)
    B = 3
    frames, text = O.synthetic_batch(cfg, B, seed=23)
    out = model(frames, text, norm_embed=True)
    ld = CLIPLoss()(out)
    ld["loss"].backward()
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ref = O.clip_forward(frames, text, pr, cfg, norm_embed=True)
    rl = O.clip_loss(ref["image_embed"], ref["text_embed"], ref["logit_scale"])
    rl["loss"].backward()
    assert rel_l2(out["image_embed"], ref["image_embed"]) < 2e-2 and rel_l2(out["text_embed"], ref["text_embed"]) < 2e-2
    assert abs(float(ld["loss"].detach()) - float(rl["loss"].detach())) < 3e-2
    for n, p in model.named_parameters():
        want = pr[n].grad
        if want is None or float(want.norm()) < 1e-7 or want.numel() == 1:
            continue
        assert p.grad is not None, n
        assert cosine(p.grad, want) > 0.99 and rel_l2(p.grad, want) < 8e-2, "%s: rel_l2 %.3e cos %.4f" % (n, rel_l2(p.grad, want), cosine(p.grad, want))
