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
