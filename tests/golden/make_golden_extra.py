"""More golden cases from the UNMODIFIED reference for the oracle (same fields as make_golden.py), written to
tests/golden/dual_encoder_extra.pt:
  * "p14"      : patch 14 (the TSF-L/14 patch size; 28 px image -> 4 patches per frame), 4 heads, depth 3, gated, batch 2
  * "fewframes": a model built for 8 frames fed 4-frame clips (timesformer.py:356-362 slices temporal_embed)
  * "wide_text": text tower 3 layers x 4 heads, context length 20

    python tests/golden/make_golden_extra.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_shim  # noqa: E402
from tests.golden import make_golden as MG  # noqa: E402
from oracle.dual_encoder import init_params, synthetic_batch  # noqa: E402

CASES = {
    "p14": (dict(img_size=28, patch_size=14, embed_dim=256, depth=3, num_heads=4, num_frames=4, ln_pre=True, text_width=128,
                 text_heads=2, text_layers=2, context_length=16, vocab_size=512, project_dim=64), dict(gated=True, norm_embed=True, batch=2, seed=11)),
    "wide_text": (dict(img_size=32, patch_size=16, embed_dim=128, depth=1, num_heads=2, num_frames=2, ln_pre=True, text_width=256,
                       text_heads=4, text_layers=3, context_length=20, vocab_size=700, project_dim=32), dict(gated=False, norm_embed=True, batch=4, seed=12)),
}


def run_fewframes():
    """8-frame model, 4-frame input."""
    from lavila.models.loss import CLIPLoss
    cfg = dict(img_size=32, patch_size=16, embed_dim=128, depth=2, num_heads=2, num_frames=8, ln_pre=True, text_width=128,
               text_heads=2, text_layers=1, context_length=16, vocab_size=512, project_dim=64)
    params = init_params(cfg, seed=13)
    model = MG.build_reference(cfg, params, False)
    frames, text = synthetic_batch(cfg, 3, seed=1234, frames=4)
    out = model(frames, text, norm_embed=True)
    ld = CLIPLoss()(out)
    ld["loss"].backward()
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    return {"cfg": cfg, "gated": False, "norm_embed": True, "batch": 3, "param_seed": 13, "input_seed": 1234, "frames": 4,
            "param_checksum": MG.param_checksum(params), "frames_checksum": float(frames.double().sum()), "text": text.clone(),
            "image_embed": out["image_embed"].detach().clone(), "text_embed": out["text_embed"].detach().clone(),
            "logit_scale": out["logit_scale"].detach().clone(), "loss": ld["loss"].detach().clone(),
            "clip_acc": ld["clip_acc"].detach().clone(), "grads": MG.summarise_grads(grads)}


def main():
    assert reference_shim.install(), "reference not found"
    torch.manual_seed(0)
    torch.set_num_threads(4)
    out = {k: MG.run_case(cfg, **kw) for k, (cfg, kw) in CASES.items()}
    out["fewframes"] = run_fewframes()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dual_encoder_extra.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k, c in out.items():
        print(k, "loss", float(c["loss"]), "acc", float(c["clip_acc"]))


if __name__ == "__main__":
    main()
