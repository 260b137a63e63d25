"""Two more narrator goldens from the UNMODIFIED reference -> tests/golden/narrator_extra.pt:
  "p14_freq3": patch-14 video encoder (TSF-L/14's patch size), 3 decoder layers, cross-attention every 3rd layer
               (VCLM_OPENAI_TIMESFORMER_LARGE_336PX_GPT2_XL uses cross_attn_freq = 3, models.py:1170), 3 pooling heads
  "freq1"    : cross-attention in every layer (the *_GPT2 factories, models.py:918 / :1107)

    python tests/golden/make_golden_narrator_extra.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_shim  # noqa: E402
from oracle.narrator import init_narrator_params  # noqa: E402
from oracle.dual_encoder import synthetic_batch  # noqa: E402
from tests.golden.make_golden_narrator import build_reference  # noqa: E402

CASES = {
    "p14_freq3": dict(visual=dict(img_size=28, patch_size=14, embed_dim=192, depth=2, num_heads=3, num_frames=2, ln_pre=True),
                      n_embd=192, n_head=3, n_layer=3, cross_attn_freq=3, vocab_size=300, n_positions=24, num_img_queries=6, pool_heads=3),
    "freq1": dict(visual=dict(img_size=32, patch_size=16, embed_dim=128, depth=1, num_heads=2, num_frames=3, ln_pre=True),
                  n_embd=128, n_head=2, n_layer=2, cross_attn_freq=1, vocab_size=400, n_positions=16, num_img_queries=5, pool_heads=2),
}


def main():
    assert reference_shim.install()
    out = {}
    for name, cfg in CASES.items():
        params = init_narrator_params(cfg, seed=3)
        model = build_reference(cfg, params)
        vcfg = dict(cfg["visual"], context_length=8, vocab_size=8)
        frames, _ = synthetic_batch(vcfg, 2, seed=91)
        text = torch.randint(0, cfg["vocab_size"], (2, 10), generator=torch.Generator().manual_seed(8))
        with torch.no_grad():
            tokens = model.encode_image(frames)
            res = model(frames, text)
        out[name] = {"cfg": cfg, "frames_seed": 91, "param_seed": 3, "text": text, "image_tokens": tokens,
                     "logits": res["text_tokens_logits"], "labels": res["labels"],
                     "param_checksum": {k: float(v.double().sum()) for k, v in params.items()}}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "narrator_extra.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
