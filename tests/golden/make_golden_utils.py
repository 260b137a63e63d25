"""Golden vectors of the UNMODIFIED reference lavila/models/utils.py (inflate_positional_embeds, remap_keys).

    python tests/golden/make_golden_utils.py          # writes tests/golden/utils_small.pt
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_shim  # noqa: E402


def clip_visual_state(layers, width, g):
    sd = {"class_embedding": torch.randn(width, generator=g), "positional_embedding": torch.randn(5, width, generator=g),
          "conv1.weight": torch.randn(width, 3, 4, 4, generator=g), "proj": torch.randn(width, 8, generator=g)}
    for n in ("ln_pre", "ln_post"):
        sd[n + ".weight"], sd[n + ".bias"] = torch.randn(width, generator=g), torch.randn(width, generator=g)
    for i in range(layers):
        p = "transformer.resblocks.%d." % i
        sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"] = torch.randn(3 * width, width, generator=g), torch.randn(3 * width, generator=g)
        sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"] = torch.randn(width, width, generator=g), torch.randn(width, generator=g)
        for ln in ("ln_1", "ln_2"):
            sd[p + ln + ".weight"], sd[p + ln + ".bias"] = torch.randn(width, generator=g), torch.randn(width, generator=g)
        sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"] = torch.randn(4 * width, width, generator=g), torch.randn(4 * width, generator=g)
        sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"] = torch.randn(width, 4 * width, generator=g), torch.randn(width, generator=g)
    return sd


def main():
    assert reference_shim.install(), "reference not found"
    from lavila.models.utils import inflate_positional_embeds, remap_keys
    g = torch.Generator().manual_seed(0)
    out = {"inflate": [], "seed": 0}
    for t_ckpt, t_model, fix in ((4, 16, "bilinear"), (4, 16, "interp"), (4, 8, "zeros"), (16, 4, "bilinear"), (4, 4, "bilinear")):
        ckpt = {"visual.temporal_embed": torch.randn(1, t_ckpt, 12, generator=g), "visual.pos_embed": torch.randn(1, 5, 12, generator=g)}
        cur = {"visual.temporal_embed": torch.zeros(1, t_model, 12), "visual.pos_embed": torch.zeros(1, 5, 12)}
        src = ckpt["visual.temporal_embed"].clone()
        res = inflate_positional_embeds(cur, dict(ckpt), num_frames=t_model, load_temporal_fix=fix)
        out["inflate"].append({"t_ckpt": t_ckpt, "t_model": t_model, "fix": fix, "input": src,
                               "output": res["visual.temporal_embed"].clone()})
    sd = clip_visual_state(2, 8, torch.Generator().manual_seed(1))
    rem = remap_keys({k: v.clone() for k, v in sd.items()}, transformer_layers=2)
    out["remap"] = {"layers": 2, "width": 8, "keys": list(rem.keys()), "shapes": {k: tuple(v.shape) for k, v in rem.items()},
                    "checksums": {k: float(v.double().sum()) for k, v in rem.items()}}
    # state_dict contract of the reference CLIP_HF + DistilBERT (models.py:176-290) at toy size
    import torch.nn as nn
    from transformers import DistilBertConfig, DistilBertModel
    from lavila.models.models import CLIP_HF
    from lavila.models.openai_model import QuickGELU
    from lavila.models.timesformer import SpaceTimeTransformer
    vis = SpaceTimeTransformer(img_size=32, patch_size=16, embed_dim=64, depth=1, num_heads=1, num_frames=2, time_init='zeros',
                               attention_style='frozen-in-time', ln_pre=True, act_layer=QuickGELU)
    vis.head = vis.pre_logits = vis.fc = nn.Identity()
    ref = CLIP_HF(embed_dim=16, vision_width=64, vision_model=vis, text_width=32,
                  text_model=DistilBertModel(DistilBertConfig(vocab_size=100, dim=32, n_layers=1, n_heads=2, hidden_dim=64,
                                                              max_position_embeddings=16)),
                  text_use_cls_token=True, text_is_regressive=False)
    out["clip_hf_state"] = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "utils_small.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
