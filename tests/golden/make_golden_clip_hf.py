"""Golden vectors of the UNMODIFIED reference CLIP_HF (lavila/models/models.py:176-290) with a toy DistilBERT text tower
-> tests/golden/clip_hf_small.pt: inputs, the DistilBERT state_dict (random init, stored: HF initialisers are not ours to
regenerate), outputs for norm_embed True, loss and gradients (CLIPLoss, world size 1).

    python tests/golden/make_golden_clip_hf.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_shim  # noqa: E402
from oracle.dual_encoder import init_params, synthetic_batch  # noqa: E402

CFG = dict(img_size=32, patch_size=16, embed_dim=128, depth=2, num_heads=2, num_frames=4, ln_pre=True, text_width=128, text_heads=2,
           text_layers=1, context_length=16, vocab_size=512, project_dim=64)
BERT = dict(vocab_size=200, dim=128, n_layers=2, n_heads=2, hidden_dim=256, max_position_embeddings=32, dropout=0.0,
            attention_dropout=0.0)


def text_inputs(B=3, L=12, seed=6):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1, BERT["vocab_size"], (B, L), generator=g)
    mask = torch.ones(B, L, dtype=torch.int64)
    mask[1, 8:] = 0
    mask[2, 5:] = 0
    return ids, mask


def main():
    assert reference_shim.install(), "reference not found"
    import torch.nn as nn
    from transformers import DistilBertConfig, DistilBertModel
    from lavila.models.loss import CLIPLoss
    from lavila.models.models import CLIP_HF
    from lavila.models.openai_model import QuickGELU
    from lavila.models.timesformer import SpaceTimeTransformer
    torch.manual_seed(0)
    params = init_params(CFG, seed=41)
    vis = SpaceTimeTransformer(img_size=CFG["img_size"], patch_size=CFG["patch_size"], embed_dim=CFG["embed_dim"], depth=CFG["depth"],
                               num_heads=CFG["num_heads"], num_frames=CFG["num_frames"], time_init="zeros",
                               attention_style="frozen-in-time", ln_pre=True, act_layer=QuickGELU)
    vis.head = vis.pre_logits = vis.fc = nn.Identity()
    bert = DistilBertModel(DistilBertConfig(**BERT)).eval()
    model = CLIP_HF(embed_dim=CFG["project_dim"], vision_width=CFG["embed_dim"], vision_model=vis, text_width=BERT["dim"],
                    text_model=bert, text_use_cls_token=True, text_is_regressive=False)
    res = model.visual.load_state_dict({k[len("visual."):]: v for k, v in params.items() if k.startswith("visual.")}, strict=False)
    assert not res.unexpected_keys
    with torch.no_grad():
        model.image_projection.copy_(params["image_projection"])
    frames, _ = synthetic_batch(CFG, 3, seed=1234)
    ids, mask = text_inputs()
    out = model(frames, ids, mask=mask, norm_embed=True)
    ld = CLIPLoss()(out)
    ld["loss"].backward()
    grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    keep = ["text_projection", "image_projection", "logit_scale", "visual.blocks.0.attn.qkv.weight", "visual.blocks.1.mlp.fc2.weight",
            "visual.temporal_embed", "textual.embeddings.word_embeddings.weight", "textual.transformer.layer.1.ffn.lin2.weight"]
    gold = {"cfg": CFG, "bert": BERT, "param_seed": 41, "bert_state": {k: v.clone() for k, v in bert.state_dict().items()},
            "text_projection": model.text_projection.detach().clone(),
            "image_embed": out["image_embed"].detach().clone(), "text_embed": out["text_embed"].detach().clone(),
            "loss": ld["loss"].detach().clone(), "clip_acc": ld["clip_acc"].detach().clone(), "grads": {k: grads[k] for k in keep}}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "clip_hf_small.pt")
    torch.save(gold, path)
    print("wrote", path, os.path.getsize(path), "loss", float(ld["loss"]))


if __name__ == "__main__":
    main()
