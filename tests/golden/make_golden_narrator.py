"""Golden vectors for the narrator path from the UNMODIFIED reference (VCLM_HF + gated GPT2LMHeadModel + coca pool).

    python tests/golden/make_golden_narrator.py     # writes tests/golden/narrator_small.pt
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_shim  # noqa: E402
from oracle.narrator import SMALL_NARRATOR, init_narrator_params  # noqa: E402
from oracle.dual_encoder import synthetic_batch  # noqa: E402


def build_reference(cfg, params):
    import torch.nn as nn
    from transformers import GPT2Config
    from lavila.models.gpt2_gated import GPT2LMHeadModel as Gated, augment_gpt2_config
    from lavila.models.narrator import VCLM_HF
    from lavila.models.openai_model import QuickGELU
    from lavila.models.timesformer import SpaceTimeTransformer
    v = cfg["visual"]
    vis = SpaceTimeTransformer(img_size=v["img_size"], patch_size=v["patch_size"], embed_dim=v["embed_dim"], depth=v["depth"],
                               num_heads=v["num_heads"], num_frames=v["num_frames"], time_init="zeros",
                               attention_style="frozen-in-time", ln_pre=True, act_layer=QuickGELU)
    vis.head = nn.Identity(); vis.pre_logits = nn.Identity(); vis.fc = nn.Identity()
    gcfg = GPT2Config(vocab_size=cfg["vocab_size"], n_positions=cfg["n_positions"], n_embd=cfg["n_embd"],
                      n_layer=cfg["n_layer"], n_head=cfg["n_head"])
    dec = Gated(augment_gpt2_config(gcfg, cross_attn_freq=cfg["cross_attn_freq"], gated_xattn=True))
    model = VCLM_HF(vision_width=v["embed_dim"], vision_model=vis, text_width=cfg["n_embd"], text_decoder=dec,
                    num_img_queries=cfg["num_img_queries"], dim_head=64, heads=cfg["pool_heads"])
    sd = model.state_dict()
    missing = [k for k in sd if k not in params and not k.endswith(("attn.bias", "attn.masked_bias", "crossattention.bias",
                                                                    "crossattention.masked_bias", ".beta"))]
    assert not missing, missing
    res = model.load_state_dict(params, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    model.text_decoder.lm_head.weight = model.text_decoder.transformer.wte.weight   # tie (see SURVEY 8c caveat)
    return model.eval()


def main():
    assert reference_shim.install()
    cfg = SMALL_NARRATOR
    params = init_narrator_params(cfg, seed=0)
    model = build_reference(cfg, params)
    vcfg = dict(cfg["visual"], context_length=8, vocab_size=8)
    frames, _ = synthetic_batch(vcfg, 2, seed=77)
    g = torch.Generator().manual_seed(5)
    text = torch.randint(0, cfg["vocab_size"], (2, 12), generator=g)
    with torch.no_grad():
        tokens = model.encode_image(frames)
        out = model(frames, text)
        dec = model.text_decoder(text[:, :5].contiguous(), encoder_hidden_states=tokens).logits
    torch.manual_seed(0)

    class Tok:
        bos_token_id = eos_token_id = cfg["vocab_size"] - 1
        pad_token_id = 0
    ids, ppl = model.generate(tokens, Tok(), max_text_length=6, top_p=0.95, temperature=0.7, num_return_sequences=2)
    gold = {"cfg": cfg, "frames_seed": 77, "text": text, "image_tokens": tokens, "logits": out["text_tokens_logits"],
            "labels": out["labels"], "logits_prefix5": dec, "gen_shape": tuple(ids.shape), "ppl_shape": tuple(ppl.shape),
            "param_checksum": {k: float(v.double().sum()) for k, v in params.items()}}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "narrator_small.pt")
    torch.save(gold, path)
    print("wrote", path, os.path.getsize(path), "gen", ids.shape, ppl)


if __name__ == "__main__":
    main()
