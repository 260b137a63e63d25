"""Narrator host logic on the CPU with the kernel test doubles (tests/ops_doubles.py): VCLM_HF.encode_image / forward, the gated
GPT-2's cross-attention schedule (every 1st / 2nd / 3rd layer), KV-cached incremental decoding and the sampling loop, against
golden vectors of the unmodified reference.  The CUDA-graph replay path needs a GPU (tests/test_gpu_narrator.py)."""
import os
from types import SimpleNamespace

import pytest
import torch

from oracle import narrator as ON
from oracle.dual_encoder import synthetic_batch
from tests import ops_doubles
from tests.util import cosine, rel_l2

SMALL = torch.load(os.path.join(os.path.dirname(__file__), "golden", "narrator_small.pt"), weights_only=False)
EXTRA = torch.load(os.path.join(os.path.dirname(__file__), "golden", "narrator_extra.pt"), weights_only=False)
CASES = {"small": dict(SMALL, param_seed=0), "p14_freq3": EXTRA["p14_freq3"], "freq1": EXTRA["freq1"]}


def _build(cfg, params):
    from lavila_b200.models.gpt2_gated import GPT2LMHeadModel, augment_gpt2_config
    from lavila_b200.models.narrator import VCLM_HF
    from lavila_b200.models.timesformer import QuickGELU, SpaceTimeTransformer
    v = cfg["visual"]
    vis = SpaceTimeTransformer(img_size=v["img_size"], patch_size=v["patch_size"], embed_dim=v["embed_dim"], depth=v["depth"],
                               num_heads=v["num_heads"], num_frames=v["num_frames"], time_init="zeros", ln_pre=True,
                               act_layer=QuickGELU)
    vis.head = torch.nn.Identity()
    vis.pre_logits = torch.nn.Identity()
    g = SimpleNamespace(vocab_size=cfg["vocab_size"], n_positions=cfg["n_positions"], n_embd=cfg["n_embd"], n_layer=cfg["n_layer"],
                        n_head=cfg["n_head"], layer_norm_epsilon=1e-5, activation_function="gelu_new")
    dec = GPT2LMHeadModel(augment_gpt2_config(g, cross_attn_freq=cfg["cross_attn_freq"], gated_xattn=True))
    m = VCLM_HF(vision_width=v["embed_dim"], vision_model=vis, text_width=cfg["n_embd"], text_decoder=dec,
                num_img_queries=cfg["num_img_queries"], dim_head=64, heads=cfg["pool_heads"])
    res = m.load_state_dict(params, strict=False)
    assert not res.unexpected_keys
    return m.eval()


def _setup(case, monkeypatch):
    ops_doubles.install(monkeypatch)
    c = CASES[case]
    cfg = c["cfg"]
    model = _build(cfg, ON.init_narrator_params(cfg, seed=c["param_seed"]))
    frames, _ = synthetic_batch(dict(cfg["visual"], context_length=8, vocab_size=8), 2, seed=c["frames_seed"])
    return c, cfg, model, frames


@pytest.mark.parametrize("case", list(CASES))
def test_narrator_host_logic_matches_reference_golden(case, monkeypatch):
    c, cfg, m, frames = _setup(case, monkeypatch)
    tok = m.encode_image(frames)
    assert rel_l2(tok, c["image_tokens"]) < 2e-2 and cosine(tok, c["image_tokens"]) > 0.999
    out = m(frames, c["text"])
    assert torch.equal(out["labels"], c["labels"])
    assert rel_l2(out["text_tokens_logits"], c["logits"]) < 3e-2 and cosine(out["text_tokens_logits"], c["logits"]) > 0.999


@pytest.mark.parametrize("case", ["small", "p14_freq3"])
def test_kv_cached_decoding_host_logic(case, monkeypatch):
    """Prefill + one-position-at-a-time decoding through the per-layer KV caches == the full-prefix forward; generate() gives the
    same ids with and without the cache (same doubles, same arithmetic order -> exact)."""
    c, cfg, m, frames = _setup(case, monkeypatch)
    tok = m.encode_image(frames)
    ids = c["text"][:, :6].contiguous()
    full = m.text_decoder(ids, encoder_hidden_states=tok).logits
    cache, ctx = {"max_len": 8}, {}
    got = [m.text_decoder(ids[:, :3].contiguous(), encoder_hidden_states=tok, ctx_kv_cache=ctx, self_kv_cache=cache, past_len=0).logits]
    for t in range(3, 6):
        got.append(m.text_decoder(ids[:, t:t + 1].contiguous(), encoder_hidden_states=tok, ctx_kv_cache=ctx, self_kv_cache=cache,
                                  past_len=t).logits)
    assert rel_l2(torch.cat(got, 1), full) < 1e-5
    # the device-scalar ("dyn") step used by the CUDA-graph path computes the same thing
    dyn = {"pos_idx": torch.tensor([5]), "lk_dev": torch.tensor([6], dtype=torch.int32)}
    step = m.text_decoder(ids[:, 5:6].contiguous(), encoder_hidden_states=tok, ctx_kv_cache=ctx, self_kv_cache=cache, dyn=dyn).logits
    assert rel_l2(step[:, 0], full[:, 5]) < 1e-5
    t = SimpleNamespace(bos_token_id=cfg["vocab_size"] - 1, eos_token_id=cfg["vocab_size"] - 1, pad_token_id=0)
    outs = []
    for use in (False, True):
        torch.manual_seed(5)
        outs.append(m.generate(tok, t, max_text_length=7, top_p=0.95, temperature=0.7, num_return_sequences=2, use_kv_cache=use))
    assert torch.equal(outs[0][0], outs[1][0])
    assert outs[0][0].shape == (4, 7) and outs[0][0].dtype == torch.int64 and bool(torch.isfinite(outs[0][1]).all())


def test_use_half_is_accepted(monkeypatch):
    """main_infer_narrator.py:155-170 `--use-half`: model.half() + half-precision frames must work.  Here the parameters stay fp32
    (the kernels own the precision) and half inputs are widened on entry: same tokens up to the fp16 rounding of the frames."""
    c, cfg, m, frames = _setup("small", monkeypatch)
    ref = m.encode_image(frames)
    assert m.half() is m and all(p.dtype == torch.float32 for p in m.parameters())
    tok = m.encode_image(frames.half())
    assert tok.dtype == torch.float32 and rel_l2(tok, ref) < 5e-3
