"""The narrator oracle (oracle/narrator.py) against golden vectors from the unmodified reference.  CPU only."""
import os

import torch

from oracle import narrator as ON
from oracle.dual_encoder import synthetic_batch

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "narrator_small.pt"), weights_only=False)
TOL = dict(rtol=3e-4, atol=3e-5)


def _setup():
    cfg = GOLD["cfg"]
    p = ON.init_narrator_params(cfg, seed=0)
    for k, v in GOLD["param_checksum"].items():
        assert abs(float(p[k].double().sum()) - v) <= 1e-6 * max(1.0, abs(v)), k
    vcfg = dict(cfg["visual"], context_length=8, vocab_size=8)
    frames, _ = synthetic_batch(vcfg, 2, seed=GOLD["frames_seed"])
    return cfg, p, frames


def test_encode_image_matches_reference():
    cfg, p, frames = _setup()
    tok = ON.vclm_encode_image(frames, p, cfg)
    torch.testing.assert_close(tok, GOLD["image_tokens"], **TOL)


def test_teacher_forced_logits_match_reference():
    cfg, p, frames = _setup()
    out = ON.vclm_forward(frames, GOLD["text"], p, cfg)
    torch.testing.assert_close(out["text_tokens_logits"], GOLD["logits"], rtol=1e-3, atol=1e-4)
    assert torch.equal(out["labels"], GOLD["labels"])
    tok = ON.vclm_encode_image(frames, p, cfg)
    pre = ON.gpt2_lm_logits(GOLD["text"][:, :5], tok, p, cfg)
    torch.testing.assert_close(pre, GOLD["logits_prefix5"], rtol=1e-3, atol=1e-4)


def test_warper_matches_transformers():
    from transformers.generation.logits_process import TemperatureLogitsWarper, TopPLogitsWarper
    torch.manual_seed(0)
    logits = torch.randn(5, 300) * 3
    ref = TopPLogitsWarper(top_p=0.95, min_tokens_to_keep=1)(None, TemperatureLogitsWarper(0.7)(None, logits.clone()))
    got = ON.warp_logits(logits, temperature=0.7, top_p=0.95)
    assert torch.equal(torch.isinf(got), torch.isinf(ref))          # the kept set is an index operation: exact
    torch.testing.assert_close(got[~torch.isinf(got)], ref[~torch.isinf(ref)])


EXTRA = torch.load(os.path.join(os.path.dirname(__file__), "golden", "narrator_extra.pt"), weights_only=False)


def _extra_case(name):
    c = EXTRA[name]
    cfg = c["cfg"]
    p = ON.init_narrator_params(cfg, seed=c["param_seed"])
    for k, v in c["param_checksum"].items():
        assert abs(float(p[k].double().sum()) - v) <= 1e-6 * max(1.0, abs(v)), k
    frames, _ = synthetic_batch(dict(cfg["visual"], context_length=8, vocab_size=8), 2, seed=c["frames_seed"])
    return c, cfg, p, frames


def test_narrator_oracle_extra_geometries():
    """Patch-14 encoder + cross-attention every 3rd decoder layer, and cross-attention in every layer
    (tests/golden/make_golden_narrator_extra.py)."""
    for name in ("p14_freq3", "freq1"):
        c, cfg, p, frames = _extra_case(name)
        torch.testing.assert_close(ON.vclm_encode_image(frames, p, cfg), c["image_tokens"], **TOL)
        out = ON.vclm_forward(frames, c["text"], p, cfg)
        torch.testing.assert_close(out["text_tokens_logits"], c["logits"], rtol=1e-3, atol=1e-4)
        assert torch.equal(out["labels"], c["labels"])
