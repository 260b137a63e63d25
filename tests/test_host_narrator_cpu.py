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


# ----------------------------------------------------------------------------------------------------------------------
# beam_sample / group_beam_search (narrator.py:149-366): the product's candidate selection + its BeamSearchScorer
# (lavila_b200/models/beam_search.py) on the kernel test doubles, against the reference's own decoding code run with the oracle
# scorer (tests/golden/make_golden_beam.py).
BEAM = torch.load(os.path.join(os.path.dirname(__file__), "golden", "narrator_beam.pt"), weights_only=False)


def _beam_setup(monkeypatch):
    ops_doubles.install(monkeypatch)
    cfg = BEAM["cfg"]
    model = _build(cfg, ON.init_narrator_params(cfg, seed=0))
    frames, _ = synthetic_batch(dict(cfg["visual"], context_length=8, vocab_size=8), BEAM["batch"], seed=BEAM["frames_seed"])
    tok = model.encode_image(frames)
    assert rel_l2(tok, BEAM["image_tokens"]) < 2e-2
    t = SimpleNamespace(bos_token_id=cfg["vocab_size"] - 1, eos_token_id=cfg["vocab_size"] - 1, pad_token_id=0)
    return model, tok, t


@pytest.mark.parametrize("case", [k for k, v in BEAM["cases"].items() if v["kind"] == "group"])
def test_group_beam_search_matches_reference(case, monkeypatch):
    """Deterministic (top-k): the ids are the reference's, the scores agree to fp32 round-off of the doubles."""
    model, tok, t = _beam_setup(monkeypatch)
    c = BEAM["cases"][case]
    seq, sc = model.group_beam_search(tok, t, **c["kw"])
    assert seq.dtype == torch.int64 and tuple(seq.shape) == tuple(c["sequences"].shape)
    assert torch.equal(seq, c["sequences"]), (seq, c["sequences"])
    assert torch.allclose(sc, c["scores"], atol=5e-3, rtol=1e-3)      # the doubles round GEMM operands to bf16 like the kernels


@pytest.mark.parametrize("case", [k for k, v in BEAM["cases"].items() if v["kind"] == "sample"])
def test_beam_sample_real_decoder_under_the_same_seed(case, monkeypatch):
    """With the real decoder the joint scores differ from the reference's by the bf16 operand rounding the doubles model, which
    can move an occasional multinomial draw: shapes / BOS / finiteness always, and at least half of the returned sequences
    identical to the reference's (the exact check of the host logic is the stub-decoder test below)."""
    model, tok, t = _beam_setup(monkeypatch)
    c = BEAM["cases"][case]
    torch.manual_seed(c["seed"])
    seq, sc = model.beam_sample(tok, t, **c["kw"])
    assert tuple(seq.shape) == tuple(c["sequences"].shape) and bool((seq[:, 0] == t.bos_token_id).all())
    assert bool(torch.isfinite(sc).all())
    same = (seq == c["sequences"]).all(dim=1)
    assert int(same.sum()) * 2 >= same.numel(), (seq, c["sequences"])
    assert torch.allclose(sc[same], c["scores"][same], atol=2e-2, rtol=1e-3)


@pytest.mark.parametrize("case", [k for k, v in BEAM["cases"].items() if v["kind"].startswith("stub")])
def test_beam_decoding_host_logic_is_exact_on_a_stub_decoder(case, monkeypatch):
    """Decoder replaced by a table look-up that is an exact function of the ids (same stub as in the golden run): every id and
    every score of beam_sample (same torch seed) and group_beam_search must equal the reference's bit for bit."""
    from tests.golden.make_golden_beam import StubDecoder
    model, tok, t = _beam_setup(monkeypatch)
    model.text_decoder = StubDecoder(BEAM["cfg"]["vocab_size"])
    c = BEAM["cases"][case]
    if c["kind"] == "stub_sample":
        torch.manual_seed(c["seed"])
        seq, sc = model.beam_sample(tok, t, **c["kw"])
    else:
        seq, sc = model.group_beam_search(tok, t, **c["kw"])
    assert torch.equal(seq, c["sequences"]), (seq, c["sequences"])
    assert torch.equal(sc, c["scores"])


def test_beam_scorer_equals_oracle_scorer_on_random_steps():
    """The product's vectorised scorer and the oracle's loop restatement agree step by step on random candidate streams
    (including EOS hits inside and outside the first group_size ranks, finished batch elements and the final ranking)."""
    from lavila_b200.models.beam_search import BeamSearchScorer as P
    from oracle.beam_scorer import BeamSearchScorer as O
    g = torch.Generator().manual_seed(3)
    for trial in range(6):
        B, nb, groups, keep, V, eos = 3, 4, (1 if trial % 2 else 2), (1 if trial < 3 else 2), 11, 10
        gs = nb // groups
        a = P(B, nb, "cpu", length_penalty=1.0 + 0.5 * (trial % 3), num_beam_groups=groups, num_beam_hyps_to_keep=keep)
        b = O(B, nb, "cpu", length_penalty=1.0 + 0.5 * (trial % 3), num_beam_groups=groups, num_beam_hyps_to_keep=keep)
        ids = torch.full((B * nb, 1), eos, dtype=torch.long)
        scores = torch.zeros(B * nb)
        for step in range(7):
            new_last = torch.zeros(B * nb, dtype=torch.long)
            for grp in range(groups):
                rows = (torch.arange(B).view(-1, 1) * nb + torch.arange(grp * gs, (grp + 1) * gs).view(1, -1)).reshape(-1)
                cs, _ = torch.sort(torch.randn(B, 2 * gs, generator=g) - step, descending=True, dim=1)
                ct = torch.randint(0, V, (B, 2 * gs), generator=g)
                ci = torch.randint(0, gs, (B, 2 * gs), generator=g)
                # keep enough non-EOS candidates for the beam to refill
                ct[:, -gs:] = torch.randint(0, V - 1, (B, gs), generator=g)
                ra = a.process(ids[rows], cs, ct, ci, pad_token_id=0, eos_token_id=eos)
                rb = b.process(ids[rows], cs, ct, ci, pad_token_id=0, eos_token_id=eos)
                for k in ("next_beam_scores", "next_beam_tokens", "next_beam_indices"):
                    assert torch.equal(ra[k], rb[k]), (trial, step, k)
                scores[rows] = ra["next_beam_scores"]
                ids[rows] = ids[rows][ra["next_beam_indices"]]
                new_last[rows] = ra["next_beam_tokens"]
            ids = torch.cat([ids, new_last.unsqueeze(-1)], dim=-1)
            assert bool(a.is_done) == bool(b.is_done)
        fa = a.finalize(ids, scores, None, None, max_length=9, pad_token_id=0, eos_token_id=eos)
        fb = b.finalize(ids, scores, None, None, max_length=9, pad_token_id=0, eos_token_id=eos)
        assert torch.equal(fa["sequences"], fb["sequences"]) and torch.allclose(fa["sequence_scores"], fb["sequence_scores"])
