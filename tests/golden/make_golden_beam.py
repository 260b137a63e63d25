"""Golden vectors for VCLM_HF.beam_sample / group_beam_search: the UNMODIFIED reference's decoding code
(lavila/models/narrator.py:149-366) run on CPU with the small narrator of make_golden_narrator.py.

    python tests/golden/make_golden_beam.py      # writes tests/golden/narrator_beam.pt

The one thing the reference cannot supply here is `transformers.BeamSearchScorer` (transformers==4.27 is pinned in
requirements.txt:8; the installed 5.5 dropped the class): oracle/beam_scorer.py -- a restatement of the 4.27 scorer -- is
injected under that name BEFORE the reference is imported, so every line of the candidate selection, the warpers, the group
bookkeeping and the decoder is the reference's own.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import beam_scorer  # noqa: E402
from oracle import reference_shim  # noqa: E402
from oracle.narrator import SMALL_NARRATOR, init_narrator_params  # noqa: E402
from oracle.dual_encoder import synthetic_batch  # noqa: E402
from tests.golden.make_golden_narrator import build_reference  # noqa: E402


class StubDecoder(torch.nn.Module):
    """logits[b, t, :] = table[ids[b, t]] + 0.1 * table[(ids[b, t-1] + t) % V]: an exact function of the ids (fixed generator)."""

    def __init__(self, vocab, seed=5):
        super().__init__()
        self.table = torch.randn(vocab, vocab, generator=torch.Generator().manual_seed(seed)) * 2.0

    def forward(self, input_ids, encoder_hidden_states=None, **kw):
        from types import SimpleNamespace
        t = torch.arange(input_ids.shape[1], device=input_ids.device).view(1, -1)
        prev = torch.cat([input_ids[:, :1], input_ids[:, :-1]], dim=1)
        tab = self.table.to(input_ids.device)
        return SimpleNamespace(logits=tab[input_ids] + 0.1 * tab[(prev + t) % tab.shape[0]])


def main():
    assert reference_shim.install()
    import transformers
    for mod in (transformers, sys.modules["transformers"]):        # replaces the shim's placeholder class
        mod.__dict__["BeamSearchScorer"] = beam_scorer.BeamSearchScorer
    cfg = SMALL_NARRATOR
    params = init_narrator_params(cfg, seed=0)
    model = build_reference(cfg, params)
    import lavila.models.narrator as RN
    assert RN.BeamSearchScorer is beam_scorer.BeamSearchScorer
    vcfg = dict(cfg["visual"], context_length=8, vocab_size=8)
    frames, _ = synthetic_batch(vcfg, 3, seed=77)

    class Tok:
        bos_token_id = eos_token_id = cfg["vocab_size"] - 1
        pad_token_id = 0
    with torch.no_grad():
        tokens = model.encode_image(frames)
    gold = {"cfg": cfg, "frames_seed": 77, "batch": 3, "image_tokens": tokens, "cases": {}}
    # group beam search is deterministic (top-k): ids and scores are pinned
    for name, kw in {"gbs_6x3": dict(num_beams=6, num_beam_groups=3, num_return_sequences=1, max_text_length=9),
                     "gbs_4x2_ret2_lp2": dict(num_beams=4, num_beam_groups=2, num_return_sequences=2, max_text_length=8, length_penalty=2.0),
                     "gbs_4x1_topp": dict(num_beams=4, num_beam_groups=1, num_return_sequences=1, max_text_length=7, top_p=0.9, temperature=0.8)}.items():
        seq, sc = model.group_beam_search(tokens, Tok(), **kw)
        gold["cases"][name] = {"kind": "group", "kw": kw, "sequences": seq, "scores": sc}
        print(name, tuple(seq.shape), sc.tolist())
    # beam_sample draws with torch.multinomial: pinned under a fixed torch seed (CPU generator)
    for name, kw, seed in [("bs_3", dict(num_beams=3, num_return_sequences=1, max_text_length=8, top_p=0.95, temperature=0.7), 11),
                           ("bs_2_ret2", dict(num_beams=2, num_return_sequences=2, max_text_length=7, top_k=20), 12)]:
        torch.manual_seed(seed)
        seq, sc = model.beam_sample(tokens, Tok(), **kw)
        gold["cases"][name] = {"kind": "sample", "kw": kw, "seed": seed, "sequences": seq, "scores": sc}
        print(name, tuple(seq.shape), sc.tolist())
    # The same two procedures on a STUB decoder whose logits are an exact function of the ids (a fixed table): the joint
    # scores are then bit-identical for the reference and the product, so the multinomial draws and every id must agree
    # exactly -- this pins the host logic (expansion, warping of the joint scores, sampling, sorting, scorer, finalize)
    # independently of kernel numerics.
    model.text_decoder = StubDecoder(cfg["vocab_size"])
    for name, kw, seed in [("stub_bs_3_ret2", dict(num_beams=3, num_return_sequences=2, max_text_length=10, top_p=0.9, temperature=0.9), 21),
                           ("stub_bs_4", dict(num_beams=4, num_return_sequences=1, max_text_length=12, top_k=15, length_penalty=0.5), 22)]:
        torch.manual_seed(seed)
        seq, sc = model.beam_sample(tokens, Tok(), **kw)
        gold["cases"][name] = {"kind": "stub_sample", "kw": kw, "seed": seed, "sequences": seq, "scores": sc}
        print(name, tuple(seq.shape), sc.tolist())
    seq, sc = model.group_beam_search(tokens, Tok(), num_beams=6, num_beam_groups=2, num_return_sequences=2, max_text_length=12)
    gold["cases"]["stub_gbs_6x2_ret2"] = {"kind": "stub_group", "kw": dict(num_beams=6, num_beam_groups=2, num_return_sequences=2, max_text_length=12),
                                          "sequences": seq, "scores": sc}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "narrator_beam.pt")
    torch.save(gold, path)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
