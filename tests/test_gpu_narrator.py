"""GPU parity of the narrator inference path (VCLM_HF: TimeSformer features -> attention pooling -> gated GPT-2) against
the golden vectors of the unmodified reference and the oracle.  Tolerance: bf16 operands / fp32 accumulation vs fp32
reference -> rel-L2 <= 2e-2, cosine >= 0.999 (logits 3e-2: 2 decoder layers + LM head over bf16 activations)."""
import os
from types import SimpleNamespace

import pytest
import torch

from oracle import narrator as ON
from oracle.dual_encoder import synthetic_batch
from tests.util import assert_close_bf16, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "narrator_small.pt"), weights_only=False)


def build(cfg, params):
    from lavila_b200.models.gpt2_gated import GPT2LMHeadModel, augment_gpt2_config
    from lavila_b200.models.narrator import VCLM_HF
    from lavila_b200.models.timesformer import SpaceTimeTransformer, QuickGELU
    v = cfg["visual"]
    vis = SpaceTimeTransformer(img_size=v["img_size"], patch_size=v["patch_size"], embed_dim=v["embed_dim"], depth=v["depth"],
                               num_heads=v["num_heads"], num_frames=v["num_frames"], time_init="zeros", ln_pre=True,
                               act_layer=QuickGELU)
    vis.head = torch.nn.Identity()
    vis.pre_logits = torch.nn.Identity()
    g = SimpleNamespace(vocab_size=cfg["vocab_size"], n_positions=cfg["n_positions"], n_embd=cfg["n_embd"],
                        n_layer=cfg["n_layer"], n_head=cfg["n_head"], layer_norm_epsilon=1e-5, activation_function="gelu_new")
    dec = GPT2LMHeadModel(augment_gpt2_config(g, cross_attn_freq=cfg["cross_attn_freq"], gated_xattn=True))
    m = VCLM_HF(vision_width=v["embed_dim"], vision_model=vis, text_width=cfg["n_embd"], text_decoder=dec,
                num_img_queries=cfg["num_img_queries"], dim_head=64, heads=cfg["pool_heads"])
    res = m.load_state_dict(params, strict=False)
    assert not res.unexpected_keys
    assert all(k.endswith(("attn.bias", "attn.masked_bias", "crossattention.bias", "crossattention.masked_bias", ".beta"))
               for k in res.missing_keys), res.missing_keys
    return m.to(DEV).eval()


def _setup():
    cfg = GOLD["cfg"]
    p = ON.init_narrator_params(cfg, seed=0)
    vcfg = dict(cfg["visual"], context_length=8, vocab_size=8)
    frames, _ = synthetic_batch(vcfg, 2, seed=GOLD["frames_seed"])
    return cfg, build(cfg, p), frames.to(DEV)


def test_encode_image_matches_reference_golden():
    cfg, m, frames = _setup()
    tok = m.encode_image(frames)
    assert_close_bf16(tok, GOLD["image_tokens"], "image tokens")


def test_logits_match_reference_golden():
    cfg, m, frames = _setup()
    out = m(frames, GOLD["text"].to(DEV))
    assert torch.equal(out["labels"].cpu(), GOLD["labels"])
    assert_close_bf16(out["text_tokens_logits"], GOLD["logits"], "teacher-forced logits", rel=3e-2, cos=0.999)
    tok = m.encode_image(frames)
    pre = m.text_decoder(GOLD["text"][:, :5].contiguous().to(DEV), encoder_hidden_states=tok).logits
    assert_close_bf16(pre, GOLD["logits_prefix5"], "prefix logits", rel=3e-2, cos=0.999)
    last = m.text_decoder(GOLD["text"][:, :5].contiguous().to(DEV), encoder_hidden_states=tok, last_only=True).logits
    assert rel_l2(last[:, 0], pre[:, -1]) < 1e-5     # last-position shortcut == full LM head, same kernels


def test_generate_runs_like_reference():
    cfg, m, frames = _setup()
    tok = m.encode_image(frames)
    t = SimpleNamespace(bos_token_id=cfg["vocab_size"] - 1, eos_token_id=cfg["vocab_size"] - 1, pad_token_id=0)
    torch.manual_seed(0)
    ids, ppl = m.generate(tok, t, max_text_length=6, top_p=0.95, temperature=0.7, num_return_sequences=2)
    assert tuple(ids.shape) == GOLD["gen_shape"] and tuple(ppl.shape) == GOLD["ppl_shape"]
    assert ids.dtype == torch.int64 and bool((ids[:, 0] == t.bos_token_id).all())
    assert bool(torch.isfinite(ppl).all())
    # the oracle's warper keeps the same token set as the transformers warper used by generate (index op: exact)
    logits = m.text_decoder(ids[:, :3].contiguous(), encoder_hidden_states=tok.repeat_interleave(2, 0), last_only=True).logits[:, -1]
    w = m._get_logits_warper(top_p=0.95, temperature=0.7, num_beams=1)(ids[:, :3], logits.clone())
    assert torch.equal(torch.isinf(w), torch.isinf(ON.warp_logits(logits, 0.7, 0.95)))


def test_kv_cached_decoding_equals_full_prefix():
    """Incremental decoding (self-attention KV cache) gives the logits of the reference algorithm (full
    prefix re-forward, narrator.py:118-121) position by position, and generate() samples the same ids with and without it."""
    cfg, m, frames = _setup()
    tok = m.encode_image(frames)
    ids = GOLD["text"][:, :6].contiguous().to(DEV)
    full = m.text_decoder(ids, encoder_hidden_states=tok).logits                      # [B, 6, V]
    cache, ctx = {"max_len": 8}, {}
    # prefill 3 positions at once, then one position at a time
    got = [m.text_decoder(ids[:, :3].contiguous(), encoder_hidden_states=tok, ctx_kv_cache=ctx, self_kv_cache=cache, past_len=0).logits]
    for t in range(3, 6):
        got.append(m.text_decoder(ids[:, t:t + 1].contiguous(), encoder_hidden_states=tok, ctx_kv_cache=ctx, self_kv_cache=cache,
                                  past_len=t).logits)
    got = torch.cat(got, 1)
    assert rel_l2(got, full) < 1e-5, rel_l2(got, full)
    t = SimpleNamespace(bos_token_id=cfg["vocab_size"] - 1, eos_token_id=cfg["vocab_size"] - 1, pad_token_id=0)
    outs = []
    for use in (False, True):
        torch.manual_seed(5)
        outs.append(m.generate(tok, t, max_text_length=8, top_p=0.95, temperature=0.7, num_return_sequences=3, use_kv_cache=use))
    assert torch.equal(outs[0][0], outs[1][0])
    assert rel_l2(outs[1][1], outs[0][1]) < 1e-4
    # the CUDA-graph state (KV buffers, captured step) is reused for the next batch of clips: same answers as a cold run
    assert m._decode_states and next(iter(m._decode_states.values()))["graph"] is not None
    tok2 = m.encode_image(frames.flip(0).contiguous())
    outs2 = []
    for use in (True, False):
        torch.manual_seed(6)
        outs2.append(m.generate(tok2, t, max_text_length=8, top_p=0.95, temperature=0.7, num_return_sequences=3, use_kv_cache=use))
    assert torch.equal(outs2[0][0], outs2[1][0])
    assert rel_l2(outs2[0][1], outs2[1][1]) < 1e-4


@pytest.mark.parametrize("temperature,top_p", [(0.7, 0.95), (1.0, 0.9), (1.3, 0.5), (0.7, 0.999)])
def test_fused_top_p_filter_keeps_the_library_token_set(temperature, top_p):
    """lv_top_p_filter (csrc/sampling.cu) vs transformers' TemperatureLogitsWarper + TopPLogitsWarper on real decoder logits, a
    peaked synthetic distribution and a distribution with exact ties: same -inf mask (index op), same scaled logits elsewhere."""
    from lavila_b200 import ops
    cfg, m, frames = _setup()
    tok = m.encode_image(frames)
    ids = GOLD["text"][:, :4].contiguous().to(DEV)
    real = m.text_decoder(ids, encoder_hidden_states=tok, last_only=True).logits[:, -1, :].float()
    g = torch.Generator(device="cpu").manual_seed(3)
    peaked = (torch.randn(5, 50257, generator=g) * 4.0).to(DEV)
    ties = torch.randint(-3, 4, (3, 4001), generator=g).float().to(DEV)          # many exactly equal logits
    for x in (real, peaked, ties):
        ref = m._get_logits_warper(top_p=top_p, temperature=temperature, num_beams=1)(None, x.clone())
        got = ops.top_p_filter_(x.clone().contiguous(), temperature, top_p)
        kept_ref, kept_got = ~torch.isinf(ref), ~torch.isinf(got)
        if x is ties:
            # which of the tied tokens at the threshold go is the sort's choice: same COUNT per row, same kept values
            assert torch.equal(kept_ref.sum(1), kept_got.sum(1))
            assert torch.equal(torch.sort(ref, dim=1).values, torch.sort(got, dim=1).values)
        else:
            # The threshold token is decided by comparing a cumulative fp32 sum with 1 - top_p; the library sums in sorted order,
            # the kernel per radix bin, so a token whose own probability is below the round-off of that sum (top_p = 0.999: the
            # boundary sits among tokens of probability ~1e-7) may fall on the other side: at most one token per row, and only
            # such a negligible one.  At the script's top_p = 0.95 the masks are identical.
            diff = kept_ref ^ kept_got
            assert int(diff.sum(1).max()) <= (1 if top_p > 0.99 else 0), (int(kept_ref.sum()), int(kept_got.sum()))
            probs = torch.softmax(x / temperature, dim=-1)
            assert float(probs[diff].max() if bool(diff.any()) else 0.0) < 1e-5
            both = kept_ref & kept_got
            assert torch.equal(ref[both], got[both])


def test_beam_decoding_on_the_kernels():
    """beam_sample / group_beam_search (narrator.py:149-366) on the CUDA path.  Host logic is pinned exactly on CPU
    (tests/test_host_narrator_cpu.py); here: shapes / dtypes, and the returned score of every open-ended sequence equals the
    length-normalised sum of its token log-probabilities recomputed by teacher forcing through the same kernels."""
    cfg, m, frames = _setup()
    tok = m.encode_image(frames)
    t = SimpleNamespace(bos_token_id=cfg["vocab_size"] - 1, eos_token_id=cfg["vocab_size"] - 1, pad_token_id=0)
    L = 8
    seq, sc = m.group_beam_search(tok, t, num_beams=4, num_beam_groups=2, num_return_sequences=1, max_text_length=L)
    assert seq.dtype == torch.int64 and seq.shape[0] == tok.shape[0] and seq.shape[1] <= L and sc.shape == (tok.shape[0],)
    assert bool((seq[:, 0] == t.bos_token_id).all()) and bool(torch.isfinite(sc).all())
    checked = 0
    for b in range(seq.shape[0]):
        row = seq[b]
        if row.shape[0] == L and not bool(((row[1:] == t.eos_token_id) | (row[1:] == t.pad_token_id)).any()):
            logits = m.text_decoder(row[None, :-1].contiguous(), encoder_hidden_states=tok[b:b + 1]).logits[0]
            lp = torch.log_softmax(logits.float(), dim=-1).gather(1, row[1:, None]).sum()
            assert abs(float(lp) / L - float(sc[b])) < 2e-2 * max(1.0, abs(float(sc[b]))), (float(lp) / L, float(sc[b]))
            checked += 1
    assert checked >= 1
    torch.manual_seed(3)
    seq2, sc2 = m.beam_sample(tok, t, num_beams=3, num_return_sequences=2, max_text_length=7, top_p=0.95, temperature=0.7)
    assert seq2.shape[0] == 2 * tok.shape[0] and seq2.shape[1] <= 7 and bool(torch.isfinite(sc2).all())


@pytest.mark.parametrize("B,H,Lq,Lk,mqa,causal", [(2, 3, 5, 40, False, False), (2, 25, 77, 256, False, False),
                                                   (3, 4, 77, 77, False, True), (2, 12, 256, 785, True, False),
                                                   (1, 2, 1, 1, False, True), (2, 2, 130, 130, False, True),
                                                   # >= 64 query rows: the tcgen05 key-tiled kernel (flash_tc.cu)
                                                   (2, 3, 64, 64, False, True), (2, 5, 200, 300, False, False),
                                                   (1, 25, 128, 1025, True, False), (3, 25, 77, 256, False, False),
                                                   (2, 4, 129, 257, False, True), (1, 2, 300, 16, False, False)])
def test_flash_attention(B, H, Lq, Lk, mqa, causal):
    from lavila_b200 import ops
    torch.manual_seed(1)
    q = torch.randn(B, Lq, H, 64, device=DEV).bfloat16()
    hk = 1 if mqa else H
    k = torch.randn(B, Lk, hk, 64, device=DEV).bfloat16()
    v = torch.randn(B, Lk, hk, 64, device=DEV).bfloat16()
    kv = torch.cat((k, v), dim=2).contiguous()            # [B, Lk, 2*hk, 64]: k heads then v heads
    out = torch.zeros(B, Lq, H, 64, device=DEV, dtype=torch.bfloat16)
    ops.flash_attn_fwd(q, kv, kv.view(B * Lk, -1)[:, hk * 64:], out, B, H, Lq, Lk, q_rows=Lq, kv_rows=Lk, ld_q=H * 64,
                       ld_kv=2 * hk * 64, ld_out=H * 64, kv_head_stride=0 if mqa else 64, causal=causal, scale=0.125)
    qf, kf, vf = q.float().permute(0, 2, 1, 3), k.float().permute(0, 2, 1, 3), v.float().permute(0, 2, 1, 3)
    s = (qf @ kf.transpose(-1, -2)) * 0.125
    if causal:
        mask = torch.ones(Lk, Lk, dtype=torch.bool, device=DEV).tril()[Lk - Lq:Lk]
        s = s.masked_fill(~mask, float("-inf"))
    ref = (torch.softmax(s, dim=-1) @ vf).permute(0, 2, 1, 3)
    assert rel_l2(out, ref) < 1e-2


@pytest.mark.parametrize("M,N,K,flags", [(32, 1600, 1600, "bias"), (5, 4800, 1600, "bias"), (320, 6400, 1600, "gelu"),
                                         (32, 1600, 6400, "resid_gate"), (70, 768, 768, "sqrelu"), (32, 1600, 1600, "resid"),
                                         # several 64-row blocks of sequences (num_return_sequences > 2 at batch 32), ragged last block
                                         (128, 1600, 1600, "bias"), (320, 1600, 6400, "resid_gate"), (512, 4800, 1600, "bias"),
                                         (100, 768, 768, "sqrelu"), (68, 1600, 1600, "resid"), (320, 6400, 1600, "bias")])
def test_skinny_gemm(M, N, K, flags):
    """lv_gemm_skinny_bf16 (decode-sized GEMMs, csrc/gemm_skinny.cu) against fp32 torch on the same bf16 operands, every epilogue
    the gated GPT-2 uses; two runs are bit-identical (deterministic split-K)."""
    from lavila_b200 import ops, _lib as L
    torch.manual_seed(7)
    A = (torch.randn(M, K, device=DEV) * 0.5).to(torch.bfloat16)
    W = (torch.randn(K, N, device=DEV) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV)
    resid = torch.randn(M, N, device=DEV)
    gate = torch.tensor(0.7, device=DEV)
    acc = A.float() @ W.float() + bias
    if flags == "bias":
        fl, kw, ref, dt = 0, {}, acc, torch.bfloat16
    elif flags == "gelu":
        fl, kw, dt = L.EPI_GELU_TANH, {}, torch.bfloat16
        ref = torch.nn.functional.gelu(acc, approximate="tanh")
    elif flags == "sqrelu":
        fl, kw, ref, dt = L.EPI_SQRELU, {}, torch.relu(acc) ** 2, torch.bfloat16
    elif flags == "resid":
        fl, kw, ref, dt = L.EPI_RESID, dict(resid=resid), acc + resid, torch.float32
    else:
        fl, kw, dt = L.EPI_RESID | L.EPI_SCALE | L.EPI_SCALE_TANH, dict(resid=resid, scale=gate), torch.float32
        ref = torch.tanh(gate) * acc + resid
    outs = []
    for _ in range(2):
        out = torch.empty(M, N, device=DEV, dtype=dt)
        assert ops.gemm_skinny(A, W, M, N, K, out, flags=fl | L.EPI_BIAS, bias=bias, **kw)
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    assert_close_bf16(outs[0], ref, "skinny gemm %s" % flags, rel=1e-2)
