"""The oracle (oracle/dual_encoder.py) against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import pytest
import torch

from oracle import dual_encoder as O

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "dual_encoder_small.pt"), weights_only=False)
TOL = dict(rtol=2e-4, atol=2e-5)


def _setup(case):
    c = GOLD[case]
    cfg = c["cfg"]
    p = O.init_params(cfg, seed=c["param_seed"], gated=c["gated"])
    for k, v in c["param_checksum"].items():   # RNG drift guard: parameters are regenerated, not stored
        assert abs(float(p[k].double().sum()) - v) <= 1e-6 * max(1.0, abs(v)), k
    frames, text = O.synthetic_batch(cfg, c["batch"], seed=c["input_seed"])
    assert abs(float(frames.double().sum()) - c["frames_checksum"]) < 1e-6 * frames.numel()
    assert torch.equal(text, c["text"])          # token ids: bit exact
    return c, cfg, p, frames, text


@pytest.mark.parametrize("case", ["plain", "norm", "gated_norm"])
def test_forward_matches_reference(case):
    c, cfg, p, frames, text = _setup(case)
    out = O.clip_forward(frames, text, p, cfg, norm_embed=c["norm_embed"])
    torch.testing.assert_close(out["image_embed"], c["image_embed"], **TOL)
    torch.testing.assert_close(out["text_embed"], c["text_embed"], **TOL)
    torch.testing.assert_close(out["logit_scale"], c["logit_scale"], **TOL)
    toks = O.timesformer_features(frames.permute(0, 2, 1, 3, 4), p, cfg, cls_at_last=False)
    torch.testing.assert_close(toks[:, :5], c["visual_tokens"], **TOL)
    ld = O.clip_loss(out["image_embed"], out["text_embed"], out["logit_scale"])
    torch.testing.assert_close(ld["loss"], c["loss"], rtol=1e-4, atol=1e-4)
    assert float(ld["clip_acc"]) == float(c["clip_acc"])   # argmax / labels: exact


@pytest.mark.parametrize("case", ["norm", "gated_norm"])
def test_gradients_match_reference(case):
    c, cfg, p, frames, text = _setup(case)
    p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    out = O.clip_forward(frames, text, p, cfg, norm_embed=c["norm_embed"])
    O.clip_loss(out["image_embed"], out["text_embed"], out["logit_scale"])["loss"].backward()
    assert set(c["grads"]) <= set(p), set(c["grads"]) - set(p)
    for name, ref in c["grads"].items():
        g = p[name].grad
        assert g is not None, name
        if "full" in ref:
            torch.testing.assert_close(g, ref["full"].reshape(g.shape), rtol=2e-3, atol=2e-6, msg=lambda m: name + ": " + m)
        else:
            torch.testing.assert_close(g.flatten()[ref["idx"]], ref["sample"], rtol=2e-3, atol=2e-6, msg=lambda m: name + ": " + m)
            torch.testing.assert_close(g.norm(), ref["norm"], rtol=1e-3, atol=1e-7)


def test_multirank_loss_semantics():
    """2-rank CLIPLoss(use_vissl) == single-process loss on the concatenated batch; local-embedding gradient
    is W x the single-process one (GatherLayer all_reduce-SUM), `gather_features` path is 1 x."""
    m = GOLD["multirank"]
    W = m["world"]
    imgs = [r["image"].clone().requires_grad_(True) for r in m["ranks"]]
    txts = [r["text"].clone().requires_grad_(True) for r in m["ranks"]]
    out = O.clip_loss_multi_rank(imgs, txts, torch.tensor(14.2857))
    gi = torch.autograd.grad(out["loss"], imgs + txts)
    for r in range(W):
        ref = m["ranks"][r]
        torch.testing.assert_close(out["loss"], ref["vissl"]["loss"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(out["loss"], ref["plain"]["loss"], rtol=1e-5, atol=1e-6)
        assert float(out["clip_acc"]) == float(ref["vissl"]["acc"])
        torch.testing.assert_close(W * gi[r], ref["vissl"]["grad_image"], rtol=1e-4, atol=1e-7)
        torch.testing.assert_close(W * gi[W + r], ref["vissl"]["grad_text"], rtol=1e-4, atol=1e-7)
        torch.testing.assert_close(gi[r], ref["plain"]["grad_image"], rtol=1e-4, atol=1e-7)


def test_known_answers():
    """SURVEY.md 8(c) self-checks: identical embeddings -> acc 100; labels are arange."""
    e = torch.nn.functional.normalize(torch.randn(8, 16), dim=-1)
    out = O.clip_loss(e, e, torch.tensor(100.0))
    assert float(out["clip_acc"]) == 100.0
    assert float(out["loss"]) < 0.05


# ----------------------------------------------------------------------------------------------- SSLCLIPLoss
SSL = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ssl_loss_small.pt"), weights_only=False)


def _ssl_oracle(img, txt, gt, scale_init):
    import math
    img, txt = img.clone().requires_grad_(True), txt.clone().requires_grad_(True)
    s = torch.tensor(SSL["scale"], requires_grad=True)
    lp = torch.tensor(math.log(1 / scale_init), requires_grad=True)
    out = O.ssl_clip_loss(img, txt, s, lp, gt)
    return out, torch.autograd.grad(out["loss"], (img, txt, s, lp))


def _same(a, b):
    a, b = torch.as_tensor(a).float().reshape(-1), torch.as_tensor(b).float().reshape(-1)
    if torch.isnan(b).any():
        return bool(torch.isnan(a).all())
    return bool(torch.allclose(a, b, rtol=1e-5, atol=1e-6))


@pytest.mark.parametrize("idx", [0, 1, 2, 3])
def test_ssl_clip_loss_matches_reference_golden(idx):
    c = SSL["world1"][idx]
    out, (gi, gt, gs, gp) = _ssl_oracle(c["image"], c["text"], c["gt"], SSL["scale_init"])
    for k in ("loss", "clip_acc", "clip_acc_gt", "clip_acc_pseudo", "num_gt", "num_pseudo"):
        assert _same(out[k], c[k]), (k, out[k], c[k])
    torch.testing.assert_close(gi, c["grad_image"], rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(gt, c["grad_text"], rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(gs, c["grad_scale"], rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(gp, c["grad_scale_pseudo"], rtol=1e-4, atol=1e-7)


def test_ssl_clip_loss_multirank_semantics():
    """2-rank SSLCLIPLoss(use_vissl): every rank's loss == the single-process loss on the concatenated batch; the
    local-embedding gradient is W x the single-process one (GatherLayer), the scale gradients are 1 x."""
    r = SSL["world2"]
    W = len(r)
    img, txt, gt = torch.cat([x["image"] for x in r]), torch.cat([x["text"] for x in r]), torch.cat([x["gt"] for x in r])
    out, (gi, gtx, gs, gp) = _ssl_oracle(img, txt, gt, SSL["scale_init"])
    B = r[0]["image"].shape[0]
    for k in range(W):
        assert _same(out["loss"], r[k]["loss"]) and _same(out["clip_acc"], r[k]["clip_acc"])
        assert _same(out["clip_acc_gt"], r[k]["clip_acc_gt"]) and _same(out["clip_acc_pseudo"], r[k]["clip_acc_pseudo"])
        torch.testing.assert_close(W * gi[k * B:(k + 1) * B], r[k]["grad_image"], rtol=1e-4, atol=1e-7)
        torch.testing.assert_close(W * gtx[k * B:(k + 1) * B], r[k]["grad_text"], rtol=1e-4, atol=1e-7)
        torch.testing.assert_close(gs, r[k]["grad_scale"], rtol=1e-4, atol=1e-7)
        torch.testing.assert_close(gp, r[k]["grad_scale_pseudo"], rtol=1e-4, atol=1e-7)


# ----------------------------------------------------------------------------------------------- extra reference goldens
EXTRA = torch.load(os.path.join(os.path.dirname(__file__), "golden", "dual_encoder_extra.pt"), weights_only=False)


@pytest.mark.parametrize("case", ["p14", "wide_text", "fewframes"])
def test_oracle_matches_reference_extra_geometries(case):
    """Patch 14 / 4 heads / depth 3 / gated; a deeper, wider text tower; an 8-frame model fed 4-frame clips
    (tests/golden/make_golden_extra.py): outputs, loss, accuracy and every parameter gradient of the unmodified reference."""
    c = EXTRA[case]
    cfg = c["cfg"]
    p = O.init_params(cfg, seed=c["param_seed"], gated=c["gated"])
    for k, v in c["param_checksum"].items():
        assert abs(float(p[k].double().sum()) - v) <= 1e-6 * max(1.0, abs(v)), k
    frames, text = O.synthetic_batch(cfg, c["batch"], seed=c["input_seed"], frames=c.get("frames"))
    assert abs(float(frames.double().sum()) - c["frames_checksum"]) < 1e-6 * frames.numel()
    assert torch.equal(text, c["text"])
    p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    out = O.clip_forward(frames, text, p, cfg, norm_embed=c["norm_embed"])
    torch.testing.assert_close(out["image_embed"], c["image_embed"], **TOL)
    torch.testing.assert_close(out["text_embed"], c["text_embed"], **TOL)
    ld = O.clip_loss(out["image_embed"], out["text_embed"], out["logit_scale"])
    torch.testing.assert_close(ld["loss"], c["loss"], rtol=1e-4, atol=1e-4)
    assert float(ld["clip_acc"]) == float(c["clip_acc"])
    ld["loss"].backward()
    for name, ref in c["grads"].items():
        g = p[name].grad
        assert g is not None, name
        # fp32 summation-order noise scales with the largest entry (a few entries are near-cancelling sums): atol relative to it
        if "full" in ref:
            want = ref["full"].reshape(g.shape)
            torch.testing.assert_close(g, want, rtol=2e-3, atol=2e-3 * float(want.abs().max()) + 1e-9, msg=lambda m: name + ": " + m)
        else:
            want = ref["sample"]
            torch.testing.assert_close(g.flatten()[ref["idx"]], want, rtol=2e-3, atol=2e-3 * float(want.abs().max()) + 1e-9,
                                       msg=lambda m: name + ": " + m)


def test_oracle_block_at_tsfb_geometry_matches_reference():
    """One SpaceTimeBlock at the REAL TSF-B geometry (D = 768, 12 heads, 16 x 196 patches, tanh-gated), forward and every
    gradient, against the unmodified reference (tests/golden/make_golden_block.py; 4096-entry samples + norms).  The GPU test
    tests/test_gpu_model.py::test_block_midsize_vs_oracle compares the kernels with this same oracle at this same geometry."""
    from tests.golden.make_golden_block import block_inputs, block_params
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "block_tsfb.pt"), weights_only=False)
    geo = g["geometry"]
    params = block_params(g["param_seed"])
    for k, v in g["param_checksum"].items():
        assert abs(float(params[k].double().sum()) - v) <= 1e-6 * max(1.0, abs(v)), k
    x, dy = block_inputs(g["input_seed"])
    assert abs(float(x.double().sum()) - g["x_checksum"]) < 1e-6 * x.numel()
    p = {"b." + k: v.clone().requires_grad_(True) for k, v in params.items()}
    x.requires_grad_(True)
    y = O.space_time_block(x, p, "b.", geo["H"], geo["T"], geo["n"])
    y.backward(dy)

    def check(t, ref, name, rtol):
        got = t.detach().float().flatten()[ref["idx"]]
        scale = float(ref["sample"].abs().max())
        torch.testing.assert_close(got, ref["sample"], rtol=rtol, atol=rtol * scale, msg=lambda m: name + ": " + m)
        torch.testing.assert_close(t.detach().float().norm(), ref["norm"], rtol=1e-4, atol=0, msg=lambda m: name + " norm: " + m)

    check(y, g["y"], "y", 1e-4)
    check(x.grad, g["dx"], "dx", 1e-3)
    for k, ref in g["grads"].items():
        check(p["b." + k].grad, ref, "d" + k, 2e-3)
