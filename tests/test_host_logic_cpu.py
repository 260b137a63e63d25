"""Host-side logic that needs no GPU: the FLOP model bench.py reports against, GEMM split heuristics,
loss-module contracts, model factories' parameter naming (weight-decay grouping of main_pretrain.py:199-213)."""
import math

import pytest
import torch


def test_flop_model_matches_survey():
    import bench
    assert abs(bench.flops_per_clip_train() / 1e12 - 2.2354) < 1e-3                     # cfg 2/3: 2.236 TF / clip
    f5 = bench.flops_per_clip_train(T=32, **bench.MODELS["large336"][2])
    assert abs(f5 / 1e12 - 47.97) < 0.02                                                # cfg 5: 47.97 TF / clip
    # per-block forward figure of SURVEY 8(a) a4: 61.29 GF (N = 3137, D = 768)
    N, D, T, n = 3137, 768, 16, 196
    f_blk = 32 * N * D * D + 4 * D * ((T * n) * (T + n + 2) + 2 * N)
    assert abs(f_blk / 1e9 - 61.29) < 0.01


def test_wgrad_split_heuristic_bounds():
    from lavila_b200 import ops
    for m_out, n_in, tokens in ((768, 768, 200768), (2304, 768, 200768), (3072, 768, 200768), (768, 3072, 200768),
                                (512, 512, 4928), (256, 768, 64)):
        s = ops.wgrad_splits(m_out, n_in, tokens)
        assert 1 <= s <= (tokens + 63) // 64


def test_sslcliploss_module_contract():
    from lavila_b200.models.loss import SSLCLIPLoss, CLIPLoss
    m = SSLCLIPLoss(scale_init=0.08)
    assert list(m.state_dict().keys()) == ["logit_scale_pseudo"]                        # loss.py:140 (checkpointed, main_pretrain.py:398)
    assert abs(float(m.logit_scale_pseudo.detach()) - math.log(1 / 0.08)) < 1e-6
    assert m.logit_scale_pseudo.requires_grad
    assert not SSLCLIPLoss(freeze_scale=True).logit_scale_pseudo.requires_grad
    with pytest.raises(NotImplementedError):                                            # loss.py:167-168
        SSLCLIPLoss(world_size=2, use_vissl=False)({"image_embed": None, "text_embed": None, "logit_scale": None}, None)
    assert list(CLIPLoss().state_dict().keys()) == []


def test_get_loss_and_metric_names():
    from types import SimpleNamespace
    from lavila_b200.models import models as M
    args = SimpleNamespace(contrastive_use_vissl=True, rank=3, world_size=8)
    crit = M.get_loss("CLIP_OPENAI_TIMESFORMER_BASE", args)
    assert crit.use_vissl and crit.rank == 3 and crit.world_size == 8 and crit.cache_labels
    assert M.get_metric_names("CLIP_OPENAI_TIMESFORMER_BASE") == ["loss", "clip_loss", "clip_acc"]
    assert M.loss.SSLCLIPLoss is not None                                               # main_pretrain.py:189 reaches it this way


def test_weight_decay_grouping_names():
    """main_pretrain.py:199-213 puts a parameter in the no-decay group iff ndim < 2 or its name contains 'bias', 'ln' or 'bn':
    the mirror must expose the reference's names so the same parameters land in the same group."""
    import contextlib
    import io
    import bench
    from lavila_b200.models import models as M
    with contextlib.redirect_stdout(io.StringIO()):
        model = M.CLIP_OPENAI_TIMESFORMER_BASE(num_frames=4, project_embed_dim=256)
    groups = bench.param_groups(model)
    decay = {id(p) for p in groups[0]["params"]}
    named = dict(model.named_parameters())
    assert id(named["visual.blocks.0.attn.qkv.weight"]) in decay
    assert id(named["transformer.resblocks.0.mlp.c_fc.weight"]) in decay
    assert id(named["visual.pos_embed"]) in decay and id(named["visual.temporal_embed"]) in decay   # 3-D, no 'ln' in the name
    for n in ("visual.blocks.0.attn.qkv.bias", "visual.ln_pre.weight", "ln_final.weight", "logit_scale",
              "transformer.resblocks.0.ln_1.weight"):
        assert id(named[n]) not in decay, n
    assert id(named["visual.blocks.0.norm1.weight"]) not in decay                       # ndim 1
    assert abs(sum(p.numel() for p in model.parameters()) / 1e6 - 177.7) < 0.5


def test_narrator_factories_exist_with_reference_names():
    from lavila_b200.models import models as M
    for name in ("VCLM_OPENAI_TIMESFORMER_BASE_GPT2", "VCLM_OPENAI_TIMESFORMER_BASE_GPT2_XL", "VCLM_OPENAI_TIMESFORMER_LARGE_GPT2",
                 "VCLM_OPENAI_TIMESFORMER_LARGE_GPT2_XL", "VCLM_OPENAI_TIMESFORMER_LARGE_336PX_GPT2_XL",
                 "CLIP_OPENAI_TIMESFORMER_BASE", "CLIP_OPENAI_TIMESFORMER_LARGE", "CLIP_OPENAI_TIMESFORMER_LARGE_336PX",
                 "CLIP_OPENAI_TIMESFORMER_BASE_DISTILBERT_BASE", "CLIP_OPENAI_TIMESFORMER_LARGE_DISTILBERT_BASE",
                 "CLIP_OPENAI_TIMESFORMER_LARGE_336PX_DISTILBERT_BASE"):
        assert callable(getattr(M, name)), name


MAIN_PRETRAIN_KWARGS = dict(pretrained=None, pretrained2d=False, text_use_cls_token=False, project_embed_dim=256, gated_xattn=True,
                            random_init_gpt2=False, timesformer_gated_xattn=True, timesformer_freeze_space=False, freeze_lm_vclm=True,
                            freeze_visual_vclm=True, freeze_visual_vclm_temporal=False, num_frames=4, drop_path_rate=0.0,
                            temperature_init=0.07)


@pytest.mark.parametrize("name", ["CLIP_OPENAI_TIMESFORMER_BASE", "CLIP_OPENAI_TIMESFORMER_BASE_DISTILBERT_BASE",
                                  "VCLM_OPENAI_TIMESFORMER_BASE_GPT2"])
def test_factories_take_the_drivers_keyword_set(name):
    """main_pretrain.py:157-172 calls every factory with the same 14 keywords; each must accept them all (unknown ones swallowed)
    and return a model exposing what the driver touches next (:173-176, :199-213)."""
    import contextlib
    import io
    from lavila_b200.models import models as M
    with contextlib.redirect_stdout(io.StringIO()):
        model = getattr(M, name)(**MAIN_PRETRAIN_KWARGS)
    names = [n for n, _ in model.named_parameters()]
    assert any(n.startswith("visual.blocks.0.timeattn.") for n in names)
    assert "visual.blocks.0.alpha_timeattn" in names                       # timesformer_gated_xattn=True
    if name.startswith("CLIP"):
        assert model.logit_scale.requires_grad and abs(float(model.logit_scale.detach().exp()) - 1 / 0.07) < 1e-3
    else:
        # freeze_lm_vclm: only the cross-attention additions of the decoder stay trainable (gpt2_gated.py:1019-1029)
        dec = dict(model.text_decoder.named_parameters())
        assert dec["transformer.h.0.crossattention.q_attn.weight"].requires_grad and dec["transformer.h.0.alpha_cattn"].requires_grad
        assert not dec["transformer.h.0.attn.c_attn.weight"].requires_grad and not dec["transformer.wte.weight"].requires_grad
        # freeze_visual_vclm: the spatial weights of the video encoder are frozen, the temporal ones are not
        vis = dict(model.visual.named_parameters())
        assert not vis["blocks.0.attn.qkv.weight"].requires_grad and vis["blocks.0.timeattn.qkv.weight"].requires_grad


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver launches): ONE JSON line on stdout with the contract's keys.  Run at
    2 frames so that it takes seconds (the metric string still names the benchmark; the sample is stated in `config`)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--frames", "2"], capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "clips/s" and d["higher_is_better"] is True and d["value"] > 0
    # "reference" = the unmodified reference installed at baseline/_ref (this container, shipped to the GPU box);
    # "port" = the oracle restatement, only when that install is absent
    ref_installed = os.path.isfile(os.path.join(root, "baseline", "_ref", "lavila", "models", "models.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_installed else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_grad_pool_hands_out_aligned_zero_views_and_falls_back():
    """engine.grad_pool: the fp32 gradient buffers of one autograd node are views of ONE zero-filled allocation (64-float = 256-byte
    aligned, contiguous, the parameter's shape); requests beyond the pool (or outside one) get their own torch.zeros."""
    from lavila_b200 import engine
    ps = [torch.nn.Parameter(torch.randn(5, 7)), torch.nn.Parameter(torch.randn(3)), torch.nn.Parameter(torch.randn(70, 2))]
    with engine.grad_pool(ps):
        gs = [engine._zeros_like_param(p) for p in ps]
        extra = engine._zeros_like_param(torch.nn.Parameter(torch.randn(1000)))          # does not fit any more
    base = gs[0].untyped_storage().data_ptr()
    for g, p in zip(gs, ps):
        assert g.shape == p.shape and g.dtype == torch.float32 and g.is_contiguous() and float(g.abs().sum()) == 0.0
        assert g.untyped_storage().data_ptr() == base and (g.data_ptr() - base) % 256 == 0
    assert extra.untyped_storage().data_ptr() != base and extra.shape == (1000,)
    gs[0].add_(1.0)
    assert float(gs[1].abs().sum()) == 0.0 and float(gs[2].abs().sum()) == 0.0            # disjoint slices
    lone = engine._zeros_like_param(ps[0])                                               # outside a pool
    assert lone.untyped_storage().data_ptr() != base and not engine._GRAD_POOL
