"""Host-side logic of the N > 1 path on CPU (gloo, world_size 2): the single fused all-gather of [image | text]
must reproduce the reference's rank-ordered concatenation (loss.py:76-77 / distributed_utils.py:59-60), and the
local-row slice used by the backward must address this rank's rows."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lavila_b200.models.distributed_utils import gather_embeddings, gather_from_all
    B, E = 3, 8
    g = torch.Generator().manual_seed(10 + rank)
    img, txt = torch.randn(B, E, generator=g), torch.randn(B, E, generator=g)
    all_i, all_t = gather_embeddings(img, txt, world)
    ref_i = gather_from_all(img)          # reference semantics (GatherLayer)
    ref_t = gather_from_all(txt)
    ok = torch.equal(all_i, ref_i) and torch.equal(all_t, ref_t)
    ok = ok and torch.equal(all_i[rank * B:(rank + 1) * B], img) and torch.equal(all_t[rank * B:(rank + 1) * B], txt)
    # GatherLayer backward = all_reduce(SUM) of the stacked grads, own slice: W identical copies -> W x local grad
    x = img.clone().requires_grad_(True)
    gathered = gather_from_all(x)
    (gathered * torch.arange(world * B * E, dtype=torch.float32).view(world * B, E)).sum().backward()
    expect = world * torch.arange(world * B * E, dtype=torch.float32).view(world * B, E)[rank * B:(rank + 1) * B]
    ok = ok and torch.allclose(x.grad, expect)
    # SSLCLIPLoss: [image | text | gt] in one collective == three reference gathers (loss.py:155-157)
    from lavila_b200.models.distributed_utils import gather_embeddings_gt
    gt = torch.tensor([1., 0., 1.]) if rank == 0 else torch.tensor([0., 0., 1.])
    a_i, a_t, a_g = gather_embeddings_gt(img, txt, gt, world)
    ok = ok and torch.equal(a_i, ref_i) and torch.equal(a_t, ref_t) and torch.equal(a_g, gather_from_all(gt))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_gather_embeddings_matches_reference_gather_layer():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29541, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)


def test_cliploss_world_size_grad_scale():
    from lavila_b200.models.loss import CLIPLoss
    assert CLIPLoss(use_vissl=True, rank=1, world_size=8).world_size == 8
    import pytest
    with pytest.raises(NotImplementedError):
        CLIPLoss(local_loss=True)


# ----------------------------------------------------------------------------------------------------------------------
# The loss modules' N > 1 HOST logic (gather order, local-row slice, GatherLayer gradient scale) on CPU over gloo.  The two
# kernel entry points are replaced by torch test doubles computing the same quantities (this file tests the host side; the
# kernels themselves are tested on the GPU), and the results are compared with 2-rank goldens of the unmodified reference.
def _double_clip_loss_fwd(img, txt, scale, Ng, E, lse_i, lse_t, partial, counter, result):
    logits = scale * img @ txt.t()
    lab = torch.arange(Ng)
    lse_i.copy_(torch.logsumexp(logits, 1))
    lse_t.copy_(torch.logsumexp(logits.t(), 1))
    result[0] = (torch.nn.functional.cross_entropy(logits, lab) + torch.nn.functional.cross_entropy(logits.t(), lab)) / 2
    result[1] = 100.0 * (logits.argmax(-1) == lab).float().mean()


def _double_clip_loss_bwd(img, txt, scale, lse_i, lse_t, gout, grad_scale, scale_grad_scale, Ng, E, r0, Nl, d_img, d_txt, d_scale):
    with torch.enable_grad():      # autograd runs Function.backward with grad mode off
        i, t, s = img.clone().requires_grad_(True), txt.clone().requires_grad_(True), scale.clone().requires_grad_(True)
        logits = s * i @ t.t()
        lab = torch.arange(Ng)
        loss = (torch.nn.functional.cross_entropy(logits, lab) + torch.nn.functional.cross_entropy(logits.t(), lab)) / 2
        gi, gt, gs = torch.autograd.grad(loss, (i, t, s))
    d_img.copy_(gout * grad_scale * gi[r0:r0 + Nl])
    d_txt.copy_(gout * grad_scale * gt[r0:r0 + Nl])
    if d_scale is not None:
        d_scale += gout * scale_grad_scale * gs


def _loss_worker(rank, world, port, gold, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lavila_b200 import ops
    from lavila_b200.models.loss import CLIPLoss
    ops.clip_loss_fwd, ops.clip_loss_bwd = _double_clip_loss_fwd, _double_clip_loss_bwd
    ok = True
    r = gold["ranks"][rank]
    for key, vissl in (("vissl", True), ("plain", False)):
        img, txt = r["image"].clone().requires_grad_(True), r["text"].clone().requires_grad_(True)
        crit = CLIPLoss(use_vissl=vissl, cache_labels=True, rank=rank, world_size=world)
        ld = crit({"image_embed": img, "text_embed": txt, "logit_scale": torch.tensor(14.2857)})
        gi, gt = torch.autograd.grad(ld["loss"], (img, txt))
        ok = ok and torch.allclose(ld["loss"], r[key]["loss"], rtol=1e-5, atol=1e-6)
        ok = ok and float(ld["clip_acc"]) == float(r[key]["acc"])
        ok = ok and torch.allclose(gi, r[key]["grad_image"], rtol=1e-4, atol=1e-7)      # W x for vissl, 1 x for gather_features
        ok = ok and torch.allclose(gt, r[key]["grad_text"], rtol=1e-4, atol=1e-7)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_cliploss_two_rank_host_logic_matches_reference():
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "dual_encoder_small.pt"), weights_only=False)["multirank"]
    world = gold["world"]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_loss_worker, args=(world, 29547, gold, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)


def _ssl_worker(rank, world, port, gold, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lavila_b200 import ops
    from lavila_b200.models.loss import SSLCLIPLoss
    from tests import ops_doubles
    ops.ssl_clip_loss_fwd, ops.ssl_clip_loss_bwd = ops_doubles.ssl_clip_loss_fwd, ops_doubles.ssl_clip_loss_bwd
    r = gold["world2"][rank]
    img, txt = r["image"].clone().requires_grad_(True), r["text"].clone().requires_grad_(True)
    s = torch.tensor(gold["scale"], requires_grad=True)
    crit = SSLCLIPLoss(use_vissl=True, cache_labels=True, rank=rank, world_size=world, scale_init=gold["scale_init"])
    out = crit({"image_embed": img, "text_embed": txt, "logit_scale": s}, r["gt"])
    gi, gt, gs, gp = torch.autograd.grad(out["loss"], (img, txt, s, crit.logit_scale_pseudo))
    ok = torch.allclose(out["loss"], r["loss"], rtol=1e-5, atol=1e-6)
    for k in ("clip_acc", "clip_acc_gt", "clip_acc_pseudo"):
        ok = ok and abs(float(out[k]) - float(r[k])) < 1e-3
    ok = ok and int(out["num_gt"]) == int(r["num_gt"]) and int(out["num_pseudo"]) == int(r["num_pseudo"])
    ok = ok and torch.allclose(gi, r["grad_image"], rtol=1e-4, atol=1e-7) and torch.allclose(gt, r["grad_text"], rtol=1e-4, atol=1e-7)
    ok = ok and torch.allclose(gs, r["grad_scale"], rtol=1e-4, atol=1e-7) and torch.allclose(gp, r["grad_scale_pseudo"], rtol=1e-4, atol=1e-7)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_sslcliploss_two_rank_host_logic_matches_reference():
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ssl_loss_small.pt"), weights_only=False)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ssl_worker, args=(2, 29549, gold, ret), nprocs=2, join=True)
    assert all(ret[r] for r in range(2)), dict(ret)


def _install_doubles():
    from lavila_b200 import engine, ops
    from tests import ops_doubles
    for name in ops_doubles.DOUBLES:
        setattr(ops, name, getattr(ops_doubles, name))
    engine.SHADOW.clear()


def _toy_model_and_batch(batch):
    from oracle import dual_encoder as O
    from tests.test_host_schedule_cpu import GOLD, _build
    cfg = GOLD["norm"]["cfg"]
    model = _build(cfg, O.init_params(cfg, seed=4), gated=False)
    frames, text = O.synthetic_batch(cfg, batch, seed=8)
    return model, frames, text


PROBE = ["visual.blocks.0.attn.qkv.weight", "visual.blocks.1.mlp.fc2.bias", "visual.temporal_embed", "visual.cls_token",
         "transformer.resblocks.0.mlp.c_fc.weight", "token_embedding.weight", "image_projection", "logit_scale"]


def _ddp_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _install_doubles()
    from lavila_b200.models.loss import CLIPLoss
    model, frames, text = _toy_model_and_batch(2 * world)
    ddp = torch.nn.parallel.DistributedDataParallel(model)
    crit = CLIPLoss(use_vissl=True, cache_labels=True, rank=rank, world_size=world)
    sl = slice(2 * rank, 2 * rank + 2)
    ld = crit(ddp(frames[sl], text[sl], norm_embed=True))
    ld["loss"].backward()
    named = dict(model.named_parameters())
    ret[rank] = {"loss": ld["loss"].detach().clone(), "acc": ld["clip_acc"].detach().clone(),
                 "grads": {k: named[k].grad.clone() for k in PROBE}}
    dist.destroy_process_group()


def test_ddp_two_ranks_equal_single_process_on_the_global_batch(monkeypatch):
    """DistributedDataParallel (gloo) around the model + CLIPLoss(use_vissl) on 2 ranks with 2 clips each == one process on the 4
    clips: same loss / accuracy on every rank, and DDP's averaged parameter gradients equal the single-process gradients (the
    embedding gradient is scaled by W in the loss, DDP divides by W: SURVEY.md 8(a) a13).  Kernel wrappers = test doubles."""
    from lavila_b200.models.loss import CLIPLoss
    from tests import ops_doubles
    from tests.util import rel_l2
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ddp_worker, args=(world, 29553, ret), nprocs=world, join=True)
    ops_doubles.install(monkeypatch)
    model, frames, text = _toy_model_and_batch(2 * world)
    ld = CLIPLoss()(model(frames, text, norm_embed=True))
    ld["loss"].backward()
    named = dict(model.named_parameters())
    for r in range(world):
        assert abs(float(ret[r]["loss"]) - float(ld["loss"])) < 1e-5 and float(ret[r]["acc"]) == float(ld["clip_acc"])
        for k in PROBE:
            assert rel_l2(ret[r]["grads"][k], named[k].grad) < 2e-3, (r, k, rel_l2(ret[r]["grads"][k], named[k].grad))


# ----------------------------------------------------------------------------------------------------------------------
# ZeroRedundancyOptimizer (main_pretrain.py:215-219, the TSF-L@HR recipe): after its step every rank holds updated
# weights it did NOT step itself -- they arrive by a broadcast into `param.data`, which does not bump `param._version`.  The bf16
# GEMM-operand shadows (engine.SHADOW) must follow them: they are keyed on the optimizer-step generation too.
def _zero_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _install_doubles()
    from torch.distributed.optim import ZeroRedundancyOptimizer
    from lavila_b200 import engine
    from lavila_b200.models.loss import CLIPLoss
    model, frames, text = _toy_model_and_batch(2 * world)
    ddp = torch.nn.parallel.DistributedDataParallel(model)
    crit = CLIPLoss(use_vissl=True, cache_labels=True, rank=rank, world_size=world)
    opt = ZeroRedundancyOptimizer(model.parameters(), optimizer_class=torch.optim.AdamW, lr=1e-2, weight_decay=0.01)
    sl = slice(2 * rank, 2 * rank + 2)
    losses, stale = [], 0
    for it in range(3):
        ld = crit(ddp(frames[sl], text[sl], norm_embed=True))
        opt.zero_grad(set_to_none=True)
        ld["loss"].backward()
        opt.step()
        losses.append(float(ld["loss"]))
        for n, p in model.named_parameters():
            if p.ndim == 2 and "weight" in n:               # the GEMM weights are the shadowed ones
                if not torch.equal(engine.SHADOW.get(p).float(), p.detach().to(torch.bfloat16).float()):
                    stale += 1
    # all ranks hold identical weights after ZeRO's broadcast
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    other = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(other, flat)
    ret[rank] = {"losses": losses, "stale": stale, "same": bool(all(torch.equal(o, flat) for o in other))}
    dist.destroy_process_group()


def test_zero_redundancy_optimizer_two_ranks_refreshes_the_bf16_shadows():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_zero_worker, args=(world, 29557, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret[r]["stale"] == 0, dict(ret)
        assert ret[r]["same"]
        assert ret[r]["losses"][-1] < ret[r]["losses"][0], ret[r]["losses"]        # and it trains
    assert ret[0]["losses"] == ret[1]["losses"]


def test_shadow_follows_param_data_writes_after_invalidate(monkeypatch):
    from lavila_b200 import engine
    from tests import ops_doubles
    ops_doubles.install(monkeypatch)
    p = torch.nn.Parameter(torch.randn(8, 8))
    a = engine.SHADOW.get(p).clone()
    p.data.add_(1.0)                               # does not bump p._version
    engine.invalidate_param_caches()
    b = engine.SHADOW.get(p)
    assert not torch.equal(a, b) and torch.equal(b.float(), p.detach().to(torch.bfloat16).float())
    opt = torch.optim.SGD([p], lr=0.5)
    p.grad = torch.ones_like(p)
    g0 = engine.param_generation()
    opt.step()
    assert engine.param_generation() > g0          # the global post-step hook
    assert torch.equal(engine.SHADOW.get(p).float(), p.detach().to(torch.bfloat16).float())


def test_clip_acc_carries_no_graph(monkeypatch):
    from lavila_b200.models.loss import CLIPLoss
    from tests import ops_doubles
    ops_doubles.install(monkeypatch)
    i = torch.nn.functional.normalize(torch.randn(4, 8), dim=-1).requires_grad_(True)
    t = torch.nn.functional.normalize(torch.randn(4, 8), dim=-1).requires_grad_(True)
    ld = CLIPLoss()({"image_embed": i, "text_embed": t, "logit_scale": torch.tensor(14.0)})
    assert ld["loss"].requires_grad and not ld["clip_acc"].requires_grad and ld["clip_acc"].grad_fn is None
