"""Generate golden vectors by running the UNMODIFIED reference (facebookresearch/LaViLa @ /root/reference)
in the build container.  The reference cannot travel to the GPU box, so its outputs are committed here.

    python tests/golden/make_golden.py          # writes tests/golden/dual_encoder_small.pt (+ _gated, + multirank)

What is pinned (all fp32, CPU, torch 2.11):
  * lavila.models.timesformer.SpaceTimeTransformer  (forward, forward_features(cls_at_last=False))
  * lavila.models.models.CLIP.forward (image_embed, text_embed, logit_scale), norm_embed True/False
  * lavila.models.loss.CLIPLoss (world_size=1) loss / acc, and gradients of the loss w.r.t. every parameter
    (small tensors in full; large ones as a fixed random sample of 512 entries + L2 norm)
  * 2-rank CLIPLoss(use_vissl=True) over gloo: loss, acc and local-embedding gradients (GatherLayer semantics)
Parameters are NOT stored: they are regenerated from oracle.dual_encoder.init_params(cfg, seed) (torch CPU generator),
and a checksum of the regenerated parameters is stored to detect RNG drift.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_shim  # noqa: E402
from oracle.dual_encoder import init_params, synthetic_batch  # noqa: E402

SMALL = dict(img_size=32, patch_size=16, embed_dim=128, depth=2, num_heads=2, num_frames=4, ln_pre=True,
             text_width=128, text_heads=2, text_layers=2, context_length=16, vocab_size=512, project_dim=64)


def build_reference(cfg, params, gated):
    from lavila.models.models import CLIP
    from lavila.models.openai_model import QuickGELU
    from lavila.models.timesformer import SpaceTimeTransformer
    import torch.nn as nn
    vis = SpaceTimeTransformer(img_size=cfg["img_size"], patch_size=cfg["patch_size"], embed_dim=cfg["embed_dim"],
                               depth=cfg["depth"], num_heads=cfg["num_heads"], num_frames=cfg["num_frames"],
                               time_init="zeros", attention_style="frozen-in-time", ln_pre=True, act_layer=QuickGELU,
                               is_tanh_gating=gated)
    vis.head = nn.Identity()          # models.py:347-349
    vis.pre_logits = nn.Identity()
    vis.fc = nn.Identity()
    model = CLIP(embed_dim=cfg["project_dim"], vision_width=cfg["embed_dim"], vision_model=vis,
                 context_length=cfg["context_length"], vocab_size=cfg["vocab_size"],
                 transformer_width=cfg["text_width"], transformer_heads=cfg["text_heads"],
                 transformer_layers=cfg["text_layers"])
    missing, unexpected = model.load_state_dict(params, strict=False)
    assert not unexpected, unexpected
    assert all("attn_mask" in m for m in missing) or not missing, missing
    return model


def sample_indices(numel, k=512, seed=7):
    g = torch.Generator().manual_seed(seed + numel)
    return torch.randint(0, numel, (k,), generator=g)


def summarise_grads(named_grads):
    out = {}
    for n, g in named_grads.items():
        g = g.detach().float()
        if g.numel() <= 4096:
            out[n] = {"full": g.clone()}
        else:
            idx = sample_indices(g.numel())
            out[n] = {"idx": idx, "sample": g.flatten()[idx].clone(), "norm": g.norm().clone()}
    return out


def param_checksum(params):
    return {k: float(v.double().sum()) for k, v in params.items()}


def run_case(cfg, gated, norm_embed, batch=3, seed=0):
    from lavila.models.loss import CLIPLoss
    params = init_params(cfg, seed=seed, gated=gated)
    model = build_reference(cfg, params, gated)
    frames, text = synthetic_batch(cfg, batch, seed=1234)
    out = model(frames, text, norm_embed=norm_embed)
    crit = CLIPLoss(use_vissl=False, cache_labels=True, rank=0, world_size=1)
    ld = crit(out)
    ld["loss"].backward()
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    with torch.no_grad():
        feats = model.visual.forward_features(frames.permute(0, 2, 1, 3, 4).contiguous(), cls_at_last=False)
    return {
        "cfg": cfg, "gated": gated, "norm_embed": norm_embed, "batch": batch, "param_seed": seed, "input_seed": 1234,
        "param_checksum": param_checksum(params),
        "frames_checksum": float(frames.double().sum()), "text": text.clone(),
        "image_embed": out["image_embed"].detach().clone(), "text_embed": out["text_embed"].detach().clone(),
        "logit_scale": out["logit_scale"].detach().clone(),
        "loss": ld["loss"].detach().clone(), "clip_acc": ld["clip_acc"].detach().clone(),
        "visual_tokens": feats[:, :5].clone(),   # first 5 tokens of norm(x) for every clip
        "grads": summarise_grads(grads),
    }


def _rank_worker(rank, world, cfg, port, ret):
    import torch.distributed as dist
    reference_shim.install()
    from lavila.models.loss import CLIPLoss
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    E = cfg["project_dim"]
    img = torch.nn.functional.normalize(torch.randn(4, E, generator=g), dim=-1).requires_grad_()
    txt = torch.nn.functional.normalize(torch.randn(4, E, generator=g), dim=-1).requires_grad_()
    res = {}
    for vissl in (True, False):
        crit = CLIPLoss(use_vissl=vissl, cache_labels=True, rank=rank, world_size=world)
        ld = crit({"image_embed": img, "text_embed": txt, "logit_scale": torch.tensor(14.2857)})
        gi, gt = torch.autograd.grad(ld["loss"], (img, txt))
        res["vissl" if vissl else "plain"] = {"loss": ld["loss"].detach(), "acc": ld["clip_acc"].detach(),
                                              "grad_image": gi, "grad_text": gt}
    res["image"], res["text"] = img.detach(), txt.detach()
    ret[rank] = res
    dist.destroy_process_group()


def run_multirank(cfg, world=2):
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_rank_worker, args=(world, cfg, 29533, ret), nprocs=world, join=True)
    return {"world": world, "ranks": [ret[r] for r in range(world)]}


def main():
    assert reference_shim.install(), "reference not found"
    torch.manual_seed(0)
    torch.set_num_threads(4)
    here = os.path.dirname(os.path.abspath(__file__))
    cases = {
        "plain": run_case(SMALL, gated=False, norm_embed=False),
        "norm": run_case(SMALL, gated=False, norm_embed=True),
        "gated_norm": run_case(SMALL, gated=True, norm_embed=True),
    }
    cases["multirank"] = run_multirank(SMALL)
    path = os.path.join(here, "dual_encoder_small.pt")
    torch.save(cases, path)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k in ("plain", "norm", "gated_norm"):
        print(k, "loss", float(cases[k]["loss"]), "acc", float(cases[k]["clip_acc"]))


if __name__ == "__main__":
    main()
