"""Golden vectors of the UNMODIFIED reference SSLCLIPLoss (lavila/models/loss.py:121-217), generated in the build container.

    python tests/golden/make_golden_ssl.py          # writes tests/golden/ssl_loss_small.pt

Cases: world_size 1 (three gt_indicator patterns incl. all-human and all-pseudo) and a 2-rank use_vissl run over gloo.
Stored: inputs, loss / accuracies / counts, gradients w.r.t. the embeddings, the (exp'ed) logit_scale input and the
module's logit_scale_pseudo parameter.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_shim  # noqa: E402

E = 64
SCALE = 14.2857


def _inputs(n, seed):
    g = torch.Generator().manual_seed(seed)
    img = torch.nn.functional.normalize(torch.randn(n, E, generator=g), dim=-1)
    txt = torch.nn.functional.normalize(img + 0.7 * torch.randn(n, E, generator=g), dim=-1)   # correlated pairs
    return img, txt


def _run(crit, img, txt, gt):
    img, txt = img.clone().requires_grad_(True), txt.clone().requires_grad_(True)
    scale = torch.tensor(SCALE, requires_grad=True)
    out = crit({"image_embed": img, "text_embed": txt, "logit_scale": scale}, gt)
    gi, gt_, gs, gp = torch.autograd.grad(out["loss"], (img, txt, scale, crit.logit_scale_pseudo))
    res = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in out.items()}
    res.update(grad_image=gi, grad_text=gt_, grad_scale=gs, grad_scale_pseudo=gp)
    return res


def _rank_worker(rank, world, port, ret):
    import torch.distributed as dist
    reference_shim.install()
    from lavila.models.loss import SSLCLIPLoss
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    img, txt = _inputs(6, 300 + rank)
    gt = torch.tensor([[1., 0., 1., 1., 0., 0.], [0., 0., 1., 0., 1., 1.]][rank])
    crit = SSLCLIPLoss(use_vissl=True, cache_labels=True, rank=rank, world_size=world, scale_init=0.08)
    res = _run(crit, img, txt, gt)
    res.update(image=img, text=txt, gt=gt)
    ret[rank] = res
    dist.destroy_process_group()


def main():
    assert reference_shim.install(), "reference not found"
    from lavila.models.loss import SSLCLIPLoss
    cases = {"scale": SCALE, "scale_init": 0.08, "world1": []}
    for seed, gt in ((1, [1., 0., 1., 1., 0., 0., 1., 0., 0., 1., 1., 0.]), (2, [1.] * 8), (3, [0.] * 8),
                     (4, [0., 1., 1., 1., 1., 1., 1., 1., 1., 1.])):
        img, txt = _inputs(len(gt), seed)
        gt = torch.tensor(gt)
        crit = SSLCLIPLoss(scale_init=0.08)
        res = _run(crit, img, txt, gt)
        res.update(image=img, text=txt, gt=gt)
        cases["world1"].append(res)
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_rank_worker, args=(2, 29541, ret), nprocs=2, join=True)
    cases["world2"] = [ret[r] for r in range(2)]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ssl_loss_small.pt")
    torch.save(cases, path)
    print("wrote", path, os.path.getsize(path), "bytes")
    for c in cases["world1"]:
        print({k: (float(v) if torch.is_tensor(v) and v.numel() == 1 else None) for k, v in c.items() if "acc" in k or k == "loss"})


if __name__ == "__main__":
    main()
