"""N > 1 CUDA path of CLIPLoss: the fused NVLink gather + loss kernel (lv_clip_loss_fwd_gather over symmetric memory) against the
NCCL all_gather route and fp32 torch on the concatenated batch -- loss, accuracy and the embedding gradients (GatherLayer
semantics: W x the single-process gradient, lavila/models/distributed_utils.py:51-67).  Needs >= 2 GPUs on the box (skipped on the
single-GPU test box; `bench.py --gpus N` runs the same comparison before its timed region on every multi-GPU bench)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from lavila_b200.models.loss import CLIPLoss
    B, E = 48, 256
    g = torch.Generator().manual_seed(100 + rank)
    img = torch.nn.functional.normalize(torch.randn(B, E, generator=g), dim=-1).to(dev)
    txt = torch.nn.functional.normalize(torch.randn(B, E, generator=g), dim=-1).to(dev)
    scale = torch.tensor(14.2857, device=dev)

    def run(crit):
        i, t = img.clone().requires_grad_(True), txt.clone().requires_grad_(True)
        ld = crit({"image_embed": i, "text_embed": t, "logit_scale": scale})
        ld["loss"].backward()
        return float(ld["loss"].detach()), float(ld["clip_acc"]), i.grad, t.grad

    fused = CLIPLoss(use_vissl=True, cache_labels=True, rank=rank, world_size=world)
    la, aa, gia, gta = run(fused)
    la2, _, gia2, _ = run(fused)                       # second call: the other slot of the symmetric block
    path = fused.gather_path
    fused.check_peer_error()
    os.environ["LAVILA_B200_P2P_LOSS"] = "0"
    lb, ab, gib, gtb = run(CLIPLoss(use_vissl=True, cache_labels=True, rank=rank, world_size=world))
    os.environ["LAVILA_B200_P2P_LOSS"] = "1"
    both = torch.cat((img, txt), 1)
    buf = [torch.empty_like(both) for _ in range(world)]
    dist.all_gather(buf, both)
    allb = torch.cat(buf, 0)
    ai, at = allb[:, :E].clone().requires_grad_(True), allb[:, E:].clone().requires_grad_(True)
    logits = (scale * ai) @ at.t()
    lab = torch.arange(logits.shape[0], device=dev)
    lc = (torch.nn.functional.cross_entropy(logits, lab) + torch.nn.functional.cross_entropy(logits.t(), lab)) / 2
    lc.backward()
    sl = slice(rank * B, (rank + 1) * B)
    ret[rank] = {"path": path, "fused_vs_nccl": abs(la - lb), "fused_vs_torch": abs(la - float(lc)), "repeat": abs(la - la2),
                 "acc": (aa, ab, 100.0 * float((logits.argmax(-1) == lab).float().mean())),
                 "grad_vs_nccl": float((gia - gib).abs().max()), "grad_vs_torch": float((gia - world * ai.grad[sl]).abs().max()),
                 "gradt_vs_torch": float((gta - world * at.grad[sl]).abs().max()), "grad_repeat": float((gia - gia2).abs().max())}
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_gather_loss_two_gpus():
    import torch.multiprocessing as mp
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, 29641, ret), nprocs=world, join=True)
    for r in range(world):
        d = ret[r]
        assert d["path"] == "p2p", d
        assert d["fused_vs_nccl"] <= 1e-6 and d["fused_vs_torch"] <= 1e-5 and d["repeat"] == 0.0, d
        assert d["acc"][0] == d["acc"][1] and abs(d["acc"][0] - d["acc"][2]) < 1e-3, d
        assert d["grad_vs_nccl"] <= 1e-6 and d["grad_vs_torch"] <= 1e-5 and d["gradt_vs_torch"] <= 1e-5 and d["grad_repeat"] == 0.0, d
