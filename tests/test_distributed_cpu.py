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
