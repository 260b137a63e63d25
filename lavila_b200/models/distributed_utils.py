"""Mirror of lavila/models/distributed_utils.py (GatherLayer / gather_from_all) on torch.distributed.

The fused CLIPLoss (lavila_b200.models.loss) does not go through GatherLayer -- it gathers [image | text] with ONE
collective and needs no backward collective at all -- but the symbols are kept so code importing them keeps working,
with the reference's semantics (forward all_gather, backward all_reduce(SUM) of the stacked gradients, own slice).
"""
from typing import Tuple

import torch
import torch.distributed as dist


def convert_to_distributed_tensor(tensor: torch.Tensor) -> Tuple[torch.Tensor, str]:
    """distributed_utils.py:17-30."""
    orig_device = "cpu" if not tensor.is_cuda else "gpu"
    if dist.is_available() and dist.get_backend() == dist.Backend.NCCL and not tensor.is_cuda:
        tensor = tensor.cuda()
    return tensor, orig_device


def convert_to_normal_tensor(tensor: torch.Tensor, orig_device: str) -> torch.Tensor:
    """distributed_utils.py:33-40."""
    if tensor.is_cuda and orig_device == "cpu":
        tensor = tensor.cpu()
    return tensor


def is_distributed_training_run() -> bool:
    """distributed_utils.py:43-48."""
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class GatherLayer(torch.autograd.Function):
    """distributed_utils.py:51-67."""

    @staticmethod
    def forward(ctx, x):
        output = [torch.zeros_like(x) for _ in range(dist.get_world_size())]
        dist.all_gather(output, x)
        return tuple(output)

    @staticmethod
    def backward(ctx, *grads):
        all_gradients = torch.stack(grads)
        dist.all_reduce(all_gradients)
        return all_gradients[dist.get_rank()]


def gather_from_all(tensor: torch.Tensor) -> torch.Tensor:
    """distributed_utils.py:70-89."""
    if tensor.ndim == 0:
        tensor = tensor.unsqueeze(0)
    if is_distributed_training_run():
        tensor, orig_device = convert_to_distributed_tensor(tensor)
        gathered = GatherLayer.apply(tensor)
        gathered = [convert_to_normal_tensor(t, orig_device) for t in gathered]
    else:
        gathered = [tensor]
    return torch.cat(gathered, 0)


def gather_embeddings(image_features, text_features, world_size):
    """ONE all_gather of the concatenated [B, 2E] block instead of the reference's two (loss.py:76-77 / :32-35).
    Returns (all_image [W*B, E], all_text [W*B, E]) without autograd history."""
    B, E = image_features.shape
    local = torch.cat((image_features.detach(), text_features.detach()), dim=1).contiguous()
    buf = [torch.empty_like(local) for _ in range(world_size)]
    dist.all_gather(buf, local)
    allb = torch.stack(buf, 0).view(world_size * B, 2 * E)
    return allb[:, :E].contiguous(), allb[:, E:].contiguous()


def gather_embeddings_gt(image_features, text_features, gt_indicators, world_size):
    """SSLCLIPLoss (loss.py:155-157 issues three gather_from_all): ONE all_gather of [image | text | gt] ([B, 2E+1]).
    Returns (all_image [W*B, E], all_text [W*B, E], all_gt [W*B] float) in rank order, without autograd history."""
    B, E = image_features.shape
    gt = gt_indicators.detach().reshape(-1, 1).to(device=image_features.device, dtype=image_features.dtype)
    local = torch.cat((image_features.detach(), text_features.detach(), gt), dim=1).contiguous()
    buf = [torch.empty_like(local) for _ in range(world_size)]
    dist.all_gather(buf, local)
    allb = torch.stack(buf, 0).view(world_size * B, 2 * E + 1)
    return allb[:, :E].contiguous(), allb[:, E:2 * E].contiguous(), allb[:, 2 * E].contiguous()


class PeerEmbeddingExchange:
    """Symmetric (NVLink peer-mapped) staging blocks for the fused gather + loss kernel (csrc/clip_loss.cu): every rank owns
    2 slots x ([B x 2E fp32] + 32 flag words); all ranks hold device pointers to all blocks.  Built on
    torch.distributed._symmetric_memory (CUDA VMM + fabric handles over NVLink/NVSwitch); torch only allocates and exchanges
    the handles -- the data movement is the kernel's own peer loads."""
    PAD = 32

    def __init__(self, B, E, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass
        self.B, self.E = B, E
        self.slot_floats = B * 2 * E + self.PAD
        self.buf = symm_mem.empty(2 * self.slot_floats, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group)
        self.peers_dev = int(self.handle.buffer_ptrs_dev)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.ctrl = torch.zeros(8, device=device, dtype=torch.int32)      # word 4 = sticky peer-timeout flag (clip_loss.cu)
        self.err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.err_event = None
        self.step = 0
        # how long the kernel waits for a peer's rows: the process group's own collective timeout (what the reference's
        # NCCL all_gather would wait, distributed_utils.py:56), overridable for tests
        import os
        t = os.environ.get("LAVILA_B200_P2P_TIMEOUT_S")
        if t is not None:
            self.timeout_ms = max(1, int(float(t) * 1000))
        else:
            try:
                self.timeout_ms = max(1000, int(group._get_backend(torch.device(device)).options._timeout.total_seconds() * 1000))
            except Exception:
                self.timeout_ms = 600000
        torch.cuda.synchronize(device)
        dist.barrier(group)          # nobody polls a flag before every block is zeroed

    def next_step(self):
        self.step += 1
        return self.step

    def post_launch(self):
        """Enqueue an async D2H copy of the error word behind the kernel (no host sync)."""
        self.err_host.copy_(self.ctrl[4:5], non_blocking=True)
        self.err_event = torch.cuda.Event()
        self.err_event.record()

    def check_error(self, wait=False):
        """Raise if an earlier fused gather timed out waiting for a peer.  Called before the next launch (by then the
        training loop's `loss.item()` has drained the stream, so the query is free) and by `CLIPLoss.check_peer_error()`."""
        if self.err_event is None:
            return
        if wait:
            self.err_event.synchronize()
        if self.err_event.query():
            self.err_event = None
            bad = int(self.err_host[0])
            if bad:
                from .._lib import LavilaB200Error
                raise LavilaB200Error("fused NVLink gather + CLIPLoss: a peer rank did not publish its embeddings within %.0f s "
                                      "at exchange step %d (rank %d); the loss of that step is NaN"
                                      % (self.timeout_ms / 1e3, bad, self.rank))
