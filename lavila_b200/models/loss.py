"""CLIPLoss / SSLCLIPLoss -- mirror of lavila/models/loss.py:18-217 on the fused B200 kernels.

forward(outputs) takes the dict returned by CLIP.forward and returns {'loss', 'clip_loss', 'clip_acc'} (0-dim tensors).
Multi-GPU (world_size > 1): one all-gather of [image | text] embeddings over NCCL/NVLink, then every rank evaluates the
global N x N loss redundantly (exactly what the reference does, loss.py:76-79) in one kernel.  Backward needs no
collective: all ranks hold the same global loss, so the reference's all_reduce(SUM) of identical per-rank embedding
gradients (distributed_utils.py:64-67) equals a multiplication by world_size, applied as `grad_scale`.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .distributed_utils import gather_embeddings, gather_embeddings_gt

F32 = torch.float32


def gather_features(image_features, text_features, local_loss=False, gather_with_grad=False, rank=0, world_size=1):
    """loss.py:18-43 (kept for API parity; returns tensors that carry gradient only through the local slot unless
    gather_with_grad)."""
    import torch.distributed as dist
    import torch.distributed.nn
    if gather_with_grad:
        all_image = torch.cat(torch.distributed.nn.all_gather(image_features), dim=0)
        all_text = torch.cat(torch.distributed.nn.all_gather(text_features), dim=0)
    else:
        gi = [torch.zeros_like(image_features) for _ in range(world_size)]
        gt = [torch.zeros_like(text_features) for _ in range(world_size)]
        dist.all_gather(gi, image_features)
        dist.all_gather(gt, text_features)
        if not local_loss:
            gi[rank] = image_features
            gt[rank] = text_features
        all_image, all_text = torch.cat(gi, dim=0), torch.cat(gt, dim=0)
    return all_image, all_text


def _peer_exchange(state, B, E, dev, world_size):
    """The symmetric-memory blocks of the fused gather + loss kernel, created on first use (collective).  Disabled with
    LAVILA_B200_P2P_LOSS=0 (then: one NCCL all_gather + the single-GPU loss kernel).  The same NCCL route is taken -- on
    every rank, agreed BEFORE the symmetric-memory rendezvous -- when the global batch exceeds what the cooperative kernel
    supports (`lv_clip_loss_gather_max_rows`: ~1184 rows at E = 256 on a B200) or symmetric memory is not importable."""
    import os
    if os.environ.get("LAVILA_B200_P2P_LOSS", "1") == "0" or dev.type != "cuda":
        return None
    if state.get("xch_unavailable"):
        return None
    pool = state.setdefault("xch_pool", {})       # one exchange per (B, E): train and validation batches may differ
    x = pool.get((B, E))
    if x is None:
        if (B, E) in state.setdefault("xch_too_large", set()):
            return None
        import torch.distributed as dist
        # ---- go / no-go, agreed across ranks before anyone enters the (collective, blocking) rendezvous
        why = None
        max_rows = ops.clip_loss_gather_max_rows(E)
        if B * world_size > max_rows:
            why = "global batch %d > %d rows supported by the cooperative kernel" % (B * world_size, max_rows)
        else:
            try:
                import importlib
                importlib.import_module("torch.distributed._symmetric_memory")
                from .distributed_utils import PeerEmbeddingExchange
            except Exception as e:
                why = "torch symmetric memory unavailable (%r)" % (e,)
        ok = torch.tensor([0 if why else 1], device=dev, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if dist.get_rank() == 0:
                print("lavila_b200: fused NVLink gather not used for B=%d E=%d (%s); CLIPLoss uses one NCCL all_gather"
                      % (B, E, why or "a peer rank declined"))
            state["xch_too_large"].add((B, E))
            return None
        err = None
        try:
            x = PeerEmbeddingExchange(B, E, dev)
        except Exception as e:   # no peer access / no symmetric-memory support on this box
            x, err = None, e
        ok = torch.tensor([0 if x is None else 1], device=dev, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)     # all ranks take the same path
        if int(ok.item()) == 0:
            if dist.get_rank() == 0:
                print("lavila_b200: symmetric-memory peer exchange unavailable (%r); CLIPLoss uses one NCCL all_gather" % (err,))
            state["xch_unavailable"] = True
            pool.clear()
            state["xch"] = None
            return None
        pool[(B, E)] = x
    state["xch"] = x                               # the exchange used by the latest call (inspected by tools / tests)
    return x


class _ClipLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, text, logit_scale, rank, world_size, grad_scale, state):
        image = image.contiguous().float()
        text = text.contiguous().float()
        B, E = image.shape
        dev = image.device
        scale = logit_scale.detach().reshape(1).contiguous().float()
        Ng = B * world_size
        lse_i = torch.empty(Ng, device=dev, dtype=F32)
        lse_t = torch.empty(Ng, device=dev, dtype=F32)
        partial = torch.empty(2 * Ng, device=dev, dtype=F32)
        result = torch.empty(2, device=dev, dtype=F32)
        xch = _peer_exchange(state, B, E, dev, world_size) if world_size > 1 else None
        state["path"] = "p2p" if xch is not None else ("nccl" if world_size > 1 else "local")
        if xch is not None:
            # ONE kernel: publish own rows -> pull the peers' rows over NVLink -> global loss (csrc/clip_loss.cu)
            xch.check_error()          # a peer timeout of an earlier step raises here (LavilaB200Error), never silently
            all_i = torch.empty(Ng, E, device=dev, dtype=F32)
            all_t = torch.empty(Ng, E, device=dev, dtype=F32)
            ops.clip_loss_fwd_gather(image, text, xch.peers_dev, rank, world_size, B, xch.next_step(), all_i, all_t, scale, E,
                                     lse_i, lse_t, partial, xch.ctrl, result, timeout_ms=xch.timeout_ms)
            xch.post_launch()
        else:
            if world_size > 1:
                all_i, all_t = gather_embeddings(image, text, world_size)     # NCCL all_gather (LAVILA_B200_P2P_LOSS=0)
            else:
                all_i, all_t = image, text
            if state.get("counter") is None or state["counter"].device != dev:
                state["counter"] = torch.zeros(1, device=dev, dtype=torch.int32)
            ops.clip_loss_fwd(all_i, all_t, scale, Ng, E, lse_i, lse_t, partial, state["counter"], result)
        ctx.saved = (all_i, all_t, scale, lse_i, lse_t)
        ctx.meta = (B, E, Ng, rank, world_size, grad_scale)
        loss, acc = result[0], result[1]
        ctx.mark_non_differentiable(acc)           # clip_acc is computed under no_grad in the reference (loss.py:113-116)
        return loss, acc

    @staticmethod
    def backward(ctx, gloss, gacc):
        all_i, all_t, scale, lse_i, lse_t = ctx.saved
        ctx.saved = None
        B, E, Ng, rank, world_size, grad_scale = ctx.meta
        dev = all_i.device
        d_i = torch.empty(B, E, device=dev, dtype=F32)
        d_t = torch.empty(B, E, device=dev, dtype=F32)
        d_s = torch.zeros(1, device=dev, dtype=F32)
        g = gloss.reshape(1).contiguous().float()
        # Only this rank's B rows are differentiated (2B CTAs instead of 2*W*B).  The logit_scale gradient is the double sum
        # over (local image rows) x (all texts) times W: the W ranks' shares add up to W x the full sum, and DDP's mean over
        # ranks (the only way a world_size > 1 gradient is consumed, main_pretrain.py:180) gives exactly the reference's
        # value, whose ranks each hold the identical full sum (loss.py:78, distributed_utils.py:64-67).
        ops.clip_loss_bwd(all_i, all_t, scale, lse_i, lse_t, g, grad_scale, float(world_size), Ng, E, rank * B, B, d_i, d_t, d_s)
        return d_i, d_t, d_s.reshape(()), None, None, None, None


class CLIPLoss(nn.Module):
    """loss.py:46-118."""

    def __init__(self, use_vissl=False, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0,
                 world_size=1):
        super().__init__()
        self.use_vissl = use_vissl
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        self.prev_num_logits = 0
        self.labels = {}
        self._state = {}
        if local_loss:
            raise NotImplementedError("local_loss=True is never enabled by the reference's get_loss (models.py:295-300)")

    def forward(self, outputs):
        image_features = outputs['image_embed']
        text_features = outputs['text_embed']
        logit_scale = outputs['logit_scale']
        # embedding-gradient scale: GatherLayer / gather_with_grad sum W identical copies; gather_features keeps 1x
        grad_scale = float(self.world_size) if (self.world_size > 1 and (self.use_vissl or self.gather_with_grad)) else 1.0
        loss, acc = _ClipLossFn.apply(image_features, text_features, logit_scale, self.rank, self.world_size,
                                      grad_scale, self._state)
        return {'loss': loss, 'clip_loss': loss, 'clip_acc': acc}

    @property
    def gather_path(self):
        """'p2p' (fused NVLink gather + loss kernel), 'nccl' (all_gather + loss kernel) or 'local' (world size 1): the route
        the latest forward took."""
        return self._state.get("path")

    def check_peer_error(self):
        """Block until the latest fused-gather launch has finished and raise LavilaB200Error if a peer timed out."""
        x = self._state.get("xch")
        if x is not None:
            x.check_error(wait=True)


class _SSLClipLossFn(torch.autograd.Function):
    """SSLCLIPLoss core (loss.py:148-213) on the gathered batch: one forward kernel (logits with the per-pair scale,
    both log-sum-exps, arg-max, the three accuracies) and one backward kernel; as for CLIPLoss, no backward collective."""

    @staticmethod
    def forward(ctx, image, text, logit_scale, scale_pseudo, gt, rank, world_size, grad_scale, state):
        image = image.contiguous().float()
        text = text.contiguous().float()
        B, E = image.shape
        dev = image.device
        gt_f = gt.detach().reshape(-1).to(device=dev, dtype=F32)
        if world_size > 1:
            all_i, all_t, all_g = gather_embeddings_gt(image, text, gt_f, world_size)
        else:
            all_i, all_t, all_g = image, text, gt_f
        gt_i = all_g.round().to(torch.int32).contiguous()
        Ng = all_i.shape[0]
        scale = logit_scale.detach().reshape(1).contiguous().float()
        scale_p = scale_pseudo.detach().reshape(1).contiguous().float()
        lse_i = torch.empty(Ng, device=dev, dtype=F32)
        lse_t = torch.empty(Ng, device=dev, dtype=F32)
        partial = torch.empty(2 * Ng, device=dev, dtype=F32)
        result = torch.empty(6, device=dev, dtype=F32)
        if state.get("counter") is None or state["counter"].device != dev:
            state["counter"] = torch.zeros(1, device=dev, dtype=torch.int32)
        ops.ssl_clip_loss_fwd(all_i, all_t, scale, scale_p, gt_i, Ng, E, lse_i, lse_t, partial, state["counter"], result)
        ctx.saved = (all_i, all_t, scale, scale_p, gt_i, lse_i, lse_t)
        ctx.meta = (B, E, Ng, rank, world_size, grad_scale)
        stats = result[1:]
        ctx.mark_non_differentiable(stats)
        return result[0], stats

    @staticmethod
    def backward(ctx, gloss, gstats):
        all_i, all_t, scale, scale_p, gt_i, lse_i, lse_t = ctx.saved
        ctx.saved = None
        B, E, Ng, rank, world_size, grad_scale = ctx.meta
        dev = all_i.device
        d_i = torch.empty(Ng, E, device=dev, dtype=F32)
        d_t = torch.empty(Ng, E, device=dev, dtype=F32)
        d_s = torch.zeros(2, device=dev, dtype=F32)
        g = gloss.reshape(1).contiguous().float()
        ops.ssl_clip_loss_bwd(all_i, all_t, scale, scale_p, gt_i, lse_i, lse_t, g, grad_scale, 1.0, Ng, E, 0, Ng, d_i, d_t, d_s)
        sl = slice(rank * B, (rank + 1) * B)
        return d_i[sl], d_t[sl], d_s[0].reshape(()), d_s[1].reshape(()), None, None, None, None, None


class SSLCLIPLoss(nn.Module):
    """loss.py:121-217: dual-temperature contrastive loss over human (gt_indicators = 1) and pseudo-narrated (0) pairs,
    selected by main_pretrain.py:189-197 when --metadata-aux is given (the LaViLa recipe, docs/PRETRAIN.md:109-122)."""

    def __init__(self, use_vissl=False, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1,
                 scale_init=0.08, freeze_scale=False):
        super().__init__()
        self.use_vissl = use_vissl
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        self.logit_scale_pseudo = nn.Parameter(torch.ones([]) * np.log(1 / scale_init))
        if freeze_scale:
            self.logit_scale_pseudo.requires_grad = False
        self.prev_num_logits = 0
        self.labels = {}
        self._state = {}

    def forward(self, outputs, gt_indicators):
        if self.world_size > 1 and not self.use_vissl:
            raise NotImplementedError   # loss.py:167-168
        logit_scale_pseudo = self.logit_scale_pseudo.exp()
        grad_scale = float(self.world_size) if self.world_size > 1 else 1.0   # GatherLayer semantics (use_vissl)
        loss, st = _SSLClipLossFn.apply(outputs['image_embed'], outputs['text_embed'], outputs['logit_scale'],
                                        logit_scale_pseudo, gt_indicators, self.rank, self.world_size, grad_scale,
                                        self._state)
        # the reference returns the two counts as CPU LongTensors of shape [1] (loss.py:210); that forces one D2H read
        counts = st[3:5].round().to(torch.int64).cpu()
        return {'loss': loss, 'clip_loss': loss, 'num_gt': counts[0:1], 'num_pseudo': counts[1:2],
                'clip_acc': st[0], 'clip_acc_gt': st[1], 'clip_acc_pseudo': st[2]}
