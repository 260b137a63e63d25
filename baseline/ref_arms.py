"""The reference arms of bench.py: the UNMODIFIED reference (baseline/_ref, facebookresearch/LaViLa) driven through its own
public API -- `lavila.models.models.CLIP_OPENAI_TIMESFORMER_BASE(...)`, `models.get_loss(...)` -- with the body of its own
`train()` loop (main_pretrain.py:486-530) around it.  None of lavila_b200's models, kernels or engine is on this path.

    reference_step_fn(device='cuda', amp='bf16' | 'fp16' | 'off')   -> step(frames, tokens) -> loss (python float)

* GPU ("eager") arm: `torch.autocast('cuda', bf16)` (the dtype BASELINE names) or the reference's literal mode
  (`torch.cuda.amp.autocast()` = fp16 + `GradScaler`, main_pretrain.py:223,490), AdamW with the reference's weight-decay
  grouping (main_pretrain.py:199-213).
* CPU arm: the same modules on the host cores, fp32 (amp off: CUDA autocast does nothing on CPU tensors).
"""
import argparse
import contextlib
import os
import sys

from . import ref_shim


def _random_openai_clip_vit_b16(*a, **k):
    """Stands in for `load_openai_clip('ViT-B/16', 'cpu')` (lavila/models/models.py:329 downloads the checkpoint; no
    network here): the reference's own OpenAI-CLIP module class with the ViT-B/16 hyper-parameters, randomly initialised."""
    from lavila.models.openai_model import CLIP as OpenAICLIP
    m = OpenAICLIP(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=16,
                   context_length=77, vocab_size=49408, transformer_width=512, transformer_heads=8, transformer_layers=12)
    return m, None


def build_reference_clip(num_frames=16, project_embed_dim=256):
    assert ref_shim.install(), "reference not installed: run python baseline/install_ref.py where /root/reference exists"
    from lavila.models import models as RM
    RM.load_openai_clip = _random_openai_clip_vit_b16
    with contextlib.redirect_stdout(sys.stderr):
        model = RM.CLIP_OPENAI_TIMESFORMER_BASE(num_frames=num_frames, project_embed_dim=project_embed_dim)
    return model, RM


def reference_step_fn(device, amp="bf16", num_frames=16, world_size=1, rank=0, seed=0):
    import torch
    torch.manual_seed(seed)
    model, RM = build_reference_clip(num_frames)
    # zero-init trap: the reference zero-initialises the time attention; randomise like our arm does so both
    # arms do the same arithmetic on non-degenerate values
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "timeattn" in n or "temporal_embed" in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    model.to(device)
    args = argparse.Namespace(model="CLIP_OPENAI_TIMESFORMER_BASE", contrastive_use_vissl=True, local_loss=False,
                              gather_with_grad=False, rank=rank, world_size=world_size, metadata_aux=None)
    criterion = RM.get_loss("CLIP_OPENAI_TIMESFORMER_BASE", args, tokenizer=None).to(device)
    # main_pretrain.py:199-213
    p_wd, p_non_wd = [], []
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (p_non_wd if (p.ndim < 2 or 'bias' in n or 'ln' in n or 'bn' in n) else p_wd).append(p)
    optimizer = torch.optim.AdamW([{"params": p_wd, "weight_decay": 0.01}, {"params": p_non_wd, "weight_decay": 0}],
                                  lr=3e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    on_gpu = torch.device(device).type == "cuda"
    use_scaler = on_gpu and amp == "fp16"
    scaler = torch.amp.GradScaler("cuda", enabled=use_scaler)          # main_pretrain.py:223
    model.train()

    def autocast():
        if not on_gpu or amp == "off":
            return contextlib.nullcontext()
        return torch.autocast("cuda", dtype=torch.bfloat16 if amp == "bf16" else torch.float16)

    def step(frames, tokens, use_checkpoint=False):
        """main_pretrain.py:486-530 with update_freq = 1 and no grad clipping (the documented recipes)."""
        with autocast():
            outputs = model(frames, tokens, use_checkpoint=use_checkpoint, norm_embed=True)
            loss_dict = criterion(outputs)
            loss = loss_dict["loss"]
        lv = loss.item()                                   # :503 (host sync, as in the reference)
        scaler.scale(loss).backward()
        scaler.step(optimizer)
        scaler.update()
        model.zero_grad(set_to_none=True)
        model.logit_scale.data.clamp_(0, 4.6052)
        return lv

    step.model = model
    return step


def host_threads():
    """Threads a CPU arm may use: scheduler affinity, capped by the cgroup CPU quota and by 32."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(per))))
    except Exception:
        pass
    return max(1, min(n, 32))
