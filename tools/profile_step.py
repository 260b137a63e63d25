"""Profiling harness: N warm-up training steps, then one step (or only its forward / backward) inside a
cudaProfilerStart/Stop range, for `ncu --profile-from-start off`.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv \
        python tools/profile_step.py --batch 64 --range step
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--range", default="step", choices=["step", "fwd", "bwd"])
    ap.add_argument("--model", default="base", choices=list(bench.MODELS))
    a = ap.parse_args()
    from lavila_b200.models import models as M
    from lavila_b200.models.loss import CLIPLoss
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    factory, img_size, _ = bench.MODELS[a.model]
    model = getattr(M, factory)(num_frames=a.frames, project_embed_dim=256)
    bench.randomise_zero_init(model)
    model.to(dev)
    crit = CLIPLoss(use_vissl=True, rank=0, world_size=1)
    opt = torch.optim.AdamW(bench.param_groups(model), lr=3e-5, weight_decay=0.01)
    fr, tx = bench.make_batch(a.batch, a.frames, 1234, img_size)
    fr, tx = fr.to(dev), tx.to(dev)
    rt = torch.cuda.cudart()

    def step(profile):
        opt.zero_grad(set_to_none=True)
        if profile in ("step", "fwd"):
            rt.cudaProfilerStart()
        out = model(fr, tx, norm_embed=True)
        ld = crit(out)
        if profile == "fwd":
            torch.cuda.synchronize()
            rt.cudaProfilerStop()
        if profile == "bwd":
            torch.cuda.synchronize()
            rt.cudaProfilerStart()
        ld["loss"].backward()
        if profile == "bwd":
            torch.cuda.synchronize()
            rt.cudaProfilerStop()
        opt.step()
        model.logit_scale.data.clamp_(0, 4.6052)
        if profile == "step":
            torch.cuda.synchronize()
            rt.cudaProfilerStop()
        return ld["loss"]

    for _ in range(a.warmup):
        step(None)
    torch.cuda.synchronize()
    print("loss", float(step(a.range)))


if __name__ == "__main__":
    main()
