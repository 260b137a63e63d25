"""In-situ kernel timeline of training steps (CUPTI through torch.profiler): per-kernel device time inside a real step
(warm L2, power-capped clocks) and the GPU idle time between kernels.  Complements the ncu launch list, whose
per-launch times are cold-cache and serialised.

    python tools/gpu_step_timeline.py [--batch 64] [--steps 3]
"""
import argparse
import collections
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    from lavila_b200.models import models as M
    from lavila_b200.models.loss import CLIPLoss
    from torch.profiler import profile, ProfilerActivity
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = M.CLIP_OPENAI_TIMESFORMER_BASE(num_frames=a.frames, project_embed_dim=256)
    bench.randomise_zero_init(model)
    model.to(dev)
    crit = CLIPLoss(use_vissl=True, rank=0, world_size=1)
    opt = torch.optim.AdamW(bench.param_groups(model), lr=3e-5, weight_decay=0.01)
    fr, tx = bench.make_batch(a.batch, a.frames, 1234)
    fr, tx = fr.to(dev), tx.to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        ld = crit(model(fr, tx, norm_embed=True))
        ld["loss"].backward()
        opt.step()
        model.logit_scale.data.clamp_(0, 4.6052)

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    spans = sorted((e.time_range.start, e.time_range.end, e.name) for e in evs)
    t0, t1 = spans[0][0], max(s[1] for s in spans)
    busy, cur_end = 0.0, t0
    for s, e, _ in spans:   # union of kernel intervals
        if e <= cur_end:
            continue
        busy += e - max(s, cur_end)
        cur_end = e
    wall = t1 - t0
    print("steps %d  wall %.2f ms/step  busy %.2f ms/step  idle %.2f ms/step (%.1f%%)  launches/step %d" % (
        a.steps, wall / a.steps / 1e3, busy / a.steps / 1e3, (wall - busy) / a.steps / 1e3, 100 * (wall - busy) / wall,
        len(spans) // a.steps))
    agg = collections.OrderedDict()
    for s, e, name in spans:
        short = re.sub(r"^void ", "", re.sub(r"\(.*", "", name))
        v = agg.setdefault(short, [0, 0.0])
        v[0] += 1
        v[1] += e - s
    print("%-86s %6s %10s %7s %9s" % ("kernel (in situ)", "count", "ms/step", "share", "avg us"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print("%-86s %6d %10.3f %6.1f%% %9.1f" % (k[:86], v[0] // a.steps, v[1] / a.steps / 1e3, 100 * v[1] / busy, v[1] / v[0]))


if __name__ == "__main__":
    main()
