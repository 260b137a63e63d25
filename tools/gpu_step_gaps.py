"""How much of a training step is GPU-idle between kernels?  torch.profiler (CUPTI) over 3 steps: sum of kernel durations vs span."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from lavila_b200.models import models as M
from lavila_b200.models.loss import CLIPLoss

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = M.CLIP_OPENAI_TIMESFORMER_BASE(num_frames=16, project_embed_dim=256)
bench.randomise_zero_init(model)
model.to(dev)
crit = CLIPLoss(use_vissl=True, rank=0, world_size=1)
opt = torch.optim.AdamW(bench.param_groups(model), lr=3e-5, weight_decay=0.01)
fr, tx = bench.make_batch(64, 16, 1234, 224)
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
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
busy = sum(e.time_range.end - e.time_range.start for e in ev)
gaps = []
last_end = ev[0].time_range.end
for e in ev[1:]:
    if e.time_range.start > last_end:
        gaps.append(e.time_range.start - last_end)
    last_end = max(last_end, e.time_range.end)
gaps.sort(reverse=True)
print("span %.2f ms for 3 steps, kernel time %.2f ms, idle %.2f ms (%.1f %%), %d kernels, %d gaps; largest gaps (us): %s" % (
    (t1 - t0) / 1e3, busy / 1e3, sum(gaps) / 1e3, 100 * sum(gaps) / (t1 - t0), len(ev), len(gaps), [round(g, 1) for g in gaps[:10]]))
small = [g for g in gaps if g < 20]
print("gaps < 20 us: %d, total %.2f ms, mean %.2f us" % (len(small), sum(small) / 1e3, sum(small) / max(1, len(small))))
