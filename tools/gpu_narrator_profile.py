"""Per-kernel time table of the narrator's decoding (torch.profiler / CUPTI), R sequences per clip, eager KV-cached loop."""
import os
import sys
from types import SimpleNamespace
import torch
os.environ["LAVILA_B200_DECODE_GRAPH"] = "0"
sys.path.insert(0, ".")
from lavila_b200.models import models as M

R = int(sys.argv[1]) if len(sys.argv) > 1 else 10
L = int(sys.argv[2]) if len(sys.argv) > 2 else 24
dev = torch.device("cuda", 0)
torch.manual_seed(0)
tok = SimpleNamespace(bos_token_id=50256, eos_token_id=50256, pad_token_id=0)
model = M.VCLM_OPENAI_TIMESFORMER_LARGE_GPT2_XL(gated_xattn=True, num_frames=4).to(dev).eval()
with torch.no_grad():
    for n, p in model.named_parameters():
        if "alpha" in n:
            p.fill_(0.5)
clips = torch.randn(32, 3, 4, 224, 224, device=dev)
t = model.encode_image(clips)
model.generate(t, tok, max_text_length=8, top_p=0.95, temperature=0.7, num_return_sequences=R, early_stopping=False)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    model.generate(t, tok, max_text_length=L, top_p=0.95, temperature=0.7, num_return_sequences=R, early_stopping=False)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    if e.device_type == torch.autograd.DeviceType.CUDA or getattr(e, "device_time_total", 0) > 0:
        rows.append((e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows if not r[2].startswith("aten::") and not r[2].startswith("cuda"))
print("R=%d, %d decoding steps; device time of kernels %.1f ms -> %.2f ms/step" % (R, L - 1, tot / 1e3, tot / 1e3 / (L - 1)))
for us, cnt, key in rows[:45]:
    print("%9.1f us %6d  %7.2f us/step  %s" % (us, cnt, us / (L - 1), key[:110]))
