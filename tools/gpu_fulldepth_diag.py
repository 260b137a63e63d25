"""Per-parameter gradient agreement (norm ratio, rel-L2, cosine) of the full-depth TSF-B model vs the fp32 oracle on the GPU."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import dual_encoder as O
from tests.util import cosine, rel_l2
from tests.test_gpu_model import build_clip
from lavila_b200.models.loss import CLIPLoss

DEV = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 12
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
gated = (sys.argv[3] == "1") if len(sys.argv) > 3 else True
cfg = dict(O.tsf_base_config(num_frames=16), depth=depth)
params = O.init_params(cfg, seed=2, gated=gated)
frames, text = O.synthetic_batch(cfg, B, seed=4321)
model = build_clip(cfg, params, gated=gated)
out = model(frames.to(DEV), text.to(DEV), norm_embed=True)
ld = CLIPLoss()(out)
ld["loss"].backward()
pr = {k: v.to(DEV).clone().requires_grad_(True) for k, v in params.items()}
ref = O.clip_forward(frames.to(DEV), text.to(DEV), pr, cfg, norm_embed=True)
rl = O.clip_loss(ref["image_embed"], ref["text_embed"], ref["logit_scale"])
rl["loss"].backward()
print("loss %.5f vs %.5f; image rel %.3e text rel %.3e" % (float(ld["loss"]), float(rl["loss"]),
      rel_l2(out["image_embed"], ref["image_embed"]), rel_l2(out["text_embed"], ref["text_embed"])))
for name, p in model.named_parameters():
    g, gr = p.grad.float(), pr[name].grad.float()
    if float(gr.norm()) < 1e-12:
        continue
    print("%-46s ratio %.4f rel %.3e cos %.6f |ref| %.3e" % (name, float(g.norm() / gr.norm()), rel_l2(g, gr), cosine(g, gr), float(gr.norm())))
