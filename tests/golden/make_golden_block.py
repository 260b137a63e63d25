"""ONE SpaceTimeBlock of the UNMODIFIED reference at the real TSF-B geometry (D = 768, 12 heads, 16 frames x 196 patches,
N = 3137 tokens, tanh-gated), forward + every gradient, on seeded inputs -> tests/golden/block_tsfb.pt.
Tensors are too large to commit (9.6 MB each): a fixed random sample of 4096 entries + the L2 norm per tensor is stored.
Parameters and inputs are regenerated from seeds by `block_params` / `block_inputs` (checksums stored).

    python tests/golden/make_golden_block.py
"""
import os
import sys
from functools import partial

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_shim  # noqa: E402

D, H, T, n = 768, 12, 16, 196
N = 1 + T * n
NAMES = ["norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias",
         "timeattn.qkv.weight", "timeattn.qkv.bias", "timeattn.proj.weight", "timeattn.proj.bias", "alpha_timeattn",
         "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias",
         "norm3.weight", "norm3.bias"]


def block_params(seed=31):
    g = torch.Generator().manual_seed(seed)
    shapes = {"qkv.weight": (3 * D, D), "qkv.bias": (3 * D,), "proj.weight": (D, D), "proj.bias": (D,), "fc1.weight": (4 * D, D),
              "fc1.bias": (4 * D,), "fc2.weight": (D, 4 * D), "fc2.bias": (D,)}
    p = {}
    for name in NAMES:
        leaf = name.split(".", 1)[1] if "." in name and not name.startswith("norm") else name
        if name.startswith("norm"):
            p[name] = (1.0 + 0.1 * torch.randn(D, generator=g)) if name.endswith("weight") else 0.05 * torch.randn(D, generator=g)
        elif name == "alpha_timeattn":
            p[name] = torch.tensor(0.5)
        else:
            p[name] = torch.randn(*shapes[leaf], generator=g) * 0.02
    return p


def block_inputs(seed=32):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(2, N, D, generator=g), torch.randn(2, N, D, generator=g)   # B = 2: at B = 1 the reference's in-place `q *= scale` (timesformer.py:113) hits a view error under torch 2.x


def sample(t, k=4096, seed=5):
    t = t.detach().float().flatten()
    idx = torch.randint(0, t.numel(), (min(k, t.numel()),), generator=torch.Generator().manual_seed(seed + t.numel()))
    return {"idx": idx, "sample": t[idx].clone(), "norm": t.norm().clone()}


def main():
    assert reference_shim.install(), "reference not found"
    from lavila.models.openai_model import QuickGELU
    from lavila.models.timesformer import SpaceTimeBlock
    torch.set_num_threads(8)
    blk = SpaceTimeBlock(D, H, qkv_bias=True, act_layer=QuickGELU, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
                         time_init="zeros", attention_style="frozen-in-time", is_tanh_gating=True)
    params = block_params()
    res = blk.load_state_dict(params, strict=True)
    x, dy = block_inputs()
    x.requires_grad_(True)
    y = blk(x, 'b (f n) d', '(b f) n d', 'b (f n) d', '(b n) f d', time_n=n, space_f=T)
    y.backward(dy)
    out = {"geometry": dict(D=D, H=H, T=T, n=n), "param_seed": 31, "input_seed": 32,
           "param_checksum": {k: float(v.double().sum()) for k, v in params.items()},
           "x_checksum": float(x.detach().double().sum()), "y": sample(y), "dx": sample(x.grad),
           "grads": {k: sample(v.grad) for k, v in blk.named_parameters()}}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "block_tsfb.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes;", res, "|y|", float(y.norm()))


if __name__ == "__main__":
    main()
