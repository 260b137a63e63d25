"""Fused CLS-query space attention (lv_space_attn_{fwd,bwd}_tc_cls) vs the separate passes: values and time."""
import sys
import torch
sys.path.insert(0, ".")
from lavila_b200 import ops

B, H, T, n = (int(x) for x in sys.argv[1:5]) if len(sys.argv) > 4 else (4, 12, 16, 196)
MODE = int(sys.argv[5]) if len(sys.argv) > 5 else 0       # 0 = space, 1 = time
D = 64 * H
N = 1 + T * n
M = B * N
torch.manual_seed(0)
dev = "cuda"
qkv = torch.randn(M, 3 * D, device=dev).bfloat16()
qkv[:, :D] *= 2.0
dout = torch.randn(M, D, device=dev).bfloat16()


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20))


def ref_fwd():
    out = torch.zeros(M, D, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(M, H, device=dev)
    ops.group_attn_fwd(qkv, out, lse, MODE, B, H, T=T, n=n)
    ops.cls_attn_fwd(qkv, out, lse, B, H, N)
    return out, lse


def fused_fwd():
    out = torch.zeros(M, D, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(M, H, device=dev)
    (ops.time_attn_fwd_cls if MODE else ops.space_attn_fwd_cls)(qkv, out, lse, B, H, T, n)
    return out, lse


def ref_bwd(o, l):
    dqkv = torch.full((M, 3 * D), 7.0, device=dev, dtype=torch.bfloat16)
    dcls = torch.zeros(B, H, 2, 64, device=dev)
    ops.group_attn_bwd(qkv, o, l, dout, dqkv, dcls, 0, MODE, B, H, T=T, n=n)
    ops.cls_attn_bwd(qkv, o, dout, l, dqkv, dcls, B, H, N, accumulate=True)
    ops.cls_kv_finalize(dcls, dqkv, B, H, N)
    return dqkv


def fused_bwd(o, l):
    dqkv = torch.full((M, 3 * D), 7.0, device=dev, dtype=torch.bfloat16)
    (ops.time_attn_bwd_cls if MODE else ops.space_attn_bwd_cls)(qkv, o, l, dout, dqkv, B, H, T, n)
    return dqkv


o_r, l_r = ref_fwd()
o_f, l_f = fused_fwd()
torch.cuda.synchronize()
cls = torch.arange(B, device=dev) * N
print("fwd out rel %.3e (cls rows %.3e)  lse max|diff| %.3e (cls rows %.3e)" % (
    rel(o_f, o_r), rel(o_f[cls], o_r[cls]), float((l_f - l_r).abs().max()), float((l_f[cls] - l_r[cls]).abs().max())))
g_r = ref_bwd(o_r, l_r)
g_f = fused_bwd(o_r, l_r)
torch.cuda.synchronize()
for nm, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
    print("bwd %s rel %.3e   cls-row rel %.3e" % (nm, rel(g_f[:, sl], g_r[:, sl]), rel(g_f[cls][:, sl], g_r[cls][:, sl])))
ok = rel(o_f, o_r) < 5e-3 and rel(o_f[cls], o_r[cls]) < 1e-2 and rel(g_f, g_r) < 2e-2 and rel(g_f[cls], g_r[cls]) < 3e-2
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("fwd separate", ref_fwd), ("fwd fused", fused_fwd), ("bwd separate", lambda: ref_bwd(o_r, l_r)), ("bwd fused", lambda: fused_bwd(o_r, l_r))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-13s %.3f ms" % (name, e0.elapsed_time(e1) / 5))
print("RESULT", "PASS" if ok else "FAIL")
