import sys, torch
sys.path.insert(0, ".")
from lavila_b200 import ops, _lib as L
B, H, T, n = 16, 12, 16, 196
D = 64 * H; N = 1 + T * n; M = B * N
torch.manual_seed(0)
qkv = torch.randn(M, 3 * D, device="cuda").bfloat16()
out = torch.zeros(M, D, device="cuda", dtype=torch.bfloat16); lse = torch.zeros(M, H, device="cuda")
ops.group_attn_fwd(qkv, out, lse, 0, B, H, T=T, n=n); ops.cls_attn_fwd(qkv, out, lse, B, H, N)
dout = torch.randn(M, D, device="cuda").bfloat16()
dqkv = torch.zeros(M, 3 * D, device="cuda", dtype=torch.bfloat16); dcls = torch.zeros(B, H, 2, 64, device="cuda")
dbg = torch.zeros(512, device="cuda", dtype=torch.int64)
ops.group_attn_bwd(qkv, out, lse, dout, dqkv, dcls, 0, 0, B, H, T=T, n=n)
L.check(L.lib().lv_debug_set_buffer(dbg.data_ptr()))
ops.group_attn_bwd(qkv, out, lse, dout, dqkv, dcls, 0, 0, B, H, T=T, n=n)
torch.cuda.synchronize()
L.lib().lv_debug_set_buffer(None)
d = dbg.cpu().view(-1, 64)
for it in range(1, 3):
    r = d[it]; t0 = int(r[0])
    f = lambda i: (int(r[i]) - t0) if int(r[i]) else None
    print("group", it, "M: load_wait_done", f(1))
    for s in range(4):
        print("  step", s, "M: A-issued", f(2 + s * 4), "pds-ready", f(3 + s * 4), "B-issued", f(4 + s * 4),
              "| E: sdp-ready", f(21 + s * 4), "mma3-ready", f(22 + s * 4), "elem-done", f(23 + s * 4), "arrived", f(24 + s * 4))
    print("  E: prep-done", f(20), "final mma3-ready", f(40), "epilogue-done", f(41))
    print("  next group M start:", int(d[it + 1][0]) - t0)
