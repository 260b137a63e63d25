"""tcgen05 key-tiled attention forward vs the mma.sync kernels: timing at the narrator / TSF-L shapes (run twice: LAVILA_B200_FLASH_TC=0/1)."""
import sys
import torch
sys.path.insert(0, ".")
from lavila_b200 import ops

dev = "cuda"
torch.manual_seed(0)


def time_it(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# narrator shapes: CoCa pool (MQA 256 q x 1025 keys, 25 heads, 32 clips), cross-attention prefill (77 x 256, 25 heads, 320 sequences)
for name, B, H, Lq, Lk, mqa, causal in [("coca pool", 32, 25, 256, 1025, True, False), ("xattn prefill", 320, 25, 77, 256, False, False),
                                         ("self prefill", 320, 25, 77, 77, False, True)]:
    hk = 1 if mqa else H
    q = torch.randn(B * Lq, H * 64, device=dev).bfloat16()
    kv = torch.randn(B * Lk, 2 * hk * 64, device=dev).bfloat16()
    out = torch.zeros(B * Lq, H * 64, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.flash_attn_fwd(q, kv, kv[:, hk * 64:], out, B, H, Lq, Lk, q_rows=Lq, kv_rows=Lk, ld_q=H * 64, ld_kv=2 * hk * 64,
                                    ld_out=H * 64, kv_head_stride=0 if mqa else 64, causal=causal)
    ms = time_it(fn)
    fl = 4.0 * B * H * Lq * Lk * 64 * (0.5 if causal else 1.0)
    print("%-14s B%d H%d %dx%d: %.3f ms  %.1f TF/s" % (name, B, H, Lq, Lk, ms, fl / ms / 1e9))

# TSF-L space attention forward: 16 heads, n = 256 (224 px, 4 frames, 32 clips) and n = 576 (336 px, 32 frames, 4 clips)
for name, B, T, n in [("tsf-l 224 4f", 32, 4, 256), ("tsf-l 336 32f", 4, 32, 576)]:
    H, D = 16, 1024
    N = 1 + T * n
    qkv = torch.randn(B * N, 3 * D, device=dev).bfloat16()
    out = torch.zeros(B * N, D, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B * N, H, device=dev)
    fn = lambda: ops.group_attn_fwd(qkv, out, lse, 0, B, H, T=T, n=n)
    ms = time_it(fn)
    fl = 4.0 * B * T * H * n * (n + 1) * 64
    print("%-14s B%d T%d n%d: %.3f ms  %.1f TF/s" % (name, B, T, n, ms, fl / ms / 1e9))
