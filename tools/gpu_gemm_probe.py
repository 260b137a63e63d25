"""GPU probe for lv_gemm_bf16: one variant per process (a device trap must not poison the others).

    python tools/gpu_gemm_probe.py <variant> [M N K]
variants: nt, nn (B MN-major), tn (A,B MN-major), epi (all epilogue flags), splitk, perf
"""
import ctypes
import sys
import time

import torch

sys.path.insert(0, ".")
from lavila_b200 import _lib as L  # noqa: E402


import os
TWO_CTA = os.environ.get("LV_2CTA", "0") == "1"


def gemm(A, a_mn, B, b_mn, M, N, K, flags=0, out=None, out2=None, bias=None, resid=None, aux=None, scale=None,
         k_splits=1):
    e = L.LvGemmEpilogue()
    e.flags = flags
    e.out, e.ldo = out.data_ptr(), out.stride(0)
    if out2 is not None:
        e.out2, e.ldo2 = out2.data_ptr(), out2.stride(0)
    if bias is not None:
        e.bias = bias.data_ptr()
    if resid is not None:
        e.resid, e.ldr = resid.data_ptr(), resid.stride(0)
    if aux is not None:
        e.aux, e.ldaux = aux.data_ptr(), aux.stride(0)
    if scale is not None:
        e.scale_ptr = scale.data_ptr()
    st = torch.cuda.current_stream().cuda_stream
    fn = L.lib().lv_gemm_bf16_2cta if TWO_CTA else L.lib().lv_gemm_bf16
    rc = fn(A.data_ptr(), A.stride(0), a_mn, B.data_ptr(), B.stride(0), b_mn, M, N, K, k_splits, ctypes.byref(e), st)
    L.check(rc, "lv_gemm_bf16")


def report(name, got, ref):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs().max().item()
    rel = ((got - ref).norm() / ref.norm().clamp_min(1e-12)).item()
    print("%-28s max_abs_err %.4e  rel_l2 %.4e  ref_absmax %.3f  %s" % (name, err, rel, ref.abs().max().item(),
                                                                        "OK" if rel < 2e-2 else "MISMATCH"))
    return rel < 2e-2


def main():
    variant = sys.argv[1]
    M, N, K = (int(x) for x in sys.argv[2:5]) if len(sys.argv) >= 5 else (1000, 768, 768)
    torch.manual_seed(0)
    dev = "cuda"
    ok = True
    if variant == "nt":
        A = torch.randn(M, K, device=dev).bfloat16()
        B = torch.randn(N, K, device=dev).bfloat16()
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        gemm(A, 0, B, 0, M, N, K, 0, out)
        torch.cuda.synchronize()
        ok &= report("nt bf16 %dx%dx%d" % (M, N, K), out, A.float() @ B.float().t())
        out32 = torch.empty(M, N, device=dev, dtype=torch.float32)
        gemm(A, 0, B, 0, M, N, K, L.EPI_OUT_F32, out32)
        torch.cuda.synchronize()
        ok &= report("nt f32", out32, A.float() @ B.float().t())
    elif variant == "nn":
        A = torch.randn(M, K, device=dev).bfloat16()
        Bt = torch.randn(K, N, device=dev).bfloat16()  # stored [K][N]
        out = torch.empty(M, N, device=dev, dtype=torch.float32)
        gemm(A, 0, Bt, 1, M, N, K, L.EPI_OUT_F32, out)
        torch.cuda.synchronize()
        ok &= report("nn (B MN-major)", out, A.float() @ Bt.float())
    elif variant == "tn":
        At = torch.randn(K, M, device=dev).bfloat16()  # stored [K][M]
        Bt = torch.randn(K, N, device=dev).bfloat16()
        out = torch.empty(M, N, device=dev, dtype=torch.float32)
        gemm(At, 1, Bt, 1, M, N, K, L.EPI_OUT_F32, out)
        torch.cuda.synchronize()
        ok &= report("tn (A,B MN-major)", out, At.float().t() @ Bt.float())
    elif variant == "tk":
        At = torch.randn(K, M, device=dev).bfloat16()
        B = torch.randn(N, K, device=dev).bfloat16()
        out = torch.empty(M, N, device=dev, dtype=torch.float32)
        gemm(At, 1, B, 0, M, N, K, L.EPI_OUT_F32, out)
        torch.cuda.synchronize()
        ok &= report("tk (A MN-major)", out, At.float().t() @ B.float().t())
    elif variant == "splitk":
        At = torch.randn(K, M, device=dev).bfloat16() * 0.1
        Bt = torch.randn(K, N, device=dev).bfloat16() * 0.1
        out = torch.ones(M, N, device=dev, dtype=torch.float32)
        gemm(At, 1, Bt, 1, M, N, K, L.EPI_ATOMIC, out, k_splits=7)
        torch.cuda.synchronize()
        ok &= report("splitk atomic", out, 1 + At.float().t() @ Bt.float())
    elif variant == "epi":
        A = torch.randn(M, K, device=dev).bfloat16() * 0.2
        B = torch.randn(N, K, device=dev).bfloat16() * 0.2
        bias = torch.randn(N, device=dev)
        resid = torch.randn(M, N, device=dev)
        aux = torch.randn(M, N, device=dev).bfloat16()
        scale = torch.tensor(0.7, device=dev)
        base = A.float() @ B.float().t() + bias
        # quickgelu
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        gemm(A, 0, B, 0, M, N, K, L.EPI_BIAS | L.EPI_QUICKGELU, out, out2=pre, bias=bias)
        torch.cuda.synchronize()
        ok &= report("bias+quickgelu pre", pre, base)
        hb = base.bfloat16().float()
        ok &= report("bias+quickgelu act", out, hb * torch.sigmoid(1.702 * hb))
        # dquickgelu
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        gemm(A, 0, B, 0, M, N, K, L.EPI_DQUICKGELU, out, aux=aux)
        torch.cuda.synchronize()
        h = aux.float()
        s = torch.sigmoid(1.702 * h)
        ok &= report("dquickgelu", out, (A.float() @ B.float().t()) * (s * (1 + 1.702 * h * (1 - s))))
        # scale(tanh) + resid, f32 out + bf16 copy
        out = torch.empty(M, N, device=dev, dtype=torch.float32)
        cp = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        gemm(A, 0, B, 0, M, N, K, L.EPI_BIAS | L.EPI_SCALE | L.EPI_SCALE_TANH | L.EPI_RESID | L.EPI_OUT_F32 | L.EPI_COPY_BF16,
             out, out2=cp, bias=bias, resid=resid, scale=scale)
        torch.cuda.synchronize()
        ref = base * torch.tanh(scale) + resid
        ok &= report("bias+tanh-scale+resid f32", out, ref)
        ok &= report("  bf16 copy", cp, ref)
    elif variant == "perf":
        F = L
        cases = [
            (200768, 2304, 768, 0, 0, 1, F.EPI_BIAS, "qkv fwd +bias"),
            (200768, 768, 768, 0, 0, 1, F.EPI_BIAS | F.EPI_RESID | F.EPI_OUT_F32, "proj fwd +bias+resid f32"),
            (200768, 3072, 768, 0, 0, 1, F.EPI_BIAS | F.EPI_QUICKGELU, "fc1 fwd +bias+gelu(2 outs)"),
            (200768, 768, 3072, 0, 0, 1, F.EPI_BIAS | F.EPI_RESID | F.EPI_OUT_F32, "fc2 fwd +bias+resid f32"),
            (200768, 768, 2304, 0, 1, 1, 0, "qkv dgrad"),
            (200768, 3072, 768, 0, 1, 1, F.EPI_DQUICKGELU, "fc2 dgrad *dgelu"),
            (200768, 768, 3072, 0, 1, 1, 0, "fc1 dgrad"),
            (2304, 768, 200768, 1, 1, 8, F.EPI_ATOMIC | F.EPI_OUT_F32, "qkv wgrad"),
            (3072, 768, 200768, 1, 1, 8, F.EPI_ATOMIC | F.EPI_OUT_F32, "fc1 wgrad"),
            (768, 3072, 200768, 1, 1, 8, F.EPI_ATOMIC | F.EPI_OUT_F32, "fc2 wgrad"),
            (768, 768, 200768, 1, 1, 16, F.EPI_ATOMIC | F.EPI_OUT_F32, "proj wgrad"),
        ]
        for (m, n, k, am, bm, ks, fl, nm) in cases:
            A = (torch.randn((k, m) if am else (m, k), device=dev) * 0.1).bfloat16()
            B = (torch.randn((k, n) if bm else (n, k), device=dev) * 0.1).bfloat16()
            f32 = bool(fl & F.EPI_OUT_F32)
            out = torch.zeros(m, n, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
            kw = {}
            if fl & F.EPI_BIAS:
                kw["bias"] = torch.randn(n, device=dev)
            if fl & F.EPI_RESID:
                kw["resid"] = torch.randn(m, n, device=dev)
            if fl & F.EPI_QUICKGELU:
                kw["out2"] = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
            if fl & F.EPI_DQUICKGELU:
                kw["aux"] = torch.randn(m, n, device=dev).bfloat16()
            for _ in range(2):
                gemm(A, am, B, bm, m, n, k, fl, out, k_splits=ks, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                gemm(A, am, B, bm, m, n, k, fl, out, k_splits=ks, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            At = A.t() if am else A
            Bt = B if bm else B.t()
            for _ in range(2):
                At @ Bt
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                At @ Bt
            e1.record()
            torch.cuda.synchronize()
            ms_t = e0.elapsed_time(e1) / 5
            print("%-28s %7dx%5dx%7d  lv %.3f ms  %6.1f TF/s | torch(plain mm) %.3f ms %6.1f TF/s" % (
                nm, m, n, k, ms, 2.0 * m * n * k / ms / 1e9, ms_t, 2.0 * m * n * k / ms_t / 1e9))
    print("RESULT", variant, "PASS" if ok else "FAIL")


if __name__ == "__main__":
    main()
