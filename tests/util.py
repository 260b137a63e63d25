import torch


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def cosine(a, b):
    a, b = a.detach().float().cpu().flatten(), b.detach().float().cpu().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


def assert_close_bf16(got, ref, name, rel=2e-2, cos=0.999):
    """Tolerance for a bf16-operand / fp32-accumulate kernel against the fp32 oracle:
    relative L2 error <= 2e-2 and cosine similarity >= 0.999."""
    r, c = rel_l2(got, ref), cosine(got, ref)
    assert torch.isfinite(got.detach().float()).all(), name + ": non-finite values"
    assert r <= rel and c >= cos, "%s: rel_l2=%.4e (tol %.1e) cosine=%.6f (tol %.4f)" % (name, r, rel, c, cos)
    return r, c
