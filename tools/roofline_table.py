"""Per-kernel roofline table from an `ncu --metrics gpu__time_duration.sum` launch list of ONE training step at TSF-B 16f x 224^2,
batch 64 (tools/profile_step.py): algorithmic bytes / FLOPs per launch divided by the measured
duration of the FULL-SIZE launches of each kernel (the longest launches of that name: the text tower and the CLS-only tail
reuse the same kernels on tiny shapes), against MEASURED_PEAKS.json.  The durations are ncu's cold-cache serialised ones
(SM clocks not power-capped for a single short kernel): fractions are upper-ish bounds of the in-step fractions.

    python tools/roofline_table.py profiles/launches_r01_final2_step_b64.csv > profiles/roofline_table_r01.md
"""
import collections
import csv
import json
import os
import re
import sys

B, T, n, D, H = 64, 16, 196, 768, 12
N = 1 + T * n
M = B * N
u = M * D * 2                     # one bf16 [M, D] tensor: 308 MB
GROUPS = B * T * H                # space groups
ATT_F = GROUPS * 2 * 2 * (n + 1) * n * 64    # QK^T + PV, MAC = 2


def model():
    """kernel-name regex -> (bound, algorithmic bytes, algorithmic flops, what)."""
    return [
        (r"space_attn_fwd_tc", ("tensor+SFU", 4 * u, ATT_F, "q,k,v read + o written; 2 GEMMs of 196 x 197 x 64 per group (round 2: + the fused CLS query row)")),
        (r"space_attn_bwd_tc", ("tensor+SFU", 8 * u, 2.5 * ATT_F, "q,k,v,o,do read + dq,dk,dv written; 5 GEMMs per group")),
        (r"time_attn_fwd", ("hbm", 4 * u, 0, "q,k,v read + o written (17 keys per group; round 2: + the fused CLS query partials)")),
        (r"time_attn_bwd", ("hbm", 8 * u, 0, "q,k,v,o,do read + dq,dk,dv written")),
        (r"cls_attn_fwd", ("hbm", 2 * u, 0, "k,v of every token read once")),
        (r"cls_attn_bwd", ("hbm", 6 * u, 0, "k,v read; dk,dv read-modify-written")),

        (r"ln_fwd_kernel", ("hbm", 3 * u, 0, "fp32 x read, bf16 y written")),
        (r"ln_bwd_kernel<6, 1, 2, 1>", ("hbm", 10 * u, 0, "bf16 dy + fp32 x + 2 fp32 adds read, fp32 + bf16 dx written")),
        (r"ln_bwd_kernel<6, 2, 1, 1>", ("hbm", 8 * u, 0, "bf16 dy + fp32 x + 1 fp32 add read, fp32 + bf16 dx written")),
        (r"ln_bwd_kernel<6, 2, 0, 1>", ("hbm", 6 * u, 0, "bf16 dy + fp32 x read, fp32 + bf16 dx written")),
    ]


def main(path):
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm, tf = float(peaks.get("hbm_gbs", 6581.6)), float(peaks.get("bf16_tflops_sustained", 1422.2))
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.defaultdict(list)
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r.get("Metric Unit", "ns"), 1e-3)
        per[re.sub(r"^void ", "", re.sub(r"\(.*", "", r["Kernel Name"]))].append(v)
    print("| kernel | full-size launches | avg us | bound | algorithmic traffic / work | achieved | fraction of measured peak |")
    print("|---|---|---|---|---|---|---|")
    for pat, (bound, nbytes, flops, what) in model():
        names = [k for k in per if re.search(pat, k)]
        if not names:
            continue
        ts = sorted(sum((per[k] for k in names), []), reverse=True)
        big = [t for t in ts if t > 0.5 * ts[0]]       # the full-size launches
        avg = sum(big) / len(big)
        cells = []
        if nbytes:
            gbs = nbytes / avg / 1e3
            cells.append("%.0f GB/s (%.2f of %.0f)" % (gbs, gbs / hbm, hbm))
        if flops:
            tfs = flops / avg / 1e6
            cells.append("%.0f TFLOP/s (%.2f of %.0f)" % (tfs, tfs / tf, tf))
        trf = ("%.2f GB" % (nbytes / 1e9) if nbytes else "-") + ((", %.0f GFLOP" % (flops / 1e9)) if flops else "") + " -- " + what
        print("| `%s` | %d | %.1f | %s | %s | %s | %s |" % (names[0].split("::")[-1] if len(names) == 1 else pat, len(big), avg, bound, trf,
                                                        cells[0].split(" (")[0] if cells else "-",
                                                        "; ".join(c.split("(")[1].rstrip(")") for c in cells) if cells else "-"))


if __name__ == "__main__":
    main(sys.argv[1])
