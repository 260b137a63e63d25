"""Summarise an .ncu-rep (read here, no GPU): one line per captured launch with the metrics the roofline needs."""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conf"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "st_long"),
    ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "st_bar"),
    ("smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "st_short"),
    ("smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "st_lg"),
    ("smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "st_math"),
    ("smsp__warp_issue_stalled_wait_per_warp_active.pct", "st_wait"),
    ("smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "st_mio"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(" | ".join(["kernel", "grid"] + [m[1] for m in METRICS]))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0][-48:]
        vals = []
        for m, _ in METRICS:
            if m in idx:
                v = r[idx[m]]
                u = units[idx[m]]
                try:
                    f = float(v.replace(",", ""))
                    if u in ("byte", "Kbyte", "Mbyte", "Gbyte"):
                        f *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
                        v = "%.1fMB" % (f / 1e6)
                    elif u in ("ns", "us", "ms", "s", "usecond", "msecond", "nsecond", "second"):
                        f *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}[u]
                        v = "%.1fus" % f
                    else:
                        v = "%.4g" % f
                except ValueError:
                    pass
                vals.append(v)
            else:
                vals.append("-")
        print(" | ".join([name, r[idx["Grid Size"]]] + vals))


if __name__ == "__main__":
    main(sys.argv[1])
