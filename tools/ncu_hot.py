"""Top stall locations (source page) of one kernel in an .ncu-rep: python tools/ncu_hot.py rep [regex] [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + rx] if rx else [])
out = subprocess.run(cmd, capture_output=True, text=True).stdout
secs, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": [], "hdr": None}; secs.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = row; continue
    cur["rows"].append(row)
s = secs[0]
ix = {n: i for i, n in enumerate(s["hdr"])}
tot = sum(int(r[ix["# Samples"]] or 0) for r in s["rows"])
print(s["name"][:100], "samples", tot)
order = sorted(range(len(s["rows"])), key=lambda i: -int(s["rows"][i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = s["rows"][i]
    print("%5d %5.1f%%  #%-5d %s" % (int(r[ix["# Samples"]]), 100.0 * int(r[ix["# Samples"]]) / max(tot, 1), i, r[ix["Source"]][:110]))
