"""Per-source-line executed-instruction and stall-sample shares of a kernel from an `ncu --import-source on` report.
Usage: python tools/src_lines.py rep.ncu-rep [min_pct]   (joins the SASS view back to CUDA lines through -lineinfo)"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; minp = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; agg = collections.OrderedDict(); cur = None
for r in rows:
    if "Source" in r and "# Samples" in r:
        hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or len(r) < len(hdr): continue
    src = r[hdr["Source"]]
    try:
        s = int(r[hdr["# Samples"]] or 0); n = int(r[hdr["Instructions Executed"]] or 0)
    except ValueError:
        continue
    key = r[hdr.get("Line No", 0)] if "Line No" in hdr else src
    agg[(key, src.strip()[:110])] = (s, n)
ts = sum(v[0] for v in agg.values()) or 1; tn = sum(v[1] for v in agg.values()) or 1
for (k, src), (s, n) in agg.items():
    if 100 * s / ts >= minp or 100 * n / tn >= minp:
        print(f"{k:>5} samp {100*s/ts:5.1f}% exec {100*n/tn:5.1f}%  {src}")
