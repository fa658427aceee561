"""Per-opcode executed-instruction mix and stall samples from `ncu -i rep --page source --csv --print-source sass`.
Usage: python tools/sass_mix.py rep.ncu-rep [top_n]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); samples = collections.Counter(); stalls = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
lines = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.rstrip(";")
    n = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    ops[op.split(".")[0]] += n; samples[op.split(".")[0]] += s
    for c in stall_cols:
        stalls[c] += int(r[ix[c]] or 0)
    lines.append((s, n, src, {c: int(r[ix[c]] or 0) for c in stall_cols if int(r[ix[c]] or 0) > 0}))
tot = sum(ops.values()); ts = sum(samples.values())
print(f"total warp instructions {tot}, samples {ts}")
for op, n in ops.most_common(topn):
    print(f"{op:12s} {n:10d} {100*n/tot:5.1f}%   samples {100*samples[op]/max(ts,1):5.1f}%")
print("stall totals:", {k: f"{100*v/max(ts,1):.1f}%" for k, v in stalls.most_common(10)})
print("hottest instructions:")
for s, n, src, st in sorted(lines, key=lambda t: -t[0])[:topn]:
    print(f"{100*s/max(ts,1):5.1f}% exec {n:8d}  {src[:70]:70s} {st}")
