"""Executed warp instructions and stall samples per SOURCE LINE of a kernel (needs -lineinfo + --import-source on).
Usage: python tools/sass_by_line.py rep.ncu-rep [top]"""
import csv, io, os, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
recs = []; fname = "?"; ix = None
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "File Path": fname = os.path.basename(r[1]); continue
    if r[0] == "Line No": ix = {h: i for i, h in reversed(list(enumerate(r)))}; continue
    if ix is None or len(r) < len(ix) or not r[0].isdigit(): continue
    def num(v):
        try: return int(v)
        except ValueError: return 0
    n = num(r[ix["Instructions Executed"]]); s = num(r[ix["# Samples"]])
    recs.append((n, s, fname, int(r[0]), r[1].strip()[:100]))
tn = sum(r[0] for r in recs); ts = sum(r[1] for r in recs)
print(f"total executed warp-instr {tn}  samples {ts}")
for n, s, f, ln, src in sorted(recs, reverse=True)[:top]:
    print(f"{100*n/max(tn,1):5.1f}% exec {100*s/max(ts,1):5.1f}% smp  {f}:{ln}  {src}")
