"""Stall-reason breakdown per opcode (and per instruction-index range) from an ncu report's SASS source page.
Usage: python tools/sass_stalls_by_op.py rep.ncu-rep OPC[,OPC...] [lo hi]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; opcs = set(sys.argv[2].split(",")); lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0; hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); n = 0; ex = 0; allsamp = 0
for i, r in enumerate(rows[2:]):
    if len(r) < len(hdr): continue
    allsamp += int(r[ix["# Samples"]] or 0)
    if not (lo <= i < hi): continue
    src = r[ix["Source"]].strip().split()
    if not src: continue
    op = (src[1] if src[0].startswith("@") and len(src) > 1 else src[0]).split(".")[0].rstrip(";")
    if "*" not in opcs and op not in opcs: continue
    n += int(r[ix["# Samples"]] or 0); ex += int(r[ix["Instructions Executed"]] or 0)
    for c in stall_cols: tot[c] += int(r[ix[c]] or 0)
print(f"samples {n} ({100*n/allsamp:.1f}% of all), executed {ex}")
print({k: f"{100*v/max(n,1):.1f}%" for k, v in tot.most_common(8)})
