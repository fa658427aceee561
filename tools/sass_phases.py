"""Cumulative stall-sample profile along the SASS of a kernel: prints one line per run of ~N instructions with its share
of samples and executed instructions, plus the first/last opcode, to attribute time to code regions.
Usage: python tools/sass_phases.py rep.ncu-rep [chunk]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
recs = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    recs.append((r[ix["Source"]].strip(), int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0)))
ts = sum(s for _, s, _ in recs); tn = sum(n for _, _, n in recs)
for i in range(0, len(recs), chunk):
    blk = recs[i:i + chunk]
    s = sum(x[1] for x in blk); n = sum(x[2] for x in blk)
    ops = collections.Counter((x[0].split()[1] if x[0].startswith("@") else x[0].split()[0]).split(".")[0] for x in blk if x[0])
    top = ",".join(f"{k}{v}" for k, v in ops.most_common(4))
    if s or n:
        print(f"{i:5d} samples {100*s/ts:5.1f}%  exec {100*n/tn:5.1f}%  {top}")
