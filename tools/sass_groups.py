"""Execution profile of a kernel's SASS grouped into runs of equal execution count (= loop nests): share of executed
warp instructions and of stall samples per run, with the opcode mix.  Usage: python tools/sass_groups.py rep.ncu-rep"""
import collections, csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
hdr = None; recs = []
for r in csv.reader(txt.splitlines()):
    if "Source" in r and "# Samples" in r:
        hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or len(r) < len(hdr): continue
    try:
        recs.append((r[hdr["Source"]].strip(), int(r[hdr["# Samples"]] or 0), int(r[hdr["Instructions Executed"]] or 0)))
    except ValueError:
        pass
ts = sum(o[1] for o in recs) or 1; tn = sum(o[2] for o in recs) or 1
print(f"{len(recs)} SASS instructions, {tn} warp instructions executed, {ts} samples")
i = 0
while i < len(recs):
    j = i
    while j + 1 < len(recs) and abs(recs[j + 1][2] - recs[i][2]) <= 0.15 * max(recs[i][2], 1): j += 1
    blk = recs[i:j + 1]; s = sum(x[1] for x in blk); n = sum(x[2] for x in blk)
    ops = collections.Counter((x[0].split()[1] if x[0].startswith("@") else x[0].split()[0]).split(".")[0] for x in blk if x[0])
    if n / tn > 0.01 or s / ts > 0.01:
        print(f"{i:4d}-{j:4d} x{recs[i][2]:>9d}  exec {100*n/tn:5.1f}%  samples {100*s/ts:5.1f}%  " + ",".join(f"{k}{v}" for k, v in ops.most_common(6)))
    i = j + 1
