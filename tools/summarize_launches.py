"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name, and every
launch of OUR kernels (dd::*).  Usage: python tools/summarize_launches.py launches.csv out.md"""
import collections
import csv
import sys


def main(path, out):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    for r in rd:
        if len(r) > iv and r[im] == "gpu__time_duration.sum":
            rows.append((int(r[iid]), r[ik], float(r[iv].replace(",", ""))))
    unit_ns = True
    tot = sum(t for _, _, t in rows)
    by = collections.defaultdict(lambda: [0, 0.0])
    for _, k, t in rows:
        name = k.split("(")[0]
        by[name][0] += 1
        by[name][1] += t
    OURS = ("dd::", "kmeans_", "rownorm_", "class_mean", "partial_reduce", "energy_", "class_sort", "cfg_ddim", "affine_", "add_noise",
            "bicubic_", "image_to_uint8", "agglo_", "normalize_rows")   # ncu prints some names without the namespace
    ours = [(i, k, t) for i, k, t in rows if any(o in k for o in OURS)]
    t_ours = sum(t for _, _, t in ours)
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path})\n\n")
        f.write(f"launches: {len(rows)}, total gpu__time_duration: {tot/1e6:.3f} ms (cold-cache, serialised)\n\n")
        f.write(f"OUR kernels (dd::*): {len(ours)} launches, {t_ours/1e3:.1f} us = {100*t_ours/tot:.3f} % of the step\n\n")
        f.write("## time share by kernel (top 25)\n\n| kernel | launches | total us | share % | avg us |\n|---|---:|---:|---:|---:|\n")
        for name, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:25]:
            f.write(f"| `{name[:90]}` | {n} | {t/1e3:.1f} | {100*t/tot:.2f} | {t/1e3/n:.2f} |\n")
        f.write("\n## our kernels, per name\n\n| kernel | launches | total us | share % | avg us |\n|---|---:|---:|---:|---:|\n")
        for name, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
            if "dd::" in name:
                f.write(f"| `{name[:110]}` | {n} | {t/1e3:.1f} | {100*t/tot:.4f} | {t/1e3/n:.2f} |\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
