"""Summarise `ncu --set full` reports (.ncu-rep) into one markdown table of the metrics the roofline needs.
Usage: python tools/summarize_ncu.py out.md rep1.ncu-rep [rep2 ...]   (runs `ncu -i ... --page raw --csv`)"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("smsp__inst_executed.sum", "warp instructions"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"), ("lts__t_bytes.sum", "L2 bytes"),
        ("sm__cycles_elapsed.avg.per_second", "sm clock"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %")]


def one(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {h: (v, u) for h, v, u in zip(hdr, r, units)}
        out.append(d)
    return out


def main(out, reps):
    with open(out, "w") as f:
        f.write("# ncu --set full summaries (`--clock-control none`, B200)\n\n")
        for rep in reps:
            for d in one(rep):
                f.write(f"## `{d['Kernel Name'][0][:150]}`\n\nsource report: `{rep.split('/')[-1]}` (scratch, not committed)\n\n| metric | value |\n|---|---|\n")
                for key, label in WANT:
                    if key in d and d[key][0] != "":
                        f.write(f"| {label} (`{key}`) | {d[key][0]} {d[key][1]} |\n")
                try:
                    tr = float(d["dram__bytes_read.sum"][0].replace(",", "")) + float(d["dram__bytes_write.sum"][0].replace(",", ""))
                    f.write(f"| **dram traffic (read+write)** | {tr:.3f} {d['dram__bytes_read.sum'][1]} |\n")
                except Exception:
                    pass
                f.write("\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
