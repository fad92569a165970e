"""Compact per-kernel summary of an `ncu --set full` report: `ncu -i X.ncu-rep --page raw --csv | ncu_summary.py`."""
import csv
import sys

WANT = [
    ("time_us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("regs", "launch__registers_per_thread"),
    ("dram_rd", "dram__bytes_read.sum"),
    ("dram_wr", "dram__bytes_write.sum"),
    ("dram_%peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("fp64_pipe_%", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("dmma_pipe_%", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"),
    ("issue_%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warps_%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("L2_hit_%", "lts__t_sector_hit_rate.pct"),
]


def main():
    rows = list(csv.reader(sys.stdin))
    h, units = rows[0], rows[1]
    cols = [(n, h.index(m)) for n, m in WANT if m in h]
    kn = h.index("Kernel Name")
    print("| kernel | " + " | ".join(f"{n} [{units[i]}]" if units[i] else n for n, i in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in rows[2:]:
        name = r[kn].split("(")[0].replace("void ", "")
        vals = []
        for n, i in cols:
            try:
                v = float(r[i].replace(",", ""))
                vals.append(f"{v:.4g}")
            except ValueError:
                vals.append(r[i])
        print(f"| {name} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
