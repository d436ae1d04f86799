#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` output into the few columns the profiles/ summaries quote.

    python tools/ncu_summary.py gpurun_out/r02_cfg3_kernels.raw.csv [out.csv]

Prints a markdown table (one row per launch; launches of the same kernel are kept apart so that cold first
launches can be told from the rest) and, if given, writes the same columns as CSV.
"""
import csv
import sys

COLS = [
    ("Kernel Name", "kernel"),
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_dim_x", "cluster"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__inst_executed.sum", "inst"),
]


def load(path):
    rows = list(csv.reader(open(path, newline="")))
    # skip ncu's "==PROF==" chatter if it was captured into the file
    while rows and (not rows[0] or "Kernel Name" not in rows[0]):
        rows.pop(0)
    header, units, data = rows[0], rows[1], rows[2:]
    idx = {name: i for i, name in enumerate(header)}
    out = []
    for r in data:
        if len(r) < len(header):
            continue
        row = {}
        for name, short in COLS:
            if name in idx:
                u = units[idx[name]]
                row[short] = (r[idx[name]], u)
        out.append(row)
    return out


def fmt(v, u):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v if len(v) < 70 else v[:67] + "..."
    if u in ("byte", "Kbyte", "Mbyte", "Gbyte"):
        x *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        return f"{x / 1e6:.2f} MB"
    if u in ("ns", "us", "ms", "s", "usecond", "nsecond", "msecond", "second"):
        x *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}[u]
        return f"{x:.1f} us"
    if u == "%":
        return f"{x:.1f}"
    return f"{x:.0f}" if x == int(x) else f"{x:.2f}"


def main():
    rows = load(sys.argv[1])
    if not rows:
        print("no launches in", sys.argv[1])
        return
    cols = [short for _, short in COLS if short in rows[0]]
    print("| " + " | ".join(cols) + " |")
    print("|" + "---|" * len(cols))
    for r in rows:
        print("| " + " | ".join(fmt(*r[c]) for c in cols) + " |")
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(cols)
            for r in rows:
                w.writerow([fmt(*r[c]) for c in cols])


if __name__ == "__main__":
    main()
