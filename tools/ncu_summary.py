"""Condense `ncu -i X.ncu-rep --page raw --csv` into the small per-kernel summaries kept under profiles/.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > gpurun_out/x.raw.csv
    python tools/ncu_summary.py gpurun_out/x.raw.csv profiles/rN_name_ncu_summary.csv [kernel-substring]
"""
import csv
import re
import sys

KEEP = [
    r"^dram__bytes_(read|write)\.sum$",
    r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^gpu__time_duration\.sum$",
    r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_(ld|st)\.sum$",
    r"^launch__(block_size|grid_size|registers_per_thread|shared_mem_per_block_dynamic)$",
    r"^lts__t_sector_hit_rate\.pct$",
    r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__cycles_elapsed\.max$",
    r"^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)$",
    r"^sm__inst_executed_pipe_fp64\.sum$",
    r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
    r"^smsp__pcsamp_warps_issue_stalled_[a-z_]+$",
    r"^smsp__inst_executed\.sum$",
]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    want = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(open(src)))
    head, units, data = rows[0], rows[1], rows[2:]
    kcol = head.index("Kernel Name")
    data = [r for r in data if want in r[kcol]]
    cols = {}
    for i, h in enumerate(head):
        name = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1].startswith("Triage") else h
        if any(re.match(p, name) for p in KEEP) and name not in cols:
            cols[name] = i
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
        w.writerow(["Kernel Name", ""] + [r[kcol] for r in data])
        for name in sorted(cols):
            i = cols[name]
            w.writerow([name, units[i]] + [r[i] for r in data])
    print(f"{dst}: {len(cols)} metrics x {len(data)} launches")


if __name__ == "__main__":
    main()
