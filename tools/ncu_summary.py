#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): headline metrics + top stall lines.  Used to write profiles/*.md."""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "gpc__cycles_elapsed.max", "sm__cycles_active.max",
        "sm__cycles_active.avg", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "sm__inst_executed_pipe_lsu.sum"]


def main(path, top=14):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== kernel ==")
        for k in KEYS:
            if k in h:
                i = h.index(k)
                print(f"{k:72s} {r[i]} {units[i]}")
        st = [(float(r[i].replace(",", "")), hh) for i, hh in enumerate(h) if "pcsamp_warps_issue_stalled" in hh and "not_issued" not in hh and r[i]]
        tot = sum(v for v, _ in st) or 1
        print("stall samples:", ", ".join(f"{n.split('stalled_')[1]} {100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:7]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hi = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
    if not hi:
        return
    h = rows[hi[0]]
    si, ci = h.index("Source"), h.index("# Samples")
    data = []
    for r in rows[hi[0] + 1:]:
        try:
            data.append((float(r[ci]), r[si][:100]))
        except Exception:
            pass
    tot = sum(d[0] for d in data) or 1
    print(f"top SASS lines by samples (total {tot:.0f}):")
    for v, s in sorted(data, reverse=True)[:top]:
        print(f"  {100 * v / tot:5.1f}%  {s}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 14)
