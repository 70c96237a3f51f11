#!/usr/bin/env python
"""Per-phase instruction / stall-sample split of one kernel from an .ncu-rep captured with --import-source on:
the SASS listing is cut at block barriers (BAR.SYNC), mbarrier waits and EXIT.
  python tools/ncu_phases.py gpurun_out/x.ncu-rep [kernel-index]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kernels, cur = [], None
    for row in rows:
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}
            kernels.append(cur)
        elif cur is not None and row and row[0].startswith("0x"):
            cur["rows"].append(row)
    k = kernels[which]
    r = k["rows"]
    tot_i = sum(int(x[5]) for x in r)
    tot_s = sum(int(x[2]) for x in r)
    warps = int(r[0][5])
    print(k["name"][:100])
    print(f"warp instructions {tot_i}  ({tot_i / warps:.0f} per warp, {warps} warps)  stall samples {tot_s}")
    seg_i = seg_s = 0
    start = 0
    for i, x in enumerate(r):
        seg_i += int(x[5])
        seg_s += int(x[2])
        if any(t in x[1] for t in ("BAR.SYNC", "SYNCS.PHASECHK", "EXIT")) or i == len(r) - 1:
            print(f"{start:5d}-{i:5d}  {seg_i / warps:8.1f} instr/warp ({100 * seg_i / tot_i:5.1f}%)  samples {100 * seg_s / max(tot_s, 1):5.1f}%   ends: {x[1].strip()[:48]}")
            seg_i = seg_s = 0
            start = i + 1


if __name__ == "__main__":
    main()
