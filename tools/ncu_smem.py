#!/usr/bin/env python
"""Shared-memory wavefronts vs the conflict-free ideal per SASS instruction of every kernel in an .ncu-rep
(ncu --set full --import-source on).  python tools/ncu_smem.py x.ncu-rep [top-N]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    name, rows, hdr = None, [], None

    def flush():
        if name is None or not rows:
            return
        iw, ii = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
        items = []
        for x in rows:
            try:
                w, i = int(x[iw]), int(x[ii])
            except (ValueError, IndexError):
                continue
            if w:
                items.append((w, i, x[1].strip()[:64]))
        tw, ti = sum(a for a, _, _ in items), sum(b for _, b, _ in items)
        print(f"{name[:90]}\n  shared wavefronts {tw}  ideal {ti}  excess {100 * (tw - ti) / max(tw, 1):.0f}%")
        for a, b, c in sorted(items, key=lambda t: t[1] - t[0])[:top]:
            if a > b:
                print(f"    {a:9d} vs {b:9d}  {c}")

    for row in csv.reader(out.splitlines()):
        if not row:
            continue
        if row[0] == "Kernel Name":
            flush()
            name, rows = row[1], []
        elif row[0] == "Address":
            hdr = row
        elif row[0].startswith("0x"):
            rows.append(row)
    flush()


if __name__ == "__main__":
    main()
