#!/usr/bin/env python
"""Warp instructions and stall samples per CUDA source line of the first kernel of an .ncu-rep captured with
--import-source on (ncu --page source --print-source cuda,sass).  python tools/ncu_lines.py x.ncu-rep [top-N]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    want = sys.argv[3] if len(sys.argv) > 3 else None      # substring of the function name (default: the first kernel)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    fpath, fn, first_fn = None, None, None
    lines = {}
    for row in csv.reader(out.splitlines()):
        if not row:
            continue
        if row[0] == "File Path":
            fpath = row[1].split("/")[-1]
        elif row[0] == "Function Name":
            fn = row[1]
            if want is None or want in fn:
                first_fn = first_fn or fn
        elif row[0].isdigit() and fn == first_fn and len(row) > 7:
            key = (fpath, int(row[0]))
            e = lines.setdefault(key, [0, 0, row[1].strip()])
            e[0] += int(row[7]) if row[7].isdigit() else 0
            e[1] += int(row[4]) if row[4].isdigit() else 0
    tot_i = sum(e[0] for e in lines.values())
    tot_s = sum(e[1] for e in lines.values())
    print(first_fn, "instr", tot_i, "samples", tot_s)
    for (f, ln), e in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{f}:{ln:<5d} {100 * e[0] / tot_i:5.1f}% instr {100 * e[1] / max(tot_s, 1):5.1f}% stall  {e[2][:110]}")


if __name__ == "__main__":
    main()
