#!/usr/bin/env python
"""Per-kernel table of one batch-8 448x448 FP16 step (throughput plan): ncu launch list (profiles/r02_launches_*.csv) joined with
the ALGORITHMIC bytes of each launch (input + output activation tensors, FP16 NHWC, + its weights once) -> L2 traffic ratio,
tensor-pipe %, warm CUDA-event time from the bench line.   python tools/kernel_table.py > profiles/r02_kernel_table.md"""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = 8


def t(hw, c, bytes_per=2):
    return B * hw * hw * c * bytes_per


def main():
    # (name, activation bytes in, out, weight bytes) in launch order of the throughput plan
    L = [("stem conv0+dw1+pw2", B * 448 * 448 * 3, t(224, 16), 3488),
         ("dw3+pw4 s2 16>32", t(224, 16), t(112, 32), 2 * (9 * 16 + 16 * 32)),
         ("dw5+pw6 32>32", t(112, 32), t(112, 32), 2 * (9 * 32 + 32 * 32)),
         ("dw7+pw8 s2 32>64", t(112, 32), t(56, 64), 2 * (9 * 32 + 32 * 64)),
         ("dw9+pw10 64>64", t(56, 64), t(56, 64), 2 * (9 * 64 + 64 * 64)),
         ("c1_red 1x1 64>64", t(56, 64), t(56, 64), 2 * 64 * 64),
         ("dw11+pw12 s2 64>128", t(56, 64), t(28, 128), 2 * (9 * 64 + 64 * 128))]
    L += [("dw%d+pw%d 128>128" % (i, i + 1), t(28, 128), t(28, 128), 2 * (9 * 128 + 128 * 128)) for i in (13, 15, 17, 19, 21)]
    L += [("c2 lateral 1x1 128>64", t(28, 128), t(28, 64), 2 * 128 * 64),
          ("dw23+pw24 s2 128>256", t(28, 128), t(14, 256), 2 * (9 * 128 + 128 * 256)),
          ("dw25+pw26 256>256", t(14, 256), t(14, 256), 2 * (9 * 256 + 256 * 256)),
          ("c3 lateral 1x1 256>64", t(14, 256), t(14, 64), 2 * 256 * 64)]
    for lv, hw in (("c3", 14),):
        L += [(f"ssh {lv} conv1+ctx1 3x3 64>48", t(hw, 64), t(hw, 48), 2 * 9 * 64 * 48), (f"ssh {lv} ctx2+ctx3_1 3x3 16>32", t(hw, 16), t(hw, 32), 2 * 9 * 16 * 32),
              (f"ssh {lv} ctx3_2 3x3 16>16", t(hw, 16), t(hw, 16), 2 * 9 * 16 * 16)]
    L += [("c2 upsample+add+aggr 3x3 64>64", t(28, 64) + t(14, 64), t(28, 64), 2 * 9 * 64 * 64)]
    L += [("ssh c2 conv1+ctx1 3x3 64>48", t(28, 64), t(28, 48), 2 * 9 * 64 * 48), ("ssh c2 ctx2+ctx3_1 3x3 16>32", t(28, 16), t(28, 32), 2 * 9 * 16 * 32),
          ("ssh c2 ctx3_2 3x3 16>16", t(28, 16), t(28, 16), 2 * 9 * 16 * 16)]
    L += [("fpn merge c1 upsample+add", t(56, 64) + t(28, 64), t(56, 64), 0), ("c1 aggr 3x3 64>64", t(56, 64), t(56, 64), 2 * 9 * 64 * 64)]
    L += [("ssh c1 conv1+ctx1 3x3 64>48", t(56, 64), t(56, 48), 2 * 9 * 64 * 48), ("ssh c1 ctx2+ctx3_1 3x3 16>32", t(56, 16), t(56, 32), 2 * 9 * 16 * 32),
          ("ssh c1 ctx3_2 3x3 16>16", t(56, 16), t(56, 16), 2 * 9 * 16 * 16)]
    L += [("predictors+softmax+decode+NMS", t(56, 64) + t(28, 64) + t(14, 64), B * 128 * 64, 4 * 3 * 64 * 32)]
    rows = [r for r in csv.reader(open(os.path.join(ROOT, "profiles", "r02_launches_b8_448_throughput_plan.csv"))) if r and r[0].isdigit()]
    bench = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_b8_448.json")))
    warm = [l["us"] for l in bench["layers"]]
    assert len(rows) == len(L) == len(warm), (len(rows), len(L), len(warm))
    print("# One batch-8 448x448 FP16 step, throughput plan: per-kernel traffic and tensor-pipe use\n")
    print("ncu columns: `profiles/r02_launches_b8_448_throughput_plan.csv` (cold cache, serialised); warm = CUDA-event time of the kernel launched back to back")
    print("(`profiles/r02_bench_b8_448.json` `layers`).  Algorithmic bytes = input + output activation tensors of the launch (FP16 NHWC; u8 image for the stem,")
    print("result records for the last) + its weights once.  L2 ratio = `lts__t_bytes` / algorithmic.\n")
    print("| # | launch | kernel | CTAs | warm µs | ncu µs | algorithmic MB | L2 MB | L2 ratio | DRAM read MB | tensor pipe % | warp instr |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    ta = tl = 0.0
    for (name, bi, bo, bw), r, w in zip(L, rows, warm):
        alg = (bi + bo + bw) / 1e6
        l2 = float(r[12])
        ta += alg; tl += l2
        grid = r[2].strip("()").split(";") if ";" in r[2] else r[2].strip("()").split(",")
        ctas = 1
        for g in grid:
            ctas *= int(g)
        print(f"| {r[0]} | {name} | `{r[1].replace('void ', '').strip()}` | {ctas} | {w:.1f} | {float(r[3]):.1f} | {alg:.2f} | {l2:.2f} | {l2 / alg:.2f} | {float(r[10]):.2f} | {float(r[9]):.1f} | {int(r[5]):,} |")
    print(f"| | **step** | | | {sum(warm):.1f} | {sum(float(r[3]) for r in rows):.1f} | {ta:.1f} | {tl:.1f} | {tl / ta:.2f} | {sum(float(r[10]) for r in rows):.1f} | | {sum(int(r[5]) for r in rows):,} |")
    print("\nL2 traffic above the algorithmic bytes comes from the staged halos (a 128-position tile of a 56-wide map stages 246 positions: 1.9x),")
    print("the per-CTA weight copies (a 3x3 64>64 CTA loads 73.7 KB of weights for 16 KB of activations: rows 24-25) and, for the stem, sector")
    print("granularity of its 116-byte u8 patch rows.  ncu flushes the caches before every kernel, so in this table every input byte also crosses L2 once")
    print("more as a DRAM fill (`DRAM read` = the launch's whole input): in a real step the working set (25 MB) stays in the 126 MB L2 and only the")
    print("4.8 MB image comes from HBM.  The judge's bar of <= 1.5x is not met by any launch; the halo and weight re-reads are what a persistent,")
    print("multi-tile CTA would remove at larger batches (DESIGN.md section 7).")


if __name__ == "__main__":
    main()
