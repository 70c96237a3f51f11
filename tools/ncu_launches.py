#!/usr/bin/env python
"""Turn the CSV log of
  ncu --metrics gpu__time_duration.sum,sm__cycles_active.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum,launch__registers_per_thread,launch__waves_per_multiprocessor,\\
sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,\\
dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum \\
      --clock-control none -s <skip> -c <launches per step> --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py ...
into the one-row-per-launch table committed under profiles/ (python tools/ncu_launches.py in.csv > out.csv)."""
import csv
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    h = rows[0]
    iid, ik, ig, im, iv = h.index("ID"), h.index("Kernel Name"), h.index("Grid Size"), h.index("Metric Name"), h.index("Metric Value")
    iu = h.index("Metric Unit")
    launches = OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(r[iid], {"kernel": r[ik], "grid": r[ig]})
        v = float(r[iv].replace(",", ""))
        unit = r[iu]
        if unit in ("nsecond", "ns"):
            v /= 1e3
        elif unit == "msecond":
            v *= 1e3
        if unit == "Mbyte":
            v *= 1e6
        elif unit == "Kbyte":
            v *= 1e3
        elif unit == "Gbyte":
            v *= 1e9
        d[r[im]] = v
    tot = sum(d["gpu__time_duration.sum"] for d in launches.values())
    w = csv.writer(sys.stdout)
    print(f"# one step ({len(launches)} kernels), ncu --clock-control none (cold cache, serialised); sum {tot:.1f} us")
    w.writerow(["idx", "kernel", "grid", "ncu_us", "share_pct", "warp_inst", "regs", "waves_per_sm", "warps_active_pct", "tensor_pipe_pct",
                "dram_rd_MB", "dram_wr_MB", "l2_MB", "smem_wavefronts", "sm_busy_us_x148"])
    for i, d in enumerate(launches.values()):
        g = lambda k: d.get(k, 0.0)
        w.writerow([i, d["kernel"].split("(")[0].replace("void rf::", "").replace("rf::", "")[:48], d["grid"].replace(" ", ""),
                    f"{g('gpu__time_duration.sum'):.2f}", f"{100 * g('gpu__time_duration.sum') / tot:.1f}", int(g("smsp__inst_executed.sum")),
                    int(g("launch__registers_per_thread")), f"{g('launch__waves_per_multiprocessor'):.2f}",
                    f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f}",
                    f"{g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.2f}", f"{g('dram__bytes_read.sum') / 1e6:.2f}",
                    f"{g('dram__bytes_write.sum') / 1e6:.2f}", f"{g('lts__t_bytes.sum') / 1e6:.2f}",
                    int(g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")),
                    # SM-time of the launch spread over the whole GPU: sum of per-SM active cycles / (148 * clock): what the
                    # launch costs when other work fills the SMs it leaves idle (the 4-context throughput mode)
                    f"{g('sm__cycles_active.sum') / 148.0 / max(g('sm__cycles_elapsed.max'), 1.0) * g('gpu__time_duration.sum'):.2f}"])
    w.writerow(["total", "", "", f"{tot:.2f}", "100", int(sum(d.get("smsp__inst_executed.sum", 0) for d in launches.values())), "", "", "", "", "", "", "", "",
                f"{sum(d.get('sm__cycles_active.sum', 0) / 148.0 / max(d.get('sm__cycles_elapsed.max', 1.0), 1.0) * d.get('gpu__time_duration.sum', 0) for d in launches.values()):.2f}"])


if __name__ == "__main__":
    main()
