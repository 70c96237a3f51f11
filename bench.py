#!/usr/bin/env python
"""bench.py -- faces/sec of the RetinaFace mnet25 detect path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one pass of the hot path (u8 images -> conv0..SSH -> fused heads+decode -> NMS)
over one batch of synthetic S-real input (the golden photo letter-boxed to the
network size, element i rolled by 8*i pixels so every image has faces but distinct content).

Printed JSON (rank 0, one line):
  value      faces/s, device-timed: inputs already resident in HBM (a ring of batches larger than
             2x L2 so that no step finds its input in L2), CUDA events on the library's stream,
             max over ranks, whole job (all GPUs).
  e2e        the same metric through the public C ABI with HOST (pinned) images: every step copies its own
             images H2D and reads its own faces back D2H inside the timed region -- rf_submit_batch /
             rf_collect_batch with several batches in flight (throughput mode); `e2e.blocking` is the same
             with one blocking rf_detect_batch per step (latency mode).
  roofline   dominant kernel: algorithmic bytes (layer-granular) / CUDA-event time of
             that kernel launched K times on the library's stream, vs MEASURED_PEAKS.json.
  cpu_baseline  the oracle (cv2.dnn FP32 forward of the same caffemodel through a generated prototxt +
             oracle/postproc.c) timed on the host cores on a bounded sample (rank 0, N=1 only).

--impl reference runs only that CPU arm (the reference's own CPU path cannot be built here:
BVLC Caffe / OpenCV C++ / TensorRT are absent -- DESIGN.md), on the same config/metric/unit.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "mnet25_fp16_b8_448": dict(model="mnet25", precision="fp16", batch=8, h=448, w=448),
    "mnet25_fp32_b8_448": dict(model="mnet25", precision="fp32", batch=8, h=448, w=448),
    "mnet25_fp16_b1_448": dict(model="mnet25", precision="fp16", batch=1, h=448, w=448),
    "mnet25_fp16_b32_448": dict(model="mnet25", precision="fp16", batch=32, h=448, w=448),
    # configs[2]: INT8 with the reference's TensorRT calibration table
    "mnet0517_int8_b32_448": dict(model="mnet-deconv-0517", precision="int8", batch=32, h=448, w=448),
    "mnet0517_int8_b8_448": dict(model="mnet-deconv-0517", precision="int8", batch=8, h=448, w=448),
    "mnet0517_fp16_b32_448": dict(model="mnet-deconv-0517", precision="fp16", batch=32, h=448, w=448),
    # configs[3]: large input / many-anchor NMS stress
    "mnet25_fp16_b8_1280x896": dict(model="mnet25", precision="fp16", batch=8, h=896, w=1280),
}
DEFAULT_WORKLOAD = "mnet25_fp16_b8_448"
SCORE_THR, NMS_THR = 0.9, 0.4  # main.cpp:43, RetinaFace.h:66


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], src="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, src="fallback")


def make_batches(wl, count, rank):
    """`count` distinct S-real batches (u8 [count][B][H][W][3])."""
    import cv2
    from oracle.inputs import letterbox_bgr_u8
    img = cv2.imread(os.path.join(GOLD, "data", "img.jpg"))
    base = letterbox_bgr_u8(img, wl["h"], wl["w"])
    out = np.empty((count, wl["batch"], wl["h"], wl["w"], 3), dtype=np.uint8)
    for s in range(count):
        for i in range(wl["batch"]):
            out[s, i] = np.roll(base, 8 * (i + wl["batch"] * (s + count * rank)), axis=1)
    return out


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed region: NVML every 2 ms when pynvml can open the device (the
    timed region of the default run is ~30 ms), else one `nvidia-smi` query per 100 ms.  Rows have nvidia-smi's layout."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_ev = gpu, [], threading.Event()
        self.source = "nvidia-smi"

    def _nvml_loop(self) -> bool:
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = int(vis.split(",")[self.gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.gpu
            dev = nv.nvmlDeviceGetHandleByIndex(idx)
            mx = nv.nvmlDeviceGetMaxClockInfo(dev, nv.NVML_CLOCK_SM)
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = [nv.nvmlClocksThrottleReasonHwSlowdown, nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    nv.nvmlClocksThrottleReasonSwThermalSlowdown, nv.nvmlClocksThrottleReasonSwPowerCap]
            nv.nvmlDeviceGetClockInfo(dev, nv.NVML_CLOCK_SM)          # probe once before committing to this source
        except Exception:
            return False
        self.source = "nvml"
        while not self.stop_ev.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(dev, nv.NVML_CLOCK_SM)
                r = int(get_reasons(dev))
                self.rows.append([str(sm), str(mx)] + ["Active" if r & b else "Not Active" for b in bits])
            except Exception:
                pass
            self.stop_ev.wait(0.002)
        return True

    def run(self):
        if self._nvml_loop():
            return
        while not self.stop_ev.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self.stop_ev.wait(0.1)

    def summary(self):
        self.stop_ev.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(self.rows), source=self.source)


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle; the reference's own CPU-Caffe path is unbuildable here)
# ------------------------------------------------------------------------------------------------
class CpuPath:
    def __init__(self, wl, threads):
        import cv2
        import tempfile
        from oracle import topology
        from oracle.postproc import PostprocOracle
        cv2.setNumThreads(threads)
        self.cv2, self.topology, self.wl = cv2, topology, wl
        d = tempfile.mkdtemp()
        p = os.path.join(d, "oracle.prototxt")
        open(p, "w").write(topology.to_prototxt(wl["h"], wl["w"], wl["batch"]))
        self.net = cv2.dnn.readNetFromCaffe(p, os.path.join(GOLD, "weights", wl["model"] + ".caffemodel"))
        self.post = PostprocOracle()
        self.threads = threads

    def step(self, batch_u8):
        """RetinaFace::detect on the CPU: u8->f32 RGB planar, forward, decode+NMS.  Returns #faces."""
        x = np.ascontiguousarray(batch_u8[..., ::-1].transpose(0, 3, 1, 2), dtype=np.float32)
        self.net.setInput(x)
        outs = self.net.forward(self.topology.OUTPUT_BLOBS)
        faces = 0
        for i in range(batch_u8.shape[0]):
            r = self.post.postprocess([o[i] for o in outs], self.wl["h"], self.wl["w"], SCORE_THR, NMS_THR)
            faces += len(r["faces"])
        return faces


def best_cpu_threads(wl, batch):
    """cv2.dnn does not scale to every core of a big host: time one batch at a few thread counts and keep the
    fastest, so the CPU arm runs 'with all the host threads it can use' rather than with all that exist."""
    ncpu = os.cpu_count() or 1
    cands = sorted({t for t in (ncpu, ncpu // 2, 64, 32, 16, 8) if 1 <= t <= ncpu}, reverse=True)
    best, best_t = cands[0], float("inf")
    for t in cands:
        cpu = CpuPath(wl, t)
        cpu.step(batch)
        t0 = time.perf_counter()
        cpu.step(batch)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = t, dt
    return best


def cpu_measure(wl, batches, budget_s, min_steps=2):
    threads = best_cpu_threads(wl, batches[0])
    cpu = CpuPath(wl, threads)
    cpu.step(batches[0])  # warm-up
    t0 = time.perf_counter()
    faces = steps = 0
    while (time.perf_counter() - t0 < budget_s or steps < min_steps) and steps < 10_000:
        faces += cpu.step(batches[steps % len(batches)])
        steps += 1
    dt = time.perf_counter() - t0
    return dict(value=faces / dt, unit="faces/s", cores=threads, kind="port",
                sample=f"{steps} batches of {wl['batch']} images ({steps * wl['batch']} images, {dt:.1f} s): cv2.dnn FP32 forward "
                       f"of {wl['model']}.caffemodel + oracle/postproc.c decode/NMS, {threads} threads",
                images_per_s=steps * wl["batch"] / dt), dt, steps


# ------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else this process (or NCCL / a library) prints was
    redirected to stderr at start-up, so stdout carries exactly one line."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)            # e.g. "NCCL version ..." banners must not precede the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=0, help="execution contexts of the engine (0 = library default 2)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = args.steps, max(args.warmup, 3)
    config = dict(workload=args.workload, model=wl["model"] + ".caffemodel (reference weights)", precision=wl["precision"],
                  batch_per_gpu=wl["batch"], global_batch=wl["batch"] * world, input=f"{wl['w']}x{wl['h']}",
                  score_thr=SCORE_THR, nms_thr=NMS_THR, parallelism=f"dp{world}", input_data="S-real: golden photo letter-boxed, rolled 8*i px")

    if args.impl == "reference":
        # the reference arm: CPU path, rank 0 only
        if rank != 0:
            return
        sample = make_batches(wl, 4, 0)
        cpu = CpuPath(wl, best_cpu_threads(wl, sample[0]))
        for _ in range(min(W, 3)):
            cpu.step(sample[0])
        # each step = one batch; bound the run to a few minutes
        t0 = time.perf_counter()
        faces = 0
        steps = 0
        for s in range(K):
            faces += cpu.step(sample[s % len(sample)])
            steps += 1
            if time.perf_counter() - t0 > 150:
                break
        dt = time.perf_counter() - t0
        v = faces / dt
        line = dict(metric="faces/sec (end-to-end detect)", value=v, unit="faces/s", n_gpus=args.gpus, steps=steps, warmup=W,
                    ms_per_step=dt / steps * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                    data="synthetic", impl="reference", config=config, images_per_s=steps * wl["batch"] / dt,
                    cpu_baseline=dict(value=v, unit="faces/s", cores=cpu.threads, kind="port",
                                      sample=f"{steps} steps x {wl['batch']} images: cv2.dnn FP32 forward + oracle/postproc.c "
                                             "(the reference's CPU-Caffe path cannot be built: Caffe/OpenCV C++ absent)"),
                    e2e=dict(value=v, unit="faces/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        emit(line)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from retinaface_b200 import RF_PREC_FP16, RF_PREC_FP32, RF_PREC_INT8, Engine
    prec = {"fp16": RF_PREC_FP16, "fp32": RF_PREC_FP32, "int8": RF_PREC_INT8}[wl["precision"]]
    B, H, Wd = wl["batch"], wl["h"], wl["w"]
    eng = Engine(os.path.join(GOLD, "weights", wl["model"] + ".caffemodel"), H, Wd, precision=prec, max_batch=B,
                 max_faces=128, device=local, streams=args.streams,
                 int8_table=os.path.join(GOLD, "weights", wl["model"] + ".table.int8") if prec == RF_PREC_INT8 else None)
    stream = torch.cuda.ExternalStream(eng.stream_ptr(), device=local)
    img_bytes = B * H * Wd * 3
    l2_bytes = 126 * 2**20
    ring = max(4, min(256, -(-2 * l2_bytes // img_bytes)))     # input ring > 2 x L2
    host = make_batches(wl, ring, rank)
    pinned = torch.from_numpy(host).pin_memory()
    dev = pinned.to(f"cuda:{local}", non_blocking=False)       # device-resident inputs for `value`
    faces = np.empty((B, eng.max_faces, 15), dtype=np.float32)
    counts = np.zeros(B, dtype=np.int32)
    pin_np = pinned.numpy()

    def e2e_step(slot):
        imgs = [pin_np[slot, i] for i in range(B)]
        return eng.detect_batch(imgs, SCORE_THR, NMS_THR)

    # faces per ring slot (deterministic; also the warm-up of both paths)
    faces_per_slot = np.zeros(ring, dtype=np.int64)
    for s in range(ring):
        faces_per_slot[s] = sum(len(f) for f in e2e_step(s))
    # all-gather buffer for N>1: the fixed-size per-image detection records

    def dev_tensor(ptr, nbytes):
        class _W:  # minimal __cuda_array_interface__ carrier
            pass
        w = _W()
        w.__cuda_array_interface__ = dict(shape=(nbytes,), typestr="|u1", data=(ptr, False), version=2)
        return torch.as_tensor(w, device=f"cuda:{local}")

    eng.detect_device(B, SCORE_THR, NMS_THR, dev[0].data_ptr())
    eng.synchronize()
    views = {}      # (dets ptr) -> (det view, count view, external stream) of one execution context

    def device_step(slot):
        dptr, cptr = eng.detect_device(B, SCORE_THR, NMS_THR, dev[slot].data_ptr())
        if world > 1:
            # the one exchange of the path: all-gather of this step's fixed-size detection records,
            # issued on the stream the step ran on (each execution context has its own output buffers)
            if dptr not in views:
                views[dptr] = (dev_tensor(dptr, B * eng.max_faces * 64), dev_tensor(cptr, B * 4),
                               torch.cuda.ExternalStream(eng.last_stream_ptr(), device=local),
                               torch.empty(world * B * eng.max_faces * 64, dtype=torch.uint8, device=f"cuda:{local}"),
                               torch.empty(world * B * 4, dtype=torch.uint8, device=f"cuda:{local}"))
            dv, cv, st, gd, gc = views[dptr]
            with torch.cuda.stream(st):
                dist.all_gather_into_tensor(gd, dv)
                dist.all_gather_into_tensor(gc, cv)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed value ---------------------------------------------------------------
    for i in range(W):
        device_step(i % ring)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for i in range(K):
        device_step((W + i) % ring)
    eng.fence()                     # stream (context 0) now follows the work queued on every context
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    dev_faces = int(sum(faces_per_slot[(W + i) % ring] for i in range(K)))

    # ---- end-to-end, blocking call: rf_detect_batch (host pinned in, host faces out), one batch at a time ----
    for i in range(W):
        e2e_step(i % ring)
    barrier()
    t0 = time.perf_counter()
    blk_faces = 0
    for i in range(K):
        out = e2e_step((W + i) % ring)
        blk_faces += sum(len(f) for f in out)
    barrier()
    blk_s = time.perf_counter() - t0
    # ---- end-to-end, pipelined: rf_submit_batch / rf_collect_batch, 3 batches in flight; every step still
    #      copies its own input H2D from pinned memory and reads its own faces back ---------------------------
    fbuf = np.empty((B, eng.max_faces, 15), dtype=np.float32)
    cbuf = np.zeros(B, dtype=np.int32)

    from retinaface_b200.capi import PIPELINE_DEPTH as depth

    def pipelined(nsteps, first):
        inflight, nfaces = [], 0
        for i in range(nsteps):
            if len(inflight) == depth:
                _, c = eng.collect(inflight.pop(0), fbuf, cbuf)
                nfaces += int(c.sum())
            slot = (first + i) % ring
            inflight.append(eng.submit([pin_np[slot, j] for j in range(B)], SCORE_THR, NMS_THR))
        while inflight:
            _, c = eng.collect(inflight.pop(0), fbuf, cbuf)
            nfaces += int(c.sum())
        return nfaces

    pipelined(W, 0)
    barrier()
    t0 = time.perf_counter()
    e2e_faces = pipelined(K, W)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.summary()

    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, float(dev_faces), float(e2e_faces), blk_s, float(blk_faces)], dtype=torch.float64,
                         device=f"cuda:{local}")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, e2e_s, blk_s = float(tmax[0]), float(tmax[1]), float(tmax[4])
        dev_faces, e2e_faces, blk_faces = int(tsum[2]), int(tsum[3]), int(tsum[5])

    line = None
    if rank == 0:
        pk = peaks()
        value = dev_faces / (dev_ms * 1e-3)
        # ---- roofline of the dominant kernel (direct launches, CUDA events on the library stream) ----
        prof = eng.profile_layers(B, iters=max(10, min(K, 100)))
        tot = sum(p["ms"] for p in prof)
        top = max(prof, key=lambda p: p["ms"])
        ach_gbs = top["bytes"] / (top["ms"] * 1e-3) / 1e9
        ach_tf = top["flops"] / (top["ms"] * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tp):            # dram__bytes_read+write of the dominant kernel from the committed ncu --set full capture
            tj = json.load(open(tp))
            if tj.get("workload") == args.workload and tj.get("kernel") == top["name"]:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        roof = dict(bound="hbm", kernel=top["name"], achieved=ach_gbs, peak=pk["hbm_gbs"], unit="GB/s", frac=ach_gbs / pk["hbm_gbs"],
                    traffic=traffic, peak_source=pk["src"], kernel_ms=top["ms"], kernel_share_of_step=top["ms"] / tot,
                    tensor_tflops=ach_tf, tensor_frac=ach_tf / pk["bf16_tflops"],
                    step_algorithmic_gb=sum(p["bytes"] for p in prof) / 1e9, step_algorithmic_gflop=sum(p["flops"] for p in prof) / 1e9,
                    step_sum_of_kernels_ms=tot)
        line = dict(metric="faces/sec (end-to-end detect)", value=value, unit="faces/s", n_gpus=world, steps=K, warmup=W,
                    ms_per_step=dev_ms / K, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype={RF_PREC_FP16: "f16", RF_PREC_FP32: "f32", RF_PREC_INT8: "s8"}[prec], data="synthetic", config=dict(config, execution_contexts=args.streams or 4, l2_policy=f"input ring of {ring} batches = {ring * img_bytes / 2**20:.0f} MiB > 2x L2; activations reused in place"),
                    images_per_s=K * B * world / (dev_ms * 1e-3), clocks=clocks,
                    e2e=dict(value=e2e_faces / e2e_s, unit="faces/s", h2d_bytes_per_step=img_bytes,
                             d2h_bytes_per_step=B * 4 + B * eng.max_faces * 64, images_per_s=K * B * world / e2e_s,
                             ms_per_step=e2e_s / K * 1e3,
                             timing=f"host wall clock around K rf_submit_batch/rf_collect_batch steps, {depth} batches in flight",
                             blocking=dict(value=blk_faces / blk_s, ms_per_step=blk_s / K * 1e3, images_per_s=K * B * world / blk_s,
                                           note="one blocking rf_detect_batch per step (latency mode)")),
                    gpu_launches=K * eng.launches_per_batch(B), launches_per_step=eng.launches_per_batch(B), roofline=roof,
                    layers=[dict(name=p["name"], us=round(p["ms"] * 1e3, 2)) for p in prof])
    eng_close = eng.close
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb, _, _ = cpu_measure(wl, host[:4], args.cpu_seconds)
        line["cpu_baseline"] = cb
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        emit(line)
    eng_close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
