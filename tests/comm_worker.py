"""One rank of the fused-exchange test (tests/test_gpu_round2.py): a process with its own handle; `world` of them share ONE
GPU (separate CUDA contexts, time-sliced -- ranks that shared a process could deadlock on hardware-queue aliasing: rank B's
kernels queued behind rank A's spinning wait).  Rendezvous through files in a scratch directory.
usage: python comm_worker.py <rank> <world> <scratch dir> <device>"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def wait_for(paths, timeout=120.0):
    t0 = time.time()
    while not all(os.path.exists(p) for p in paths):
        if time.time() - t0 > timeout:
            raise TimeoutError(f"rendezvous: {paths}")
        time.sleep(0.01)


def put(path, data: bytes):
    with open(path + ".tmp", "wb") as f:
        f.write(data)
    os.replace(path + ".tmp", path)


def main():
    rank, world, scratch, device = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
    import cv2
    from conftest import GOLDEN, caffemodel
    from oracle.inputs import letterbox_bgr_u8
    from retinaface_b200 import RF_PREC_FP16, Engine
    from retinaface_b200.multigpu import unpack_gathered
    B = 4
    img = cv2.imread(os.path.join(GOLDEN, "data", "img.jpg"))
    inp = letterbox_bgr_u8(img, 448, 448)
    mine = [np.ascontiguousarray(np.roll(inp, 100 * rank + 12 * i, axis=1)) for i in range(B)]
    eng = Engine(caffemodel("mnet25"), 448, 448, precision=RF_PREC_FP16, max_batch=B, max_faces=32, device=device)
    local = eng.detect_batch(mine, 0.9, 0.4)
    np.savez(os.path.join(scratch, f"local_{rank}.tmp.npz"), *local)
    os.replace(os.path.join(scratch, f"local_{rank}.tmp.npz"), os.path.join(scratch, f"local_{rank}.npz"))
    put(os.path.join(scratch, f"blob_{rank}.bin"), eng.comm_export(rank, world))
    wait_for([os.path.join(scratch, f"blob_{r}.bin") for r in range(world)] + [os.path.join(scratch, f"local_{r}.npz") for r in range(world)])
    eng.comm_init([open(os.path.join(scratch, f"blob_{r}.bin"), "rb").read() for r in range(world)])
    everyone = []
    for r in range(world):
        z = np.load(os.path.join(scratch, f"local_{r}.npz"))
        everyone.append([z[f"arr_{i}"] for i in range(B)])
    bad = 0
    for step in range(8):            # more steps than the ring is deep in contexts; two steps in flight
        tickets = [eng.submit(mine, 0.9, 0.4, allgather=True) for _ in range(2)]
        for t in tickets:
            faces, counts = eng.collect(t)
            per = unpack_gathered(faces, counts, world, B, world * B)
            for r in range(world):
                for i in range(B):
                    want = everyone[r][i].reshape(-1, 15)
                    if per[r * B + i].shape != want.shape or not np.array_equal(per[r * B + i], want):
                        bad += 1
    # a plain (local) step still works on a comm-enabled handle
    again = eng.detect_batch(mine, 0.9, 0.4)
    bad += sum(0 if np.array_equal(a, b) else 1 for a, b in zip(again, local))
    eng.close()
    print(f"rank {rank}: {bad} mismatches, {sum(len(x) for x in local)} local faces", flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
