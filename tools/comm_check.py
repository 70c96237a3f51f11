#!/usr/bin/env python
"""Multi-GPU check of the exchange fused into the NMS kernel (csrc/comm.cu), one process per GPU under torchrun:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/comm_check.py
Every rank detects its own shard; rf_submit_batch_allgather / rf_collect_batch_allgather must hand EVERY rank the faces of ALL
ranks (compared with what each rank reports for itself through torch.distributed), for both bootstrap paths (blobs through the
caller's process group; blobs through NCCL inside the library) and for the device-resident entry point."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import cv2
    import torch
    import torch.distributed as dist
    from oracle.inputs import letterbox_bgr_u8
    from retinaface_b200 import RF_PREC_FP16, Engine
    from retinaface_b200.capi import nccl_unique_id
    from retinaface_b200.multigpu import init_comm, unpack_gathered
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = 4
    img = cv2.imread(os.path.join(ROOT, "tests", "golden", "data", "img.jpg"))
    inp = letterbox_bgr_u8(img, 448, 448)
    mine = [np.ascontiguousarray(np.roll(inp, 40 * rank + 8 * i, axis=1)) for i in range(B)]
    ok = True
    for mode in ("caller", "nccl"):
        eng = Engine(os.path.join(ROOT, "tests/golden/weights/mnet25.caffemodel"), 448, 448, precision=RF_PREC_FP16, max_batch=B, max_faces=32, device=local)
        local_faces = eng.detect_batch(mine, 0.9, 0.4)
        everyone = [None] * world
        dist.all_gather_object(everyone, [f.tolist() for f in local_faces])
        if mode == "caller":
            init_comm(eng, dist, rank, world, local)
        else:
            uid = [nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            eng.comm_init_nccl(uid[0], rank, world)
        for step in range(12):
            tickets = [eng.submit(mine, 0.9, 0.4, allgather=True) for _ in range(3)]       # three steps in flight
            for t in tickets:
                faces, counts = eng.collect(t)
                per = unpack_gathered(faces, counts, world, B, world * B)
                for r in range(world):
                    for i in range(B):
                        want = np.asarray(everyone[r][i], dtype=np.float32).reshape(-1, 15)
                        if per[r * B + i].shape != want.shape or not np.array_equal(per[r * B + i], want):
                            ok = False
                            print(f"rank {rank} mode {mode} step {step}: mismatch for rank {r} image {i}", flush=True)
        dev = torch.from_numpy(np.stack(mine)).cuda()
        dptr, cptr = eng.detect_device_allgather(B, 0.9, 0.4, dev.data_ptr())
        eng.synchronize()
        # counts of all ranks, read back from this rank's window
        import ctypes
        cnt = torch.empty(world * B, dtype=torch.int32, device="cuda")
        ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(cnt.data_ptr()), ctypes.c_void_p(cptr), world * B * 4, 3)
        want_counts = [len(everyone[r][i]) for r in range(world) for i in range(B)]
        if cnt.cpu().tolist() != want_counts:
            ok = False
            print(f"rank {rank} mode {mode}: device counts {cnt.cpu().tolist()} != {want_counts}", flush=True)
        eng.close()
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("COMM CHECK", "PASS" if int(flag[0]) else "FAIL", f"(world {world})", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag[0]) else 1)


if __name__ == "__main__":
    main()
