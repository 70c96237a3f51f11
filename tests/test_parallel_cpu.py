"""CPU: the N>1 plumbing (contiguous sharding + all-gather of detection records) with the gloo backend,
world_size 2, rendezvous on 127.0.0.1."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from retinaface_b200.parallel import gather_detections, shard_range, unpack


def test_shard_range_covers_everything():
    for total in (0, 1, 7, 8, 9, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, max_faces, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = -(-total // world)
    lo, hi = shard_range(total, rank, world)
    dets = torch.zeros(per, max_faces, 16)
    counts = torch.zeros(per, dtype=torch.int32)
    for j, img in enumerate(range(lo, hi)):      # image i holds (i % 5) faces whose score encodes (i, k)
        k = img % 5
        counts[j] = k
        for f in range(k):
            dets[j, f, 0] = img + f / 10.0
    g_d, g_c = gather_detections(dets, counts)
    if rank == 0:
        # re-pack into global image order: rank r contributed its first (hi-lo) rows
        rows = []
        for r in range(world):
            a, b = shard_range(total, r, world)
            rows += list(range(r * per, r * per + (b - a)))
        q.put((g_d.numpy()[rows], g_c.numpy()[rows]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_allgather_of_detection_records_gloo_world2():
    total, max_faces, world = 7, 6, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, max_faces, q)) for r in range(world)]
    for p in procs:
        p.start()
    d, c = q.get(timeout=90)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    faces = unpack(d, c, total)
    for img in range(total):
        assert len(faces[img]) == img % 5
        for f in range(img % 5):
            assert abs(faces[img][f, 0] - (img + f / 10.0)) < 1e-6
