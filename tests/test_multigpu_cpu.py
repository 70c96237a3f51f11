"""CPU: the host side of the N>1 path (retinaface_b200/multigpu.py) with the gloo backend, rendezvous on 127.0.0.1:
contiguous sharding, the blob exchange that precedes rf_comm_init, and the un-padding of gathered rows.  (The exchange of
the detection records itself runs inside the NMS kernel: tests/test_gpu_parity.py::test_comm_*.)"""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from retinaface_b200.capi import COMM_BLOB_BYTES
from retinaface_b200.multigpu import exchange_blobs, init_comm, shard_range, unpack_gathered


def test_shard_range_covers_everything():
    for total in (0, 1, 7, 8, 9, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_unpack_gathered_skips_the_padding_of_every_rank():
    """world 3, 7 images, max_batch 3: ranks own 3 / 2 / 2 images, so rows 5 and 8 are padding BETWEEN / BEHIND ranks
    (taking the first `total` rows is wrong whenever world >= 3 and total % world != 0)."""
    world, mb, total, mf = 3, 3, 7, 4
    faces = np.full((world * mb, mf, 15), -1.0, dtype=np.float32)
    counts = np.full(world * mb, 99, dtype=np.int32)          # padding rows carry garbage on purpose
    for r in range(world):
        lo, hi = shard_range(total, r, world)
        for i, img in enumerate(range(lo, hi)):
            k = img % (mf + 1)
            counts[r * mb + i] = k
            faces[r * mb + i, :k, 0] = img + np.arange(k) / 10.0
    out = unpack_gathered(faces, counts, world, mb, total)
    assert len(out) == total
    for img in range(total):
        k = img % (mf + 1)
        assert out[img].shape == (k, 15)
        assert np.allclose(out[img][:, 0], img + np.arange(k) / 10.0)
    with pytest.raises(ValueError):
        unpack_gathered(faces, counts, world, 2, total)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeEngine:
    """Stands in for capi.Engine on a CPU box: records what the comm set-up hands to the library."""
    def __init__(self, rank):
        self.rank, self.got = rank, None

    def comm_export(self, rank, world):
        assert rank == self.rank
        return bytes([rank]) * COMM_BLOB_BYTES

    def comm_init(self, blobs):
        self.got = blobs


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = _FakeEngine(rank)
    init_comm(eng, dist, rank, world, None)
    again = exchange_blobs(bytes([100 + rank]) * COMM_BLOB_BYTES, dist, world)
    q.put((rank, eng.got, again))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("world", [2, 3])
def test_blob_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=90) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, blobs, again in got:
        assert [b[0] for b in blobs] == list(range(world)) and all(len(b) == COMM_BLOB_BYTES for b in blobs)
        assert [b[0] for b in again] == [100 + r for r in range(world)]
