"""Host side of the multi-GPU path above the C ABI's rf_comm_* entry points.

Images are independent: the batch is split contiguously over the ranks (one process per GPU, weights replicated) and the
only exchange is the all-gather of the per-image detection records, which librf_b200 fuses into its NMS kernel (peer
stores over NVLink, csrc/comm.cu).  What is left for the host: who owns which images (`shard_range`), the one-time
exchange of the 128-byte window blobs (`init_comm`, through whatever process group the caller has -- NCCL on the GPUs,
gloo in the CPU tests), and turning the gathered [world * max_batch] rows back into global image order (`unpack_gathered`).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from .capi import COMM_BLOB_BYTES


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split: rank r owns images [lo, hi).  The first (total % world) ranks get one extra."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def exchange_blobs(blob: bytes, dist, world: int, device=None) -> List[bytes]:
    """all_gather of every rank's window blob, in rank order.  `dist` is torch.distributed with an initialised default
    group; `device` is the CUDA device of this rank for the NCCL backend (None: CPU tensors, gloo)."""
    import torch
    assert len(blob) == COMM_BLOB_BYTES
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8)
    if device is not None:
        mine = mine.to(device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().numpy().tobytes()) for t in out]


def init_comm(engine, dist, rank: int, world: int, local_device: int | None = None) -> None:
    """rf_comm_export on every rank, blob exchange through `dist`, rf_comm_init."""
    import torch
    blob = engine.comm_export(rank, world)
    dev = torch.device("cuda", local_device) if (local_device is not None and dist.get_backend() == "nccl") else None
    engine.comm_init(exchange_blobs(blob, dist, world, dev))


def unpack_gathered(faces: np.ndarray, counts: np.ndarray, world: int, max_batch: int, total: int) -> List[np.ndarray]:
    """faces [world * max_batch, max_faces, 15], counts [world * max_batch] as rf_collect_batch_allgather returns them (rank
    r's image i at row r * max_batch + i; a rank with fewer than max_batch images leaves padding rows behind its own) ->
    one (k, 15) array per GLOBAL image 0..total-1 for the contiguous split of shard_range."""
    out: List[np.ndarray] = []
    for r in range(world):
        lo, hi = shard_range(total, r, world)
        if hi - lo > max_batch:
            raise ValueError(f"rank {r} owns {hi - lo} images, more than max_batch {max_batch}")
        for i in range(hi - lo):
            row = r * max_batch + i
            out.append(np.asarray(faces[row, :counts[row]]))
    assert len(out) == total
    return out
