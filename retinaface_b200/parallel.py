"""Multi-GPU plumbing of the detect path: images are independent, so the batch
is split contiguously over ranks (one process per GPU, weights replicated) and the ONLY exchange is an
all-gather of the fixed-size per-image detection records after NMS.  torch.distributed is the plumbing
(NCCL on GPUs; gloo in the CPU tests) -- there is no collective inside the network itself.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

DET_FLOATS = 16  # rf_det: 15 FaceDetectInfo floats + int32 anchor index (bit-cast), 64 bytes


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split: rank r owns images [lo, hi).  The first (total % world) ranks get one extra."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_detections(local_dets, local_counts, group=None):
    """all_gather of fixed-size records.  local_dets: torch tensor [B_local, max_faces, 16] (float32 or raw
    bytes viewed as such), local_counts: [B_local] int32.  Every rank must pass the same B_local (pad the
    last shard).  Returns (dets [world*B_local, max_faces, 16], counts [world*B_local]) in rank order,
    i.e. global image order for the contiguous split of shard_range."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    b = local_dets.shape[0]
    dets = torch.empty((world * b,) + tuple(local_dets.shape[1:]), dtype=local_dets.dtype, device=local_dets.device)
    counts = torch.empty((world * b,), dtype=local_counts.dtype, device=local_counts.device)
    dist.all_gather_into_tensor(dets, local_dets.contiguous(), group=group)     # concatenation along dim 0, rank order
    dist.all_gather_into_tensor(counts, local_counts.contiguous(), group=group)
    return dets, counts


def unpack(dets: np.ndarray, counts: np.ndarray, total: int) -> List[np.ndarray]:
    """Host view: per-image (k,15) face arrays for the first `total` images (drops shard padding)."""
    out = []
    for i in range(total):
        out.append(np.asarray(dets[i, :counts[i], :15]))
    return out
