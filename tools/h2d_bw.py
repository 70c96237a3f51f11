#!/usr/bin/env python
"""Host->device copy bandwidth of this box from pinned memory, in the transfer sizes bench.py's e2e leg uses
(one batch = 8 x 448 x 448 x 3 bytes): the ceiling of any end-to-end images/s number.  Prints one JSON line."""
import json

import torch


def main():
    out = {}
    for mb in (4.816896, 19.267584, 256.0):
        n = int(mb * 1e6)
        src = torch.empty(n, dtype=torch.uint8).pin_memory()
        dst = torch.empty(n, dtype=torch.uint8, device="cuda")
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(5):
                dst.copy_(src, non_blocking=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(10, int(2e9 / n))
            e0.record()
            for _ in range(reps):
                dst.copy_(src, non_blocking=True)
            e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[f"{mb:.1f}MB"] = {"ms": ms, "GBps": n / ms / 1e6, "images_per_s_448": n / 602112 / (ms * 1e-3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
