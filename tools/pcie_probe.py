#!/usr/bin/env python
"""PCIe device->host ceiling of the box, to put bench.py's end-to-end number in context: a plain contiguous copy into
pinned memory, and the strided pattern tgx_generate_host uses (2 rows of every trajectory = 16 KB runs, pitch 112 KB)."""
import json
import time

import torch

dev = torch.device("cuda", 0)
n, row = 8192, 1024
src = torch.empty((n, 14, row), dtype=torch.float64, device=dev).normal_()
dst = torch.empty((n, 14, row), dtype=torch.float64).pin_memory()
out = {}


def timed(fn, nbytes, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


out["contiguous_d2h_GBps"] = timed(lambda: dst.copy_(src, non_blocking=True), src.numel() * 8)


def strided():
    for q in range(5):
        dst[:, 3 * q:3 * q + 2].copy_(src[:, 3 * q:3 * q + 2], non_blocking=True)


out["strided_16KB_runs_d2h_GBps"] = timed(strided, n * 10 * row * 8)
h = torch.empty((n, 128), dtype=torch.uint8).pin_memory()
d = torch.empty((n, 128), dtype=torch.uint8, device=dev)
out["h2d_small_GBps"] = timed(lambda: d.copy_(h, non_blocking=True), n * 128)
print(json.dumps(out))
