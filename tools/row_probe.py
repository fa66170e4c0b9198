#!/usr/bin/env python
"""Probe: eval bandwidth vs row stride (channel/partition balance) for the config-2 circles."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trajectory_generator_ros2_b200 import workloads  # noqa: E402
from trajectory_generator_ros2_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = torch.device("cuda", 0)
eng = Engine(0)
d_params = eng.upload_params(workloads.circles_cfg2(n))
counts, _ = eng.count(d_params)
total = int(counts.sum(dtype=torch.int64))
eng.plan(d_params, want_outputs=False)
eng.plan(d_params, want_outputs=False)
buf = torch.empty(n * 14 * 1024 + 4096, dtype=torch.float64, device=dev)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


print("lib:", os.environ.get("TGX_LIB", "default"))
for row in (1024, 1008, 1004, 1012, 1016, 1020):
    out = buf[: n * 14 * row].view(n, 14, row)
    ms = timed(lambda: eng.eval(out))
    print(f"row stride {row:5d}: {ms:8.3f} ms  {112 * total / ms / 1e6:8.1f} GB/s  {total / ms / 1e6:7.2f} Gsamples/s")
eng.close()
