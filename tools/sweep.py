#!/usr/bin/env python
"""GPU-side tuning sweep (run under gpurun): kernel shapes x layouts for tgx_eval, plus write-bandwidth probes.

Prints one line per configuration: eval ms, achieved GB/s on 112 B/sample.  Not part of the product path.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trajectory_generator_ros2_b200 import abi, workloads  # noqa: E402
from trajectory_generator_ros2_b200.engine import Engine  # noqa: E402


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


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
    workload = sys.argv[2] if len(sys.argv) > 2 else "circles_cfg2"
    dev = torch.device("cuda", 0)
    eng = Engine(0)
    params = getattr(workloads, workload)(n)
    d_params = eng.upload_params(params)
    counts, _ = eng.count(d_params)
    total = int(counts.sum(dtype=torch.int64))
    row = (int(counts.max()) + 1023) // 1024 * 1024
    out = torch.empty((n, 14, row), dtype=torch.float64, device=dev)
    nbytes = out.numel() * 8
    # ---- write-only ceilings -------------------------------------------------------------------------------
    ms = timed(lambda: out.zero_())
    print(f"torch zero_ (memset)        : {ms:8.3f} ms  {nbytes / ms / 1e6:8.1f} GB/s over {nbytes / 1e9:.1f} GB")
    ms = timed(lambda: out.fill_(1.5))
    print(f"torch fill_ (store kernel)  : {ms:8.3f} ms  {nbytes / ms / 1e6:8.1f} GB/s")
    half = out.view(-1)[: out.numel() // 2]
    other = out.view(-1)[out.numel() // 2: out.numel() // 2 * 2]
    ms = timed(lambda: other.copy_(half))
    print(f"torch copy_ (read+write)    : {ms:8.3f} ms  {2 * half.numel() * 8 / ms / 1e6:8.1f} GB/s (r+w bytes)")
    # ---- plan cost --------------------------------------------------------------------------------------------
    ms = timed(lambda: eng.count(d_params))
    print(f"tgx_count                   : {ms:8.3f} ms")
    ms = timed(lambda: eng.plan(d_params, want_outputs=False))
    print(f"tgx_plan                    : {ms:8.3f} ms   tiles {eng._lib.tgx_plan_tiles(eng._h)} segs {eng._lib.tgx_plan_segments(eng._h)} (single-replay, two-replay plans so far: {eng.plan_path_counts()})")
    eng.set_slab_planning(False)
    ms = timed(lambda: eng.plan(d_params, want_outputs=False))
    print(f"tgx_plan (two-replay path)  : {ms:8.3f} ms")
    eng.set_slab_planning(True)
    eng.set_plan_mode(True)
    ms = timed(lambda: eng.plan(d_params, want_outputs=False))
    print(f"tgx_plan (exact ramps)      : {ms:8.3f} ms   paths {eng.plan_path_counts()}")
    eng.set_plan_mode(False)
    # ---- eval shapes --------------------------------------------------------------------------------------------
    for shift, spt in ((9, 2), (9, 4), (10, 4)):
        eng.set_tuning(shift, spt)
        eng.plan(d_params, want_outputs=False)
        for slabs in (False, True):
            eng.set_slab_planning(slabs)
            eng.plan(d_params, want_outputs=False)
            eng.plan(d_params, want_outputs=False)
            ms = timed(lambda: eng.eval(out))
            print(f"eval tile={1 << shift:5d} spt={spt} threads={(1 << shift) // spt:4d} {'slab plan         ' if slabs else 'exact-offset plan  '}: {ms:8.3f} ms  "
                  f"{112 * total / ms / 1e6:8.1f} GB/s  {total / ms / 1e6:7.2f} Gsamples/s")
    eng.close()


if __name__ == "__main__":
    main()
