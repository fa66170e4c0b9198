#!/usr/bin/env python
"""Probe: can tgx_plan of chunk c+1 hide under tgx_eval of chunk c on a second (high-priority) stream?"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from trajectory_generator_ros2_b200 import workloads  # noqa: E402
from trajectory_generator_ros2_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
engs = [Engine(0), Engine(0)]
params = workloads.circles_cfg2(n)
d_params = engs[0].upload_params(params)
rows = n // chunks
out = torch.empty((n, 14, 1024), dtype=torch.float64, device=dev)


def serial():
    for c in range(chunks):
        engs[0].plan(d_params[c * rows:(c + 1) * rows], want_outputs=False)
        engs[0].eval(out[c * rows:(c + 1) * rows])


def pipelined(s_eval, s_plan):
    # plan chunk 0 up front; then eval(c) on s_eval while plan(c+1) runs on s_plan with the other engine
    done_plan = [torch.cuda.Event() for _ in range(chunks)]
    done_eval = [torch.cuda.Event() for _ in range(chunks)]
    with torch.cuda.stream(s_plan):
        engs[0].plan(d_params[0:rows], want_outputs=False)
        done_plan[0].record()
    for c in range(chunks):
        e = engs[c & 1]
        with torch.cuda.stream(s_eval):
            s_eval.wait_event(done_plan[c])
            e.eval(out[c * rows:(c + 1) * rows])
            done_eval[c].record()
        if c + 1 < chunks:
            with torch.cuda.stream(s_plan):
                if c >= 1:
                    s_plan.wait_event(done_eval[c - 1])     # engine (c+1)&1 tables are free once eval(c-1) is done
                engs[(c + 1) & 1].plan(d_params[(c + 1) * rows:(c + 2) * rows], want_outputs=False)
                done_plan[c + 1].record()


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


print(f"n={n} chunks={chunks}")
print(f"serial plan+eval          : {timed(serial):8.3f} ms")
for pri_name, pe, pp in (("equal priority", 0, 0), ("plan high priority", 0, -1), ("eval high priority", -1, 0)):
    s_eval = torch.cuda.Stream(priority=pe)
    s_plan = torch.cuda.Stream(priority=pp)
    ms = timed(lambda: pipelined(s_eval, s_plan))
    print(f"pipelined ({pri_name:18s}): {ms:8.3f} ms")
engs[0].plan(d_params, want_outputs=False)
print(f"eval only (one launch)    : {timed(lambda: engs[0].eval(out)):8.3f} ms")
print(f"plan only (whole batch)   : {timed(lambda: engs[0].plan(d_params, want_outputs=False)):8.3f} ms")
