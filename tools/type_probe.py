"""Where does config 3's evaluation lose its 6 %?  Evaluation bandwidth of single-type batches on one B200:

    PYTHONPATH=. python tools/type_probe.py

Measured (round 2): config-2-shaped batches, Circle 7.06 TB/s, Figure8 7.07 TB/s (the atan2 of its yaw costs nothing);
config 3's own lines / circles / figure-eights, each alone: 6.55 / 6.62 / 6.73 TB/s.  The type does not matter; what
does is that config 3's trajectories (two speed goals, up to a dozen segments) do not fit the phase records and are
evaluated from segment tables (two dependent rounds of loads per tile)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from trajectory_generator_ros2_b200 import abi, workloads
from trajectory_generator_ros2_b200.engine import Engine
e = Engine(0)
n = 1 << 19
for name, ty in (("circle", abi.TGX_CIRCLE), ("figure8", abi.TGX_FIGURE8)):
    p = workloads.circles_cfg2(n)
    p["type"] = ty
    d = e.upload_params(p)
    out = torch.empty((n, 14, 1024), dtype=torch.float64, device=d.device)
    for _ in range(3):
        e.plan(d, want_outputs=False); e.eval(out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        pl = e.plan(d, want_outputs=False)
        a.record(); e.eval(out); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(name, "eval ms", np.median(ts), "TB/s", 112 * pl.total_samples / np.median(ts) / 1e9)
# lines of similar length
p = workloads.mixed_cfg3(n)
for name, ty in (("cfg3 lines only", abi.TGX_LINE), ("cfg3 circles only", abi.TGX_CIRCLE), ("cfg3 figure8 only", abi.TGX_FIGURE8)):
    q = p[p["type"] == ty]
    q = abi.concat([q] * (n // len(q)))
    d = e.upload_params(q)
    c, _ = e.count(d)
    row = (int(c.max()) + 1023) // 1024 * 1024
    out = torch.empty((len(q), 14, row), dtype=torch.float64, device=d.device)
    for _ in range(3):
        e.plan(d, want_outputs=False); e.eval(out)
    ts = []
    for _ in range(5):
        pl = e.plan(d, want_outputs=False)
        a.record(); e.eval(out); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(name, "n", len(q), "row", row, "samples", pl.total_samples, "eval ms", np.median(ts), "TB/s", 112 * pl.total_samples / np.median(ts) / 1e9, "tiles", pl.tiles)
    del out
