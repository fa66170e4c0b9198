"""Which trajectories of BASELINE.json config 3's mix keep a batch off the phase-record planning path?

Run on a GPU box (python tools/phase_probe.py).  Plans the lines, the circles and the figure-eights of a 20 000-trajectory
draw separately, then in blocks, and bisects a block that was planned with segment tables down to the trajectory that
does not fit a PhaseRec (tgx_internal.cuh).  Used in round 2 to size the PhaseExt rows: with 12 segments per record 3 of
8 078 circles and 3 of 5 938 figure-eights (two goal speeds, a slow first one: the hold starts at ~0.004 rad and crosses
eight binades of theta) kept whole batches on the table path; with 20 none does.
"""
import sys; sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from trajectory_generator_ros2_b200 import abi, workloads
from trajectory_generator_ros2_b200.engine import Engine
e = Engine(0)
def took(params):
    e.set_phase_planning(True)
    d = e.upload_params(params)
    e.plan(d); 
    b = e.phase_plan_count
    pl = e.plan(d)
    return e.phase_plan_count - b, int(pl.counts.max()), int(pl.counts.min())
mix = workloads.mixed_cfg3(20000)
for t, name in ((abi.TGX_LINE, 'lines'), (abi.TGX_CIRCLE, 'circles'), (abi.TGX_FIGURE8, 'fig8')):
    sub = mix[mix['type'] == t]
    print(name, len(sub), took(sub))
    bad = 0
    for lo in range(0, len(sub), 500):
        r = took(sub[lo:lo+500])
        if r[0] != 1:
            bad += 1
            # bisect
            for i in range(lo, min(lo+500, len(sub)), 50):
                r2 = took(sub[i:i+50])
                if r2[0] != 1:
                    for j in range(i, min(i+50, len(sub))):
                        r3 = took(sub[j:j+1])
                        if r3[0] != 1:
                            print('  misfit', name, j, r3, sub[j:j+1]); break
                    break
            if bad > 2: break
print('mixed all', took(mix))
