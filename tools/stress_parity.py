#!/usr/bin/env python
"""Randomised parity sweep on a B200 (not part of the test suite; run under gpurun):

    python tools/stress_parity.py [first_seed] [n_seeds]

For every seed: a mixed circle / line / figure-eight batch and a polyline batch are planned three times (exact offsets,
fixed slices with the class-sorted replay, then — the mixed batch — phase records, one CTA per trajectory), evaluated through
the TMA and the vector-store path, packed into records by both record paths, and compared with the CPU oracle (counts
exact, samples within the parity tolerances); the three plans must write the same bytes."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import Oracle                     # noqa: E402
from parity import assert_samples_close           # noqa: E402
from trajectory_generator_ros2_b200 import abi, workloads   # noqa: E402
from trajectory_generator_ros2_b200.engine import Engine    # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
count = int(sys.argv[2]) if len(sys.argv) > 2 else 4
eng, orc = Engine(0), Oracle()
lim = abi.make_limits(box=(-3.0, 3.0, -3.0, 3.0, 0.5, 2.2))
checked = 0
for seed in range(first, first + count):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(300, 3000))
    for family in ("classic", "polyline"):
        if family == "classic":
            params = workloads.mixed_cfg3(n, seed=seed)
            planner = eng.plan
        else:
            params = eng.finalize_polyline(workloads.polyline_mix(n, seed=seed).copy())
            planner = eng.plan_polyline
        d = eng.upload_params(params)
        o_counts, o_status = orc.count_batch(params)
        outs = []
        for rep in range(3):
            plan = planner(d)
            counts = plan.counts.cpu().numpy()
            assert (counts == o_counts).all(), (seed, family, rep, "counts")
            assert (plan.status.cpu().numpy().view(np.uint32) == o_status).all(), (seed, family, rep, "status")
            cap = int((counts.max() + 3) // 4 * 4)
            for tma in (True, False):
                eng.set_store_path(tma)
                out = torch.full((n, abi.TGX_NCHAN, cap), float("nan"), dtype=torch.float64, device=d.device)
                eng.eval(out)
                outs.append(out)
            eng.set_store_path(True)
            torch.cuda.synchronize()
            assert torch.equal(torch.nan_to_num(outs[-1], nan=-7.0), torch.nan_to_num(outs[-2], nan=-7.0)), \
                (seed, family, rep, "store paths differ")
            want = eng.pack_goals(outs[-1], plan.counts, lim)
            got = eng.eval_records(n, cap, lim)
            torch.cuda.synchronize()
            assert torch.equal(got, want), (seed, family, rep, "record paths differ")
        for rep in (1, 2):
            assert torch.equal(torch.nan_to_num(outs[2 * rep], nan=-7.0), torch.nan_to_num(outs[0], nan=-7.0)), \
                (seed, family, rep, "the planning path changed the samples")
        host = outs[-2].cpu().numpy()
        for i in rng.choice(n, size=40, replace=False):
            ref = (orc.generate(params[i:i + 1])[0] if family == "classic" else orc.polyline_generate(params[i:i + 1])[0])
            assert_samples_close(host[i, :, :counts[i]], ref, f"seed {seed} {family}[{i}]")
            n4 = min((counts[i] + 3) // 4 * 4, cap)
            assert (host[i, :, counts[i]:n4] == 0).all() and np.isnan(host[i, :, n4:]).all()
            checked += 1
        del outs
    print(f"seed {seed}: n = {n} ok (phase plans so far: {eng.phase_plan_count})", flush=True)
print(f"stress parity ok: {count} seeds, {checked} trajectories compared with the oracle sample by sample")
