"""CPU: synthetic workloads (BASELINE.json configs) and the multi-GPU host logic on the gloo backend."""
import os
import socket

import numpy as np
import pytest

from trajectory_generator_ros2_b200 import abi, workloads


def test_cfg2_recipe_counts(oracle):
    p = workloads.circles_cfg2(20000)
    counts, status = oracle.count_batch(p)
    assert counts.min() >= 1000 and counts.max() <= 1001 and (status == 0).all()
    assert (p["t_traj"] >= 1.4).all()


def test_workloads_are_shard_invariant():
    for fn in (workloads.circles_cfg2, workloads.mixed_cfg3, workloads.montecarlo_cfg4, workloads.polyline_mix,
               workloads.letters_T):
        full = fn(200000)
        part = fn(200000, lo=65000, hi=140001)
        assert full[65000:140001].tobytes() == part.tobytes()
        assert fn(10, lo=3, hi=3).shape == (0,)


def test_philox_known_answers_and_montecarlo_stream(oracle):
    """The counter-based generator behind tgx_fill_montecarlo (config 5 draws its shards on the device): Philox4x32-10
    against the Random123 known-answer vectors, shard invariance, and the config-4 distribution."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = workloads.philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]
        assert tuple(int(x) for x in got) == want
    full = workloads.montecarlo_philox(100000)
    assert full[40000:70001].tobytes() == workloads.montecarlo_philox(100000, lo=40000, hi=70001).tobytes()
    assert workloads.montecarlo_philox(100000, seed=1238)[:100].tobytes() != full[:100].tobytes()
    assert 0.2 <= full["r"].min() and full["r"].max() < 5.0 and abs(full["r"].mean() - 2.6) < 0.03
    v = full["v_goals"][:, 0]
    assert 0.2 <= v.min() and v.max() < 8.0 and abs(v.mean() - 4.1) < 0.05 and (full["t_traj"] >= 0.5).all()
    counts, status = oracle.count_batch(full[:4000])
    assert (status == 0).all() and 900 < counts.mean() < 1300


def test_cfg3_mix_and_line_feasibility(oracle):
    p = workloads.mixed_cfg3(30000)
    frac = np.bincount(p["type"], minlength=3) / len(p)
    assert abs(frac[0] - 0.4) < 0.02 and abs(frac[1] - 0.3) < 0.02 and abs(frac[2] - 0.3) < 0.02
    lines = p[p["type"] == abi.TGX_LINE][:500]
    for i in range(0, len(lines), 25):
        assert oracle.line_d2(lines[i:i + 1]) > 0
    counts, status = oracle.count_batch(p[:3000])
    assert (status == 0).all() and counts.min() > 50 and counts.max() < 2400


def test_polyline_workloads(oracle):
    """Row f2 workloads: all seven shapes present, ~1000 samples each, every record accepted; the T batch is the shape
    default.yaml ships."""
    p = workloads.polyline_mix(7000)
    kinds = np.bincount(p["type"], minlength=11)
    assert (kinds[:4] == 0).all() and (kinds[4:] > 700).all()
    counts, status = oracle.count_batch(p)
    assert (status == 0).all() and counts.min() >= 800 and counts.max() <= 1202
    t = workloads.letters_T(3000)
    counts, status = oracle.count_batch(t)
    assert (t["type"] == abi.TGX_T).all() and (status == 0).all() and counts.min() >= 990 and counts.max() <= 1011
    for k in abi.POLYLINE_TYPES:
        n, st = oracle.count(workloads.default_polyline(k))
        assert st == 0 and n in (8000, 8001)
    tr = workloads.fleet_transitions(400)
    lens = [len(oracle.transition(tr[i:i + 1], box=(-5, 5, -5, 5, 0, 5))[0]) for i in range(0, 400, 7)]
    assert min(lens) > 50 and max(lens) < 2700


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, ret):
    import torch
    import torch.distributed as dist
    from oracle_lib import Oracle
    from trajectory_generator_ros2_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_bounds(n, rank, world)
        params = workloads.montecarlo_cfg4(n, lo=lo, hi=hi)           # every rank draws only its own shard
        lim = abi.make_limits(**workloads.MONTECARLO_LIMITS)
        flags = Oracle().feasibility_batch(params, lim, nthreads=2)[0]   # stand-in for the per-rank GPU result
        full = sharding.gather_flags(torch.from_numpy(flags), n)
        total = sharding.count_feasible(torch.from_numpy(flags))
        ret[rank] = (full.numpy().copy(), total, (lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1001), (3, 500)])
def test_gloo_shard_and_gather(oracle, world, n):
    """world-size 2/3 on CPU: contiguous shards + the optional all-gather of feasibility flags reproduce the
    single-process result (the flags themselves come from the oracle here; the GPU path is tested with -m gpu)."""
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, ret), nprocs=world, join=True)
    lim = abi.make_limits(**workloads.MONTECARLO_LIMITS)
    expect = oracle.feasibility_batch(workloads.montecarlo_cfg4(n), lim)[0]
    spans = []
    for r in range(world):
        full, total, span = ret[r]
        np.testing.assert_array_equal(full, expect)
        assert total == int(expect.sum())
        spans.append(span)
    assert spans[0][0] == 0 and spans[-1][1] == n
