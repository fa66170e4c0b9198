"""GPU parity tests for the constant-speed polyline family (Square, Rectangle, Reciprocating,
Bounce, M, I, T), through the C-ABI, against the CPU oracle.

Counts, status bits, the leg structure (what the per-sample index_msgs strings are a function of) and POSITIONS are
bit-exact (frac = i / steps; p = start + frac * (end - start) is evaluated with the reference's own roundings);
velocity and yaw go through the device's atan2 / cos / sin of the leg heading and are held to the tolerances in parity.py.
"""
import numpy as np
import pytest

from parity import assert_samples_close, merge_errors
from trajectory_generator_ros2_b200 import abi, workloads
from trajectory_generator_ros2_b200 import trajectories as T

pytestmark = pytest.mark.gpu


def gpu_polyline(engine, params, capacity=None, plane_major=False):
    """tgx_polyline_finalize_host + tgx_plan_polyline + tgx_eval on device tensors."""
    import torch
    params = engine.finalize_polyline(np.ascontiguousarray(params).copy())
    d_params = engine.upload_params(params)
    plan = engine.plan_polyline(d_params, want_legs=True)
    counts = plan.counts.cpu().numpy()
    status = plan.status.cpu().numpy().view(np.uint32)
    n = len(params)
    cap = capacity if capacity is not None else max(4, int((counts.max(initial=0) + 3) // 4 * 4))
    shape = (abi.TGX_NCHAN, n, cap) if plane_major else (n, abi.TGX_NCHAN, cap)
    out = torch.full(shape, float("nan"), dtype=torch.float64, device=d_params.device)
    engine.eval(out, plane_major=plane_major)
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    if plane_major:
        o = np.ascontiguousarray(o.transpose(1, 0, 2))
    legs = plan.legs.cpu().numpy().view(abi.LEGS_DTYPE).reshape(n)
    assert plan.total_samples == int(counts.sum())
    return o, counts, status, legs


def check_polyline_batch(engine, oracle, params, what, check_msgs_every=1, **kw):
    out, counts, status, legs = gpu_polyline(engine, params, **kw)
    o_counts, o_status = oracle.count_batch(params)
    np.testing.assert_array_equal(counts, o_counts, err_msg=f"{what}: sample counts")
    np.testing.assert_array_equal(status, o_status, err_msg=f"{what}: status bits")
    worst = {}
    for i in range(len(params)):
        n = counts[i]
        if n == 0:
            continue
        ref, st, leg_of, msgs = oracle.polyline_generate(params[i:i + 1])
        assert ref.shape[1] == n
        got = out[i, :, :n]
        merge_errors(worst, assert_samples_close(got, ref, f"{what}[{i}]"))
        # positions: bit for bit (-0.0 == +0.0)
        assert np.array_equal(got[abi.PX:abi.PZ + 1] + 0.0, ref[abi.PX:abi.PZ + 1] + 0.0), f"{what}[{i}]: positions"
        # the row's last 32-byte sector is completed with zeros, nothing else is written (tgx.h: tgx_layout)
        n4 = min((n + 3) // 4 * 4, out.shape[2])
        assert (out[i, :, n:n4] == 0).all() and np.isnan(out[i, :, n4:]).all(), f"{what}[{i}]: padding was written"
        assert int(legs[i]["n"]) == n
        if i % check_msgs_every == 0:
            t = int(params["type"][i])
            assert abi.polyline_index_msgs(t, legs[i]) == msgs, f"{what}[{i}]: index_msgs"
    return worst


@pytest.mark.parametrize("kind", abi.POLYLINE_TYPES)
def test_default_yaml_shapes(engine, oracle, kind):
    """config/default.yaml's parameters for every shape of the family (traj_type: T is the shipped default)."""
    p = workloads.default_polyline(kind)
    out, counts, status, legs = gpu_polyline(engine, p)
    expect = 8001 if kind in (abi.TGX_SQUARE, abi.TGX_RECTANGLE, abi.TGX_RECIPROCATING) else 8000
    assert counts[0] == expect and status[0] == 0
    check_polyline_batch(engine, oracle, p, f"default {abi.TYPE_NAMES[kind]}")


def test_golden_polyline_fixtures(engine):
    """The committed outputs of the unmodified reference (tests/golden/reference_golden_polyline.json)."""
    import golden_util
    for c in golden_util.load_polyline()["cases"]:
        p = c["params"]
        out, counts, status, legs = gpu_polyline(engine, p)
        assert counts[0] == c["n"] and status[0] == c["status"], c["name"]
        for k, ref in c["sample_values"].items():
            got = out[0, :, k]
            assert_samples_close(got[:, None], ref[:, None], f"{c['name']}[{k}]")
            assert golden_util.same_bits(got[:3], ref[:3]), f"{c['name']}[{k}] position bits"
        if c["n"]:
            assert abi.polyline_index_msgs(c["type"], legs[0]) == c["msgs"], c["name"]
            g = c["stop"]
            frm = np.full(abi.TGX_NCHAN, np.nan)
            if g["from_k"] in c["sample_values"]:
                frm = c["sample_values"][g["from_k"]]
                _, scounts, _, _ = engine.stop_host(p, frm, 0)
                assert scounts[0] == g["n"], c["name"]


def test_random_mix(engine, oracle):
    worst = check_polyline_batch(engine, oracle, workloads.polyline_mix(420), "polyline mix", check_msgs_every=7)
    assert worst["pos_abs"] == 0.0


def test_round_parameters_hit_ceil_boundaries(engine, oracle):
    """"Round" YAML-style parameters put distance / (v*dt) on integers, where one ulp of the rotated waypoints decides
    steps = ceil(...): counts and leg boundaries must still match exactly for every orientation."""
    parts = []
    for ori in (0.0, 0.5, 1.0, np.pi / 2, np.pi / 4, -2.0, 3.0):
        for v in (0.5, 1.0, 2.0):
            for dt in (0.01, 0.02):
                parts += [abi.square_params(1.8, 2.0, 0.0, 0.0, ori, [v], 12.0, 0.4, dt),
                          abi.rectangle_params(1.8, 2.0, 4.0, 0.5, -0.5, ori, [v], 12.0, 0.4, dt),
                          abi.letter_params(abi.TGX_M, 0.0, 0.0, 3.0, 4.0, 1.8, [v], 12.0, ori, dt),
                          abi.letter_params(abi.TGX_I, 1.0, 1.0, 3.0, 4.0, 1.8, [v], 12.0, ori, dt),
                          abi.letter_params(abi.TGX_T, 0.0, 0.0, 3.0, 4.0, 1.8, [v], 12.0, ori, dt),
                          abi.bounce_params(0.0, 0.0, 4.0, 1.0, [v], 12.0, ori, dt),
                          abi.reciprocating_params(1.8, [0.0, -3.0, 1.8], [0.0, 3.0, 1.8], [v], 1.5, 1.0, 12.0, dt)]
    check_polyline_batch(engine, oracle, abi.concat(parts), "round polyline", check_msgs_every=3)


def test_edge_cases(engine, oracle):
    """Empty and tiny trajectories, a leg shorter than one step, laps that end exactly on a leg boundary, rejected
    parameters and a trajectory of the other family."""
    sq = abi.square_params
    parts = [
        sq(1.8, 2.0, 0, 0, 0.0, [1.0], 0.0, 0.4, 0.01),            # t_traj = 0: only the start sample
        sq(1.8, 2.0, 0, 0, 0.0, [1.0], -1.0, 0.4, 0.01),
        sq(1.8, 2.0, 0, 0, 0.0, [1.0], 0.005, 0.4, 0.01),          # one step
        sq(1.8, 0.001, 0, 0, 0.3, [1.0], 1.0, 0.4, 0.01),          # every side is a single step
        abi.bounce_params(0, 0, 1.0, 1.004, [1.0], 0.5, 0.0, 0.01),  # legs of two samples
        abi.bounce_params(0, 0, 1.0, 3.0, [1.0], 0.0, 0.0, 0.01),    # empty
        abi.reciprocating_params(1.8, [0, 0, 1.8], [1, 0, 1.8], [1.0], 1.0, 1.0, 1.02, 0.01),   # H == steps + 2
        abi.reciprocating_params(1.8, [0, 0, 1.8], [1, 0, 1.8], [1.0], 1.0, 1.0, 1.01, 0.01),   # cut before the flip
        abi.reciprocating_params(1.8, [0, 0, 1.8], [1, 0, 1.8], [1.0], 1.0, 1.0, 2.5, 0.01),
        abi.reciprocating_params(1.8, [0, 0, 1.8], [0, 0, 1.8], [1.0], 1.0, 1.0, 2.5, 0.01),    # A == B: rejected
        abi.letter_params(abi.TGX_T, 0, 0, 3.0, 4.0, 1.8, [0.0], 5.0, 0.0, 0.01),               # v = 0: rejected
        abi.letter_params(abi.TGX_M, 0, 0, 3.0, 0.0, 1.8, [1.0], 5.0, 0.0, 0.01),               # width = 0: rejected
        abi.letter_params(abi.TGX_I, 0, 0, 3.0, 4.0, 1.8, [1.0], 30.0, 1.0, 0.01),              # several laps, both ways
        abi.letter_params(abi.TGX_T, 0, 0, 3.0, 4.0, 1.8, [], 5.0, 0.0, 0.01),                  # v_goals empty -> 1.0
    ]
    params = abi.concat(parts)
    check_polyline_batch(engine, oracle, params, "polyline edge")
    # the wrong planner: no samples, a status bit, nothing written
    circ = workloads.default_circle()
    out, counts, status, legs = gpu_polyline(engine, abi.concat([circ, parts[2]]))
    assert counts[0] == 0 and status[0] == abi.ST_WRONG_PLANNER and counts[1] > 0
    import torch
    d = engine.upload_params(abi.concat([parts[2], circ]))
    plan = engine.plan(d)
    assert plan.counts.cpu().numpy().tolist()[0] == 0
    assert plan.status.cpu().numpy().view(np.uint32)[0] == abi.ST_WRONG_PLANNER
    assert plan.counts.cpu().numpy()[1] == 25001


def test_ragged_batch_uses_the_tile_list(engine, oracle):
    """One long trajectory among short ones: slab addressing would launch mostly empty CTAs, so the plan switches to a
    scanned work list; results are identical."""
    short = workloads.polyline_mix(64, seed=99)
    long_ = workloads.default_polyline(abi.TGX_T)
    check_polyline_batch(engine, oracle, abi.concat([short[:32], long_, short[32:]]), "ragged polyline",
                         check_msgs_every=8)


def test_layouts_and_truncation(engine, oracle):
    p = workloads.polyline_mix(40, seed=5)
    a = check_polyline_batch(engine, oracle, p, "plane-major polyline", plane_major=True, check_msgs_every=40)
    assert a["pos_abs"] == 0.0
    # capacity below the sample count: the tail is not written
    out, counts, status, legs = gpu_polyline(engine, p, capacity=512)
    assert (counts > 512).all()
    ref, _, _, _ = oracle.polyline_generate(p[3:4])
    assert np.array_equal(out[3, :3, :512] + 0.0, ref[:3, :512] + 0.0)


def test_device_trig_without_host_libm(engine, oracle):
    """Without TGX_POLY_TRIG_GIVEN the planner uses the device's cos / sin of the orientation: with orientation 0 (the
    shipped default) the waypoints are exact either way; otherwise positions stay within the position tolerance
    whenever the step counts agree."""
    import torch
    p = abi.concat([workloads.default_polyline(k) for k in abi.POLYLINE_TYPES])
    d = engine.upload_params(p)                       # flag not set
    plan = engine.plan_polyline(d)
    counts = plan.counts.cpu().numpy()
    o_counts, _ = oracle.count_batch(p)
    np.testing.assert_array_equal(counts, o_counts)
    cap = int((counts.max() + 3) // 4 * 4)
    out = torch.full((len(p), abi.TGX_NCHAN, cap), float("nan"), dtype=torch.float64, device=d.device)
    engine.eval(out)
    host = out.cpu().numpy()
    for i in range(len(p)):
        ref, _, _, _ = oracle.polyline_generate(p[i:i + 1])
        assert np.array_equal(host[i, :3, :counts[i]] + 0.0, ref[:3] + 0.0)


def test_feasibility_on_polyline_plans(engine, oracle):
    import torch
    p = engine.finalize_polyline(workloads.polyline_mix(300, seed=3).copy())
    lim = abi.make_limits(box=(-5, 5, -5, 5, 0, 5), v_max=2.0, a_max=6.0)
    d = engine.upload_params(p)
    engine.plan_polyline(d, limits=lim)
    flags, mv, ma, status = engine.feasibility(lim, len(p))
    torch.cuda.synchronize()
    o_flags, o_mv, o_ma, o_counts, o_status = oracle.feasibility_batch(p, lim)
    np.testing.assert_array_equal(status.cpu().numpy().view(np.uint32), o_status)
    np.testing.assert_array_equal(flags.cpu().numpy(), o_flags)
    np.testing.assert_allclose(mv.cpu().numpy(), o_mv, rtol=1e-12, atol=1e-15)
    assert (ma.cpu().numpy() == 0).all() and (o_ma == 0).all()
    assert 0 < o_flags.sum() < len(p)


def test_braking_trajectories(engine, oracle):
    """generateStopTraj of the family: frozen position, speed ramp along the heading (Square.cpp:112-137 and copies),
    Bounce's 0.8-decay (Bounce.cpp:74-103).  Counts exact given the same setpoint bits."""
    import torch
    params = abi.concat([workloads.polyline_mix(140, seed=21)] +
                        [workloads.default_polyline(k) for k in abi.POLYLINE_TYPES])
    n = len(params)
    froms = np.zeros((n, abi.TGX_NCHAN))
    for i in range(n):
        ref, _, _, _ = oracle.polyline_generate(params[i:i + 1])
        froms[i] = ref[:, (ref.shape[1] * (1 + i % 5)) // 7]
    out, counts, status, phases = engine.stop_host(params, froms, 0, want_phases=True)
    cap = max(4, int((counts.max() + 3) // 4 * 4))
    out, counts, status, phases = engine.stop_host(params, froms, cap, want_phases=True)
    some = 0
    for i in range(n):
        ref, st, oph = oracle.stop(params[i:i + 1], froms[i])
        assert counts[i] == ref.shape[1], (i, int(params["type"][i]))
        assert (int(status[i]) & ~abi.ST_TRUNCATED) == st
        t = int(params["type"][i])
        assert abi.phases_to_index_msgs(t, phases[i], stop_traj=True) == abi.phases_to_index_msgs(t, oph, stop_traj=True)
        if counts[i]:
            some += 1
            assert_samples_close(out[i, :, :counts[i]], ref, f"polyline stop[{i}]")
            # the speeds themselves are replayed exactly
            if t == abi.TGX_BOUNCE:
                assert np.array_equal(out[i, abi.VZ, :counts[i]] + 0.0, ref[abi.VZ] + 0.0)
    assert some > n // 2


def test_host_calls_route_mixed_batches(engine, oracle):
    """tgx_generate_host_legs on a batch mixing both families (and tgx_count_host): every trajectory goes to its own
    planner, constants are filled on the host unless a Bounce is present."""
    a = workloads.mixed_cfg3(48)
    b = workloads.polyline_mix(48, seed=8)
    for params in (abi.concat([a[:24], b[:24], a[24:], b[24:]]),
                   abi.concat([b[i:i + 1] for i in range(48) if b["type"][i] != abi.TGX_BOUNCE] + [a])):
        counts, status = engine.count_host(params)
        o_counts, o_status = oracle.count_batch(params)
        np.testing.assert_array_equal(counts, o_counts)
        np.testing.assert_array_equal(status, o_status)
        cap = int((counts.max() + 3) // 4 * 4)
        out, counts2, status2, phases, legs = engine.generate_host_legs(params, cap)
        np.testing.assert_array_equal(counts2, o_counts)
        np.testing.assert_array_equal(status2, o_status)
        for i in range(len(params)):
            t = int(params["type"][i])
            if abi.is_polyline(t):
                ref, _, _, msgs = oracle.polyline_generate(params[i:i + 1])
                assert abi.polyline_index_msgs(t, legs[i]) == msgs
            else:
                ref, _, oph = oracle.generate(params[i:i + 1])
                assert abi.phases_to_index_msgs(t, phases[i]) == abi.phases_to_index_msgs(t, oph)
            assert_samples_close(out[i, :, :counts[i]], ref, f"host mixed[{i}]")


def test_python_mirror_classes(engine, oracle):
    """The drop-in classes (same constructor arguments as Square.hpp ... T.hpp): generateTraj appends and keys
    index_msgs by sample index, generateStopTraj replaces, trajectoryInsideBounds, create<Shape>Goal."""
    trajs = [
        (T.Square(1.8, 2.0, 0.0, 0.0, 0.3, [1.0, 2.0], 12.0, 0.4, 0.01, engine=engine), abi.TGX_SQUARE),
        (T.Rectangle(1.8, 2.0, 4.0, 0.0, 0.0, 0.0, [1.0], 12.0, 0.4, 0.01, engine=engine), abi.TGX_RECTANGLE),
        (T.Reciprocating(1.8, [0, -3, 1.8], [0, 3, 1.8], [1.0], 1.5, 1.0, 12.0, 0.01, engine=engine),
         abi.TGX_RECIPROCATING),
        (T.Bounce(0.0, 0.0, 4.0, 1.0, [1.0], 12.0, 0.0, 0.01, engine=engine), abi.TGX_BOUNCE),
        (T.M(0.0, 0.0, 3.0, 4.0, 1.8, [1.0], 12.0, 0.2, 0.01, engine=engine), abi.TGX_M),
        (T.I(0.0, 0.0, 3.0, 4.0, 1.8, [1.0], 12.0, 0.0, 0.01, engine=engine), abi.TGX_I),
        (T.T(0.0, 0.0, 3.0, 4.0, 1.8, [1.0, 2.0, 2.0], 12.0, 0.0, 0.01, engine=engine), abi.TGX_T),
    ]
    for traj, kind in trajs:
        ref, _, _, msgs = oracle.polyline_generate(traj.params)
        goals = [T.Goal()]                      # generateTraj appends
        index_msgs = {}
        traj.generateTraj(goals, index_msgs)
        assert len(goals) == 1 + ref.shape[1]
        assert index_msgs == {k + 1: m for k, m in msgs.items()}
        got = np.stack([g.channels() for g in goals[1:]], axis=1)
        assert_samples_close(got, ref, traj.shape)
        assert goals[1].frame_id == "world" and goals[1].power
        # braking from the middle
        k = len(goals) // 2
        sref, _, sph = oracle.stop(traj.params, goals[k].channels())
        idx = traj.generateStopTraj(goals, index_msgs, k)
        assert idx == 0 and len(goals) == sref.shape[1]
        assert index_msgs == abi.phases_to_index_msgs(kind, sph, stop_traj=True)
        if goals:
            assert_samples_close(np.stack([g.channels() for g in goals], axis=1), sref, traj.shape + " stop")
        for box in ((-5, 5, -5, 5, -5, 5), (-1, 1, -1, 1, 0, 3)):
            assert traj.trajectoryInsideBounds(*box) == oracle.inside_bounds(traj.params, box)
    sq = trajs[0][0]
    g = sq.createSquareGoal(1.0, 2.0, 1.5, -0.4, 0.7)
    np.testing.assert_allclose([g.p.x, g.p.y, g.p.z, g.v.x, g.v.y, g.a.x, g.a.y, g.psi],
                               [1.0, 2.0, 1.8, 1.5 * np.cos(0.7), 1.5 * np.sin(0.7), -0.4 * np.cos(0.7),
                                -0.4 * np.sin(0.7), 0.7], rtol=1e-14)
    g = trajs[3][0].createBounceGoal(0.1, 0.2, 2.5, -1.0, 0.3)
    assert (g.p.x, g.p.y, g.p.z, g.v.x, g.v.y, g.v.z, g.psi) == (0.1, 0.2, 2.5, 0.0, 0.0, -1.0, 0.3)


def test_full_size_T_batch(engine, oracle):
    """Row f2 at full size: 1 Mi T trajectories (the shape default.yaml ships) x ~1000 samples, 117.5 GB of planes.
    Counts against the oracle for every trajectory; for every sample p.z = alt, a = j = dpsi = 0, |v| = v_goal; the batch
    evaluated as 16 independent shards reproduces its slices bit for bit; positions of one trajectory in 4096 are
    bit-identical to the oracle's."""
    import torch
    n = 1 << 20
    params = engine.finalize_polyline(workloads.letters_T(n).copy())
    d = engine.upload_params(params)
    torch.cuda.empty_cache()                        # blocks cached by earlier tests count as used otherwise
    free, _ = torch.cuda.mem_get_info()
    if free < 135 * (1 << 30):
        pytest.skip("needs ~125 GB of free device memory")
    plan = engine.plan_polyline(d)
    counts = plan.counts
    o_counts, o_status = oracle.count_batch(params, nthreads=32)
    np.testing.assert_array_equal(counts.cpu().numpy(), o_counts)
    np.testing.assert_array_equal(plan.status.cpu().numpy().view(np.uint32), o_status)
    out = torch.full((n, 14, 1024), float("nan"), dtype=torch.float64, device=d.device)
    engine.eval(out)
    torch.cuda.synchronize()
    k = torch.arange(1024, device=d.device)[None, :]
    step = 1 << 15
    zero = torch.zeros((), dtype=torch.float64, device=d.device)
    for lo in range(0, n, step):
        sl = slice(lo, lo + step)
        o = out[sl]
        valid = k < counts[sl][:, None]
        alt = torch.from_numpy(params["alt"][sl].copy()).to(d.device)[:, None]
        vg = torch.from_numpy(params["poly_v_goal"][sl].copy()).to(d.device)[:, None]
        assert bool(((o[:, abi.PZ] == alt) | ~valid).all())
        for ch in (abi.VZ, abi.AX, abi.AY, abi.AZ, abi.JX, abi.JY, abi.JZ, abi.DPSI):
            assert bool(((o[:, ch] == 0) | ~valid).all())
        speed = torch.hypot(o[:, abi.VX], o[:, abi.VY])
        assert float(torch.where(valid, (speed - vg).abs() / vg, zero).max()) < 1e-15
    buf = torch.empty((n // 16, 14, 1024), dtype=torch.float64, device=d.device)
    for s in range(16):
        sl = slice(s * (n // 16), (s + 1) * (n // 16))
        p2 = engine.plan_polyline(d[sl])
        assert torch.equal(p2.counts, counts[sl])
        engine.eval(buf)
        assert torch.equal(buf[:, :, :988], out[sl][:, :, :988]), "a shard must reproduce its slice bit for bit"
    sub = np.arange(0, n, 4096)
    host = out[torch.from_numpy(sub).to(d.device)].cpu().numpy()
    for j, i in enumerate(sub):
        ref = oracle.polyline_generate(params[i:i + 1])[0]
        got = host[j, :, :o_counts[i]]
        assert_samples_close(got, ref, f"full size T[{i}]")
        assert np.array_equal(got[:3] + 0.0, ref[:3] + 0.0), f"full size T[{i}]: positions"
