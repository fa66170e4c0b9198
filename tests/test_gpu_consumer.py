"""GPU: the consumer side of the path and the drop-in proven with the reference's own node.

  * tgx_pack_goals (clamp to the room bounds + pack to 128-byte records) against the oracle restatement of
    TrajectoryGenerator.cpp:557 + :602-604, which tests/test_node_oracle.py pins to the unmodified node: bit-exact;
  * tgx_generate_records_host end to end;
  * the UNMODIFIED TrajectoryGenerator.cpp built against the GPU-backed drop-in classes (tests/cpp/bin/libnodegpu.so)
    flown through the same scripted missions as the same node with the reference's classes (oracle/_ref/libnoderef.so):
    the two published streams must agree tick for tick.
"""
import os

import numpy as np
import pytest

import golden_util
import node_lib
from parity import assert_samples_close
from trajectory_generator_ros2_b200 import abi, workloads

pytestmark = pytest.mark.gpu

BOX = (-1.5, 2.5, -2.0, 1.0, 0.5, 2.0)


def oracle_rows(oracle, params, cap):
    """Oracle samples of a batch as [n, 14, cap] (NaN padded) + counts."""
    n = len(params)
    out = np.full((n, abi.TGX_NCHAN, cap), np.nan)
    counts = np.zeros(n, dtype=np.int32)
    for i in range(n):
        p = params[i:i + 1]
        s = oracle.polyline_generate(p)[0] if abi.is_polyline(int(p["type"][0])) else oracle.generate(p)[0]
        counts[i] = s.shape[1]
        out[i, :, :min(cap, s.shape[1])] = s[:, :cap]
    return out, counts


@pytest.mark.parametrize("plane_major", [False, True])
def test_pack_goals_is_bit_exact(engine, oracle, plane_major):
    import torch
    params = abi.concat([workloads.mixed_cfg3(40), workloads.polyline_mix(40, seed=4)])
    cap = 1280
    host, counts = oracle_rows(oracle, params, cap)
    n = len(params)
    dev = torch.device("cuda", engine.device)
    planes = torch.from_numpy(np.ascontiguousarray(host.transpose(1, 0, 2)) if plane_major else host).to(dev)
    d_counts = torch.from_numpy(counts).to(dev)
    for box in (BOX, None):
        lim = abi.make_limits(box=box) if box else None
        rec = engine.pack_goals(planes, d_counts, lim, plane_major=plane_major)
        torch.cuda.synchronize()
        got = rec.cpu().numpy().view(abi.RECORD_DTYPE).reshape(n, cap)
        for i in range(n):
            m = min(int(counts[i]), cap)
            want = oracle.pack_goals(host[i, :, :m], traj=i, box=box)
            if counts[i] > cap:
                want["last"][:] = 0                       # the trajectory's real last sample is beyond the capacity
            assert got[i, :m].tobytes() == want.tobytes(), (i, box)
            assert not got[i, m:].view(np.uint8).any(), "records beyond the count must not be written"
        if box:
            assert sum(int((got[i, :counts[i]]["clamped"] != 0).sum()) for i in range(n)) > 1000


def test_pack_goals_packed_offsets_and_capacity(engine, oracle):
    """Ragged batch into a densely packed record array (exclusive scan of the counts), and a record capacity below the
    sample count."""
    import torch
    params = workloads.mixed_cfg3(64, seed=77)
    host, counts = oracle_rows(oracle, params, 1536)
    dev = torch.device("cuda", engine.device)
    planes = torch.from_numpy(host).to(dev)
    d_counts = torch.from_numpy(counts).to(dev)
    offs = np.concatenate([[0], np.cumsum(counts[:-1], dtype=np.int64)])
    total = int(counts.sum())
    records = torch.zeros((total, 128), dtype=torch.uint8, device=dev)
    engine.pack_goals(planes, d_counts, abi.make_limits(box=BOX), records=records, rec_capacity=1536,
                      rec_offset=torch.from_numpy(offs).to(dev))
    torch.cuda.synchronize()
    got = records.cpu().numpy().view(abi.RECORD_DTYPE).reshape(total)
    for i in range(len(params)):
        want = oracle.pack_goals(host[i, :, :counts[i]], traj=i, box=BOX)
        assert got[offs[i]:offs[i] + counts[i]].tobytes() == want.tobytes(), i
    rec = engine.pack_goals(planes, d_counts, None, rec_capacity=300)
    torch.cuda.synchronize()
    got = rec.cpu().numpy().view(abi.RECORD_DTYPE).reshape(len(params), 300)
    for i in (0, 5, 63):
        m = min(300, counts[i])
        want = oracle.pack_goals(host[i, :, :m], traj=i)
        if counts[i] > 300:
            want["last"][:] = 0
        assert got[i, :m].tobytes() == want.tobytes()


def test_generate_records_host(engine, oracle):
    """plan + eval + pack + D2H of the records through the host-buffer call, mixed families."""
    params = abi.concat([workloads.mixed_cfg3(30, seed=5), workloads.polyline_mix(30, seed=6)])
    counts, _ = engine.count_host(params)
    cap = int((counts.max() + 3) // 4 * 4)
    lim = abi.make_limits(box=BOX)
    rec, counts2, status = engine.generate_records_host(params, cap, lim)
    o_counts, o_status = oracle.count_batch(params)
    np.testing.assert_array_equal(counts2, o_counts)
    host, _ = oracle_rows(oracle, params, cap)
    for i in range(len(params)):
        m = int(counts2[i])
        want = oracle.pack_goals(host[i, :, :m], traj=i, box=BOX)
        got = rec[i, :m]
        assert_samples_close(abi.records_to_channels(got), abi.records_to_channels(want), f"records[{i}]")
        for f in ("traj", "k", "power", "mode_xy", "mode_z", "last"):
            np.testing.assert_array_equal(got[f], want[f], err_msg=f"{i}:{f}")
        # the clamp verdict may differ only where the reference sample sits within the tolerance of a wall
        diff = got["clamped"] != want["clamped"]
        if diff.any():
            p = host[i, :3, :m]
            walls = np.array(BOX).reshape(3, 2)
            near = (np.abs(p[:, None, :] - walls[:, :, None]) < 1e-9).any(axis=(0, 1))
            assert near[diff].all(), i


# ---- the reference's own node, flown with the drop-in classes ------------------------------------------------

needs_nodes = pytest.mark.skipif(not (os.path.exists(node_lib.NODE_REF_SO) and os.path.exists(node_lib.NODE_GPU_SO)),
                                 reason="node harness libraries not built (need /root/reference at build time)")

T_GO1, T_GO2, T_GO3 = 5, 700, 2600
ALL_TYPES = ["Circle", "Line", "Boomerang", "Figure8", "Square", "Reciprocating", "Rectangle", "Bounce", "M", "I", "T"]


def compare_streams(ref, gpu, what):
    assert ref is not None and gpu is not None, what
    assert ref.shape == gpu.shape, (what, ref.shape, gpu.shape)
    np.testing.assert_array_equal(ref[:, 0], gpu[:, 0], err_msg=f"{what}: ticks")
    np.testing.assert_array_equal(ref[:, 15:], gpu[:, 15:], err_msg=f"{what}: power / modes")
    return assert_samples_close(gpu[:, 1:15].T, ref[:, 1:15].T, what)


@needs_nodes
@pytest.mark.parametrize("traj_type", ALL_TYPES)
def test_unmodified_node_with_dropin_classes_full_mission(traj_type):
    """take off -> go to the start -> follow -> (trajectory ends, hover) -> END: go home and land."""
    cfg = {"traj_type": traj_type, "t_traj": 9.0, "orientation": 0.25, "center_x": 0.3, "center_y": -0.2}
    ref_node, gpu_node = node_lib.Node(node_lib.NODE_REF_SO), node_lib.Node(node_lib.NODE_GPU_SO)
    ev = [(T_GO1, node_lib.GO), (T_GO2, node_lib.GO), (T_GO3, node_lib.GO), (7000, node_lib.LAND)]
    start = (0.4, -0.6, 0.0, 0.2)
    ref = ref_node.run(cfg, ev, 10500, start)
    gpu = gpu_node.run(cfg, ev, 10500, start)
    e = compare_streams(ref, gpu, f"node mission {traj_type}")
    assert ref[-1, 15] == 0, "the mission must end on the ground with the motors off"
    assert e["pos_abs"] < 1e-9


@needs_nodes
@pytest.mark.parametrize("traj_type", ["Circle", "Figure8", "Line", "Boomerang", "T", "Bounce", "Reciprocating"])
def test_unmodified_node_with_dropin_classes_end_pressed_while_following(traj_type):
    """END while following: the node replaces its goal vector by generateStopTraj's (TrajectoryGenerator.cpp:514-517).
    The braking count depends on the bits of the sample being braked from, which the drop-in reproduces within
    tolerance only — so the braking lengths may differ by one step; the streams are compared up to the shorter one and
    must agree again once both hover."""
    cfg = {"traj_type": traj_type, "t_traj": 9.0}
    ref_node, gpu_node = node_lib.Node(node_lib.NODE_REF_SO), node_lib.Node(node_lib.NODE_GPU_SO)
    t_end = T_GO3 + 420
    ev = [(T_GO1, node_lib.GO), (T_GO2, node_lib.GO), (T_GO3, node_lib.GO), (t_end, node_lib.LAND)]
    ref = ref_node.run(cfg, ev, t_end + 900)
    gpu = gpu_node.run(cfg, ev, t_end + 900)
    assert ref is not None and gpu is not None
    r_before, g_before = ref[ref[:, 0] < t_end], gpu[gpu[:, 0] < t_end]
    compare_streams(r_before, g_before, f"{traj_type} before END")

    def braking(rows):          # ticks from END until the speed reaches zero for good (the hover goal)
        after = rows[rows[:, 0] >= t_end]
        moving = np.nonzero(np.abs(after[:, 4:7]).sum(axis=1) > 0)[0]
        return after, (moving[-1] + 1 if len(moving) else 0)

    r_after, r_n = braking(ref)
    g_after, g_n = braking(gpu)
    assert abs(r_n - g_n) <= 1, (traj_type, r_n, g_n)
    m = min(r_n, g_n)
    assert m > 10
    assert_samples_close(g_after[:m, 1:15].T, r_after[:m, 1:15].T, f"{traj_type} braking")
    # both end hovering at (almost) the same place
    assert np.abs(ref[-1, 1:4] - gpu[-1, 1:4]).max() < 1e-6 and (ref[-1, 4:13] == 0).all() and (gpu[-1, 4:13] == 0).all()


@needs_nodes
def test_node_refuses_the_same_configurations():
    """readParameters() fails for the same parameter files on both builds when the shape does not fit the room
    (trajectoryInsideBounds, TrajectoryGenerator.cpp:419-422).  (Parameter files the node rejects BEFORE it has
    constructed traj_ — accel <= 0, an unknown traj_type — make the reference dereference a null traj_ at :71, and a Line
    too short for its speed runs into Line.cpp:76-79's exit(1): those end the process on both builds and are not
    flown here.)"""
    ref_node, gpu_node = node_lib.Node(node_lib.NODE_REF_SO), node_lib.Node(node_lib.NODE_GPU_SO)
    tight = {"x_min": -1.0, "x_max": 1.0, "y_min": -1.0, "y_max": 1.0}
    for cfg in ({"traj_type": "Circle", **tight}, {"traj_type": "T", **tight},
                {"traj_type": "Square", "side_length": 3.0, **tight}, {"traj_type": "Bounce", "z_max": 3.0},
                {"traj_type": "Reciprocating", **tight}, {"traj_type": "M", "orientation": 0.7, "x_max": 2.2}):
        assert ref_node.run(cfg, [], 3) is None, cfg
        assert gpu_node.run(cfg, [], 3) is None, cfg
    ok = {"traj_type": "Square", "side_length": 1.5, **tight}
    assert ref_node.run(ok, [], 3) is not None and gpu_node.run(ok, [], 3) is not None


# ---- node-side transitions ------------------------------------------------------------------

def random_transitions(n, seed=11):
    rng = np.random.default_rng(seed)
    t = np.zeros(n, dtype=abi.TRANSITION_DTYPE)
    kind = rng.integers(0, 3, n)
    t["kind"] = kind
    t["dt"] = rng.choice([0.01, 0.02, 0.005], n)
    t["start"] = rng.uniform(-4, 4, (n, 3))
    t["start"][:, 2] = np.where(kind == abi.TR_TAKEOFF, rng.uniform(0.0, 0.3, n), rng.uniform(1.0, 2.5, n))
    t["start_v"] = np.where(rng.random((n, 2)) < 0.5, 0.0, rng.uniform(-0.5, 0.5, (n, 2)))
    t["start_psi"] = rng.uniform(-3.5, 3.5, n)
    t["dest"] = rng.uniform(-4, 4, (n, 3))
    t["dest"][:, 2] = np.where(kind == abi.TR_LANDING, rng.uniform(0.0, 0.2, n), rng.uniform(1.0, 2.5, n))
    t["dest_yaw"] = rng.uniform(-3.5, 3.5, n)
    t["vel"] = np.where(kind == abi.TR_GOTO, rng.uniform(0.2, 1.0, n), rng.uniform(0.2, 0.5, n))
    t["vel_yaw"] = np.where(kind == abi.TR_LANDING, rng.uniform(0.03, 0.1, n), rng.uniform(0.1, 0.5, n))
    t["dist_thresh"] = rng.uniform(0.05, 0.4, n)
    t["yaw_thresh"] = rng.uniform(0.05, 0.3, n)
    t["ticks"] = np.where((kind == abi.TR_GOTO) & (rng.random(n) < 0.3), rng.integers(1, 3000, n), 0)
    return t


def test_transitions_are_bit_exact(engine, oracle):
    """tgx_transitions against the oracle restatement that tests/test_node_oracle.py pins to the unmodified node."""
    t = random_transitions(600)
    # some rejected records and one that cannot end within the guard
    t["dt"][5] = 0.0
    t["vel"][6] = -1.0
    t["kind"][7] = 9
    # destinations beyond the walls can never be reached: the vehicle pushes against the wall until the guard fires
    box = (-3.5, 3.5, -3.0, 3.8, 0.0, 2.2)
    lim = abi.make_limits(box=box)
    guard = 6000
    engine.set_max_samples(guard)
    try:
        _, counts, status = engine.transitions_host(t, 0, lim)
        cap = int(counts.max())
        rec, counts2, status2 = engine.transitions_host(t, cap, lim)
    finally:
        engine.set_max_samples(abi.DEFAULT_MAX_SAMPLES)
    np.testing.assert_array_equal(counts, counts2)
    assert cap == guard and (status2 & abi.ST_TOO_LONG).any()
    kinds_seen = set()
    for i in range(len(t)):
        want, st = oracle.transition(t[i:i + 1], traj=i, box=box, max_samples=guard)
        assert counts[i] == len(want) and status2[i] == st, (i, counts[i], len(want), status2[i], st)
        assert rec[i, :len(want)].tobytes() == want.tobytes(), i
        assert not rec[i, len(want):].view(np.uint8).any()
        if len(want):
            kinds_seen.add(int(t["kind"][i]))
    assert kinds_seen == {0, 1, 2} and status2[5] == abi.ST_BAD_PARAM and status2[7] == abi.ST_BAD_PARAM
    # capacity below the tick count: the head is written, the status says so
    engine.set_max_samples(guard)
    try:
        rec, c3, s3 = engine.transitions_host(t[:50], 64, lim)
        rec4, c4, s4 = engine.transitions_host(t[:40], cap, None)
    finally:
        engine.set_max_samples(abi.DEFAULT_MAX_SAMPLES)
    for i in range(50):
        want, st = oracle.transition(t[i:i + 1], traj=i, box=box, max_samples=guard)
        m = min(64, len(want))
        assert rec[i, :m].tobytes() == want[:m].tobytes()
        assert bool(s3[i] & abi.ST_TRUNCATED) == (len(want) > 64)
    # no box: nothing is saturated
    for i in range(40):
        want, st = oracle.transition(t[i:i + 1], traj=i, box=None, max_samples=guard)
        assert rec4[i, :len(want)].tobytes() == want.tobytes() and (want["clamped"] == 0).all()


def test_transitions_large_fleet(engine, oracle):
    """A fleet of several thousand vehicles: every vehicle's records land in its own row, bit for bit."""
    t = workloads.fleet_transitions(6000)
    box = (-5.0, 5.0, -5.0, 5.0, 0.0, 5.0)
    lim = abi.make_limits(box=box)
    _, counts, status = engine.transitions_host(t, 0, lim)
    cap = int((counts.max() + 3) // 4 * 4)
    rec, counts2, status2 = engine.transitions_host(t, cap, lim)
    np.testing.assert_array_equal(counts, counts2)
    assert (status2 == 0).all()
    for i in list(range(0, 6000, 61)) + [5999]:
        want, st = oracle.transition(t[i:i + 1], traj=i, box=box)
        assert counts2[i] == len(want) and rec[i, :len(want)].tobytes() == want.tobytes(), i
        assert not rec[i, len(want):].view(np.uint8).any()


@needs_nodes
def test_transitions_reproduce_the_node_mission(engine, oracle):
    """The whole mission of the unmodified node outside TRAJ_FOLLOWING, from the GPU: take-off, the trip to the start
    of the trajectory, the trip home and the landing, bit for bit."""
    from test_node_oracle import mission_transitions, params_for, oracle_samples, transition
    ref_node = node_lib.Node(node_lib.NODE_REF_SO)
    start_pose = (0.7, -1.1, 0.0, 0.4)
    cfg = {"traj_type": "Figure8", "t_traj": 4.0, "z_min": 0.0}
    s = oracle_samples(oracle, params_for("Figure8", cfg))
    n = s.shape[1]
    t_land = T_GO3 + n + 60
    ev = [(T_GO1, node_lib.GO), (T_GO2, node_lib.GO), (T_GO3, node_lib.GO), (t_land, node_lib.LAND)]
    rows = ref_node.run(cfg, ev, t_land + 6000, start=start_pose)
    takeoff, goto, home, c = mission_transitions(rows, cfg, start_pose, s[:, 0], T_GO1, T_GO2, T_GO3, t_land)
    lim = abi.make_limits(box=[c[k] for k in ("x_min", "x_max", "y_min", "y_max", "z_min", "z_max")])
    batch = np.concatenate([takeoff, goto, home])
    rec, counts, status = engine.transitions_host(batch, 4096, lim)
    assert (status == 0).all()

    def check(r, t0, what):
        got = rows[(rows[:, 0] >= t0) & (rows[:, 0] < t0 + len(r))]
        assert len(got) == len(r) and golden_util.same_bits(got[:, 1:15].T, abi.records_to_channels(r)), what
        np.testing.assert_array_equal(got[:, 15], r["power"])

    check(rec[0, :counts[0]], T_GO1, "take-off")
    check(rec[1, :counts[1]], T_GO2, "go to start")
    check(rec[2, :counts[2]], t_land, "go home")
    last = rec[2, counts[2] - 1]
    landing = transition(abi.TR_LANDING, 0.01, last["p"], last["v"][:2], last["psi"], [0, 0, start_pose[2]], 0.0,
                         c["vel_land_fast"], c["vel_land_slow"])
    rec2, counts2, status2 = engine.transitions_host(landing, 4096, lim)
    check(rec2[0, :counts2[0]], t_land + counts[2], "landing")
    assert rec2[0, counts2[0] - 1]["power"] == 0


def test_plane_major_host_layout(engine, oracle):
    """tgx_set_host_layout(1): the host buffer is [14][n][capacity]; same values as the default layout, with and
    without the host-side fill of the constant planes, across chunk boundaries."""
    params = abi.concat([workloads.circles_cfg2(700), workloads.mixed_cfg3(300)])
    cap = 1536
    ref_out, ref_counts, ref_status, _ = engine.generate_host(params, cap)
    try:
        for fill in (True, False):
            engine.set_host_fill(fill)
            engine.set_host_layout(True)
            out, counts, status, _ = engine.generate_host(params, cap)
            assert out.shape == (abi.TGX_NCHAN, len(params), cap)
            np.testing.assert_array_equal(counts, ref_counts)
            np.testing.assert_array_equal(status, ref_status)
            for i in range(0, len(params), 37):
                m = counts[i]
                assert np.array_equal(out[:, i, :m], ref_out[i, :, :m]), (fill, i)
    finally:
        engine.set_host_fill(True)
        engine.set_host_layout(False)


@pytest.mark.parametrize("tuning", [(10, 4), (9, 4), (9, 2)])
def test_eval_records_equals_eval_then_pack(engine, oracle, tuning):
    """The fused kernel (records straight from the evaluation kernels through the swizzled shared-memory transpose) is
    byte-identical to tgx_eval followed by tgx_pack_goals, for both families, braking plans, packed offsets and a record
    capacity below the sample count, in every kernel shape."""
    import torch
    engine.set_tuning(*tuning)
    try:
        lim = abi.make_limits(box=BOX)
        batches = {
            "classic": (abi.concat([workloads.mixed_cfg3(90, seed=9), workloads.default_circle()]), engine.plan),
            "polyline": (engine.finalize_polyline(abi.concat([workloads.polyline_mix(90, seed=10),
                                                               workloads.default_polyline(abi.TGX_I)]).copy()),
                         engine.plan_polyline),
        }
        for name, (params, planner) in batches.items():
            d = engine.upload_params(params)
            plan = planner(d)
            counts = plan.counts
            n = len(params)
            cap = int((int(counts.max()) + 3) // 4 * 4)
            planes = torch.full((n, abi.TGX_NCHAN, cap), float("nan"), dtype=torch.float64, device=d.device)
            engine.eval(planes)
            want = engine.pack_goals(planes, counts, lim)
            got = engine.eval_records(n, cap, lim)
            torch.cuda.synchronize()
            assert torch.equal(got, want), (name, tuning)
            # no box, smaller capacity
            want = engine.pack_goals(planes, counts, None, rec_capacity=700)
            got = engine.eval_records(n, 700, None)
            # the truncated trajectory's `last` flag: pack sees the full count too, so both agree
            assert torch.equal(got, want), (name, tuning, "cap 700")
            # densely packed rows
            c = counts.cpu().numpy()
            offs = torch.from_numpy(np.concatenate([[0], np.cumsum(c[:-1], dtype=np.int64)])).to(d.device)
            flat_w = torch.zeros((int(c.sum()), 128), dtype=torch.uint8, device=d.device)
            flat_g = torch.zeros_like(flat_w)
            engine.pack_goals(planes, counts, lim, records=flat_w, rec_capacity=cap, rec_offset=offs)
            engine.eval_records(n, cap, lim, records=flat_g, rec_offset=offs)
            torch.cuda.synchronize()
            assert torch.equal(flat_g, flat_w), (name, tuning, "packed")
            # offsets that point outside the buffer (negative, beyond the end, 2^32 + a valid row: the TMA row coordinate
            # is 32 bits and must not wrap into the buffer) write nothing; a row that runs past the end is cut there
            total = int(c.sum())
            guard = 4096
            big = torch.full((total + guard, 128), 0xab, dtype=torch.uint8, device=d.device)
            bad = offs.clone()
            bad[1] = -5
            bad[2] = total + 7
            bad[3] = (1 << 32) + int(offs[3])
            bad[4] = total - 10                                   # the last 10 records of the buffer, then the end
            bad[n - 1] = -1                                       # (so that nothing else writes those 10 records)
            big[:total] = 0
            engine.eval_records(n, cap, lim, records=big[:total], rec_offset=bad)
            torch.cuda.synchronize()
            assert bool((big[total:] == 0xab).all()), (name, tuning, "wrote past the end of the record buffer")
            # every other trajectory wrote its usual rows, the five displaced ones left theirs untouched
            keep = torch.ones(total, dtype=torch.bool, device=d.device)
            for i in (1, 2, 3, 4, n - 1):
                keep[int(offs[i]):int(offs[i]) + int(c[i])] = False
                assert bool((big[int(offs[i]):min(int(offs[i]) + int(c[i]), total - 10)] == 0).all()), (name, tuning, i)
            assert torch.equal(big[:total][keep], flat_w[keep]), (name, tuning, "bad offsets disturbed other rows")
            assert int(c[4]) > 10 and int(c[n - 1]) > 10
            assert torch.equal(big[total - 10:total], flat_w[int(offs[4]):int(offs[4]) + 10]), (name, tuning, "cut row")
        # braking plans (segment tables incl. the frozen-position records of the polyline family)
        params = abi.concat([workloads.mixed_cfg3(40, seed=12), batches["polyline"][0][:40]])
        froms = np.zeros((len(params), abi.TGX_NCHAN))
        for i in range(len(params)):
            p = params[i:i + 1]
            smp = oracle.polyline_generate(p)[0] if abi.is_polyline(int(p["type"][0])) else oracle.generate(p)[0]
            froms[i] = smp[:, smp.shape[1] // 2]
        d = engine.upload_params(params)
        plan = engine.plan_stop(d, torch.from_numpy(froms).to(d.device))
        cap = max(4, int((int(plan.counts.max()) + 3) // 4 * 4))
        planes = torch.full((len(params), abi.TGX_NCHAN, cap), float("nan"), dtype=torch.float64, device=d.device)
        engine.eval(planes)
        want = engine.pack_goals(planes, plan.counts, lim)
        got = engine.eval_records(len(params), cap, lim)
        torch.cuda.synchronize()
        assert torch.equal(got, want), ("stop", tuning)
    finally:
        engine.set_tuning(10, 4)
