"""CPU: the plain-C oracle against the UNMODIFIED reference compiled behind stub headers (oracle/_ref), bit for bit,
on random parameters.  Skipped where oracle/_ref is not built (no /root/reference at build time)."""
import numpy as np
import pytest

from trajectory_generator_ros2_b200 import abi, workloads


@pytest.mark.parametrize("name,n", [("circles_cfg2", 400), ("mixed_cfg3", 900), ("montecarlo_cfg4", 300)])
def test_batches_bit_exact(oracle, reference, name, n):
    params = getattr(workloads, name)(n)
    o_out, o_counts, o_status = oracle.generate_batch(params, 2400)
    r_out, r_counts, r_status = reference.generate_batch(params, 2400)
    np.testing.assert_array_equal(o_counts, r_counts)
    np.testing.assert_array_equal(o_status, r_status)
    assert (o_counts <= 2400).all()
    ob, rb = (o_out + 0.0).view(np.uint64), (r_out + 0.0).view(np.uint64)
    assert (ob == rb).all()


def test_index_msgs_and_stop_bit_exact(oracle, reference):
    params = workloads.mixed_cfg3(120)
    rng = np.random.default_rng(7)
    for i in range(len(params)):
        p = params[i:i + 1]
        o, ost, oph = oracle.generate(p)
        r, rst, rmsgs = reference.generate(p)
        assert ost == rst
        if ost & abi.ST_FATAL_MASK:
            continue
        assert o.shape == r.shape
        t = int(p["type"][0])
        assert abi.phases_to_index_msgs(t, oph) == rmsgs
        k = int(rng.integers(0, o.shape[1]))
        so, sost, soph = oracle.stop(p, o[:, k])
        sr, srst, srmsgs = reference.stop(p, r[:, k])
        assert so.shape == sr.shape
        assert ((so + 0.0).view(np.uint64) == (sr + 0.0).view(np.uint64)).all()
        assert abi.phases_to_index_msgs(t, soph, stop_traj=True) == srmsgs


def test_vgoals_of_any_length_bit_exact(oracle, reference):
    """The reference loops over a std::vector of any length (Circle.cpp:43, Figure8.cpp:43) and accepts any non-zero
    radius: more than 8 goal speeds (continuation records, tgx.h TGX_VGOALS_MORE), none at all, r < 0."""
    rng = np.random.default_rng(5)
    for K in (0, 1, 8, 9, 12, 16, 17, 30, 64):
        for kind in (abi.TGX_CIRCLE, abi.TGX_FIGURE8):
            v = list(np.sort(rng.uniform(0.3, 2.8, K)))
            if K == 12:
                v[5] = v[4] * 0.5                           # a goal below the current speed: warning, skipped ramp
            r = rng.uniform(0.8, 3.0) * (-1.0 if K in (1, 12) else 1.0)
            p = abi.circle_params(1.5, r, 0.3, -0.2, v, rng.uniform(0.05, 0.6), 1.3, 0.01, kind=kind)
            assert len(p) == abi.orbit_records(K)
            o, ost, oph = oracle.generate(p)
            rr, rst, rmsgs = reference.generate(p)
            assert ost == rst and o.shape == rr.shape and o.shape[1] >= 1
            assert ((o + 0.0).view(np.uint64) == (rr + 0.0).view(np.uint64)).all()
            assert abi.phases_to_index_msgs(kind, oph) == rmsgs
            assert len(rmsgs) == (2 * K + 2 if K else 1) - (1 if K == 12 else 0)     # a skipped ramp shares its key
            k = o.shape[1] // 2
            so, _, soph = oracle.stop(p, o[:, k])
            sr, _, srmsgs = reference.stop(p, rr[:, k])
            assert so.shape == sr.shape and ((so + 0.0).view(np.uint64) == (sr + 0.0).view(np.uint64)).all()
            assert abi.phases_to_index_msgs(kind, soph, stop_traj=True) == srmsgs


def test_line_edge_cases_bit_exact(oracle, reference):
    """The Line shapes the phase-record planner has special segments for (tests/test_gpu_parity.py holds the GPU to the
    oracle on the same ones): no cruise at all (d2 < 0: the trajectory ends short of B and the node would exit, status
    LINE_END_NOT_B), a ramp-up or a ramp-down of a single clamped step, a cruise longer than one rebase interval, A == B."""
    cases = [
        abi.line_params(1.0, [0, 0, 1], [0.3, 0.1, 1], [2.0], 1.0, 1.0, 0.01),
        abi.line_params(1.0, [0, 0, 1], [3, 4, 1], [1.0], 200.0, 1.0, 0.01),
        abi.line_params(1.0, [0, 0, 1], [3, 4, 1], [1.0], 1.0, 200.0, 0.01),
        abi.line_params(1.0, [1, 1, 1], [-20, 1, 1], [1.0], 0.7, 0.9, 0.01),
        abi.line_params(0.5, [2, -1, 0.5], [2, -1, 0.5], [1.0], 1.0, 1.0, 0.01),
    ]
    rng = np.random.default_rng(23)
    for _ in range(40):                                  # any direction, length, speed and pair of accelerations
        A = rng.uniform(-4, 4, 3)
        B = A + rng.uniform(-6, 6, 3)
        cases.append(abi.line_params(rng.uniform(0.5, 2), A, B, [rng.uniform(0.3, 2.5)], rng.uniform(0.3, 3),
                                     rng.uniform(0.3, 3), 0.01))
    for p in cases:
        o, ost, oph = oracle.generate(p)
        r, rst, rmsgs = reference.generate(p)
        assert ost == rst
        assert o.shape == r.shape and o.shape[1] >= 2
        if ost & abi.ST_FATAL_MASK:
            # the reference exit(1)s at Line.cpp:71-79, before it would overwrite the last sample with B (:81-82): there is
            # no published trajectory to compare; every sample before the last one still is the reference's
            assert ost & abi.ST_LINE_END_NOT_B
            assert ((o[:, :-1] + 0.0).view(np.uint64) == (r[:, :-1] + 0.0).view(np.uint64)).all()
            continue
        assert ((o + 0.0).view(np.uint64) == (r + 0.0).view(np.uint64)).all()
        assert abi.phases_to_index_msgs(abi.TGX_LINE, oph) == rmsgs


def test_bounds_and_feasibility_match(oracle, reference):
    params = abi.concat([workloads.montecarlo_cfg4(500), workloads.mixed_cfg3(300)])
    lim = abi.make_limits(**workloads.MONTECARLO_LIMITS)
    for i in range(0, len(params), 7):
        assert oracle.inside_bounds(params[i:i + 1], lim.box[:]) == reference.inside_bounds(params[i:i + 1], lim.box[:])
    of = oracle.feasibility_batch(params, lim)
    rf = reference.feasibility_batch(params, lim)
    for a, b in zip(of, rf):
        np.testing.assert_array_equal(a, b)
    assert 0 < of[0].sum() < len(params)


def test_optimisation_level_does_not_change_the_reference(tmp_path, oracle):
    """The reference's CMake sets no optimisation flag; -O0 and -O2 builds of the oracle agree bit for bit."""
    import os
    import subprocess
    from oracle_lib import ORACLE_DIR, Oracle
    so = tmp_path / "liboracle_O0.so"
    subprocess.run(["gcc", "-std=c11", "-O0", "-ffp-contract=off", "-fPIC", "-pthread", "-shared", "-o", str(so),
                    os.path.join(ORACLE_DIR, "traj_oracle.c"), "-lm"], check=True)
    o0 = Oracle(str(so))
    for p in (workloads.default_circle(), workloads.default_figure8(), workloads.default_line()):
        assert o0.fnv(o0.generate(p)[0]) == oracle.fnv(oracle.generate(p)[0])
