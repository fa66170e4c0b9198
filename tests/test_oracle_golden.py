"""CPU: the plain-C oracle against the committed golden vectors of the unmodified reference, and against the
known answers the survey recorded from the reference.  Bit-exact."""
import numpy as np
import pytest

import golden_util
from trajectory_generator_ros2_b200 import abi, workloads

GOLDEN = golden_util.load()


@pytest.mark.parametrize("case", GOLDEN["cases"], ids=[c["name"] for c in GOLDEN["cases"]])
def test_oracle_matches_reference_golden(oracle, case):
    p = case["params"]
    s, st, ph = oracle.generate(p)
    assert st == case["status"]
    if not st & abi.ST_FATAL_MASK:
        assert s.shape[1] == case["n"]
    if st & abi.ST_FATAL_MASK:
        # the reference calls exit(1) at this point (Line.cpp:76-79): what it had produced up to then is moot; the
        # engine and the oracle define the outcome as "all N samples, last one forced to B, status bit set"
        return
    assert abi.phases_to_index_msgs(case["type"], ph) == case["msgs"]
    assert f"{oracle.fnv(s):016x}" == case["fnv1a64"]
    for k, vals in case["sample_values"].items():
        assert golden_util.same_bits(s[:, k], vals), (case["name"], k)
    v = np.sqrt((s[abi.VX:abi.VZ + 1] ** 2).sum(0)).max()
    a = np.sqrt((s[abi.AX:abi.AZ + 1] ** 2).sum(0)).max()
    assert v.hex() == case["max_v"] and a.hex() == case["max_a"]
    assert oracle.inside_bounds(p, workloads.MONTECARLO_LIMITS["box"]) == case["inside_bounds"]
    g = case["stop"]
    ss, sst, sph = oracle.stop(p, s[:, g["from_k"]])
    assert ss.shape[1] == g["n"]
    assert f"{oracle.fnv(ss):016x}" == g["fnv1a64"]
    assert abi.phases_to_index_msgs(case["type"], sph, stop_traj=True) == g["msgs"]
    if g["n"]:
        assert golden_util.same_bits(ss[:, -1], g["last_values"])


def test_survey_known_answers_circle(oracle):
    """SURVEY.md §8c, KAT Circle (default.yaml): values the survey probe read off the reference, 17 digits."""
    s, st, ph = oracle.generate(workloads.default_circle())
    assert s.shape[1] == 25001 and st == 0
    assert sorted(abi.phases_to_index_msgs(0, ph)) == [0, 250, 8250, 8500, 16500, 24500, 25000]
    k1 = [3.3999999997647055, 3.9999999999077279e-05, 1.8, -4.7058823528326211e-08, 0.0039999999997231833, 0,
          -4.7058823526155088e-06, -5.5363321798030833e-11, 0, 6.513331976238923e-14, -5.536332179547659e-09, 0,
          1.5708080915007789, 0.0011764705882352942]
    assert golden_util.same_bits(s[:, 1], k1)
    k12500 = [-0.54951194115888291, -3.3552997819157375, 1.8, 1.9737057540680809, -0.32324231832875466, 0,
              0.19014254019338511, 1.16100338474593, 0, -0.68294316749760586, 0.11184855305493242, 0,
              50.103149267983106, 0.58823529411764708]
    assert golden_util.same_bits(s[:, 12500], k12500)
    assert s[abi.PX, 25000] == 1.2075333488138387 and s[abi.PY, 25000] == 3.1783428404598575
    assert s[abi.PSI, 25000] == 122.15903162087876 and s[abi.DPSI, 25000] == 0
    v = np.sqrt((s[3:6] ** 2).sum(0)); a = np.sqrt((s[6:9] ** 2).sum(0)); j = np.sqrt((s[9:12] ** 2).sum(0))
    assert v.max() == 2.0 and a.max() == 1.1764705882352944 and j.max() == 0.69204152249134965
    ss, _, _ = oracle.stop(workloads.default_circle(), s[:, 12500])
    assert ss.shape[1] == 500


def test_survey_known_answers_figure8_and_line(oracle):
    s, st, ph = oracle.generate(workloads.default_figure8())
    assert s.shape[1] == 25001
    assert golden_util.same_bits(s[[0, 1, 3, 4, 12, 13], 1],
                                 [3.9999999999077279e-05, 3.9999999996309115e-05, 0.0039999999997231833,
                                  0.003999999998892733, 0.78539816329364198, 0.0011764705882352942])
    assert golden_util.same_bits(s[[0, 1, 3, 4, 6, 7, 12], 12500],
                                 [-3.3552997819157375, 0.54228744009720398, -0.32324231832875466, -1.895514403641452,
                                  1.16100338474593, -0.75057085134561108, -1.739701678684624])
    v = np.sqrt((s[3:6] ** 2).sum(0)); a = np.sqrt((s[6:9] ** 2).sum(0))
    assert v.max() == 2.828427099985388 and a.max() == 2.4999999998036522
    assert (s[9:12] == 0).all()
    assert oracle.stop(workloads.default_figure8(), s[:, 12500])[0].shape[1] == 481

    s, st, ph = oracle.generate(workloads.default_line())
    assert s.shape[1] == 685 and st == 0
    msgs = abi.phases_to_index_msgs(1, ph)
    assert sorted(msgs) == [0, 67, 584, 684]
    assert msgs[67] == "Line traj: reached 1.000000 m/s, keeping constant v for 5.166667 s"
    assert golden_util.same_bits(s[[0, 1, 3, 4, 6, 7], 1],
                                 [9.184850993605148e-21, -2.9998499999999999, 9.184850993605148e-19,
                                  0.014999999999999999, 9.1848509936051484e-17, 1.5])
    assert golden_util.same_bits(s[[0, 1, 3, 4], 342],
                                 [1.8930896382919695e-16, 0.091649999999987797, 6.123233995736766e-17, 1])
    assert s[0, 684] == 0 and s[1, 684] == 3 and s[7, 684] == -1
    assert (s[abi.PSI] == 1.5707963267948966).all()
    assert oracle.stop(workloads.default_line(), s[:, 342])[0].shape[1] == 100
    short = abi.line_params(1.8, [0, -3, 1.8], [0, -2.5, 1.8], [1.0], 1.5, 1.0, 0.01)
    assert not oracle.inside_bounds(short, [-5, 5, -5, 5, -5, 5])
    assert oracle.inside_bounds(workloads.default_line(), [-5, 5, -5, 5, -5, 5])


def test_survey_count_traps(oracle):
    """(v 2, a 0.3, t 10) -> ramp 667, hold 1001; (v 3, a 0.1, t 1) -> ramp 3001, hold 100; (v 1, a 0.3, t 5) -> 334, 501."""
    for (v, a, t), (ramp, hold) in (((2.0, 0.3, 10.0), (667, 1001)), ((3.0, 0.1, 1.0), (3001, 100)),
                                    ((1.0, 0.3, 5.0), (334, 501))):
        n, st = oracle.count(abi.circle_params(1.5, 2.0, 0, 0, [v], t, a, 0.01))
        assert n == 1 + ramp + hold + ramp and st == 0
    n, st = oracle.count(abi.circle_params(1.8, 3.4, 0, 0, [2.0, 1.0], 1.0, 0.4, 0.01))
    assert n == 1201 and st == abi.ST_VGOALS_NOT_INCREASING


def test_oracle_rejects_bad_parameters(oracle):
    good = workloads.default_circle()
    for field, val in (("accel", 0.0), ("r", 0.0), ("dt", 0.0), ("n_vgoals", -1), ("n_vgoals", 65), ("type", 5)):
        q = good.copy()
        q[field] = val
        n, st = oracle.count(q)
        assert n == -1 and st == abi.ST_BAD_PARAM
    # nine goal speeds need a continuation record (tgx.h: TGX_VGOALS_MORE): without one the record is rejected
    q = abi.concat([good, good])
    q["n_vgoals"][0] = 9
    counts, status = oracle.count_batch(q)
    assert counts[0] == 0 and status[0] == abi.ST_BAD_PARAM and counts[1] > 0
    # what the reference accepts although it is unusual: an empty vector (the start sample alone), a negative radius
    q = good.copy()
    q["n_vgoals"] = 0
    assert oracle.count(q) == (1, 0)
    q = good.copy()
    q["r"] = -3.4
    assert oracle.count(q) == oracle.count(good)
    n, st = oracle.count(abi.circle_params(1, 1, 0, 0, [1.0], 1e9, 1.0, 0.01), max_samples=10000)
    assert n == -1 and st == abi.ST_TOO_LONG
