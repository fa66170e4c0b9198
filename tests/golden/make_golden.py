#!/usr/bin/env python
"""Generate tests/golden/reference_golden.json from the UNMODIFIED reference (oracle/_ref/libtrajref.so).

Run in the container that has /root/reference:   make -C oracle && python tests/golden/make_golden.py
The reference ships no tests or golden vectors of its own; these vectors are outputs of its own
sources (Circle.cpp, Line.cpp, Figure8.cpp compiled behind the stub headers in compat/ros2_stubs with
-O2 -ffp-contract=off, g++ 13.3, glibc 2.39) and pin the oracle wherever /root/reference is absent (the GPU box).
All doubles are stored as C99 hex-float strings, i.e. bit-exactly.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle_lib import Oracle, Reference  # noqa: E402
from trajectory_generator_ros2_b200 import abi, workloads  # noqa: E402


def hx(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def record_params(p):
    return np.ascontiguousarray(p).view(np.uint8).tobytes().hex()


def cases():
    out = []
    out.append(("default_circle", workloads.default_circle(), 12500))
    out.append(("default_figure8", workloads.default_figure8(), 12500))
    out.append(("default_line", workloads.default_line(), 342))
    out.append(("circle_decreasing_vgoals", abi.circle_params(1.8, 3.4, 0, 0, [2.0, 1.0], 1.0, 0.4, 0.01), 600))
    out.append(("circle_trap_hold_1001", abi.circle_params(1.5, 2.0, 0.1, -0.2, [2.0], 10.0, 0.3, 0.01), 700))
    out.append(("circle_trap_ramp_3001", abi.circle_params(1.5, 2.0, 0.1, -0.2, [3.0], 1.0, 0.1, 0.01), 3050))
    out.append(("circle_trap_hold_501", abi.circle_params(1.5, 2.0, 0.1, -0.2, [1.0], 5.0, 0.3, 0.01), 400))
    out.append(("figure8_five_goals", abi.figure8_params(1.8, 2.0, 1, -1, [0.5, 1.0, 1.5, 2.0, 2.5], 1.5, 1.0, 0.01), 777))
    out.append(("circle_eight_goals", abi.circle_params(1.0, 1.0, 0, 0, [0.3] * 8, 0.25, 2.0, 0.01), 100))
    out.append(("line_d2_negative", abi.line_params(1.8, [0, -3, 1.8], [0, -2.5, 1.8], [1.0], 1.5, 1.0, 0.01), 50))
    out.append(("line_diagonal_fast", abi.line_params(1.0, [-4.25, -3.5, 1.0], [4.5, 4.25, 1.0], [3.0], 1.5, 1.0, 0.01), 300))
    out.append(("line_quadrant3", abi.line_params(1.0, [1, 1, 1.0], [-2, 0.5, 1.0], [0.7], 0.9, 0.6, 0.01), 200))
    out.append(("line_z_differs", abi.line_params(1.0, [0, 0, 0.5], [0, 2, 2.5], [0.5], 1.0, 1.0, 0.02), 100))
    out.append(("default_boomerang", abi.boomerang_params(1.8, [0, -3, 1.8], [0, 3, 1.8], [1.0], 1.5, 1.0, 0.01), 1000))
    out.append(("boomerang_diagonal_fast", abi.boomerang_params(1.0, [-4.25, -3.5, 1.0], [4.5, 4.25, 1.0], [3.0], 1.5, 1.0, 0.01), 900))
    out.append(("boomerang_quadrant3", abi.boomerang_params(1.0, [1, 1, 1.0], [-2, 0.5, 1.0], [0.7], 0.9, 0.6, 0.01), 700))
    out.append(("boomerang_d2_negative", abi.boomerang_params(1.8, [0, -3, 1.8], [0, -2.5, 1.8], [1.0], 1.5, 1.0, 0.01), 50))
    c2 = workloads.circles_cfg2(8)
    for i in range(8):
        out.append((f"cfg2_circle_{i}", c2[i:i + 1].copy(), 500))
    c3 = workloads.mixed_cfg3(24)
    for i in range(24):
        out.append((f"cfg3_mixed_{i}", c3[i:i + 1].copy(), 100))
    c4 = workloads.montecarlo_cfg4(8)
    for i in range(8):
        out.append((f"cfg4_montecarlo_{i}", c4[i:i + 1].copy(), 300))
    return out


def main():
    assert Reference.available(), "build oracle/_ref first (make -C oracle)"
    ref, orc = Reference(), Oracle()
    box = workloads.MONTECARLO_LIMITS["box"]
    golden = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref (unmodified reference sources)",
              "toolchain": "g++ 13.3 -std=c++17 -O2 -ffp-contract=off, glibc 2.39", "cases": []}
    for name, p, k_stop in cases():
        s, st, msgs = ref.generate(p)
        n = s.shape[1]
        ks = sorted({0, 1, n // 3, n // 2, n - 2, n - 1} & set(range(n)))
        v = np.sqrt((s[abi.VX:abi.VZ + 1] ** 2).sum(0))
        a = np.sqrt((s[abi.AX:abi.AZ + 1] ** 2).sum(0))
        k_stop = min(k_stop, n - 1)
        ss, sst, smsgs = ref.stop(p, s[:, k_stop])
        golden["cases"].append({
            "name": name, "params_hex": record_params(p), "type": int(p["type"][0]),
            "n": int(n), "status": int(st), "index_msgs": {str(k): m for k, m in sorted(msgs.items())},
            "fnv1a64": f"{orc.fnv(s):016x}",
            "samples": {str(k): hx(s[:, k]) for k in ks},
            "max_v": float(v.max()).hex(), "max_a": float(a.max()).hex(),
            "inside_bounds": bool(ref.inside_bounds(p, box)),
            "stop": {"from_k": int(k_stop), "n": int(ss.shape[1]), "fnv1a64": f"{orc.fnv(ss):016x}",
                     "index_msgs": {str(k): m for k, m in sorted(smsgs.items())},
                     "last": hx(ss[:, -1]) if ss.shape[1] else []},
        })
    path = os.path.join(HERE, "reference_golden.json")
    with open(path, "w") as f:
        json.dump(golden, f, indent=1)
    print(f"wrote {path}: {len(golden['cases'])} cases, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
