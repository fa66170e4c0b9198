#!/usr/bin/env python
"""Generate tests/golden/reference_golden_polyline.json from the UNMODIFIED reference (oracle/_ref/libtrajref.so):
the constant-speed polyline family (Square, Rectangle, Reciprocating, Bounce, M, I, T).

Run in the container that has /root/reference:   make -C oracle && python tests/golden/make_golden_polyline.py
Same conventions as make_golden.py (hex-float doubles, FNV-1a-64 over the 14 channels).  The reference announces every
sample in index_msgs, so the messages are stored run-length encoded: [first sample index, text] per run of equal text.
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
from make_golden import hx, record_params  # noqa: E402


def runs(msgs: dict):
    out, prev = [], None
    for k in sorted(msgs):
        if msgs[k] != prev:
            out.append([int(k), msgs[k]])
            prev = msgs[k]
    return out


def cases():
    out = [(f"default_{abi.TYPE_NAMES[k]}", workloads.default_polyline(k)) for k in abi.POLYLINE_TYPES]
    out.append(("square_rotated_round", abi.square_params(1.8, 2.0, 0.0, 0.0, 0.5, [1.0], 12.0, 0.4, 0.01)))
    out.append(("rectangle_rotated_round", abi.rectangle_params(1.8, 2.0, 4.0, 0.5, -0.5, 1.0, [2.0], 12.0, 0.4, 0.02)))
    out.append(("M_rotated", abi.letter_params(abi.TGX_M, 0.3, -0.2, 3.0, 4.0, 1.8, [1.0], 30.0, np.pi / 4, 0.01)))
    out.append(("I_rotated", abi.letter_params(abi.TGX_I, 1.0, 1.0, 3.0, 4.0, 1.8, [0.5], 30.0, -2.0, 0.01)))
    out.append(("T_rotated", abi.letter_params(abi.TGX_T, 0.0, 0.0, 3.0, 4.0, 1.8, [2.0], 30.0, 3.0, 0.01)))
    out.append(("bounce_up_first", abi.bounce_params(0.5, -0.5, 1.0, 2.5, [0.7], 9.0, 0.3, 0.01)))
    out.append(("reciprocating_diagonal", abi.reciprocating_params(1.2, [-1, -2, 1.2], [2.5, 1.0, 1.2], [1.3], 1.5, 0.8, 10.0, 0.01)))
    out.append(("reciprocating_cut_leg", abi.reciprocating_params(1.8, [0, 0, 1.8], [1, 0, 1.8], [1.0], 1.0, 1.0, 1.01, 0.01)))
    out.append(("square_t_traj_zero", abi.square_params(1.8, 2.0, 0, 0, 0.0, [1.0], 0.0, 0.4, 0.01)))
    mix = workloads.polyline_mix(28)
    for i in range(28):
        out.append((f"polyline_mix_{i}", mix[i:i + 1].copy()))
    return out


def main():
    assert Reference.available(), "build oracle/_ref first (make -C oracle)"
    ref, orc = Reference(), Oracle()
    box = workloads.MONTECARLO_LIMITS["box"]
    golden = {"generator": "tests/golden/make_golden_polyline.py",
              "source": "oracle/_ref (unmodified reference sources)",
              "toolchain": "g++ 13.3 -std=c++17 -O2 -ffp-contract=off, glibc 2.39", "cases": []}
    for name, p in cases():
        s, st, msgs = ref.generate(p)
        n = s.shape[1]
        r = runs(msgs)
        # samples at the start, the end, and either side of a few leg boundaries
        ks = {0, 1, 2, n // 3, n // 2, n - 2, n - 1}
        for k, _ in r[:12]:
            ks |= {k - 1, k, k + 1}
        ks = sorted(ks & set(range(n)))
        k_stop = n // 3 if n else 0
        if n:
            ss, sst, smsgs = ref.stop(p, s[:, k_stop])
        else:
            ss, smsgs = np.zeros((abi.TGX_NCHAN, 0)), {}
        golden["cases"].append({
            "name": name, "family": "polyline", "params_hex": record_params(p), "type": int(p["type"][0]),
            "n": int(n), "status": int(st), "n_msgs": len(msgs), "msg_runs": r,
            "fnv1a64": f"{orc.fnv(s):016x}",
            "samples": {str(k): hx(s[:, k]) for k in ks},
            "inside_bounds": bool(ref.inside_bounds(p, box)),
            "stop": {"from_k": int(k_stop), "n": int(ss.shape[1]), "fnv1a64": f"{orc.fnv(ss):016x}",
                     "index_msgs": {str(k): m for k, m in sorted(smsgs.items())},
                     "last": hx(ss[:, -1]) if ss.shape[1] else []},
        })
    path = os.path.join(HERE, "reference_golden_polyline.json")
    with open(path, "w") as f:
        json.dump(golden, f, indent=1)
    print(f"wrote {path}: {len(golden['cases'])} cases, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
