"""Load tests/golden/reference_golden.json (outputs of the unmodified reference, see make_golden.py)."""
import json
import os

import numpy as np

from trajectory_generator_ros2_b200 import abi

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.json")


def load():
    with open(PATH) as f:
        g = json.load(f)
    for c in g["cases"]:
        c["params"] = np.frombuffer(bytes.fromhex(c["params_hex"]), dtype=abi.PARAMS_DTYPE).copy()
        c["sample_values"] = {int(k): np.array([float.fromhex(x) for x in v]) for k, v in c["samples"].items()}
        c["msgs"] = {int(k): m for k, m in c["index_msgs"].items()}
        c["stop"]["msgs"] = {int(k): m for k, m in c["stop"]["index_msgs"].items()}
        c["stop"]["last_values"] = np.array([float.fromhex(x) for x in c["stop"]["last"]])
    return g


def same_bits(a, b) -> bool:
    """Bit equality with -0.0 == +0.0 (the checksum canonicalises the sign of zero the same way)."""
    a = np.asarray(a, dtype=np.float64) + 0.0
    b = np.asarray(b, dtype=np.float64) + 0.0
    return bool((a.view(np.uint64) == b.view(np.uint64)).all())


POLY_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden_polyline.json")


def expand_runs(runs, n_msgs):
    """[[first index, text], ...] -> {index: text} for indices 0 .. n_msgs-1 (see make_golden_polyline.py)."""
    out = {}
    for i, (k, m) in enumerate(runs):
        end = runs[i + 1][0] if i + 1 < len(runs) else n_msgs
        for q in range(k, end):
            out[q] = m
    return out


def load_polyline():
    with open(POLY_PATH) as f:
        g = json.load(f)
    for c in g["cases"]:
        c["params"] = np.frombuffer(bytes.fromhex(c["params_hex"]), dtype=abi.PARAMS_DTYPE).copy()
        c["sample_values"] = {int(k): np.array([float.fromhex(x) for x in v]) for k, v in c["samples"].items()}
        if c["n"] == 0 and c["n_msgs"]:
            c["msgs"] = {int(k): m for k, m in c["msg_runs"]}      # index_msgs[size() - 1] on an empty vector: key -1
        else:
            c["msgs"] = expand_runs(c["msg_runs"], c["n_msgs"])
        c["stop"]["msgs"] = {int(k): m for k, m in c["stop"]["index_msgs"].items()}
        c["stop"]["last_values"] = np.array([float.fromhex(x) for x in c["stop"]["last"]])
    return g
