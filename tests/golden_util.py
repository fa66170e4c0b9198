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
