"""ctypes driver for the in-process node harness (oracle/node_shim.cpp): the UNMODIFIED reference node behind a fake
rclcpp::Node, either with the reference's own trajectory classes (oracle/_ref/libnoderef.so, the oracle) or with this
repo's GPU-backed drop-in classes (tests/cpp/bin/libnodegpu.so).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NODE_REF_SO = os.path.join(ROOT, "oracle", "_ref", "libnoderef.so")
NODE_GPU_SO = os.path.join(ROOT, "tests", "cpp", "bin", "libnodegpu.so")

GO, LAND, KILL = 4, 2, 6      # QuadFlightMode (TrajectoryGenerator.cpp:437-440)

# config/default.yaml (:4-72) of the reference, key for key
DEFAULT_YAML = {
    "alt": 1.8, "pub_freq": 100.0, "traj_type": "T",
    "T_length": 3.0, "T_width": 4.0, "I_length": 3.0, "I_width": 4.0, "M_length": 3.0, "M_width": 4.0,
    "Az": 4.0, "Bz": 1.0,
    "side_length": 2.0, "square_accel": 0.4, "orientation": 0.0,
    "side_a": 2.0, "side_b": 4.0, "rectangle_accel": 0.4,
    "r": 3.4, "center_x": 0.0, "center_y": 0.0, "v_goals": [1.0, 2.0, 2.0], "t_traj": 80.0, "circle_accel": 0.4,
    "Ax": 0.0, "Ay": -3.0, "Bx": 0.0, "By": 3.0, "v_line": 1.0, "line_accel": 1.5, "line_decel": 1.0,
    "vel_initpos": 0.4, "vel_take": 0.3, "vel_land_fast": 0.35, "vel_land_slow": 0.04, "vel_yaw": 0.2,
    "dist_thresh": 0.3, "yaw_thresh": 0.2, "margin_takeoff_outside_bounds": 0.05,
    "x_min": -5.0, "x_max": 5.0, "y_min": -5.0, "y_max": 5.0, "z_min": -5.0, "z_max": 5.0,
}

ROW = 18   # tick, 14 channels, power, mode_xy, mode_z


def config_text(overrides=None) -> bytes:
    cfg = dict(DEFAULT_YAML)
    cfg.update(overrides or {})
    lines = []
    for k, v in cfg.items():
        if isinstance(v, (list, tuple)):
            lines.append(f"{k}=" + ",".join(repr(float(x)) for x in v))
        elif isinstance(v, str):
            lines.append(f"{k}={v}")
        else:
            lines.append(f"{k}={float(v)!r}")
    return ("\n".join(lines) + "\n").encode()


class Node:
    def __init__(self, path: str):
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL)   # both flavours define the same C++ symbols: keep them apart
        self.lib.node_run.restype = C.c_int64
        self.lib.node_run.argtypes = [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.c_int32, C.c_int64,
                                      C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int64]

    def run(self, overrides, events, n_ticks, start=(0.0, 0.0, 0.0, 0.0)):
        """events: [(tick, mode)].  -> rows [m, 18] of every published Goal, or None if the node refused to start."""
        ticks = np.array([e[0] for e in events], dtype=np.int32)
        modes = np.array([e[1] for e in events], dtype=np.uint8)
        st = np.array(start, dtype=np.float64)
        cap = n_ticks + 2 * len(events) + 8
        out = np.zeros((cap, ROW))
        n = self.lib.node_run(config_text(overrides), ticks.ctypes.data_as(C.POINTER(C.c_int32)),
                              modes.ctypes.data_as(C.POINTER(C.c_uint8)), len(events), n_ticks,
                              st.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)), cap)
        if n < 0:
            return None
        assert n <= cap
        return out[:n]
