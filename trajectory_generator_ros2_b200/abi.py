"""ctypes / numpy mirror of include/tgx.h (the C-ABI records, enums and status bits).

Nothing here computes anything: it only describes memory layouts so that Python callers (the host-side
classes in ``trajectories.py``, the tests and ``bench.py``) can build ``tgx_params`` arrays and read the
engine's outputs.  Field meanings cite the reference constructors they carry
(Circle.hpp:30-31, Line.hpp:30-31, Figure8.hpp:30-31).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

TGX_MAX_VGOALS = 8
TGX_NCHAN = 14
TGX_MAX_PHASES = 2 * TGX_MAX_VGOALS + 2

# enum tgx_type
TGX_CIRCLE, TGX_LINE, TGX_FIGURE8, TGX_BOOMERANG = 0, 1, 2, 3
# prefix of the index_msgs texts; Boomerang.cpp announces itself as "Line traj: ..." (Boomerang.cpp:44,55,64,94,134)
TYPE_NAMES = {TGX_CIRCLE: "Circle", TGX_LINE: "Line", TGX_FIGURE8: "Figure8", TGX_BOOMERANG: "Line"}

# enum tgx_channel
CHANNELS = ("px", "py", "pz", "vx", "vy", "vz", "ax", "ay", "az", "jx", "jy", "jz", "psi", "dpsi")
(PX, PY, PZ, VX, VY, VZ, AX, AY, AZ, JX, JY, JZ, PSI, DPSI) = range(TGX_NCHAN)

# enum tgx_status_bits
ST_VGOALS_NOT_INCREASING = 1 << 0
ST_FINAL_V_NONZERO = 1 << 1
ST_LINE_END_NOT_B = 1 << 2
ST_LINE_D2_NEGATIVE = 1 << 3
ST_OUTSIDE_BOUNDS = 1 << 4
ST_BAD_PARAM = 1 << 5
ST_VMAX_EXCEEDED = 1 << 6
ST_AMAX_EXCEEDED = 1 << 7
ST_TOO_LONG = 1 << 8
ST_TRUNCATED = 1 << 9
ST_FATAL_MASK = ST_FINAL_V_NONZERO | ST_LINE_END_NOT_B | ST_BAD_PARAM | ST_TOO_LONG
STATUS_NAMES = {
    ST_VGOALS_NOT_INCREASING: "VGOALS_NOT_INCREASING",
    ST_FINAL_V_NONZERO: "FINAL_V_NONZERO",
    ST_LINE_END_NOT_B: "LINE_END_NOT_B",
    ST_LINE_D2_NEGATIVE: "LINE_D2_NEGATIVE",
    ST_OUTSIDE_BOUNDS: "OUTSIDE_BOUNDS",
    ST_BAD_PARAM: "BAD_PARAM",
    ST_VMAX_EXCEEDED: "VMAX_EXCEEDED",
    ST_AMAX_EXCEEDED: "AMAX_EXCEEDED",
    ST_TOO_LONG: "TOO_LONG",
    ST_TRUNCATED: "TRUNCATED",
}

# enum tgx_error
TGX_OK, TGX_ERR_INVALID, TGX_ERR_CUDA, TGX_ERR_ALIGNMENT, TGX_ERR_NO_PLAN, TGX_ERR_NOMEM, TGX_ERR_CAPACITY = range(7)

# enum tgx_phase_kind
PH_ACCEL_TO, PH_REACHED, PH_DECEL, PH_STOPPED, PH_PRESSED_END = range(5)

DEFAULT_MAX_SAMPLES = 1 << 24


class OrbitParams(C.Structure):
    _fields_ = [("r", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("t_traj", C.c_double),
                ("accel", C.c_double), ("v_goals", C.c_double * TGX_MAX_VGOALS)]


class LineParams(C.Structure):
    _fields_ = [("A", C.c_double * 3), ("B", C.c_double * 3), ("a1", C.c_double), ("a3", C.c_double),
                ("v_goal", C.c_double), ("reserved", C.c_double * 4)]


class _ParamsUnion(C.Union):
    _fields_ = [("orbit", OrbitParams), ("line", LineParams)]


class Params(C.Structure):
    """struct tgx_params (128 bytes)."""
    _fields_ = [("type", C.c_int32), ("n_vgoals", C.c_int32), ("dt", C.c_double), ("alt", C.c_double),
                ("u", _ParamsUnion)]


class Limits(C.Structure):
    """struct tgx_limits."""
    _fields_ = [("box", C.c_double * 6), ("v_max", C.c_double), ("a_max", C.c_double),
                ("check_box", C.c_int32), ("reserved", C.c_int32)]


class Layout(C.Structure):
    """struct tgx_layout."""
    _fields_ = [("d_base", C.c_void_p), ("traj_stride", C.c_int64), ("chan_stride", C.c_int64),
                ("d_traj_offset", C.c_void_p), ("capacity", C.c_int64), ("channel_mask", C.c_uint32),
                ("reserved", C.c_uint32)]


class Phases(C.Structure):
    """struct tgx_phases."""
    _fields_ = [("n", C.c_int32), ("key", C.c_int32 * TGX_MAX_PHASES), ("kind", C.c_int32 * TGX_MAX_PHASES),
                ("value", C.c_double * TGX_MAX_PHASES), ("value2", C.c_double * TGX_MAX_PHASES)]


assert C.sizeof(Params) == 128, C.sizeof(Params)
assert C.sizeof(Limits) == 72, C.sizeof(Limits)
assert C.sizeof(Layout) == 48, C.sizeof(Layout)

# numpy view of tgx_params: overlapping fields mirror the C union.
PARAMS_DTYPE = np.dtype({
    "names": ["type", "n_vgoals", "dt", "alt",
              "r", "cx", "cy", "t_traj", "accel", "v_goals",
              "A", "B", "a1", "a3", "v_goal"],
    "formats": ["<i4", "<i4", "<f8", "<f8",
                "<f8", "<f8", "<f8", "<f8", "<f8", ("<f8", (TGX_MAX_VGOALS,)),
                ("<f8", (3,)), ("<f8", (3,)), "<f8", "<f8", "<f8"],
    "offsets": [0, 4, 8, 16,
                24, 32, 40, 48, 56, 64,
                24, 48, 72, 80, 88],
    "itemsize": 128,
})

PHASES_DTYPE = np.dtype({
    "names": ["n", "key", "kind", "value", "value2"],
    "formats": ["<i4", ("<i4", (TGX_MAX_PHASES,)), ("<i4", (TGX_MAX_PHASES,)),
                ("<f8", (TGX_MAX_PHASES,)), ("<f8", (TGX_MAX_PHASES,))],
    "offsets": [Phases.n.offset, Phases.key.offset, Phases.kind.offset, Phases.value.offset, Phases.value2.offset],
    "itemsize": C.sizeof(Phases),
})


def concat(parts) -> np.ndarray:
    """Concatenate tgx_params arrays byte-wise.

    np.concatenate on this dtype silently re-packs the overlapping (union) fields into a 200-byte record; going
    through raw bytes keeps the 128-byte C layout.
    """
    parts = [np.ascontiguousarray(x) for x in parts]
    for x in parts:
        assert x.dtype == PARAMS_DTYPE, x.dtype
    if not parts:
        return np.zeros(0, dtype=PARAMS_DTYPE)
    raw = np.concatenate([x.view(np.uint8).reshape(-1, 128) for x in parts], axis=0)
    return np.ascontiguousarray(raw).view(PARAMS_DTYPE).reshape(-1)


def make_limits(box=None, v_max=float("inf"), a_max=float("inf")) -> Limits:
    lim = Limits()
    if box is not None:
        for i, b in enumerate(box):
            lim.box[i] = float(b)
        lim.check_box = 1
    lim.v_max = float(v_max)
    lim.a_max = float(a_max)
    return lim


def circle_params(alt, r, cx, cy, v_goals, t_traj, accel, dt, kind=TGX_CIRCLE) -> np.ndarray:
    """One Circle (or Figure8) record from the reference constructor arguments (Circle.hpp:30-31)."""
    p = np.zeros(1, dtype=PARAMS_DTYPE)
    v_goals = list(v_goals)
    p["type"] = kind
    p["n_vgoals"] = len(v_goals)
    p["dt"], p["alt"] = dt, alt
    p["r"], p["cx"], p["cy"], p["t_traj"], p["accel"] = r, cx, cy, t_traj, accel
    p["v_goals"][0, :min(len(v_goals), TGX_MAX_VGOALS)] = v_goals[:TGX_MAX_VGOALS]
    return p


def figure8_params(alt, r, cx, cy, v_goals, t_traj, accel, dt) -> np.ndarray:
    return circle_params(alt, r, cx, cy, v_goals, t_traj, accel, dt, kind=TGX_FIGURE8)


def line_params(alt, A, B, v_goals, a1, a3, dt) -> np.ndarray:
    """One Line record from the reference constructor arguments (Line.hpp:30-31)."""
    p = np.zeros(1, dtype=PARAMS_DTYPE)
    p["type"] = TGX_LINE
    p["n_vgoals"] = 1
    p["dt"], p["alt"] = dt, alt
    p["A"][0, :] = A
    p["B"][0, :] = B
    p["a1"], p["a3"] = a1, a3
    p["v_goal"] = list(v_goals)[0]
    return p


def boomerang_params(alt, A, B, v_goals, a1, a3, dt) -> np.ndarray:
    """One Boomerang record from the reference constructor arguments (Boomerang.hpp:30-31)."""
    p = line_params(alt, A, B, v_goals, a1, a3, dt)
    p["type"] = TGX_BOOMERANG
    return p


def format_phase(type_id: int, kind: int, value: float, value2: float, stop_traj: bool = False) -> str:
    """Rebuild the reference's index_msgs text from a (kind, value, value2) triple.

    std::to_string(double) is "%f" (Circle.cpp:45,61-62,74,89; Line.cpp:44,55-56,64,84; Figure8.cpp:45,61-62,74,89;
    braking: Circle.cpp:148,160, Line.cpp:132,143, Figure8.cpp:146,158).  Figure8::generateTraj announces
    "Figure 8 traj: stopped" (with a space, Figure8.cpp:89) while its braking trajectory says "Figure8 traj: stopped".
    """
    name = TYPE_NAMES[type_id]
    if kind == PH_ACCEL_TO:
        return f"{name} traj: accelerating to {value:f} m/s"
    if kind == PH_REACHED:
        return f"{name} traj: reached {value:f} m/s, keeping constant v for {value2:f} s"
    if kind == PH_DECEL:
        return f"{name} traj: decelerating to 0 m/s"
    if kind == PH_STOPPED:
        if type_id == TGX_FIGURE8 and not stop_traj:
            return "Figure 8 traj: stopped"
        return f"{name} traj: stopped"
    if kind == PH_PRESSED_END:
        return f"{name} traj: pressed END, decelerating to 0 m/s"
    raise ValueError(f"unknown phase kind {kind}")


def phases_to_index_msgs(type_id: int, ph, stop_traj: bool = False) -> dict:
    """tgx_phases record -> {sample index: message}; later entries overwrite earlier ones at the same key."""
    out = {}
    for i in range(int(ph["n"])):
        out[int(ph["key"][i])] = format_phase(type_id, int(ph["kind"][i]), float(ph["value"][i]),
                                              float(ph["value2"][i]), stop_traj)
    return out
