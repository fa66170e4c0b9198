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
TGX_MAX_VGOALS_TOTAL = 64
TGX_VGOALS_MORE = 30          # type of a continuation record (tgx.h)


def orbit_records(k: int) -> int:
    """TGX_ORBIT_RECORDS: parameter records a Circle / Figure8 with k goal speeds occupies."""
    return 1 if k <= TGX_MAX_VGOALS else 1 + (k - 1) // TGX_MAX_VGOALS

TGX_NCHAN = 14
TGX_MAX_PHASES = 2 * TGX_MAX_VGOALS + 2

# enum tgx_type
TGX_CIRCLE, TGX_LINE, TGX_FIGURE8, TGX_BOOMERANG = 0, 1, 2, 3
TGX_SQUARE, TGX_RECTANGLE, TGX_RECIPROCATING, TGX_BOUNCE, TGX_M, TGX_I, TGX_T = 4, 5, 6, 7, 8, 9, 10
POLYLINE_TYPES = (TGX_SQUARE, TGX_RECTANGLE, TGX_RECIPROCATING, TGX_BOUNCE, TGX_M, TGX_I, TGX_T)
TGX_POLY_TRIG_GIVEN = 1
TGX_POLY_MAX_LEGS = 10
# prefix of the index_msgs texts; Boomerang.cpp announces itself as "Line traj: ..." (Boomerang.cpp:44,55,64,94,134)
TYPE_NAMES = {TGX_CIRCLE: "Circle", TGX_LINE: "Line", TGX_FIGURE8: "Figure8", TGX_BOOMERANG: "Line",
              TGX_SQUARE: "Square", TGX_RECTANGLE: "Rectangle", TGX_RECIPROCATING: "Reciprocating",
              TGX_BOUNCE: "Bounce", TGX_M: "M", TGX_I: "I", TGX_T: "T"}


def is_polyline(type_id: int) -> bool:
    return TGX_SQUARE <= int(type_id) <= TGX_T

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
ST_WRONG_PLANNER = 1 << 10
ST_FATAL_MASK = ST_FINAL_V_NONZERO | ST_LINE_END_NOT_B | ST_BAD_PARAM | ST_TOO_LONG | ST_WRONG_PLANNER
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
    ST_WRONG_PLANNER: "WRONG_PLANNER",
}

# enum tgx_error
(TGX_OK, TGX_ERR_INVALID, TGX_ERR_CUDA, TGX_ERR_ALIGNMENT, TGX_ERR_NO_PLAN, TGX_ERR_NOMEM, TGX_ERR_CAPACITY,
 TGX_ERR_COMM) = range(8)
TGX_COMM_ID_BYTES = 128

# enum tgx_phase_kind
PH_ACCEL_TO, PH_REACHED, PH_DECEL, PH_STOPPED, PH_PRESSED_END = range(5)

DEFAULT_MAX_SAMPLES = 1 << 24


class OrbitParams(C.Structure):
    _fields_ = [("r", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("t_traj", C.c_double),
                ("accel", C.c_double), ("v_goals", C.c_double * TGX_MAX_VGOALS)]


class LineParams(C.Structure):
    _fields_ = [("A", C.c_double * 3), ("B", C.c_double * 3), ("a1", C.c_double), ("a3", C.c_double),
                ("v_goal", C.c_double), ("reserved", C.c_double * 4)]


class PolylineParams(C.Structure):
    _fields_ = [("t_traj", C.c_double), ("v_goal", C.c_double), ("decel", C.c_double), ("orientation", C.c_double),
                ("cos_o", C.c_double), ("sin_o", C.c_double), ("g", C.c_double * 7)]


class _ParamsUnion(C.Union):
    _fields_ = [("orbit", OrbitParams), ("line", LineParams), ("poly", PolylineParams)]


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


TGX_NCHAN_VARYING = 10
VARYING_CHANNELS = (0, 1, 3, 4, 6, 7, 9, 10, 12, 13)      # tgx_compact_plane: px py vx vy ax ay jx jy psi dpsi
VARYING_CHANNEL_MASK = 0x36DB


class HostInfo(C.Structure):
    """struct tgx_host_info_t."""
    _fields_ = [("numa_node", C.c_int32), ("cpus_allowed", C.c_int32), ("local_ranks", C.c_int32),
                ("filler_threads", C.c_int32), ("filler_cpus", C.c_int32), ("reserved", C.c_int32 * 3)]


class Phases(C.Structure):
    """struct tgx_phases."""
    _fields_ = [("n", C.c_int32), ("key", C.c_int32 * TGX_MAX_PHASES), ("kind", C.c_int32 * TGX_MAX_PHASES),
                ("value", C.c_double * TGX_MAX_PHASES), ("value2", C.c_double * TGX_MAX_PHASES)]


class PolylineLegs(C.Structure):
    """struct tgx_polyline_legs."""
    _fields_ = [("n", C.c_int32), ("n_legs", C.c_int32), ("first_special", C.c_int32), ("last_special", C.c_int32),
                ("period", C.c_int32), ("count", C.c_int32 * TGX_POLY_MAX_LEGS), ("reserved", C.c_int32)]


class GoalRecord(C.Structure):
    """struct tgx_goal_record (128 bytes)."""
    _fields_ = [("p", C.c_double * 3), ("v", C.c_double * 3), ("a", C.c_double * 3), ("j", C.c_double * 3),
                ("psi", C.c_double), ("dpsi", C.c_double), ("traj", C.c_int32), ("k", C.c_int32),
                ("power", C.c_uint8), ("mode_xy", C.c_uint8), ("mode_z", C.c_uint8), ("clamped", C.c_uint8),
                ("last", C.c_uint8), ("reserved", C.c_uint8 * 3)]


class TransitionParams(C.Structure):
    """struct tgx_transition_params (128 bytes)."""
    _fields_ = [("kind", C.c_int32), ("ticks", C.c_int32), ("dt", C.c_double), ("start", C.c_double * 3),
                ("start_v", C.c_double * 2), ("start_psi", C.c_double), ("dest", C.c_double * 3),
                ("dest_yaw", C.c_double), ("vel", C.c_double), ("vel_yaw", C.c_double),
                ("dist_thresh", C.c_double), ("yaw_thresh", C.c_double)]


TR_TAKEOFF, TR_GOTO, TR_LANDING = 0, 1, 2
assert C.sizeof(TransitionParams) == 128, C.sizeof(TransitionParams)
TRANSITION_DTYPE = np.dtype({
    "names": ["kind", "ticks", "dt", "start", "start_v", "start_psi", "dest", "dest_yaw", "vel", "vel_yaw",
              "dist_thresh", "yaw_thresh"],
    "formats": ["<i4", "<i4", "<f8", ("<f8", (3,)), ("<f8", (2,)), "<f8", ("<f8", (3,)), "<f8", "<f8", "<f8",
                "<f8", "<f8"],
    "offsets": [0, 4, 8, 16, 40, 56, 64, 88, 96, 104, 112, 120],
    "itemsize": 128,
})
assert C.sizeof(GoalRecord) == 128, C.sizeof(GoalRecord)
assert C.sizeof(PolylineLegs) == 64, C.sizeof(PolylineLegs)
assert C.sizeof(Params) == 128, C.sizeof(Params)
assert C.sizeof(Limits) == 72, C.sizeof(Limits)
assert C.sizeof(Layout) == 48, C.sizeof(Layout)

# numpy view of tgx_params: overlapping fields mirror the C union.
PARAMS_DTYPE = np.dtype({
    "names": ["type", "n_vgoals", "dt", "alt",
              "r", "cx", "cy", "t_traj", "accel", "v_goals",
              "A", "B", "a1", "a3", "v_goal", "line_heading",
              "poly_t_traj", "poly_v_goal", "poly_decel", "orientation", "cos_o", "sin_o", "g"],
    "formats": ["<i4", "<i4", "<f8", "<f8",
                "<f8", "<f8", "<f8", "<f8", "<f8", ("<f8", (TGX_MAX_VGOALS,)),
                ("<f8", (3,)), ("<f8", (3,)), "<f8", "<f8", "<f8", "<f8",
                "<f8", "<f8", "<f8", "<f8", "<f8", "<f8", ("<f8", (7,))],
    "offsets": [0, 4, 8, 16,
                24, 32, 40, 48, 56, 64,
                24, 48, 72, 80, 88, 96,
                24, 32, 40, 48, 56, 64, 72],
    "itemsize": 128,
})

RECORD_DTYPE = np.dtype({
    "names": ["p", "v", "a", "j", "psi", "dpsi", "traj", "k", "power", "mode_xy", "mode_z", "clamped", "last"],
    "formats": [("<f8", (3,)), ("<f8", (3,)), ("<f8", (3,)), ("<f8", (3,)), "<f8", "<f8", "<i4", "<i4",
                "u1", "u1", "u1", "u1", "u1"],
    "offsets": [0, 24, 48, 72, 96, 104, 112, 116, 120, 121, 122, 123, 124],
    "itemsize": 128,
})


def records_to_channels(rec: np.ndarray) -> np.ndarray:
    """tgx_goal_record array [N] -> samples [14, N] in tgx_channel order."""
    return np.concatenate([rec["p"].T, rec["v"].T, rec["a"].T, rec["j"].T, rec["psi"][None], rec["dpsi"][None]], axis=0)


LEGS_DTYPE = np.dtype({
    "names": ["n", "n_legs", "first_special", "last_special", "period", "count"],
    "formats": ["<i4", "<i4", "<i4", "<i4", "<i4", ("<i4", (TGX_POLY_MAX_LEGS,))],
    "offsets": [0, 4, 8, 12, 16, 20],
    "itemsize": 64,
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
    """One Circle (or Figure8) from the reference constructor arguments (Circle.hpp:30-31): one record, plus the
    continuation records (TGX_VGOALS_MORE) that hold the goal speeds beyond the eighth."""
    v_goals = list(v_goals)
    p = np.zeros(orbit_records(len(v_goals)), dtype=PARAMS_DTYPE)
    p["type"] = TGX_VGOALS_MORE
    p["type"][0] = kind
    p["n_vgoals"][0] = len(v_goals)
    p["dt"], p["alt"] = dt, alt
    p["r"][0], p["cx"][0], p["cy"][0], p["t_traj"][0], p["accel"][0] = r, cx, cy, t_traj, accel
    for q in range(len(p)):
        part = v_goals[q * TGX_MAX_VGOALS:(q + 1) * TGX_MAX_VGOALS]
        p["v_goals"][q, :len(part)] = part
        if q:
            p["n_vgoals"][q] = len(part)
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


def polyline_params(kind, dt, alt, t_traj, v_goals, decel, orientation, geometry) -> np.ndarray:
    """One polyline-family record.  cos_o / sin_o are left to tgx_polyline_finalize_host (Engine.finalize_polyline)."""
    p = np.zeros(1, dtype=PARAMS_DTYPE)
    v_goals = list(v_goals)
    p["type"] = kind
    p["dt"], p["alt"] = dt, alt
    p["poly_t_traj"] = t_traj
    p["poly_v_goal"] = v_goals[0] if v_goals else 1.0          # v_goals_.empty() ? 1.0 : v_goals_[0] (Square.cpp:48)
    p["poly_decel"] = decel
    p["orientation"] = orientation
    p["g"][0, :len(geometry)] = geometry
    return p


def square_params(alt, side_length, cx, cy, orientation, v_goals, t_traj, accel, dt) -> np.ndarray:
    """Square.hpp:31-32."""
    return polyline_params(TGX_SQUARE, dt, alt, t_traj, v_goals, accel, orientation, [side_length, cx, cy])


def rectangle_params(alt, side_a, side_b, cx, cy, orientation, v_goals, t_traj, accel, dt) -> np.ndarray:
    """Rectangle.hpp."""
    return polyline_params(TGX_RECTANGLE, dt, alt, t_traj, v_goals, accel, orientation, [side_a, side_b, cx, cy])


def reciprocating_params(alt, A, B, v_goals, a1, a3, t_traj, dt) -> np.ndarray:
    """Reciprocating.hpp (a1 is stored by the class and never used)."""
    return polyline_params(TGX_RECIPROCATING, dt, alt, t_traj, v_goals, a3, 0.0, list(A) + list(B))


def bounce_params(cx, cy, Az, Bz, v_goals, t_traj, orientation, dt) -> np.ndarray:
    """Bounce.hpp (no alt: z runs between Az and Bz)."""
    return polyline_params(TGX_BOUNCE, dt, 0.0, t_traj, v_goals, 0.0, orientation, [cx, cy, Az, Bz])


def letter_params(kind, cx, cy, length, width, alt, v_goals, t_traj, orientation, dt) -> np.ndarray:
    """M.hpp / I.hpp / T.hpp."""
    return polyline_params(kind, dt, alt, t_traj, v_goals, 1.0, orientation, [cx, cy, length, width])


def polyline_leg_of(legs, k: int) -> int:
    """Leg (tgx_polyline_legs numbering) of sample k; -1 for the Square / Rectangle start sample."""
    n = int(legs["n"])
    first = int(legs["first_special"])
    if first and k == 0:
        return -1
    counts = [int(c) for c in legs["count"][:int(legs["n_legs"])]]
    if int(legs["last_special"]) and k == n - 1:
        # the yaw flip appended after the leg that t_traj cut short (Reciprocating.cpp:50-57)
        return polyline_leg_of(legs, k - 1) + 1
    m = (k - first) % int(legs["period"])
    for leg, c in enumerate(counts):
        if m < c:
            return leg
        m -= c
    raise AssertionError("inconsistent tgx_polyline_legs record")


def polyline_msg(type_id: int, leg: int, k: int, n: int) -> str:
    """The reference's index_msgs text of sample k of n (Square.cpp:61,79,88; Reciprocating.cpp:47,57; Bounce.cpp:42,50;
    M.cpp:57,65; I.cpp:65,73; T.cpp:63,71)."""
    name = TYPE_NAMES[type_id]
    last = k == n - 1
    if type_id in (TGX_SQUARE, TGX_RECTANGLE):
        if last:
            return f"{name} traj: completed"
        return f"{name} traj: starting at corner 0" if leg < 0 else f"{name} traj: moving along side {leg}"
    if type_id == TGX_RECIPROCATING:
        return ("Reciprocating: forward", "Reciprocating: yaw flip at endpoint", "Reciprocating: reverse",
                "Reciprocating: yaw flip at endpoint")[leg]
    if type_id == TGX_BOUNCE:
        if last:
            return "Bounce: completed"
        return "Bounce: ascending" if leg == 0 else "Bounce: descending"
    nseg = {TGX_M: 4, TGX_I: 5, TGX_T: 3}[type_id]
    if last:
        return f"{name} traj: completed"
    return f"{name} traj: segment {leg % nseg} {'fwd' if leg < nseg else 'rev'}"


def polyline_index_msgs(type_id: int, legs) -> dict:
    """tgx_polyline_legs record -> {sample index: message} for every sample, like the reference's map."""
    n = int(legs["n"])
    first = int(legs["first_special"])
    counts = [int(c) for c in legs["count"][:int(legs["n_legs"])]]
    out = {}
    k = 0
    if first and n > 0:
        out[0] = polyline_msg(type_id, -1, 0, n)
        k = 1
    leg = 0
    while k < n:
        for _ in range(counts[leg]):
            if k >= n:
                break
            out[k] = polyline_msg(type_id, leg, k, n)
            k += 1
        leg = (leg + 1) % len(counts)
    if int(legs["last_special"]) and n > 0:
        out[n - 1] = polyline_msg(type_id, polyline_leg_of(legs, n - 1), n - 1, n)
    return out


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
    """tgx_phases record -> {sample index: message}; later entries overwrite earlier ones at the same key.  `ph` may
    also be the ROWS of a trajectory with more than 8 goal speeds (its own and its continuation records')."""
    out = {}
    for row in (ph if isinstance(ph, np.ndarray) and ph.ndim == 1 else [ph]):
        for i in range(int(row["n"])):
            out[int(row["key"][i])] = format_phase(type_id, int(row["kind"][i]), float(row["value"][i]),
                                                   float(row["value2"][i]), stop_traj)
    return out
