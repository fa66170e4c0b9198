"""CPU: the C-ABI library loads, exports every symbol include/tgx.h declares, its records have the documented
layout, and it fails loudly (no CPU fallback) when there is no CUDA device.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from trajectory_generator_ros2_b200 import abi, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "tgx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tgx_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = engine.lib()
    names = declared_functions()
    assert len(names) >= 24, names
    for n in names:
        assert hasattr(lib, n), f"libtgx.so does not export {n}"


def test_record_layouts():
    assert C.sizeof(abi.Params) == 128 and abi.PARAMS_DTYPE.itemsize == 128
    assert abi.Params.u.offset == 24
    assert abi.OrbitParams.v_goals.offset == 40 and abi.LineParams.v_goal.offset == 64
    assert C.sizeof(abi.Layout) == 48 and C.sizeof(abi.Limits) == 72
    assert C.sizeof(abi.Phases) == abi.PHASES_DTYPE.itemsize == 440
    # numpy view and ctypes struct agree field by field
    p = abi.line_params(1.8, [1, 2, 3], [4, 5, 6], [0.7], 1.5, 1.0, 0.01)
    c = abi.Params.from_buffer_copy(p.tobytes())
    assert c.type == abi.TGX_LINE and list(c.u.line.A) == [1, 2, 3] and list(c.u.line.B) == [4, 5, 6]
    assert c.u.line.a1 == 1.5 and c.u.line.a3 == 1.0 and c.u.line.v_goal == 0.7 and c.dt == 0.01 and c.alt == 1.8
    q = abi.circle_params(1.8, 3.4, 0.5, -0.5, [1.0, 2.0], 80.0, 0.4, 0.01)
    c = abi.Params.from_buffer_copy(q.tobytes())
    assert (c.u.orbit.r, c.u.orbit.cx, c.u.orbit.cy, c.u.orbit.t_traj, c.u.orbit.accel) == (3.4, 0.5, -0.5, 80.0, 0.4)
    assert list(c.u.orbit.v_goals)[:2] == [1.0, 2.0] and c.n_vgoals == 2
    # byte-wise concat keeps the 128-byte layout (np.concatenate would re-pack the union)
    both = abi.concat([p, q])
    assert both.dtype == abi.PARAMS_DTYPE and both.tobytes() == p.tobytes() + q.tobytes()


def test_version_strerror_and_shard_range():
    lib = engine.lib()
    assert lib.tgx_version() == 100
    assert lib.tgx_strerror(abi.TGX_ERR_NO_PLAN).decode().startswith("no current plan")
    for n, world in ((10, 3), (1 << 20, 8), (7, 8), (0, 4), (100_000_000, 8)):
        spans = [engine.shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(engine.TgxError):
        engine.shard_range(10, 3, 3)


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(engine.TgxError) as ei:
        engine.Engine(0)
    assert ei.value.code == abi.TGX_ERR_CUDA


def test_index_msgs_formatting():
    assert abi.format_phase(abi.TGX_CIRCLE, abi.PH_REACHED, 1.0, 80.0) == \
        "Circle traj: reached 1.000000 m/s, keeping constant v for 80.000000 s"
    assert abi.format_phase(abi.TGX_FIGURE8, abi.PH_STOPPED, 0, 0) == "Figure 8 traj: stopped"
    assert abi.format_phase(abi.TGX_FIGURE8, abi.PH_STOPPED, 0, 0, stop_traj=True) == "Figure8 traj: stopped"
    assert abi.format_phase(abi.TGX_LINE, abi.PH_PRESSED_END, 0, 0) == "Line traj: pressed END, decelerating to 0 m/s"


def test_new_record_layouts_and_host_finalize():
    """tgx_polyline_params / tgx_polyline_legs / tgx_goal_record / tgx_transition_params, and the one host-only entry
    point that needs no GPU: tgx_polyline_finalize_host (cos / sin of the orientation from this host's libm)."""
    import math
    assert C.sizeof(abi.PolylineParams) == 13 * 8 and C.sizeof(abi.PolylineLegs) == abi.LEGS_DTYPE.itemsize == 64
    assert C.sizeof(abi.GoalRecord) == abi.RECORD_DTYPE.itemsize == 128
    assert C.sizeof(abi.TransitionParams) == abi.TRANSITION_DTYPE.itemsize == 128
    for name in ("p", "v", "a", "j", "psi", "dpsi", "traj", "k", "power", "mode_xy", "mode_z", "clamped", "last"):
        assert getattr(abi.GoalRecord, name).offset == abi.RECORD_DTYPE.fields[name][1], name
    for name in abi.TRANSITION_DTYPE.names:
        assert getattr(abi.TransitionParams, name).offset == abi.TRANSITION_DTYPE.fields[name][1], name
    p = abi.concat([abi.square_params(1.8, 2.0, 0.1, 0.2, 0.5, [1.5], 12.0, 0.4, 0.01),
                    abi.letter_params(abi.TGX_T, 0.0, 0.0, 3.0, 4.0, 1.8, [], 80.0, -2.0, 0.01),
                    abi.circle_params(1.8, 3.4, 0, 0, [1.0], 80.0, 0.4, 0.01)])
    c = abi.Params.from_buffer_copy(p[0:1].tobytes())
    assert (c.u.poly.t_traj, c.u.poly.v_goal, c.u.poly.decel, c.u.poly.orientation) == (12.0, 1.5, 0.4, 0.5)
    assert list(c.u.poly.g)[:3] == [2.0, 0.1, 0.2] and c.n_vgoals == 0
    assert p["poly_v_goal"][1] == 1.0                       # v_goals_.empty() ? 1.0 : v_goals_[0]  (T.cpp:45)
    before = p[2:3].tobytes()
    engine.Engine.finalize_polyline(p)
    assert p["cos_o"][0] == math.cos(0.5) and p["sin_o"][0] == math.sin(0.5)
    assert p["cos_o"][1] == math.cos(-2.0) and p["sin_o"][1] == math.sin(-2.0)
    assert p["n_vgoals"][0] & abi.TGX_POLY_TRIG_GIVEN and p["n_vgoals"][1] & abi.TGX_POLY_TRIG_GIVEN
    assert p[2:3].tobytes() == before, "records of the other families are left untouched"


def test_polyline_message_formatting():
    L = np.zeros(1, dtype=abi.LEGS_DTYPE)
    L["n"], L["n_legs"], L["first_special"], L["period"] = 10, 4, 1, 8
    L["count"][0, :4] = [2, 2, 2, 2]
    m = abi.polyline_index_msgs(abi.TGX_SQUARE, L[0])
    assert m[0] == "Square traj: starting at corner 0" and m[1] == m[2] == "Square traj: moving along side 0"
    assert m[6] == "Square traj: moving along side 2" and m[7] == m[8] == "Square traj: moving along side 3"
    assert m[9] == "Square traj: completed" and len(m) == 10
    L["n"], L["n_legs"], L["first_special"], L["last_special"], L["period"] = 8, 4, 0, 1, 8
    L["count"][0, :4] = [3, 1, 3, 1]
    m = abi.polyline_index_msgs(abi.TGX_RECIPROCATING, L[0])
    assert [m[k] for k in (0, 2, 3, 4, 6)] == ["Reciprocating: forward", "Reciprocating: forward",
                                               "Reciprocating: yaw flip at endpoint", "Reciprocating: reverse",
                                               "Reciprocating: reverse"]
    assert m[7] == "Reciprocating: yaw flip at endpoint"
    assert abi.polyline_msg(abi.TGX_M, 5, 3, 100) == "M traj: segment 1 rev"
    assert abi.polyline_msg(abi.TGX_I, 4, 3, 100) == "I traj: segment 4 fwd"
    assert abi.polyline_msg(abi.TGX_BOUNCE, 1, 99, 100) == "Bounce: completed"
