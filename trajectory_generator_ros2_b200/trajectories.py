"""Python mirror of the reference's trajectory classes over the C-ABI (include/tgx.h).

Same names, constructor arguments and call semantics as the reference's C++ classes
(Circle.hpp:30-31, Line.hpp:30-31, Figure8.hpp:30-31; Trajectory.hpp:33-46):

    traj = Circle(alt, r, cx, cy, v_goals, t_traj, accel, dt)
    traj.generateTraj(goals, index_msgs)            # APPENDS to goals, keys index_msgs by sample index
    pub_index = traj.generateStopTraj(goals, index_msgs, pub_index)   # REPLACES both, returns the new index (0)
    ok = traj.trajectoryInsideBounds(xmin, xmax, ymin, ymax, zmin, zmax)

Python has no reference parameters, so generateStopTraj returns the new pub_index instead of writing it through
an int&.  Every number comes from the CUDA engine; there is no Python or CPU implementation of the samplers here.
"""
from __future__ import annotations

import sys
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import abi
from .engine import Engine


@dataclass
class Vector3:
    x: float = 0.0
    y: float = 0.0
    z: float = 0.0


@dataclass
class Goal:
    """The fields of snapstack_msgs2/Goal the samplers fill (Circle.cpp:105-127)."""
    frame_id: str = ""
    p: Vector3 = field(default_factory=Vector3)
    v: Vector3 = field(default_factory=Vector3)
    a: Vector3 = field(default_factory=Vector3)
    j: Vector3 = field(default_factory=Vector3)
    psi: float = 0.0
    dpsi: float = 0.0
    power: bool = False

    @staticmethod
    def from_channels(c: Sequence[float]) -> "Goal":
        return Goal("world", Vector3(c[0], c[1], c[2]), Vector3(c[3], c[4], c[5]), Vector3(c[6], c[7], c[8]),
                    Vector3(c[9], c[10], c[11]), float(c[12]), float(c[13]), True)

    def channels(self) -> np.ndarray:
        return np.array([self.p.x, self.p.y, self.p.z, self.v.x, self.v.y, self.v.z, self.a.x, self.a.y, self.a.z,
                         self.j.x, self.j.y, self.j.z, self.psi, self.dpsi], dtype=np.float64)


class TrajectoryError(RuntimeError):
    """Raised where the reference logs an error and calls exit(1) (Circle.cpp:85-88, Line.cpp:76-79)."""


_shared_engine: Optional[Engine] = None


def shared_engine(device: int = 0) -> Engine:
    global _shared_engine
    if _shared_engine is None:
        _shared_engine = Engine(device)
    return _shared_engine


class Trajectory:
    """Abstract interface (Trajectory.hpp:24-61)."""

    shape = "?"

    def __init__(self, params: np.ndarray, engine: Optional[Engine] = None):
        self.params = params
        self.dt_ = float(params["dt"][0])
        self._engine = engine
        self.last_status = 0
        self.warnings: List[str] = []

    @property
    def engine(self) -> Engine:
        return self._engine if self._engine is not None else shared_engine()

    # -- generateTraj (Trajectory.hpp:33-35) -------------------------------------------------------------
    def generateTraj(self, goals: List[Goal], index_msgs: Dict[int, str], clock=None) -> None:
        counts, status = self.engine.count_host(self.params)
        st = int(status[0])
        if st & (abi.ST_BAD_PARAM | abi.ST_TOO_LONG):
            raise TrajectoryError(f"{self.shape} trajectory parameters rejected (status {st:#x})")
        cap = max(4, (int(counts[0]) + 3) // 4 * 4)
        out, counts, status, phases, legs = self.engine.generate_host_legs(self.params, cap)
        n, st = int(counts[0]), int(status[0])
        self.last_status = st
        base = len(goals)                                    # appends; keys are offset by the current size
        rows = out[0, :, :n].T
        goals.extend(Goal.from_channels(r) for r in rows)
        type_id = int(self.params["type"][0])
        if abi.is_polyline(type_id):
            msgs = abi.polyline_index_msgs(type_id, legs[0])
            if n == 0 and type_id != abi.TGX_RECIPROCATING:
                msgs = {-1: abi.polyline_msg(type_id, 0, -1, 0)}   # index_msgs[goals.size() - 1] on an empty vector
        else:
            msgs = abi.phases_to_index_msgs(type_id, phases[0])
        for k, msg in msgs.items():
            index_msgs[base + k] = msg
        if st & abi.ST_VGOALS_NOT_INCREASING:
            self.warnings.append("Vels are not in increasing order, ignoring vels from the first to decrease...")
        if st & abi.ST_FINAL_V_NONZERO:
            raise TrajectoryError("Error: final velocity is not zero")
        if st & abi.ST_LINE_END_NOT_B:
            raise TrajectoryError("Error: final point is not B")

    # -- generateStopTraj (Trajectory.hpp:38-41) ----------------------------------------------------------
    def generateStopTraj(self, goals: List[Goal], index_msgs: Dict[int, str], pub_index: int, clock=None) -> int:
        from14 = goals[pub_index].channels()
        _, counts, _, _ = self.engine.stop_host(self.params, from14, 0)
        cap = max(4, (int(counts[0]) + 3) // 4 * 4)
        out, counts, status, phases = self.engine.stop_host(self.params, from14, cap, want_phases=True)
        n = int(counts[0])
        self.last_status = int(status[0]) & ~abi.ST_TRUNCATED
        new_goals = [Goal.from_channels(r) for r in out[0, :, :n].T]
        new_msgs = abi.phases_to_index_msgs(int(self.params["type"][0]), phases[0], stop_traj=True)
        goals[:] = new_goals                                  # replaced, not appended (Circle.cpp:162-164)
        index_msgs.clear()
        index_msgs.update(new_msgs)
        return 0

    # -- trajectoryInsideBounds (Trajectory.hpp:44-46) -------------------------------------------------
    def trajectoryInsideBounds(self, xmin, xmax, ymin, ymax, zmin, zmax) -> bool:
        lim = abi.make_limits(box=(xmin, xmax, ymin, ymax, zmin, zmax))
        _, status = self.engine.count_host(self.params, lim)
        st = int(status[0])
        if st & abi.ST_LINE_D2_NEGATIVE:
            print("Line trajectory not feasible. Please increase accel, decrease v, or increase line length.",
                  file=sys.stderr)
        return (st & (abi.ST_OUTSIDE_BOUNDS | abi.ST_BAD_PARAM)) == 0


class Circle(Trajectory):
    shape = "Circle"

    def __init__(self, alt, r, cx, cy, v_goals, t_traj, accel, dt, engine: Optional[Engine] = None):
        super().__init__(abi.circle_params(alt, r, cx, cy, v_goals, t_traj, accel, dt), engine)

    def createCircleGoal(self, v, accel, theta) -> Goal:
        return Goal.from_channels(self.engine.sample_host(self.params, v, accel, theta))


class Figure8(Trajectory):
    shape = "Figure8"

    def __init__(self, alt, r, cx, cy, v_goals, t_traj, accel, dt, engine: Optional[Engine] = None):
        super().__init__(abi.figure8_params(alt, r, cx, cy, v_goals, t_traj, accel, dt), engine)

    def createFigure8Goal(self, v, accel, theta) -> Goal:
        return Goal.from_channels(self.engine.sample_host(self.params, v, accel, theta))


class Line(Trajectory):
    shape = "Line"

    def __init__(self, alt, A, B, v_goals, a1, a3, dt, engine: Optional[Engine] = None):
        super().__init__(abi.line_params(alt, A, B, v_goals, a1, a3, dt), engine)

    def createLineGoal(self, last_x, last_y, v, accel, theta) -> Goal:
        p = self.params.copy()
        raw = p.view(np.float64).reshape(-1, 16)
        raw[0, 12] = theta                                    # tgx_line_params.reserved[0]: the explicit heading
        return Goal.from_channels(self.engine.sample_host(p, v, accel, last_x, last_y))


class Boomerang(Line):
    """Line out and back (Boomerang.hpp:30-31); its announcements read "Line traj: ..." like the reference's."""
    shape = "Line"

    def __init__(self, alt, A, B, v_goals, a1, a3, dt, engine: Optional[Engine] = None):
        Trajectory.__init__(self, abi.boomerang_params(alt, A, B, v_goals, a1, a3, dt), engine)


class _Polyline(Trajectory):
    """Constant-speed polyline family: the public create<Shape>Goal(x, y, v, accel, heading) helper."""

    def _create_goal(self, x, y, v, accel, heading) -> Goal:
        p = self.params.copy()
        p["g"][0, 6] = heading                                # tgx_polyline_params.g[6]: the explicit heading
        return Goal.from_channels(self.engine.sample_host(p, v, accel, x, y))


class Square(_Polyline):
    shape = "Square"

    def __init__(self, alt, side_length, cx, cy, orientation, v_goals, t_traj, accel, dt, engine: Optional[Engine] = None):
        super().__init__(abi.square_params(alt, side_length, cx, cy, orientation, v_goals, t_traj, accel, dt), engine)

    def createSquareGoal(self, x, y, v, accel, heading) -> Goal:
        return self._create_goal(x, y, v, accel, heading)


class Rectangle(_Polyline):
    shape = "Rectangle"

    def __init__(self, alt, side_a, side_b, cx, cy, orientation, v_goals, t_traj, accel, dt,
                 engine: Optional[Engine] = None):
        super().__init__(abi.rectangle_params(alt, side_a, side_b, cx, cy, orientation, v_goals, t_traj, accel, dt),
                         engine)

    def createRectangleGoal(self, x, y, v, accel, heading) -> Goal:
        return self._create_goal(x, y, v, accel, heading)


class Reciprocating(_Polyline):
    shape = "Reciprocating"

    def __init__(self, alt, A, B, v_goals, a1, a3, t_traj, dt, engine: Optional[Engine] = None):
        super().__init__(abi.reciprocating_params(alt, A, B, v_goals, a1, a3, t_traj, dt), engine)

    def createReciprocatingGoal(self, x, y, v, accel, heading) -> Goal:
        return self._create_goal(x, y, v, accel, heading)


class Bounce(_Polyline):
    shape = "Bounce"

    def __init__(self, cx, cy, Az, Bz, v_goals, t_traj, orientation, dt, engine: Optional[Engine] = None):
        super().__init__(abi.bounce_params(cx, cy, Az, Bz, v_goals, t_traj, orientation, dt), engine)

    def createBounceGoal(self, x, y, z, vz, heading) -> Goal:
        p = self.params.copy()
        p["g"][0, 5] = z
        p["g"][0, 6] = heading
        return Goal.from_channels(self.engine.sample_host(p, vz, 0.0, x, y))


class M(_Polyline):
    shape = "M"

    def __init__(self, cx, cy, length, width, alt, v_goals, t_traj, orientation, dt, engine: Optional[Engine] = None):
        super().__init__(abi.letter_params(abi.TGX_M, cx, cy, length, width, alt, v_goals, t_traj, orientation, dt),
                         engine)

    def createMGoal(self, x, y, v, accel, heading) -> Goal:
        return self._create_goal(x, y, v, accel, heading)


class I(_Polyline):
    shape = "I"

    def __init__(self, cx, cy, length, width, alt, v_goals, t_traj, orientation, dt, engine: Optional[Engine] = None):
        super().__init__(abi.letter_params(abi.TGX_I, cx, cy, length, width, alt, v_goals, t_traj, orientation, dt),
                         engine)

    def createIGoal(self, x, y, v, accel, heading) -> Goal:
        return self._create_goal(x, y, v, accel, heading)


class T(_Polyline):
    shape = "T"

    def __init__(self, cx, cy, length, width, alt, v_goals, t_traj, orientation, dt, engine: Optional[Engine] = None):
        super().__init__(abi.letter_params(abi.TGX_T, cx, cy, length, width, alt, v_goals, t_traj, orientation, dt),
                         engine)

    def createTGoal(self, x, y, v, accel, heading) -> Goal:
        return self._create_goal(x, y, v, accel, heading)
