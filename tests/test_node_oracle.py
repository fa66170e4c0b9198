"""CPU: the UNMODIFIED reference node (TrajectoryGenerator.cpp behind the fake rclcpp::Node of compat/ros2_stubs,
oracle/_ref/libnoderef.so) driven tick by tick, and the oracle's restatement of its consumer path pinned to it.

  * the node's TRAJ_FOLLOWING stream (goal_ = traj_goals_[pub_index_], position saturated to the room bounds,
    TrajectoryGenerator.cpp:556-561, :602-604) == oracle generate + orc_pack_goals, bit for bit, for all 11 classes;
  * END pressed while following: the braking trajectory the node switches to == the oracle's orc_stop.
"""
import numpy as np
import pytest

import golden_util
import node_lib
from trajectory_generator_ros2_b200 import abi

if not __import__("os").path.exists(node_lib.NODE_REF_SO):
    pytest.skip("oracle/_ref/libnoderef.so not built (needs /root/reference at build time)", allow_module_level=True)

T_GO1, T_GO2, T_GO3 = 5, 700, 2600          # take off, go to the start of the trajectory, follow it
MISSION = [(T_GO1, node_lib.GO), (T_GO2, node_lib.GO), (T_GO3, node_lib.GO)]


def params_for(traj_type: str, cfg: dict) -> np.ndarray:
    """The tgx_params record the node's dispatch builds for `traj_type` (TrajectoryGenerator.cpp:176-388)."""
    c = dict(node_lib.DEFAULT_YAML)
    c.update(cfg)
    dt, alt, vg = 1.0 / c["pub_freq"], c["alt"], c["v_goals"]
    A, B = [c["Ax"], c["Ay"], alt], [c["Bx"], c["By"], alt]
    if traj_type == "Circle":
        return abi.circle_params(alt, c["r"], c["center_x"], c["center_y"], vg, c["t_traj"], c["circle_accel"], dt)
    if traj_type == "Figure8":
        return abi.figure8_params(alt, c["r"], c["center_x"], c["center_y"], vg, c["t_traj"], c["circle_accel"], dt)
    if traj_type == "Line":
        return abi.line_params(alt, A, B, [c["v_line"]], c["line_accel"], c["line_decel"], dt)
    if traj_type == "Boomerang":
        return abi.boomerang_params(alt, A, B, [c["v_line"]], c["line_accel"], c["line_decel"], dt)
    if traj_type == "Reciprocating":
        return abi.reciprocating_params(alt, A, B, [c["v_line"]], c["line_accel"], c["line_decel"], c["t_traj"], dt)
    if traj_type == "Square":
        return abi.square_params(alt, c["side_length"], c["center_x"], c["center_y"], c["orientation"], vg,
                                 c["t_traj"], c["square_accel"], dt)
    if traj_type == "Rectangle":
        return abi.rectangle_params(alt, c["side_a"], c["side_b"], c["center_x"], c["center_y"], c["orientation"], vg,
                                    c["t_traj"], c["rectangle_accel"], dt)
    if traj_type == "Bounce":
        return abi.bounce_params(c["center_x"], c["center_y"], c["Az"], c["Bz"], vg, c["t_traj"], c["orientation"], dt)
    kind = {"M": abi.TGX_M, "I": abi.TGX_I, "T": abi.TGX_T}[traj_type]
    return abi.letter_params(kind, c["center_x"], c["center_y"], c[traj_type + "_length"], c[traj_type + "_width"],
                             alt, vg, c["t_traj"], c["orientation"], dt)


def oracle_samples(oracle, p):
    if abi.is_polyline(int(p["type"][0])):
        s, st, _, _ = oracle.polyline_generate(p)
    else:
        s, st, _ = oracle.generate(p)
    return s


ALL_TYPES = ["Circle", "Line", "Boomerang", "Figure8", "Square", "Reciprocating", "Rectangle", "Bounce", "M", "I", "T"]
# a box that cuts into every shape, so that the saturation is exercised on x, y and z
TIGHT = {"x_min": -1.5, "x_max": 2.5, "y_min": -2.0, "y_max": 1.0, "z_min": 0.0, "z_max": 2.0}


@pytest.mark.parametrize("traj_type", ALL_TYPES)
def test_following_stream_is_generate_plus_pack(oracle, traj_type):
    node = node_lib.Node(node_lib.NODE_REF_SO)
    cfg = {"traj_type": traj_type, "t_traj": 12.0, "orientation": 0.3}
    # the node refuses parameters whose shape leaves the room (TrajectoryGenerator.cpp:419-422): check that first ...
    assert node.run(dict(cfg, **TIGHT), MISSION, 10) is None
    # ... then follow inside the default room, and saturate the recorded stream to the tight box in the comparison
    p = params_for(traj_type, cfg)
    s = oracle_samples(oracle, p)
    n = s.shape[1]
    rows = node.run(cfg, MISSION, T_GO3 + n + 50)
    assert rows is not None
    # The node never publishes the LAST sample: on the tick that loads traj_goals_[N-1] pub_index_ reaches size(), goal_
    # is overwritten by the hover goal at the vehicle's pose and that is published, twice (:561-572, :602-610).
    follow = rows[(rows[:, 0] >= T_GO3) & (rows[:, 0] < T_GO3 + n - 1)]
    assert len(follow) == n - 1, "one published goal per tick while following"
    box = [node_lib.DEFAULT_YAML[k] for k in ("x_min", "x_max", "y_min", "y_max", "z_min", "z_max")]
    want = oracle.pack_goals(s, traj=0, box=box)
    got = follow[:, 1:15].T
    assert golden_util.same_bits(got, abi.records_to_channels(want)[:, :n - 1]), traj_type
    assert (follow[:, 15] == 1).all() and (follow[:, 16] == 0).all() and (follow[:, 17] == 0).all()
    assert (want["power"] == 1).all() and (want["clamped"] == 0).all()
    assert want["last"][-1] == 1 and want["last"][:-1].sum() == 0
    hover = rows[rows[:, 0] == T_GO3 + n - 1]
    assert len(hover) == 2 and (hover[:, 4:13] == 0).all() and (hover[:, 3] == 1.8).all()
    assert golden_util.same_bits(hover[0, 1:3], follow[-1, 1:3])         # perfect tracking: pose = the last goal


def test_saturation_matches_the_node(oracle):
    """The node checks its shape against the room box at start-up with the same members it later saturates with
    (TrajectoryGenerator.cpp:419-422 vs :602-604), so while FOLLOWING an accepted trajectory the saturation is a safety
    net that never fires: a Circle that touches all four walls is published unchanged.  The restated saturate() itself
    (high tested first, NaN passes through) is then checked on a synthetic stream that crosses every face of a box,
    and against the node's own landing phase, where z is driven below z_min (:588-599)."""
    node = node_lib.Node(node_lib.NODE_REF_SO)
    # Circle r = 1 centred at (1, 0): bbox corners (0, -1) and (2, 1) are inside [0, 2] x [-1, 1]; samples touch the
    # bounds (cos/sin rounding puts some a hair outside: those get saturated)
    cfg = {"traj_type": "Circle", "r": 1.0, "center_x": 1.0, "center_y": 0.0, "t_traj": 6.0, "v_goals": [1.0],
           "x_min": 0.0, "x_max": 2.0, "y_min": -1.0, "y_max": 1.0, "z_min": 0.0, "z_max": 1.8, "alt": 1.8}
    p = params_for("Circle", cfg)
    s = oracle_samples(oracle, p)
    n = s.shape[1]
    rows = node.run(cfg, [(5, node_lib.GO), (700, node_lib.GO), (1500, node_lib.GO)], 1500 + n + 5, start=(1.5, 0, 0, 0))
    assert rows is not None
    follow = rows[(rows[:, 0] >= 1500) & (rows[:, 0] < 1500 + n - 1)]
    assert len(follow) == n - 1
    want = oracle.pack_goals(s, box=[0.0, 2.0, -1.0, 1.0, 0.0, 1.8])
    assert golden_util.same_bits(follow[:, 1:15].T, abi.records_to_channels(want)[:, :n - 1])
    assert want["clamped"].sum() == 0
    # the node's landing: goal z keeps decreasing by vel_land*dt and is published saturated at z_min = 0
    t_land = 1500 + n + 50
    rows = node.run(cfg, [(5, node_lib.GO), (700, node_lib.GO), (1500, node_lib.GO), (t_land, node_lib.LAND)],
                    t_land + 4000, start=(1.5, 0, 0, 0))
    assert rows[:, 3].min() == 0.0 and rows[-1, 15] == 0
    # and a synthetic stream that crosses every face of a box
    rng = np.random.default_rng(3)
    fake = rng.uniform(-3, 3, (abi.TGX_NCHAN, 500))
    fake[abi.PX, 7] = np.nan                                   # a NaN passes through saturate() untouched
    rec = oracle.pack_goals(fake, traj=9, box=[-1, 2, -2, 1, -0.5, 0.5])
    assert np.array_equal(rec["p"][:, 0][~np.isnan(fake[0])], np.clip(fake[0], -1, 2)[~np.isnan(fake[0])])
    assert np.isnan(rec["p"][7, 0]) and not (rec["clamped"][7] & 1)
    assert np.array_equal(rec["p"][:, 1], np.clip(fake[1], -2, 1)) and np.array_equal(rec["p"][:, 2], np.clip(fake[2], -0.5, 0.5))
    assert np.array_equal(rec["clamped"] & 2, np.where((fake[1] > 1) | (fake[1] < -2), 2, 0))
    assert (rec["traj"] == 9).all() and np.array_equal(rec["k"], np.arange(500)) and rec["last"].sum() == 1


@pytest.mark.parametrize("traj_type", ["Circle", "Line", "Figure8", "T", "Bounce", "Square"])
def test_end_while_following_switches_to_the_braking_trajectory(oracle, traj_type):
    """modeCB: TRAJ_FOLLOWING --END--> generateStopTraj(traj_goals_, index_msgs_, pub_index_) (TrajectoryGenerator.cpp:
    514-517); the next ticks publish the braking trajectory from its index 0."""
    node = node_lib.Node(node_lib.NODE_REF_SO)
    cfg = {"traj_type": traj_type, "t_traj": 12.0}
    p = params_for(traj_type, cfg)
    s = oracle_samples(oracle, p)
    k_end = 333
    t_end = T_GO3 + k_end
    rows = node.run(cfg, MISSION + [(t_end, node_lib.LAND)], t_end + 600)
    # at tick t_end pub_index_ == k_end: the node brakes from traj_goals_[k_end] and publishes stop[0] on that tick
    stop, _, _ = oracle.stop(p, s[:, k_end])
    m = stop.shape[1]
    assert m > 0
    # (its last sample too is replaced by the hover goal, see test_following_stream_is_generate_plus_pack)
    got = rows[(rows[:, 0] >= t_end) & (rows[:, 0] < t_end + m - 1)]
    assert len(got) == m - 1
    assert golden_util.same_bits(got[:, 1:15].T, stop[:, :m - 1]), traj_type
    before = rows[rows[:, 0] == t_end - 1][0]
    assert golden_util.same_bits(before[1:15], s[:, k_end - 1])


def test_mission_phases_take_off_go_to_start_and_land(oracle):
    """The node's own transitions around the trajectory (take-off ramp :531-548, simpleInterpolation towards the start
    :549-554 / :637-699, landing :588-599) under perfect tracking: recorded here so the stream the drop-in build has to
    reproduce is known to pass through every flight mode."""
    node = node_lib.Node(node_lib.NODE_REF_SO)
    cfg = {"traj_type": "Circle", "t_traj": 5.0}
    p = params_for("Circle", cfg)
    n = oracle_samples(oracle, p).shape[1]
    t_land = T_GO3 + n + 100
    rows = node.run(cfg, MISSION + [(t_land, node_lib.LAND)], t_land + 3500, start=(0.5, -0.5, 0.0, 0.3))
    z = rows[:, 3]
    assert rows[0, 15] == 0 and (rows[:T_GO1, 15] == 0).all()            # GROUND: power off
    assert rows[T_GO1, 15] == 1                                          # take off
    np.testing.assert_allclose(np.diff(z[T_GO1:T_GO1 + 100]), 0.3 * 0.01, rtol=1e-9)   # vel_take * dt per tick
    assert z[T_GO2 - 1] == 1.8
    k = T_GO2 + 10                                                       # moving towards the start at vel_initpos
    step = np.hypot(rows[k + 1, 1] - rows[k, 1], rows[k + 1, 2] - rows[k, 2])
    assert abs(step - 0.4 * 0.01) < 1e-12
    assert abs(rows[T_GO3 - 1, 1] - 3.4) < 1e-12 and abs(rows[T_GO3 - 1, 2]) < 1e-12   # arrived at traj_goals_[0]
    assert rows[-1, 15] == 0 and rows[-1, 3] <= 0.0                       # landed: motors off


# ---- node-side transitions: orc_transition pinned to the unmodified node -------------------------

def transition(kind, dt, start, start_v, start_psi, dest, dest_yaw, vel, vel_yaw, dist_thresh=0.0, yaw_thresh=0.0,
               ticks=0):
    t = np.zeros(1, dtype=abi.TRANSITION_DTYPE)
    t["kind"], t["ticks"], t["dt"] = kind, ticks, dt
    t["start"][0], t["start_v"][0], t["start_psi"] = start, start_v, start_psi
    t["dest"][0], t["dest_yaw"], t["vel"], t["vel_yaw"] = dest, dest_yaw, vel, vel_yaw
    t["dist_thresh"], t["yaw_thresh"] = dist_thresh, yaw_thresh
    return t


def mission_transitions(rows, cfg, start_pose, first_goal, t_go1, t_go2, t_go3, t_land):
    """The four transition phases of a mission as tgx_transition_params, their start states read off the node's own
    stream (the goal published on the tick before the phase begins), and the tick each phase starts on."""
    c = dict(node_lib.DEFAULT_YAML)
    c.update(cfg)
    dt = 1.0 / c["pub_freq"]
    row = lambda t: rows[rows[:, 0] == t][-1]
    x0, y0, z0 = start_pose[:3]
    psi0 = row(t_go1)[13]                      # quat2yaw(pose_.orientation) as the node computed it (:471)
    takeoff = transition(abi.TR_TAKEOFF, dt, [x0, y0, z0], [0, 0], psi0, [0, 0, c["alt"]], 0.0, c["vel_take"], 0.0)
    before = row(t_go2 - 1)
    goto = transition(abi.TR_GOTO, dt, before[1:4], before[4:6], before[13], first_goal[0:3], first_goal[12],
                      c["vel_initpos"], c["vel_yaw"], c["dist_thresh"], c["yaw_thresh"], ticks=t_go3 - t_go2)
    before = row(t_land - 1)
    home = transition(abi.TR_GOTO, dt, before[1:4], before[4:6], before[13], [x0, y0, c["alt"]], before[13],
                      c["vel_initpos"], c["vel_yaw"], c["dist_thresh"], c["yaw_thresh"])
    return takeoff, goto, home, c


@pytest.mark.parametrize("traj_type,start_pose", [("Circle", (0.5, -0.5, 0.0, 0.3)), ("T", (-1.0, 2.0, 0.05, -2.0)),
                                                  ("Line", (3.0, 3.0, 0.0, 3.0))])
def test_transitions_match_the_node(oracle, traj_type, start_pose):
    node = node_lib.Node(node_lib.NODE_REF_SO)
    cfg = {"traj_type": traj_type, "t_traj": 4.0, "z_min": 0.0}
    p = params_for(traj_type, cfg)
    s = oracle_samples(oracle, p)
    n = s.shape[1]
    t_land = T_GO3 + n + 60
    rows = node.run(cfg, MISSION + [(t_land, node_lib.LAND)], t_land + 6000, start=start_pose)
    takeoff, goto, home, c = mission_transitions(rows, cfg, start_pose, s[:, 0], T_GO1, T_GO2, T_GO3, t_land)
    box = [c[k] for k in ("x_min", "x_max", "y_min", "y_max", "z_min", "z_max")]

    def check(rec, t0, what):
        got = rows[(rows[:, 0] >= t0) & (rows[:, 0] < t0 + len(rec))]
        assert len(got) == len(rec), what
        assert golden_util.same_bits(got[:, 1:15].T, abi.records_to_channels(rec)), what
        np.testing.assert_array_equal(got[:, 15], rec["power"], err_msg=what)

    rec, st = oracle.transition(takeoff, box=box)
    assert st == 0 and rec["last"][-1] == 1 and 560 < len(rec) < 620     # ~1.8 m at 0.3 m/s, 100 Hz
    check(rec, T_GO1, "take-off")
    assert rec["p"][-1, 2] == c["alt"]
    rec, st = oracle.transition(goto, box=box)
    assert st == 0 and len(rec) == T_GO3 - T_GO2
    check(rec, T_GO2, "go to the start of the trajectory")
    assert golden_util.same_bits(rec["p"][-1], s[0:3, 0]) and rec["psi"][-1] == s[12, 0]
    rec, st = oracle.transition(home, box=box)
    assert st == 0 and rec["last"][-1] == 1
    check(rec, t_land, "go home")
    t_landing = t_land + len(rec)
    last = rec[-1]
    landing = transition(abi.TR_LANDING, 1.0 / c["pub_freq"], last["p"], last["v"][:2], last["psi"],
                         [0, 0, start_pose[2]], 0.0, c["vel_land_fast"], c["vel_land_slow"])
    rec, st = oracle.transition(landing, box=box)
    assert st == 0 and rec["power"][-1] == 0 and rec["power"][:-1].all()
    check(rec, t_landing, "landing")
    assert rec["p"][-1, 2] == 0.0 and (rec["clamped"][-1] & 4)            # z < 0 saturated at z_min = 0 (:602-604)
    # fast above ground + 0.4 m, slow below (:590)
    dz = -np.diff(rec["p"][:, 2])
    assert abs(dz[5] - 0.35 * 0.01) < 1e-12 and abs(dz[-5] - 0.04 * 0.01) < 1e-12


def test_transition_edge_cases(oracle):
    dt = 0.01
    # already at the destination: finished on the first tick, which still publishes the destination
    t = transition(abi.TR_GOTO, dt, [1, 2, 1.8], [0, 0], 0.5, [1, 2, 1.8], 0.5, 0.4, 0.2, 0.3, 0.2)
    rec, st = oracle.transition(t)
    assert len(rec) == 1 and st == 0 and rec["last"][0] == 1
    # yaw only, across the +-pi cut: turns the short way (wrap, :782-788)
    t = transition(abi.TR_GOTO, dt, [0, 0, 1.8], [0, 0], 3.0, [0, 0, 1.8], -3.0, 0.4, 0.2, 0.3, 0.2)
    rec, st = oracle.transition(t)
    assert (rec["dpsi"][:-2] == 0.2).all() and rec["psi"][5] > 3.0 and rec["psi"][-1] == -3.0
    # rejected parameters, and a phase that cannot end
    bad = transition(abi.TR_TAKEOFF, 0.0, [0, 0, 0], [0, 0], 0, [0, 0, 1.8], 0, 0.3, 0)
    assert oracle.transition(bad) [1] == abi.ST_BAD_PARAM
    far = transition(abi.TR_GOTO, dt, [0, 0, 1.8], [0, 0], 0, [1e9, 0, 1.8], 0, 0.4, 0.2, 0.3, 0.2)
    rec, st = oracle.transition(far, max_samples=5000)
    assert st == abi.ST_TOO_LONG and len(rec) == 5000
    # a box that the way home crosses: the saturated position feeds back into the recurrence
    t = transition(abi.TR_GOTO, dt, [0, 0, 1.8], [0, 0], 0, [3, 3, 1.8], 0, 0.4, 0.2, 0.3, 0.2, ticks=1500)
    rec, st = oracle.transition(t, box=[-5, 2, -5, 5, 0, 5])
    assert rec["p"][:, 0].max() == 2.0 and (rec["clamped"] & 1).any() and rec["p"][-1, 1] > 2.9
