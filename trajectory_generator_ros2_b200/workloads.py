"""Synthetic parameter batches for the BASELINE.json configurations.

Every generator draws trajectory i from a counter-based stream keyed by (seed, i // BLOCK), so any contiguous shard
[lo, hi) of a batch can be produced on its own rank and is identical to the same slice of the full batch: the
multi-GPU runs need no host->host or device->device parameter exchange.
"""
from __future__ import annotations

import numpy as np

from . import abi

BLOCK = 1 << 16
DT = 0.01   # pub_freq 100 Hz (default.yaml:6, TrajectoryGenerator.cpp:171-172)


def _blocks(lo: int, hi: int):
    b = lo // BLOCK
    while b * BLOCK < hi:
        s, e = max(lo, b * BLOCK), min(hi, (b + 1) * BLOCK)
        yield b, s - b * BLOCK, e - b * BLOCK
        b += 1


def _draw(seed: int, lo: int, hi: int, fill):
    """Run `fill(rng, m) -> params[m]` per BLOCK and cut the requested slice out."""
    parts = []
    for b, s, e in _blocks(lo, hi):
        rng = np.random.default_rng([seed, b])
        parts.append(fill(rng, BLOCK)[s:e])
    if not parts:
        return np.zeros(0, dtype=abi.PARAMS_DTYPE)
    return abi.concat(parts)


def _fill_circles_cfg2(rng, m, kind=abi.TGX_CIRCLE):
    p = np.zeros(m, dtype=abi.PARAMS_DTYPE)
    p["type"] = kind
    p["n_vgoals"] = 1
    p["dt"] = DT
    p["r"] = rng.uniform(0.5, 5.0, m)
    p["cx"] = rng.uniform(-2.0, 2.0, m)
    p["cy"] = rng.uniform(-2.0, 2.0, m)
    p["alt"] = rng.uniform(1.0, 2.5, m)
    v = rng.uniform(0.5, 3.0, m)
    a = rng.uniform(0.7, 2.0, m)
    p["v_goals"][:, 0] = v
    p["accel"] = a
    p["t_traj"] = 9.98 - 2.0 * v / a
    return p


def default_circle() -> np.ndarray:
    """BASELINE.json configs[0]: the single circle of config/default.yaml (:5-6, :38-44), traj_type Circle."""
    return abi.circle_params(alt=1.8, r=3.4, cx=0.0, cy=0.0, v_goals=[1.0, 2.0, 2.0], t_traj=80.0, accel=0.4, dt=DT)


def default_figure8() -> np.ndarray:
    return abi.figure8_params(alt=1.8, r=3.4, cx=0.0, cy=0.0, v_goals=[1.0, 2.0, 2.0], t_traj=80.0, accel=0.4, dt=DT)


def default_line() -> np.ndarray:
    """config/default.yaml:46-53 (Line built with z = alt for both ends, TrajectoryGenerator.cpp:283-285)."""
    return abi.line_params(alt=1.8, A=[0.0, -3.0, 1.8], B=[0.0, 3.0, 1.8], v_goals=[1.0], a1=1.5, a3=1.0, dt=DT)


def circles_cfg2(n: int, seed: int = 1234, lo: int = 0, hi: int | None = None) -> np.ndarray:
    """BASELINE.json configs[1]: random circles with ~1000 samples each (N_i in {1000, 1001}), all three phases
    present, every row fits a 1024-sample stride."""
    hi = n if hi is None else hi
    return _draw(seed, lo, hi, _fill_circles_cfg2)


def _fill_mixed_cfg3(rng, m):
    p = np.zeros(m, dtype=abi.PARAMS_DTYPE)
    kind = rng.choice([abi.TGX_CIRCLE, abi.TGX_LINE, abi.TGX_FIGURE8], size=m, p=[0.4, 0.3, 0.3])
    orb = _fill_circles_cfg2(rng, m)
    # half of the orbits get two increasing goal speeds (two ramp-ups, two holds)
    two = rng.random(m) < 0.5
    v1 = orb["v_goals"][:, 0].copy()
    v0 = v1 * rng.uniform(0.3, 0.8, m)
    a = orb["accel"]
    orb["n_vgoals"] = np.where(two, 2, 1)
    orb["v_goals"][:, 0] = np.where(two, v0, v1)
    orb["v_goals"][:, 1] = np.where(two, v1, 0.0)
    orb["t_traj"] = np.where(two, (9.98 - 2.0 * v1 / a) / 2.0, orb["t_traj"])
    p[:] = orb
    p["type"] = kind
    # lines: A, B ~ U[-4,4]^2, redrawn until the cruise segment is at least 20 % of |B - A|
    is_line = kind == abi.TGX_LINE
    idx = np.nonzero(is_line)[0]
    ln = np.zeros(len(idx), dtype=abi.PARAMS_DTYPE)
    ln["type"] = abi.TGX_LINE
    ln["n_vgoals"] = 1
    ln["dt"] = DT
    ln["alt"] = p["alt"][idx]
    v = rng.uniform(0.5, 2.0, len(idx))
    a1 = rng.uniform(0.8, 2.0, len(idx))
    a3 = rng.uniform(0.5, 1.5, len(idx))
    A = rng.uniform(-4.0, 4.0, (len(idx), 2))
    B = rng.uniform(-4.0, 4.0, (len(idx), 2))
    for _ in range(64):
        d = np.hypot(*(B - A).T)
        d2 = d - 0.5 * v * v / a1 - 0.5 * v * v / a3
        bad = d2 < 0.2 * d
        if not bad.any():
            break
        nb = int(bad.sum())
        A[bad] = rng.uniform(-4.0, 4.0, (nb, 2))
        B[bad] = rng.uniform(-4.0, 4.0, (nb, 2))
        v[bad] = rng.uniform(0.5, 2.0, nb)
    ln["A"][:, :2], ln["B"][:, :2] = A, B
    ln["A"][:, 2] = ln["alt"]
    ln["B"][:, 2] = ln["alt"]
    ln["v_goal"], ln["a1"], ln["a3"] = v, a1, a3
    p[idx] = ln
    return p


def mixed_cfg3(n: int, seed: int = 1235, lo: int = 0, hi: int | None = None) -> np.ndarray:
    """BASELINE.json configs[2]: circle 0.4 / line 0.3 / figure-eight 0.3 with one or two ramp-ups."""
    hi = n if hi is None else hi
    return _draw(seed, lo, hi, _fill_mixed_cfg3)


def _fill_montecarlo_cfg4(rng, m):
    p = _fill_circles_cfg2(rng, m)
    p["r"] = rng.uniform(0.2, 5.0, m)
    v = rng.uniform(0.2, 8.0, m)
    a = p["accel"]
    p["v_goals"][:, 0] = v
    # keep ~1000 samples per trajectory: the hold shrinks as the ramps grow, but never below 0.5 s
    p["t_traj"] = np.maximum(9.98 - 2.0 * v / a, 0.5)
    return p


def montecarlo_cfg4(n: int, seed: int = 1236, lo: int = 0, hi: int | None = None) -> np.ndarray:
    """BASELINE.json configs[3]/[4]: wide-range circles for the max-|v| / max-|a| feasibility sweep."""
    hi = n if hi is None else hi
    return _draw(seed, lo, hi, _fill_montecarlo_cfg4)


# ---- the same distribution from a counter-based generator (device side: csrc/params_gen.cu) -----------------------

def philox4x32_10(counter: np.ndarray, key) -> np.ndarray:
    """Philox4x32-10 (Salmon et al., SC'11) on an array of counters [m, 4] (uint32) with one key (k0, k1) -> [m, 4]."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = [counter[:, i].astype(np.uint64) for i in range(4)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=1).astype(np.uint32)


def _philox_uniform(x_lo: np.ndarray, x_hi: np.ndarray, a: float, b: float) -> np.ndarray:
    bits = ((x_hi.astype(np.uint64) << np.uint64(32)) | x_lo.astype(np.uint64)) >> np.uint64(11)
    u = bits.astype(np.float64) * 2.0 ** -53
    return a + (b - a) * u


def montecarlo_philox(n: int, seed: int = 1237, lo: int = 0, hi: int | None = None) -> np.ndarray:
    """BASELINE.json configs[4]: the config-4 distribution, trajectory i drawn from Philox4x32-10 with key = seed and
    counter = (i, draw).  Bit-identical to tgx_fill_montecarlo (csrc/params_gen.cu), which draws a shard on the device
    so that the 10^8-trajectory sweep needs no host->device parameter copy; this host version serves the checked
    subsets and the CPU tests."""
    hi = n if hi is None else hi
    idx = np.arange(lo, hi, dtype=np.uint64)
    m = len(idx)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    ctr = np.zeros((m, 4), dtype=np.uint32)
    ctr[:, 0] = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[:, 1] = (idx >> np.uint64(32)).astype(np.uint32)
    draws = []
    for d in range(3):
        ctr[:, 2] = d
        draws.append(philox4x32_10(ctr, key))
    a, b, c = draws
    p = np.zeros(m, dtype=abi.PARAMS_DTYPE)
    p["type"] = abi.TGX_CIRCLE
    p["n_vgoals"] = 1
    p["dt"] = DT
    p["r"] = _philox_uniform(a[:, 0], a[:, 1], 0.2, 5.0)
    p["cx"] = _philox_uniform(a[:, 2], a[:, 3], -2.0, 2.0)
    p["cy"] = _philox_uniform(b[:, 0], b[:, 1], -2.0, 2.0)
    p["alt"] = _philox_uniform(b[:, 2], b[:, 3], 1.0, 2.5)
    v = _philox_uniform(c[:, 0], c[:, 1], 0.2, 8.0)
    acc = _philox_uniform(c[:, 2], c[:, 3], 0.7, 2.0)
    p["accel"] = acc
    p["v_goals"][:, 0] = v
    p["t_traj"] = np.maximum(9.98 - (2.0 * v) / acc, 0.5)
    return p


MONTECARLO_LIMITS = dict(box=(-5.0, 5.0, -5.0, 5.0, -5.0, 5.0), v_max=5.0, a_max=6.0)


# ---- constant-speed polyline family -------------------------------------------------------------

def default_polyline(kind: int) -> np.ndarray:
    """The polyline-family trajectories of config/default.yaml (:9-36, :40-44, :46-49): traj_type T ships as the default;
    v_goals [1.0, 2.0, 2.0] (only [0] is used), t_traj 80, orientation 0, centre (0, 0), alt 1.8."""
    alt, vg, T, ori = 1.8, [1.0, 2.0, 2.0], 80.0, 0.0
    if kind == abi.TGX_SQUARE:
        return abi.square_params(alt, 2.0, 0.0, 0.0, ori, vg, T, 0.4, DT)
    if kind == abi.TGX_RECTANGLE:
        return abi.rectangle_params(alt, 2.0, 4.0, 0.0, 0.0, ori, vg, T, 0.4, DT)
    if kind == abi.TGX_RECIPROCATING:
        return abi.reciprocating_params(alt, [0.0, -3.0, alt], [0.0, 3.0, alt], [1.0], 1.5, 1.0, T, DT)
    if kind == abi.TGX_BOUNCE:
        return abi.bounce_params(0.0, 0.0, 4.0, 1.0, vg, T, ori, DT)
    return abi.letter_params(kind, 0.0, 0.0, 3.0, 4.0, alt, vg, T, ori, DT)


def _fill_polyline(rng, m, t_lo=8.0, t_hi=12.0):
    """Random polyline-family trajectories of ~1000 samples (t_traj ~ U[8, 12] s at dt = 0.01), all seven shapes."""
    p = np.zeros(m, dtype=abi.PARAMS_DTYPE)
    kind = rng.choice(np.array(abi.POLYLINE_TYPES), size=m)
    p["type"] = kind
    p["dt"] = DT
    p["alt"] = rng.uniform(1.0, 2.5, m)
    p["poly_t_traj"] = rng.uniform(t_lo, t_hi, m)
    p["poly_v_goal"] = rng.uniform(0.3, 3.0, m)
    p["poly_decel"] = rng.uniform(0.3, 2.0, m)
    p["orientation"] = np.where(rng.random(m) < 0.25, 0.0, rng.uniform(-np.pi, np.pi, m))
    cx, cy = rng.uniform(-2.0, 2.0, m), rng.uniform(-2.0, 2.0, m)
    d0, d1 = rng.uniform(0.5, 4.0, m), rng.uniform(0.5, 4.0, m)
    g = np.zeros((m, 7))
    sq = kind == abi.TGX_SQUARE
    g[sq, 0], g[sq, 1], g[sq, 2] = d0[sq], cx[sq], cy[sq]
    rc = kind == abi.TGX_RECTANGLE
    g[rc, 0], g[rc, 1], g[rc, 2], g[rc, 3] = d0[rc], d1[rc], cx[rc], cy[rc]
    rp = kind == abi.TGX_RECIPROCATING
    A = rng.uniform(-4.0, 4.0, (m, 2))
    B = A + rng.uniform(0.5, 4.0, (m, 1)) * np.stack([np.cos(p["orientation"]), np.sin(p["orientation"])], axis=1)
    g[rp, 0], g[rp, 1], g[rp, 2] = A[rp, 0], A[rp, 1], p["alt"][rp]
    g[rp, 3], g[rp, 4], g[rp, 5] = B[rp, 0], B[rp, 1], p["alt"][rp]
    bo = kind == abi.TGX_BOUNCE
    az = rng.uniform(0.5, 2.0, m)
    g[bo, 0], g[bo, 1], g[bo, 2], g[bo, 3] = cx[bo], cy[bo], az[bo], az[bo] + d0[bo]
    flip = bo & (rng.random(m) < 0.5)                 # default.yaml starts at the top (Az 4.0, Bz 1.0)
    g[flip, 2], g[flip, 3] = g[flip, 3], g[flip, 2].copy()
    lt = (kind == abi.TGX_M) | (kind == abi.TGX_I) | (kind == abi.TGX_T)
    g[lt, 0], g[lt, 1], g[lt, 2], g[lt, 3] = cx[lt], cy[lt], d0[lt], d1[lt]
    p["g"] = g
    return p


def polyline_mix(n: int, seed: int = 1238, lo: int = 0, hi: int | None = None) -> np.ndarray:
    """Row f2 workload: the seven constant-speed shapes, ~1000 samples each.  cos_o / sin_o are left unset: the
    host-buffer calls (or Engine.finalize_polyline) fill them from the host libm."""
    hi = n if hi is None else hi
    return _draw(seed, lo, hi, _fill_polyline)


def _fill_letters_T(rng, m):
    """The shipped default shape (traj_type: T, default.yaml:9) with random size / speed / pose, ~1000 samples."""
    p = np.zeros(m, dtype=abi.PARAMS_DTYPE)
    p["type"] = abi.TGX_T
    p["dt"] = DT
    p["alt"] = rng.uniform(1.0, 2.5, m)
    p["poly_t_traj"] = rng.uniform(9.9, 10.1, m)
    p["poly_v_goal"] = rng.uniform(0.3, 3.0, m)
    p["poly_decel"] = 1.0
    p["orientation"] = rng.uniform(-np.pi, np.pi, m)
    g = np.zeros((m, 7))
    g[:, 0], g[:, 1] = rng.uniform(-2.0, 2.0, m), rng.uniform(-2.0, 2.0, m)
    g[:, 2], g[:, 3] = rng.uniform(0.5, 4.0, m), rng.uniform(0.5, 4.0, m)
    p["g"] = g
    return p


def letters_T(n: int, seed: int = 1239, lo: int = 0, hi: int | None = None) -> np.ndarray:
    hi = n if hi is None else hi
    return _draw(seed, lo, hi, _fill_letters_T)


# ---- node-side transitions ------------------------------------------------------------------------

def fleet_transitions(n: int, seed: int = 1240) -> np.ndarray:
    """A fleet's worth of tgx_transition_params: per vehicle one of take-off (ground -> alt at vel_take), the trip to the
    start of its trajectory (simpleInterpolation at vel_initpos, default.yaml:56-63 speeds and thresholds) or landing
    from the hover altitude; a few hundred to a few thousand 100 Hz ticks each."""
    rng = np.random.default_rng(seed)
    t = np.zeros(n, dtype=abi.TRANSITION_DTYPE)
    kind = rng.integers(0, 3, n)
    t["kind"] = kind
    t["dt"] = DT
    alt = rng.uniform(1.0, 2.5, n)
    t["start"][:, :2] = rng.uniform(-4.0, 4.0, (n, 2))
    t["start"][:, 2] = np.where(kind == abi.TR_TAKEOFF, 0.0, alt)
    t["start_psi"] = rng.uniform(-3.1, 3.1, n)
    t["dest"][:, :2] = rng.uniform(-4.0, 4.0, (n, 2))
    t["dest"][:, 2] = np.where(kind == abi.TR_LANDING, 0.0, alt)
    t["dest_yaw"] = rng.uniform(-3.1, 3.1, n)
    t["vel"] = np.where(kind == abi.TR_TAKEOFF, 0.3, np.where(kind == abi.TR_GOTO, 0.4, 0.35))
    t["vel_yaw"] = np.where(kind == abi.TR_LANDING, 0.04, 0.2)
    t["dist_thresh"] = 0.3
    t["yaw_thresh"] = 0.2
    return t
