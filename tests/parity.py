"""Tolerances of the parity gate (BASELINE.json north_star; SURVEY.md §8d "Parity check at scale").

  position           |dp|  <= 1e-9 m (absolute, per component)
  v, a, j vectors    ||d_k|| <= 1e-8 * max(||ref_k||, 1e-3 * max_k ||ref_k||, 1e-6)
                     relative to the vector magnitude (components cross zero), and where the magnitude itself goes
                     through zero (a trajectory starts and ends at rest; the figure-eight's acceleration vanishes twice
                     per lap) relative to 0.1 % of the trajectory's own peak magnitude
  yaw (psi)          wrapped |dpsi| <= 1e-8 * max(|psi_ref|, 1)
  yaw rate (dpsi)    |d_k| <= 1e-8 * max(|ref_k|, 1e-3 * max_k |ref_k|, 1e-6)
  -0.0 == +0.0; sample counts, status bits and index_msgs keys: exact.
"""
from __future__ import annotations

import numpy as np

from trajectory_generator_ros2_b200 import abi

POS_ATOL = 1e-9
REL = 1e-8
PEAK_FLOOR = 1e-3
ABS_FLOOR = 1e-6


def sample_errors(got: np.ndarray, ref: np.ndarray) -> dict:
    """got, ref: [14, N].  Returns the worst normalised error per group (<= 1 passes) and raw maxima."""
    assert got.shape == ref.shape, (got.shape, ref.shape)
    out = {}
    dp = np.abs(got[abi.PX:abi.PZ + 1] - ref[abi.PX:abi.PZ + 1])
    out["pos_abs"] = float(dp.max(initial=0.0))
    out["pos"] = out["pos_abs"] / POS_ATOL
    for name, lo in (("v", abi.VX), ("a", abi.AX), ("j", abi.JX)):
        d = np.linalg.norm(got[lo:lo + 3] - ref[lo:lo + 3], axis=0)
        mag = np.linalg.norm(ref[lo:lo + 3], axis=0)
        mag = np.maximum(np.maximum(mag, PEAK_FLOOR * mag.max(initial=0.0)), ABS_FLOOR)
        rel = d / mag
        out[name + "_rel"] = float(rel.max(initial=0.0))
        out[name] = out[name + "_rel"] / REL
    dpsi = got[abi.PSI] - ref[abi.PSI]
    wrapped = np.abs(np.arctan2(np.sin(dpsi), np.cos(dpsi)))
    # exact arithmetic for tiny differences: arctan2(sin, cos) loses nothing near 0
    scale = np.maximum(np.abs(ref[abi.PSI]), 1.0)
    out["psi_abs"] = float(wrapped.max(initial=0.0))
    out["psi"] = float((wrapped / scale).max(initial=0.0)) / REL
    mag = np.abs(ref[abi.DPSI])
    mag = np.maximum(np.maximum(mag, PEAK_FLOOR * mag.max(initial=0.0)), ABS_FLOOR)
    dd = np.abs(got[abi.DPSI] - ref[abi.DPSI]) / mag
    out["dpsi_rel"] = float(dd.max(initial=0.0))
    out["dpsi"] = out["dpsi_rel"] / REL
    return out


def assert_samples_close(got: np.ndarray, ref: np.ndarray, what: str = ""):
    assert not np.isnan(got).any(), f"{what}: NaN in GPU samples"
    e = sample_errors(got, ref)
    bad = {k: v for k, v in e.items() if k in ("pos", "v", "a", "j", "psi", "dpsi") and v > 1.0}
    assert not bad, f"{what}: out of tolerance {bad} (all: {e})"
    return e


def merge_errors(acc: dict, e: dict) -> dict:
    for k, v in e.items():
        acc[k] = max(acc.get(k, 0.0), v)
    return acc
