"""ctypes loaders for the CPU checkers (oracle/liboracle.so and, when built, oracle/_ref/libtrajref.so).

Test infrastructure only.  ``oracle`` is the plain-C restatement; ``ref`` is the UNMODIFIED reference
compiled behind stub headers (present only if oracle/Makefile could see /root/reference at build time, or
the prebuilt .so travelled with the repo snapshot).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from trajectory_generator_ros2_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libtrajref.so")

_c_i64 = C.c_int64
_c_dp = C.POINTER(C.c_double)


def build_oracle() -> None:
    """(Re)build oracle/liboracle.so and, if /root/reference exists, oracle/_ref/libtrajref.so."""
    subprocess.run(["make", "-C", ORACLE_DIR, "--no-print-directory"], check=True, capture_output=True)


def _ptr(a, ctype=C.c_double):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ctype))


class Oracle:
    """Thin wrapper over liboracle.so."""

    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build_oracle()
        self.lib = C.CDLL(path)
        L = self.lib
        L.orc_generate.restype = _c_i64
        L.orc_generate.argtypes = [C.c_void_p, _c_dp, _c_i64, _c_i64, C.POINTER(C.c_uint32), C.c_void_p, _c_i64]
        L.orc_stop.restype = _c_i64
        L.orc_stop.argtypes = [C.c_void_p, _c_dp, _c_dp, _c_i64, _c_i64, C.POINTER(C.c_uint32), C.c_void_p, _c_i64]
        L.orc_inside_bounds.restype = C.c_int
        L.orc_inside_bounds.argtypes = [C.c_void_p, _c_dp]
        L.orc_line_d2.restype = C.c_double
        L.orc_line_d2.argtypes = [C.c_void_p]
        L.orc_generate_batch.restype = C.c_int
        L.orc_generate_batch.argtypes = [C.c_void_p, _c_i64, _c_dp, _c_i64, _c_i64, _c_i64,
                                         C.POINTER(C.c_int32), C.POINTER(C.c_uint32), _c_i64, C.c_int]
        L.orc_feasibility_batch.restype = C.c_int
        L.orc_feasibility_batch.argtypes = [C.c_void_p, _c_i64, C.c_void_p, C.POINTER(C.c_uint8), _c_dp, _c_dp,
                                            C.POINTER(C.c_int32), C.POINTER(C.c_uint32), _c_i64, C.c_int]
        L.orc_time_batch.restype = _c_i64
        L.orc_time_batch.argtypes = [C.c_void_p, _c_i64, _c_i64, C.c_int, _c_dp]
        L.orc_polyline_generate.restype = _c_i64
        L.orc_polyline_generate.argtypes = [C.c_void_p, _c_dp, _c_i64, _c_i64, C.POINTER(C.c_uint32),
                                            C.POINTER(C.c_int16), _c_i64]
        L.orc_polyline_msg.restype = C.c_int
        L.orc_polyline_msg.argtypes = [C.c_int, C.c_int, _c_i64, _c_i64, C.c_char_p, C.c_int]
        L.orc_pack_goals.restype = None
        L.orc_pack_goals.argtypes = [_c_dp, _c_i64, _c_i64, C.c_int32, _c_dp, C.c_void_p]
        L.orc_transition.restype = _c_i64
        L.orc_transition.argtypes = [C.c_void_p, C.c_int32, _c_dp, C.c_void_p, _c_i64, C.POINTER(C.c_uint32), _c_i64]
        L.orc_fnv1a64.restype = C.c_uint64
        L.orc_fnv1a64.argtypes = [_c_dp, _c_i64, C.c_uint64]

    # -- single trajectory ---------------------------------------------------------------------------
    @staticmethod
    def _rows(p: np.ndarray) -> int:
        """Records the trajectory p[0] occupies; the C side trusts that they are all there."""
        p = np.ascontiguousarray(p)
        k = int(p["n_vgoals"][0])
        rows = abi.orbit_records(k) if int(p["type"][0]) in (abi.TGX_CIRCLE, abi.TGX_FIGURE8) and \
            0 <= k <= abi.TGX_MAX_VGOALS_TOTAL else 1
        assert len(p) >= rows and p.flags.c_contiguous, "pass the continuation records together with the first one"
        return rows

    def count(self, p: np.ndarray, max_samples: int = abi.DEFAULT_MAX_SAMPLES):
        self._rows(p)
        st = C.c_uint32(0)
        n = self.lib.orc_generate(p.ctypes.data, None, 0, 0, C.byref(st), None, max_samples)
        return int(n), int(st.value)

    def generate(self, p: np.ndarray, max_samples: int = abi.DEFAULT_MAX_SAMPLES):
        """-> (samples[14, N], status, phases-record)."""
        n, st = self.count(p, max_samples)
        rows = self._rows(p)
        ph = np.zeros(rows, dtype=abi.PHASES_DTYPE)
        if n <= 0:
            return np.zeros((abi.TGX_NCHAN, 0)), st, ph[0]
        out = np.full((abi.TGX_NCHAN, n), np.nan)
        st2 = C.c_uint32(0)
        n2 = self.lib.orc_generate(p.ctypes.data, _ptr(out), n, n, C.byref(st2), ph.ctypes.data, max_samples)
        assert n2 == n
        return out, int(st2.value), (ph[0] if rows == 1 else ph)

    def polyline_generate(self, p: np.ndarray, max_samples: int = abi.DEFAULT_MAX_SAMPLES):
        """Polyline family -> (samples[14, N], status, leg_of[N] int16, index_msgs dict)."""
        st = C.c_uint32(0)
        n = self.lib.orc_polyline_generate(p.ctypes.data, None, 0, 0, C.byref(st), None, max_samples)
        if n <= 0:
            return np.zeros((abi.TGX_NCHAN, 0)), int(st.value), np.zeros(0, dtype=np.int16), {}
        out = np.full((abi.TGX_NCHAN, n), np.nan)
        legs = np.zeros(n, dtype=np.int16)
        n2 = self.lib.orc_polyline_generate(p.ctypes.data, _ptr(out), n, n, C.byref(st), _ptr(legs, C.c_int16),
                                            max_samples)
        assert n2 == n
        buf = C.create_string_buffer(128)
        msgs = {}
        t = int(p["type"][0])
        for k in range(n):
            self.lib.orc_polyline_msg(t, int(legs[k]), k, n, buf, 128)
            msgs[k] = buf.value.decode()
        return out, int(st.value), legs, msgs

    def stop(self, p: np.ndarray, from14: np.ndarray, max_samples: int = abi.DEFAULT_MAX_SAMPLES):
        from14 = np.ascontiguousarray(from14, dtype=np.float64)
        st = C.c_uint32(0)
        n = self.lib.orc_stop(p.ctypes.data, _ptr(from14), None, 0, 0, C.byref(st), None, max_samples)
        ph = np.zeros(1, dtype=abi.PHASES_DTYPE)
        if n <= 0:
            self.lib.orc_stop(p.ctypes.data, _ptr(from14), None, 0, 0, C.byref(st), ph.ctypes.data, max_samples)
            return np.zeros((abi.TGX_NCHAN, 0)), int(st.value), ph[0]
        out = np.full((abi.TGX_NCHAN, n), np.nan)
        self.lib.orc_stop(p.ctypes.data, _ptr(from14), _ptr(out), n, n, C.byref(st), ph.ctypes.data, max_samples)
        return out, int(st.value), ph[0]

    def pack_goals(self, samples: np.ndarray, traj: int = 0, box=None) -> np.ndarray:
        """samples [14, N] -> tgx_goal_record array [N]: what the node publishes while following (saturated to box)."""
        samples = np.ascontiguousarray(samples, dtype=np.float64)
        n = samples.shape[1]
        out = np.zeros(n, dtype=abi.RECORD_DTYPE)
        b = np.asarray(box, dtype=np.float64) if box is not None else None
        self.lib.orc_pack_goals(_ptr(samples), n, n, traj, _ptr(b) if b is not None else None, out.ctypes.data)
        return out

    def transition(self, t: np.ndarray, traj: int = 0, box=None, max_samples: int = abi.DEFAULT_MAX_SAMPLES):
        """One tgx_transition_params record -> (records [N] of RECORD_DTYPE, status)."""
        assert t.dtype == abi.TRANSITION_DTYPE and t.shape == (1,)
        b = np.asarray(box, dtype=np.float64) if box is not None else None
        st = C.c_uint32(0)
        n = self.lib.orc_transition(t.ctypes.data, traj, _ptr(b) if b is not None else None, None, 0, C.byref(st),
                                    max_samples)
        out = np.zeros(n, dtype=abi.RECORD_DTYPE)
        if n:
            self.lib.orc_transition(t.ctypes.data, traj, _ptr(b) if b is not None else None, out.ctypes.data, n,
                                    C.byref(st), max_samples)
        return out, int(st.value)

    def inside_bounds(self, p: np.ndarray, box) -> bool:
        b = np.asarray(box, dtype=np.float64)
        return bool(self.lib.orc_inside_bounds(p.ctypes.data, _ptr(b)))

    def line_d2(self, p: np.ndarray) -> float:
        return float(self.lib.orc_line_d2(p.ctypes.data))

    # -- batches -------------------------------------------------------------------------------------
    def count_batch(self, params: np.ndarray, nthreads: int = 8, max_samples: int = abi.DEFAULT_MAX_SAMPLES):
        n = len(params)
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.uint32)
        self.lib.orc_generate_batch(params.ctypes.data, n, None, 0, 0, 0, _ptr(counts, C.c_int32),
                                    _ptr(status, C.c_uint32), max_samples, nthreads)
        return counts, status

    def generate_batch(self, params: np.ndarray, cap: int, nthreads: int = 8,
                       max_samples: int = abi.DEFAULT_MAX_SAMPLES):
        """-> (out[n, 14, cap] (NaN padded), counts, status)."""
        n = len(params)
        out = np.full((n, abi.TGX_NCHAN, cap), np.nan)
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.uint32)
        self.lib.orc_generate_batch(params.ctypes.data, n, _ptr(out), abi.TGX_NCHAN * cap, cap, cap,
                                    _ptr(counts, C.c_int32), _ptr(status, C.c_uint32), max_samples, nthreads)
        return out, counts, status

    def feasibility_batch(self, params: np.ndarray, limits: abi.Limits, nthreads: int = 8,
                          max_samples: int = abi.DEFAULT_MAX_SAMPLES):
        n = len(params)
        flags = np.zeros(n, dtype=np.uint8)
        mv = np.zeros(n)
        ma = np.zeros(n)
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.uint32)
        self.lib.orc_feasibility_batch(params.ctypes.data, n, C.byref(limits), _ptr(flags, C.c_uint8), _ptr(mv),
                                       _ptr(ma), _ptr(counts, C.c_int32), _ptr(status, C.c_uint32), max_samples,
                                       nthreads)
        return flags, mv, ma, counts, status

    def time_batch(self, params: np.ndarray, nthreads: int, max_samples: int = abi.DEFAULT_MAX_SAMPLES):
        cs = C.c_double(0.0)
        total = self.lib.orc_time_batch(params.ctypes.data, len(params), max_samples, nthreads, C.byref(cs))
        return int(total), float(cs.value)

    def fnv(self, samples: np.ndarray) -> int:
        """FNV-1a-64 over the 14 channels in order (samples[14, N])."""
        h = 0xcbf29ce484222325
        for c in range(samples.shape[0]):
            row = np.ascontiguousarray(samples[c], dtype=np.float64)
            h = self.lib.orc_fnv1a64(_ptr(row), row.size, h)
        return int(h)


class Reference:
    """Thin wrapper over oracle/_ref/libtrajref.so (the unmodified reference)."""

    MSG_CAP = 1 << 22   # the polyline family announces every sample

    def __init__(self, path: str = REF_SO):
        self.lib = C.CDLL(path)
        L = self.lib
        L.ref_generate.restype = _c_i64
        L.ref_generate.argtypes = [C.c_void_p, _c_dp, _c_i64, _c_i64, C.POINTER(C.c_uint32), C.c_char_p, _c_i64]
        L.ref_stop.restype = _c_i64
        L.ref_stop.argtypes = [C.c_void_p, _c_dp, _c_dp, _c_i64, _c_i64, C.POINTER(C.c_uint32), C.c_char_p, _c_i64]
        L.ref_inside_bounds.restype = C.c_int
        L.ref_inside_bounds.argtypes = [C.c_void_p, _c_dp]
        L.ref_generate_batch.restype = C.c_int
        L.ref_generate_batch.argtypes = [C.c_void_p, _c_i64, _c_dp, _c_i64, _c_i64, _c_i64,
                                         C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.c_int]
        L.ref_feasibility_batch.restype = C.c_int
        L.ref_feasibility_batch.argtypes = [C.c_void_p, _c_i64, C.c_void_p, C.POINTER(C.c_uint8), _c_dp, _c_dp,
                                            C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.c_int]
        L.ref_time_batch.restype = _c_i64
        L.ref_time_batch.argtypes = [C.c_void_p, _c_i64, C.c_int, _c_dp]

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    @staticmethod
    def _parse_msgs(buf) -> dict:
        out = {}
        for line in buf.value.decode().splitlines():
            k, _, m = line.partition("\t")
            out[int(k)] = m
        return out

    def generate(self, p: np.ndarray):
        """-> (samples[14, N], status, index_msgs dict)."""
        st = C.c_uint32(0)
        n = self.lib.ref_generate(p.ctypes.data, None, 0, 0, C.byref(st), None, 0)
        if n <= 0:
            return np.zeros((abi.TGX_NCHAN, 0)), int(st.value), {}
        out = np.full((abi.TGX_NCHAN, n), np.nan)
        buf = C.create_string_buffer(self.MSG_CAP)
        n2 = self.lib.ref_generate(p.ctypes.data, _ptr(out), n, n, C.byref(st), buf, self.MSG_CAP)
        assert n2 == n
        return out, int(st.value), self._parse_msgs(buf)

    def stop(self, p: np.ndarray, from14: np.ndarray):
        from14 = np.ascontiguousarray(from14, dtype=np.float64)
        st = C.c_uint32(0)
        buf = C.create_string_buffer(self.MSG_CAP)
        n = self.lib.ref_stop(p.ctypes.data, _ptr(from14), None, 0, 0, C.byref(st), buf, self.MSG_CAP)
        if n <= 0:
            return np.zeros((abi.TGX_NCHAN, 0)), int(st.value), self._parse_msgs(buf)
        out = np.full((abi.TGX_NCHAN, n), np.nan)
        self.lib.ref_stop(p.ctypes.data, _ptr(from14), _ptr(out), n, n, C.byref(st), buf, self.MSG_CAP)
        return out, int(st.value), self._parse_msgs(buf)

    def inside_bounds(self, p: np.ndarray, box) -> bool:
        b = np.asarray(box, dtype=np.float64)
        return bool(self.lib.ref_inside_bounds(p.ctypes.data, _ptr(b)))

    def generate_batch(self, params: np.ndarray, cap: int, nthreads: int = 8):
        n = len(params)
        out = np.full((n, abi.TGX_NCHAN, cap), np.nan)
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.uint32)
        self.lib.ref_generate_batch(params.ctypes.data, n, _ptr(out), abi.TGX_NCHAN * cap, cap, cap,
                                    _ptr(counts, C.c_int32), _ptr(status, C.c_uint32), nthreads)
        return out, counts, status

    def count_batch(self, params: np.ndarray, nthreads: int = 8):
        n = len(params)
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.uint32)
        self.lib.ref_generate_batch(params.ctypes.data, n, None, 0, 0, 0, _ptr(counts, C.c_int32),
                                    _ptr(status, C.c_uint32), nthreads)
        return counts, status

    def feasibility_batch(self, params: np.ndarray, limits: abi.Limits, nthreads: int = 8):
        n = len(params)
        flags = np.zeros(n, dtype=np.uint8)
        mv = np.zeros(n)
        ma = np.zeros(n)
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.uint32)
        self.lib.ref_feasibility_batch(params.ctypes.data, n, C.byref(limits), _ptr(flags, C.c_uint8), _ptr(mv),
                                       _ptr(ma), _ptr(counts, C.c_int32), _ptr(status, C.c_uint32), nthreads)
        return flags, mv, ma, counts, status

    def time_batch(self, params: np.ndarray, nthreads: int):
        cs = C.c_double(0.0)
        total = self.lib.ref_time_batch(params.ctypes.data, len(params), nthreads, C.byref(cs))
        return int(total), float(cs.value)
