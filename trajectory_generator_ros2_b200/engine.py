"""ctypes binding of libtgx.so (include/tgx.h) plus a small torch-based convenience layer.

torch is used only as plumbing: device memory (tensors), the current CUDA stream and, in ``sharding.py``,
``torch.distributed``.  All sampling runs in the hand-written CUDA kernels behind the C-ABI; there is no Python
or CPU implementation of any sampler in this package, and importing this module fails loudly when the built
library is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TGX_LIB", os.path.join(_HERE, "libtgx.so"))   # TGX_LIB: A/B builds in tools/ experiments


class TgxError(RuntimeError):
    def __init__(self, code: int, what: str, detail: str = ""):
        self.code = code
        super().__init__(f"{what}: tgx error {code}" + (f" ({detail})" if detail else ""))


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C trajectory_generator_ros2_b200/csrc`. There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    lib.tgx_version.restype = C.c_int
    lib.tgx_strerror.restype = C.c_char_p
    lib.tgx_strerror.argtypes = [C.c_int]
    lib.tgx_last_cuda_error.restype = C.c_char_p
    lib.tgx_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.tgx_destroy.argtypes = [vp]
    lib.tgx_set_max_samples.argtypes = [vp, i64]
    lib.tgx_set_tuning.argtypes = [vp, C.c_int, C.c_int]
    lib.tgx_set_plan_mode.argtypes = [vp, C.c_int]
    lib.tgx_set_host_fill.argtypes = [vp, C.c_int]
    lib.tgx_set_host_layout.argtypes = [vp, C.c_int]
    lib.tgx_set_phase_planning.argtypes = [vp, C.c_int]
    lib.tgx_set_store_path.argtypes = [vp, C.c_int]
    lib.tgx_phase_plan_count.restype = i64
    lib.tgx_phase_plan_count.argtypes = [vp]
    lib.tgx_set_slab_planning.argtypes = [vp, C.c_int]
    lib.tgx_plan_path_counts.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    lib.tgx_scratch_bytes.restype = i64
    lib.tgx_scratch_bytes.argtypes = [vp]
    lib.tgx_launch_count.restype = i64
    lib.tgx_launch_count.argtypes = [vp]
    lib.tgx_plan_tiles.restype = i64
    lib.tgx_plan_tiles.argtypes = [vp]
    lib.tgx_plan_segments.restype = i64
    lib.tgx_plan_segments.argtypes = [vp]
    lib.tgx_count.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    lib.tgx_plan.argtypes = [vp, vp, i64, vp, vp, vp, vp, C.POINTER(i64), vp]
    lib.tgx_plan_stop.argtypes = [vp, vp, i64, vp, vp, vp, vp, C.POINTER(i64), vp]
    lib.tgx_plan_polyline.argtypes = [vp, vp, i64, vp, vp, vp, vp, C.POINTER(i64), vp]
    lib.tgx_polyline_finalize_host.argtypes = [vp, i64]
    lib.tgx_generate_host_legs.argtypes = [vp, vp, i64, vp, vp, i64, vp, vp, vp, vp]
    lib.tgx_pack_goals.argtypes = [vp, C.POINTER(abi.Layout), vp, i64, vp, vp, i64, vp, i64, vp]
    lib.tgx_eval_records.argtypes = [vp, vp, vp, i64, vp, i64, vp]
    lib.tgx_generate_records_host.argtypes = [vp, vp, i64, vp, vp, i64, vp, vp]
    lib.tgx_transitions.argtypes = [vp, vp, i64, vp, vp, i64, i64, vp, vp, vp]
    lib.tgx_transitions_host.argtypes = [vp, vp, i64, vp, vp, i64, vp, vp]
    lib.tgx_eval.argtypes = [vp, C.POINTER(abi.Layout), vp, vp, vp]
    lib.tgx_set_generate_profiling.argtypes = [vp, C.c_int]
    lib.tgx_generate_profile.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64)]
    lib.tgx_generate.argtypes = [vp, vp, i64, vp, C.POINTER(abi.Layout), vp, vp, vp, i64, C.POINTER(i64), vp]
    lib.tgx_generate_feasibility.argtypes = [vp, vp, i64, C.POINTER(abi.Limits), vp, vp, vp, vp, i64, C.POINTER(i64), vp]
    lib.tgx_feasibility.argtypes = [vp, C.POINTER(abi.Limits), vp, vp, vp, vp, vp]
    lib.tgx_count_host.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.tgx_generate_host.argtypes = [vp, vp, i64, vp, vp, i64, vp, vp, vp]
    lib.tgx_stop_host.argtypes = [vp, vp, i64, vp, vp, i64, vp, vp, vp]
    lib.tgx_shard_range.argtypes = [i64, i32, i32, C.POINTER(i64), C.POINTER(i64)]
    lib.tgx_plan_samples.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    lib.tgx_sample_host.argtypes = [vp, vp, C.c_double, C.c_double, C.c_double, C.c_double, vp]
    lib.tgx_selftest_division.argtypes = [vp, i64, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
    lib.tgx_probe_dfma.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.tgx_probe_d2h.argtypes = [vp, vp, i64, C.c_int, C.POINTER(C.c_double)]
    lib.tgx_fill_montecarlo.argtypes = [vp, C.c_uint64, i64, i64, vp, vp]
    lib.tgx_comm_last_error.restype = C.c_char_p
    lib.tgx_comm_nccl_version.argtypes = [C.POINTER(C.c_int)]
    lib.tgx_comm_unique_id.argtypes = [C.c_char_p]
    lib.tgx_comm_init_rank.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_char_p, C.c_int]
    lib.tgx_comm_init_all.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(C.c_int)]
    lib.tgx_comm_destroy.argtypes = [vp]
    lib.tgx_gather_flags.argtypes = [vp, vp, i64, vp, vp]
    lib.tgx_alloc_host.restype = vp
    lib.tgx_alloc_host.argtypes = [i64]
    lib.tgx_alloc_host_for.restype = vp
    lib.tgx_alloc_host_for.argtypes = [vp, i64]
    lib.tgx_host_info.argtypes = [vp, C.POINTER(abi.HostInfo)]
    lib.tgx_generate_host_compact.argtypes = [vp, vp, i64, vp, vp, i64, vp, vp, vp, vp]
    lib.tgx_free_host.argtypes = [vp]
    return lib


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def shard_range(n: int, rank: int, world: int):
    """Contiguous block partition of a batch (tgx_shard_range)."""
    lo, hi = C.c_int64(0), C.c_int64(0)
    rc = lib().tgx_shard_range(n, rank, world, C.byref(lo), C.byref(hi))
    if rc:
        raise TgxError(rc, "tgx_shard_range", lib().tgx_strerror(rc).decode())
    return int(lo.value), int(hi.value)


def _limits_ptr(limits: Optional[abi.Limits]):
    return C.cast(C.pointer(limits), C.c_void_p) if limits is not None else None


@dataclass
class Plan:
    """What tgx_plan returned for a batch (tensors live on the engine's device)."""
    n: int
    counts: "object"          # torch.int32 [n]
    status: "object"          # torch.int32 [n] (bit pattern of the uint32 status)
    total_samples: int
    phases: "object" = None   # torch.uint8 [n, sizeof(tgx_phases)] or None
    tiles: int = 0
    segments: int = 0
    legs: "object" = None     # torch.uint8 [n, sizeof(tgx_polyline_legs)] or None (polyline plans)


class PinnedArray:
    """A numpy view over page-locked host memory obtained from tgx_alloc_host."""

    def __init__(self, shape, dtype=np.float64, engine: "Engine" = None):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        if engine is not None:      # pages preferably from the NUMA node of the engine's GPU
            self.ptr = lib().tgx_alloc_host_for(engine._h, max(self.nbytes, 1))
        else:
            self.ptr = lib().tgx_alloc_host(max(self.nbytes, 1))
        if not self.ptr:
            raise MemoryError(f"tgx_alloc_host({self.nbytes}) failed")
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().tgx_free_host(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Comm:
    """One rank of the flag all-gather (tgx_comm: an NCCL communicator behind the C-ABI, include/tgx.h).

    torch.distributed is only the courier of the 128-byte NCCL unique id; the collective itself is issued by
    libtgx (tgx_gather_flags -> ncclAllGather) on the caller's CUDA stream.
    """

    def __init__(self, handle, world: int, rank: int, device: int):
        self._h, self.world, self.rank, self.device = handle, world, rank, device

    @staticmethod
    def _raise(rc: int, what: str):
        text = lib().tgx_comm_last_error().decode() if rc == abi.TGX_ERR_COMM else lib().tgx_strerror(rc).decode()
        raise TgxError(rc, what, text)

    @staticmethod
    def nccl_version() -> int:
        v = C.c_int(0)
        rc = lib().tgx_comm_nccl_version(C.byref(v))
        if rc:
            Comm._raise(rc, "tgx_comm_nccl_version")
        return int(v.value)

    @classmethod
    def from_torch_distributed(cls, device: int, group=None) -> "Comm":
        """One process per GPU: rank 0 draws the NCCL unique id, torch.distributed ships it, every rank joins."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        box = [None]
        if rank == 0:
            buf = C.create_string_buffer(abi.TGX_COMM_ID_BYTES)
            rc = lib().tgx_comm_unique_id(buf)
            if rc:
                cls._raise(rc, "tgx_comm_unique_id")
            box[0] = buf.raw
        dist.broadcast_object_list(box, src=0, group=group)
        h = C.c_void_p()
        rc = lib().tgx_comm_init_rank(C.byref(h), world, rank, box[0], device)
        if rc:
            cls._raise(rc, "tgx_comm_init_rank")
        return cls(h, world, rank, device)

    def gather_flags(self, local_flags, n_total: int, out=None):
        """tgx_gather_flags: this rank's uint8 shard -> the full [n_total] vector on every rank (current stream)."""
        import torch
        lo, hi = shard_range(n_total, self.rank, self.world)
        assert local_flags.dtype == torch.uint8 and local_flags.is_contiguous() and local_flags.numel() == hi - lo
        if out is None:
            out = torch.empty(n_total, dtype=torch.uint8, device=local_flags.device)
        rc = lib().tgx_gather_flags(self._h, local_flags.data_ptr(), n_total, out.data_ptr(),
                                    int(torch.cuda.current_stream().cuda_stream))
        if rc:
            self._raise(rc, "tgx_gather_flags")
        return out

    def close(self):
        if self._h:
            lib().tgx_comm_destroy(self._h)
            self._h = None


class Engine:
    """One tgx_engine bound to one GPU."""

    def __init__(self, device: int = 0):
        self._lib = lib()
        self._h = C.c_void_p()
        rc = self._lib.tgx_create(C.byref(self._h), device)
        if rc:
            raise TgxError(rc, "tgx_create", self._detail(rc))
        self.device = device

    # ------------------------------------------------------------------------------------------------
    def _detail(self, rc: int) -> str:
        s = self._lib.tgx_strerror(rc).decode()
        if rc == abi.TGX_ERR_CUDA:
            s += ": " + self._lib.tgx_last_cuda_error().decode()
        return s

    def _check(self, rc: int, what: str):
        if rc:
            raise TgxError(rc, what, self._detail(rc))

    def close(self):
        if self._h:
            self._lib.tgx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _stream() -> int:
        import torch
        return int(torch.cuda.current_stream().cuda_stream)

    def _torch_device(self):
        import torch
        return torch.device("cuda", self.device)

    # ------------------------------------------------------------------------------------------------
    def set_max_samples(self, n: int):
        self._check(self._lib.tgx_set_max_samples(self._h, n), "tgx_set_max_samples")

    def set_tuning(self, tile_shift: int, spt: int):
        self._check(self._lib.tgx_set_tuning(self._h, tile_shift, spt), "tgx_set_tuning")

    def set_plan_mode(self, exact_ramps: bool):
        self._check(self._lib.tgx_set_plan_mode(self._h, 1 if exact_ramps else 0), "tgx_set_plan_mode")

    def set_host_fill(self, fill_constants_on_host: bool):
        self._check(self._lib.tgx_set_host_fill(self._h, 1 if fill_constants_on_host else 0), "tgx_set_host_fill")

    def set_host_layout(self, plane_major: bool):
        """Host buffers of generate_host / stop_host: [n, 14, cap] (default) or plane-major [14, n, cap]."""
        self._check(self._lib.tgx_set_host_layout(self._h, 1 if plane_major else 0), "tgx_set_host_layout")
        self._host_plane_major = bool(plane_major)

    def set_store_path(self, tma: bool):
        """tgx_eval's planes through TMA (default) or always through vector stores; same bytes either way."""
        self._check(self._lib.tgx_set_store_path(self._h, 1 if tma else 0), "tgx_set_store_path")

    def set_phase_planning(self, allow: bool):
        self._check(self._lib.tgx_set_phase_planning(self._h, 1 if allow else 0), "tgx_set_phase_planning")

    @property
    def phase_plan_count(self) -> int:
        return int(self._lib.tgx_phase_plan_count(self._h))

    def set_slab_planning(self, allow: bool):
        self._check(self._lib.tgx_set_slab_planning(self._h, 1 if allow else 0), "tgx_set_slab_planning")

    def plan_path_counts(self):
        """(single-replay plans, two-replay plans) so far."""
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self._lib.tgx_plan_path_counts(self._h, C.byref(a), C.byref(b)), "tgx_plan_path_counts")
        return int(a.value), int(b.value)

    @property
    def launch_count(self) -> int:
        return int(self._lib.tgx_launch_count(self._h))

    @property
    def scratch_bytes(self) -> int:
        return int(self._lib.tgx_scratch_bytes(self._h))

    # ------------------------------------------------------------------------------------------------
    def upload_params(self, params: np.ndarray):
        """Host tgx_params array -> uint8 [n, 128] tensor on the engine's device."""
        import torch
        assert params.dtype == abi.PARAMS_DTYPE
        raw = torch.from_numpy(np.ascontiguousarray(params).view(np.uint8).reshape(len(params), 128))
        return raw.to(self._torch_device(), non_blocking=False)

    def probe_dfma(self, reps: int = 3):
        """tgx_probe_dfma -> (thread-level DFMA instructions per second, ms per probe launch)."""
        rate, ms = C.c_double(0.0), C.c_double(0.0)
        self._check(self._lib.tgx_probe_dfma(self._h, reps, C.byref(rate), C.byref(ms)), "tgx_probe_dfma")
        return float(rate.value), float(ms.value)

    def probe_d2h(self, pinned: "PinnedArray", nbytes: int, reps: int = 1) -> float:
        """tgx_probe_d2h: seconds for `reps` plain device->host copies of nbytes into a pinned buffer."""
        sec = C.c_double(0.0)
        assert nbytes <= pinned.nbytes
        self._check(self._lib.tgx_probe_d2h(self._h, pinned.ptr, nbytes, reps, C.byref(sec)), "tgx_probe_d2h")
        return float(sec.value)

    def fill_montecarlo(self, n: int, seed: int = 1237, first_index: int = 0, out=None):
        """tgx_fill_montecarlo: records [first_index, first_index + n) of the config-4 distribution drawn on the
        device (Philox4x32-10; workloads.montecarlo_philox is the bit-identical host version) -> uint8 [n, 128]."""
        import torch
        if out is None:
            out = torch.empty((n, 128), dtype=torch.uint8, device=self._torch_device())
        self._check(self._lib.tgx_fill_montecarlo(self._h, seed, first_index, n, out.data_ptr(), self._stream()),
                    "tgx_fill_montecarlo")
        return out

    def count(self, d_params, limits: Optional[abi.Limits] = None):
        """tgx_count on a device-resident parameter tensor -> (counts int32 [n], status int32 [n])."""
        import torch
        n = int(d_params.shape[0])
        counts = torch.empty(n, dtype=torch.int32, device=d_params.device)
        status = torch.empty(n, dtype=torch.int32, device=d_params.device)
        self._check(self._lib.tgx_count(self._h, d_params.data_ptr(), n, _limits_ptr(limits), counts.data_ptr(),
                                        status.data_ptr(), self._stream()), "tgx_count")
        return counts, status

    def plan(self, d_params, limits: Optional[abi.Limits] = None, want_phases: bool = False,
             want_outputs: bool = True) -> Plan:
        """tgx_plan on a device-resident parameter tensor."""
        import torch
        n = int(d_params.shape[0])
        counts = status = phases = None
        if want_outputs:
            counts = torch.empty(n, dtype=torch.int32, device=d_params.device)
            status = torch.empty(n, dtype=torch.int32, device=d_params.device)
        if want_phases:
            phases = torch.empty((n, C.sizeof(abi.Phases)), dtype=torch.uint8, device=d_params.device)
        total = C.c_int64(0)
        self._check(self._lib.tgx_plan(self._h, d_params.data_ptr(), n, _limits_ptr(limits),
                                       counts.data_ptr() if counts is not None else None,
                                       status.data_ptr() if status is not None else None,
                                       phases.data_ptr() if phases is not None else None,
                                       C.byref(total), self._stream()), "tgx_plan")
        return Plan(n, counts, status, int(total.value), phases, int(self._lib.tgx_plan_tiles(self._h)),
                    int(self._lib.tgx_plan_segments(self._h)))

    def plan_polyline(self, d_params, limits: Optional[abi.Limits] = None, want_legs: bool = False,
                      want_outputs: bool = True) -> Plan:
        """tgx_plan_polyline on a device-resident parameter tensor (Square / Rectangle / Reciprocating / Bounce / M / I / T)."""
        import torch
        n = int(d_params.shape[0])
        counts = status = legs = None
        if want_outputs:
            counts = torch.empty(n, dtype=torch.int32, device=d_params.device)
            status = torch.empty(n, dtype=torch.int32, device=d_params.device)
        if want_legs:
            legs = torch.empty((n, C.sizeof(abi.PolylineLegs)), dtype=torch.uint8, device=d_params.device)
        total = C.c_int64(0)
        self._check(self._lib.tgx_plan_polyline(self._h, d_params.data_ptr(), n, _limits_ptr(limits),
                                                counts.data_ptr() if counts is not None else None,
                                                status.data_ptr() if status is not None else None,
                                                legs.data_ptr() if legs is not None else None,
                                                C.byref(total), self._stream()), "tgx_plan_polyline")
        return Plan(n, counts, status, int(total.value), None, int(self._lib.tgx_plan_tiles(self._h)), 0, legs)

    @staticmethod
    def finalize_polyline(params: np.ndarray) -> np.ndarray:
        """tgx_polyline_finalize_host in place: cos_o / sin_o of polyline records from the host libm."""
        assert params.dtype == abi.PARAMS_DTYPE and params.flags.c_contiguous
        rc = lib().tgx_polyline_finalize_host(params.ctypes.data, len(params))
        if rc:
            raise TgxError(rc, "tgx_polyline_finalize_host", lib().tgx_strerror(rc).decode())
        return params

    def plan_stop(self, d_params, d_from, want_phases: bool = False) -> Plan:
        """tgx_plan_stop: d_from is a float64 [n, 14] tensor with the setpoints being braked from."""
        import torch
        n = int(d_params.shape[0])
        assert d_from.dtype == torch.float64 and d_from.is_contiguous() and tuple(d_from.shape) == (n, abi.TGX_NCHAN)
        counts = torch.empty(n, dtype=torch.int32, device=d_params.device)
        status = torch.empty(n, dtype=torch.int32, device=d_params.device)
        phases = torch.empty((n, C.sizeof(abi.Phases)), dtype=torch.uint8, device=d_params.device) if want_phases else None
        total = C.c_int64(0)
        self._check(self._lib.tgx_plan_stop(self._h, d_params.data_ptr(), n, d_from.data_ptr(), counts.data_ptr(),
                                            status.data_ptr(), phases.data_ptr() if phases is not None else None,
                                            C.byref(total), self._stream()), "tgx_plan_stop")
        return Plan(n, counts, status, int(total.value), phases, int(self._lib.tgx_plan_tiles(self._h)),
                    int(self._lib.tgx_plan_segments(self._h)))

    def eval(self, out, capacity: Optional[int] = None, plane_major: bool = False, max_v=None, max_a=None,
             channel_mask: int = 0, traj_offset=None):
        """tgx_eval into a float64 tensor.

        out: [n, 14, row] (trajectory-major, default) or [14, n, row] (plane_major=True).
        """
        import torch
        assert out.dtype == torch.float64 and out.is_contiguous()
        lay = abi.Layout()
        lay.d_base = out.data_ptr()
        if plane_major:
            nch, n, row = out.shape
            lay.traj_stride, lay.chan_stride = row, n * row
        else:
            n, nch, row = out.shape
            lay.traj_stride, lay.chan_stride = nch * row, row
        assert nch == abi.TGX_NCHAN
        lay.capacity = row if capacity is None else capacity
        lay.channel_mask = channel_mask
        lay.d_traj_offset = traj_offset.data_ptr() if traj_offset is not None else None
        self._check(self._lib.tgx_eval(self._h, C.byref(lay), max_v.data_ptr() if max_v is not None else None,
                                       max_a.data_ptr() if max_a is not None else None, self._stream()), "tgx_eval")

    def generate(self, d_params, out, limits: Optional[abi.Limits] = None, plane_major: bool = False,
                 want_outputs: bool = False, want_phases: bool = False, chunk: int = 0) -> Plan:
        """tgx_generate: plan + evaluate a device-resident batch, pipelined over two internal streams.

        out: float64 [n, 14, row] (or [14, n, row] with plane_major=True).  Returns a Plan (counts / status / phases only
        if asked for)."""
        import torch
        assert out.dtype == torch.float64 and out.is_contiguous()
        n = int(d_params.shape[0])
        lay = abi.Layout()
        lay.d_base = out.data_ptr()
        if plane_major:
            nch, rows, row = out.shape
            lay.traj_stride, lay.chan_stride = row, rows * row
        else:
            rows, nch, row = out.shape
            lay.traj_stride, lay.chan_stride = nch * row, row
        assert nch == abi.TGX_NCHAN and rows >= n
        lay.capacity = row
        counts = status = phases = None
        if want_outputs:
            counts = torch.empty(n, dtype=torch.int32, device=d_params.device)
            status = torch.empty(n, dtype=torch.int32, device=d_params.device)
        if want_phases:
            phases = torch.empty((n, C.sizeof(abi.Phases)), dtype=torch.uint8, device=d_params.device)
        total = C.c_int64(0)
        self._check(self._lib.tgx_generate(self._h, d_params.data_ptr(), n, _limits_ptr(limits), C.byref(lay),
                                           counts.data_ptr() if counts is not None else None,
                                           status.data_ptr() if status is not None else None,
                                           phases.data_ptr() if phases is not None else None,
                                           chunk, C.byref(total), self._stream()), "tgx_generate")
        return Plan(n, counts, status, int(total.value), phases, 0, 0)

    def set_generate_profiling(self, on: bool):
        self._check(self._lib.tgx_set_generate_profiling(self._h, 1 if on else 0), "tgx_set_generate_profiling")

    def generate_profile(self):
        """(sum of the evaluation launches' durations in ms, number of launches) since the last query."""
        ms, cnt = C.c_double(0.0), C.c_int64(0)
        self._check(self._lib.tgx_generate_profile(self._h, C.byref(ms), C.byref(cnt)), "tgx_generate_profile")
        return float(ms.value), int(cnt.value)

    def generate_feasibility(self, d_params, limits: abi.Limits, flags=None, max_v=None, max_a=None, status=None,
                             chunk: int = 0):
        """tgx_generate_feasibility -> (flags uint8 [n], max_v, max_a, status int32, total samples)."""
        import torch
        n = int(d_params.shape[0])
        dev = d_params.device
        flags = torch.empty(n, dtype=torch.uint8, device=dev) if flags is None else flags
        max_v = torch.empty(n, dtype=torch.float64, device=dev) if max_v is None else max_v
        max_a = torch.empty(n, dtype=torch.float64, device=dev) if max_a is None else max_a
        status = torch.empty(n, dtype=torch.int32, device=dev) if status is None else status
        total = C.c_int64(0)
        self._check(self._lib.tgx_generate_feasibility(self._h, d_params.data_ptr(), n, C.byref(limits),
                                                       flags.data_ptr(), max_v.data_ptr(), max_a.data_ptr(),
                                                       status.data_ptr(), chunk, C.byref(total), self._stream()),
                    "tgx_generate_feasibility")
        return flags, max_v, max_a, status, int(total.value)

    def eval_layout(self, lay: abi.Layout, max_v=None, max_a=None):
        self._check(self._lib.tgx_eval(self._h, C.byref(lay), max_v.data_ptr() if max_v is not None else None,
                                       max_a.data_ptr() if max_a is not None else None, self._stream()), "tgx_eval")

    def eval_records(self, n: int, rec_capacity: int, limits: Optional[abi.Limits] = None, records=None,
                     rec_offset=None):
        """tgx_eval_records: the current plan straight into clamped tgx_goal_record rows
        (uint8 [n, rec_capacity, 128], or a flat [total, 128] tensor addressed through rec_offset)."""
        import torch
        if records is None:
            records = torch.zeros((n, rec_capacity, 128), dtype=torch.uint8, device=self._torch_device())
        # with offsets the call takes the buffer's total record count in place of the per-trajectory stride
        stride = rec_capacity if rec_offset is None else int(records.numel() // 128)
        self._check(self._lib.tgx_eval_records(self._h, _limits_ptr(limits), records.data_ptr(), stride,
                                               rec_offset.data_ptr() if rec_offset is not None else None,
                                               rec_capacity, self._stream()), "tgx_eval_records")
        return records

    def pack_goals(self, planes, counts, limits: Optional[abi.Limits] = None, records=None, rec_capacity=None,
                   rec_offset=None, plane_major: bool = False):
        """tgx_pack_goals: SoA planes [n, 14, row] (or [14, n, row]) + counts -> tgx_goal_record tensor
        (uint8 [n, rec_capacity, 128], or a flat [total, 128] one addressed through rec_offset)."""
        import torch
        assert planes.dtype == torch.float64 and planes.is_contiguous()
        lay = abi.Layout()
        lay.d_base = planes.data_ptr()
        if plane_major:
            nch, n, row = planes.shape
            lay.traj_stride, lay.chan_stride = row, n * row
        else:
            n, nch, row = planes.shape
            lay.traj_stride, lay.chan_stride = nch * row, row
        lay.capacity = row
        cap = row if rec_capacity is None else rec_capacity
        if records is None:
            records = torch.zeros((n, cap, 128), dtype=torch.uint8, device=planes.device)
        self._check(self._lib.tgx_pack_goals(self._h, C.byref(lay), counts.data_ptr(), n, _limits_ptr(limits),
                                             records.data_ptr(), cap,
                                             rec_offset.data_ptr() if rec_offset is not None else None, cap,
                                             self._stream()), "tgx_pack_goals")
        return records

    def generate_records_host(self, params: np.ndarray, rec_capacity: int, limits: Optional[abi.Limits] = None,
                              records: Optional[np.ndarray] = None):
        """tgx_generate_records_host -> (records [n, rec_capacity] of RECORD_DTYPE, counts, status)."""
        params = np.ascontiguousarray(params)
        n = len(params)
        if records is None:
            records = np.zeros((n, rec_capacity), dtype=abi.RECORD_DTYPE)
        assert records.dtype == abi.RECORD_DTYPE and records.flags.c_contiguous and records.shape == (n, rec_capacity)
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.uint32)
        self._check(self._lib.tgx_generate_records_host(self._h, params.ctypes.data, n, _limits_ptr(limits),
                                                        records.ctypes.data, rec_capacity, counts.ctypes.data,
                                                        status.ctypes.data), "tgx_generate_records_host")
        return records, counts, status

    def transitions(self, d_tparams, rec_capacity: int, limits: Optional[abi.Limits] = None, records=None):
        """tgx_transitions on a device-resident uint8 [n, 128] tensor of tgx_transition_params ->
        (records uint8 [n, rec_capacity, 128] or None, counts int32 [n], status int32 [n])."""
        import torch
        n = int(d_tparams.shape[0])
        if records is None and rec_capacity > 0:
            records = torch.zeros((n, rec_capacity, 128), dtype=torch.uint8, device=d_tparams.device)
        counts = torch.empty(n, dtype=torch.int32, device=d_tparams.device)
        status = torch.empty(n, dtype=torch.int32, device=d_tparams.device)
        self._check(self._lib.tgx_transitions(self._h, d_tparams.data_ptr(), n, _limits_ptr(limits),
                                              records.data_ptr() if records is not None else None, rec_capacity,
                                              rec_capacity, counts.data_ptr(), status.data_ptr(), self._stream()),
                    "tgx_transitions")
        return records, counts, status

    def transitions_host(self, tparams: np.ndarray, rec_capacity: int, limits: Optional[abi.Limits] = None):
        """tgx_transitions_host -> (records [n, rec_capacity] of RECORD_DTYPE, counts, status)."""
        tparams = np.ascontiguousarray(tparams)
        assert tparams.dtype == abi.TRANSITION_DTYPE
        n = len(tparams)
        records = np.zeros((n, max(rec_capacity, 0)), dtype=abi.RECORD_DTYPE)
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.uint32)
        self._check(self._lib.tgx_transitions_host(self._h, tparams.ctypes.data, n, _limits_ptr(limits),
                                                   records.ctypes.data if rec_capacity > 0 else None, rec_capacity,
                                                   counts.ctypes.data, status.ctypes.data), "tgx_transitions_host")
        return records, counts, status

    def feasibility(self, limits: abi.Limits, n: int, flags=None, max_v=None, max_a=None, status=None):
        """tgx_feasibility on the current plan -> (flags uint8 [n], max_v, max_a, status int32)."""
        import torch
        dev = self._torch_device()
        flags = torch.empty(n, dtype=torch.uint8, device=dev) if flags is None else flags
        max_v = torch.empty(n, dtype=torch.float64, device=dev) if max_v is None else max_v
        max_a = torch.empty(n, dtype=torch.float64, device=dev) if max_a is None else max_a
        status = torch.empty(n, dtype=torch.int32, device=dev) if status is None else status
        self._check(self._lib.tgx_feasibility(self._h, C.byref(limits), flags.data_ptr(), max_v.data_ptr(),
                                              max_a.data_ptr(), status.data_ptr(), self._stream()), "tgx_feasibility")
        return flags, max_v, max_a, status

    def selftest_division(self, n: int, seed: int = 1, per_thread: int = 64) -> int:
        bad = C.c_uint64(0)
        self._check(self._lib.tgx_selftest_division(self._h, n, seed, per_thread, C.byref(bad)),
                    "tgx_selftest_division")
        return int(bad.value)

    # ---- host-buffer calls -------------------------------------------------------------------------
    def count_host(self, params: np.ndarray, limits: Optional[abi.Limits] = None):
        params = np.ascontiguousarray(params)
        n = len(params)
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.uint32)
        self._check(self._lib.tgx_count_host(self._h, params.ctypes.data, n, _limits_ptr(limits), counts.ctypes.data,
                                             status.ctypes.data), "tgx_count_host")
        return counts, status

    def generate_host_legs(self, params: np.ndarray, capacity: int, limits: Optional[abi.Limits] = None,
                           out: Optional[np.ndarray] = None):
        """tgx_generate_host_legs -> (out [n, 14, capacity], counts, status, phases, legs)."""
        params = np.ascontiguousarray(params)
        n = len(params)
        if out is None:
            out = np.full((n, abi.TGX_NCHAN, capacity), np.nan)
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == (n, abi.TGX_NCHAN, capacity)
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.uint32)
        phases = np.zeros(n, dtype=abi.PHASES_DTYPE)
        legs = np.zeros(n, dtype=abi.LEGS_DTYPE)
        self._check(self._lib.tgx_generate_host_legs(self._h, params.ctypes.data, n, _limits_ptr(limits),
                                                     out.ctypes.data, capacity, counts.ctypes.data, status.ctypes.data,
                                                     phases.ctypes.data, legs.ctypes.data), "tgx_generate_host_legs")
        return out, counts, status, phases, legs

    def generate_host(self, params: np.ndarray, capacity: int, limits: Optional[abi.Limits] = None,
                      out: Optional[np.ndarray] = None, want_phases: bool = False):
        """tgx_generate_host -> (out [n, 14, capacity], counts, status, phases or None)."""
        params = np.ascontiguousarray(params)
        n = len(params)
        shape = (abi.TGX_NCHAN, n, capacity) if getattr(self, "_host_plane_major", False) else (n, abi.TGX_NCHAN, capacity)
        if out is None:
            out = np.full(shape, np.nan)
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == shape
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.uint32)
        phases = np.zeros(n, dtype=abi.PHASES_DTYPE) if want_phases else None
        self._check(self._lib.tgx_generate_host(self._h, params.ctypes.data, n, _limits_ptr(limits), out.ctypes.data,
                                                capacity, counts.ctypes.data, status.ctypes.data,
                                                phases.ctypes.data if phases is not None else None),
                    "tgx_generate_host")
        return out, counts, status, phases

    def generate_host_compact(self, params: np.ndarray, capacity: int, limits: Optional[abi.Limits] = None,
                              out: Optional[np.ndarray] = None, want_phases: bool = False, want_legs: bool = False):
        """tgx_generate_host_compact: the 10 varying planes only -> (out [n, 10, capacity] or, plane-major,
        [10, n, capacity]; counts; status; phases or None; legs or None).  p.z = params['alt'], v.z = a.z = j.z = 0."""
        params = np.ascontiguousarray(params)
        n = len(params)
        nv = abi.TGX_NCHAN_VARYING
        shape = (nv, n, capacity) if getattr(self, "_host_plane_major", False) else (n, nv, capacity)
        if out is None:
            out = np.full(shape, np.nan)
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == shape
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.uint32)
        phases = np.zeros(n, dtype=abi.PHASES_DTYPE) if want_phases else None
        legs = np.zeros(n, dtype=abi.LEGS_DTYPE) if want_legs else None
        self._check(self._lib.tgx_generate_host_compact(
            self._h, params.ctypes.data, n, _limits_ptr(limits), out.ctypes.data, capacity, counts.ctypes.data,
            status.ctypes.data, phases.ctypes.data if phases is not None else None,
            legs.ctypes.data if legs is not None else None), "tgx_generate_host_compact")
        return out, counts, status, phases, legs

    def host_info(self) -> dict:
        """tgx_host_info: how the host-buffer calls size themselves on this machine."""
        info = abi.HostInfo()
        self._check(self._lib.tgx_host_info(self._h, C.byref(info)), "tgx_host_info")
        return {k: int(getattr(info, k)) for k in ("numa_node", "cpus_allowed", "local_ranks", "filler_threads",
                                                   "filler_cpus")}

    def sample_host(self, params: np.ndarray, v: float, accel: float, s0: float, s1: float = 0.0) -> np.ndarray:
        """tgx_sample_host: one create*Goal evaluation -> the 14 channels."""
        params = np.ascontiguousarray(params)
        out = np.zeros(abi.TGX_NCHAN)
        self._check(self._lib.tgx_sample_host(self._h, params.ctypes.data, v, accel, s0, s1, out.ctypes.data),
                    "tgx_sample_host")
        return out

    def stop_host(self, params: np.ndarray, from14: np.ndarray, capacity: int, want_phases: bool = False):
        """tgx_stop_host -> (out [n, 14, capacity], counts, status, phases or None)."""
        params = np.ascontiguousarray(params)
        n = len(params)
        from14 = np.ascontiguousarray(from14, dtype=np.float64).reshape(n, abi.TGX_NCHAN)
        out = np.full((n, abi.TGX_NCHAN, capacity), np.nan)
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.uint32)
        phases = np.zeros(n, dtype=abi.PHASES_DTYPE) if want_phases else None
        self._check(self._lib.tgx_stop_host(self._h, params.ctypes.data, n, from14.ctypes.data, out.ctypes.data,
                                            capacity, counts.ctypes.data, status.ctypes.data,
                                            phases.ctypes.data if phases is not None else None), "tgx_stop_host")
        return out, counts, status, phases
