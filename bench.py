#!/usr/bin/env python
"""bench.py — throughput of the trajectory-sampling hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" is one pass of the hot path over one batch of synthetic parameters that are already resident in HBM:
tgx_plan (strict-IEEE replay -> counts + segment/tile tables) followed by tgx_eval (the store-bound sampling
kernel).  Default workload: BASELINE.json configs[1], 1 Mi circles x ~1000 samples on each GPU (weak scaling).
Rank 0 prints ONE JSON line.  `--impl reference` times the reference's own CPU implementation (oracle/_ref, the
unmodified sources behind stub headers; falls back to the plain-C oracle port) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "traj samples/s (p,v,a,j,yaw)"
UNIT = "samples/s"
BYTES_PER_SAMPLE = 112          # 14 fp64 channels per sample
FALLBACK_HBM_GBS = 6650.0       # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def make_params(workload: str, n: int, lo: int, hi: int):
    from trajectory_generator_ros2_b200 import workloads
    if workload == "circles_cfg2":
        return workloads.circles_cfg2(n, lo=lo, hi=hi)
    if workload == "mixed_cfg3":
        return workloads.mixed_cfg3(n, lo=lo, hi=hi)
    if workload == "montecarlo_cfg4":
        return workloads.montecarlo_cfg4(n, lo=lo, hi=hi)
    if workload in POLYLINE_WORKLOADS:
        from trajectory_generator_ros2_b200.engine import Engine
        gen = workloads.polyline_mix if workload == "polyline_mix" else workloads.letters_T
        # cos / sin of the orientation from the host libm, as the reference computes them (tgx_polyline_finalize_host)
        return Engine.finalize_polyline(np.ascontiguousarray(gen(n, lo=lo, hi=hi)).copy())
    raise SystemExit(f"unknown workload {workload}")


POLYLINE_WORKLOADS = ("polyline_mix", "letters_T")


WORKLOAD_DESC = {
    "circles_cfg2": "BASELINE configs[1]: {n} circles x ~1000 samples per GPU, random r/v/centre (rng 1234), dt 0.01",
    "mixed_cfg3": "BASELINE configs[2]: {n} mixed circle/line/figure-eight with one or two ramp-ups per GPU (rng 1235)",
    "montecarlo_cfg4": "BASELINE configs[3]: {n} wide-range circles per GPU, max-|v|/|a| feasibility only (rng 1236)",
    "polyline_mix": "SURVEY 8(f2): {n} constant-speed polylines per GPU (Square/Rectangle/Reciprocating/Bounce/M/I/T, "
                    "~1000 samples each, rng 1238)",
    "letters_T": "SURVEY 8(f2): {n} T trajectories per GPU (the shape default.yaml ships), random size/speed/pose, "
                 "~1000 samples each (rng 1239)",
}


# ---- clocks ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms; only samples inside [mark_start, mark_stop] count."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=self.tmp, stderr=subprocess.DEVNULL)
            time.sleep(1.0)     # let the sampler come up before the timed region starts
        except Exception:
            self.proc = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_stop(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        try:
            with open(self.tmp.name) as f:
                for line in f:
                    parts = [x.strip() for x in line.split(",")]
                    if len(parts) < 8:
                        continue
                    try:
                        ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                        rows.append((ts, float(parts[1]), float(parts[2]), float(parts[3]), parts[4:8]))
                    except ValueError:
                        continue
        finally:
            try:
                os.unlink(self.tmp.name)
            except OSError:
                pass
        if not rows:
            return out
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.05 <= r[0] <= self.t1 + 0.05]
        used = inside if inside else rows
        reasons = set()
        for r in used:
            for nm, val in zip(names, r[4]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        out["sm_mhz"] = float(np.median([r[1] for r in used]))
        out["sm_max_mhz"] = float(max(r[2] for r in used))
        out["power_w_max"] = float(max(r[3] for r in used))
        out["samples"] = len(used)
        out["window"] = "timed region" if inside else "whole run (no sample fell inside the timed region)"
        out["reasons"] = sorted(reasons)
        return out


# ---- CPU legs ------------------------------------------------------------------------------------------------
def cpu_time_sample(params, threads: int, prefer_reference: bool = True):
    """Time the reference's CPU path on `params` -> (samples, seconds, kind)."""
    from oracle_lib import Oracle, Reference
    if prefer_reference and Reference.available():
        ref = Reference()
        t0 = time.perf_counter()
        total, _ = ref.time_batch(params, threads)
        return total, time.perf_counter() - t0, "reference"
    orc = Oracle()
    t0 = time.perf_counter()
    total, _ = orc.time_batch(params, threads)
    return total, time.perf_counter() - t0, "port"


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    n_sample = args.cpu_sample
    params = make_params(args.workload, args.n_per_gpu, 0, n_sample)
    kind = None
    for _ in range(args.warmup):
        cpu_time_sample(params[: max(1, n_sample // 8)], threads)
    total_samples, total_s = 0, 0.0
    for _ in range(args.steps):
        s, dt, kind = cpu_time_sample(params, threads)
        total_samples += s
        total_s += dt
    value = total_samples / total_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload].format(n=args.n_per_gpu)},
        "detail": {"step": f"generateTraj of the first {n_sample} trajectories of the workload into std::vector<Goal>, "
                           f"{threads} host threads"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"first {n_sample} trajectories ({total_samples // args.steps} samples) per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---- extra legs of the default run: BASELINE.json configs[0], [2], [3], [4] -------------------------------------------
# FP64 work of one circle sample on the reduction-only path (csrc/eval.cu: reduce_kernel), counted from the source like
# the 112 algorithmic bytes of the store path: speed 1 (v = vb + j dv) + j(j+1)/2 2 + angle 2 +
# sincos_orbit 21 (1 scale-and-round fma, 1 subtraction, 3 Cody-Waite fma, z = r*r, 5 + 5 polynomial fma, r*z, 2 + 2
# final fma / mul) + omega 1 + v^2/r 1 + v.x, v.y, a.x, a.y 4 + |v|^2, |a|^2 4 + two running-maximum compares 2 = 38
# FP64-pipe instructions (ncu, thread level: 38.3 per sample; the quadrant selects, index arithmetic and segment
# bookkeeping are ~35 integer / move instructions on top of it).
FP64_INSTR_PER_SAMPLE = 38


class Ctx:
    """What the extra legs share with the main leg."""
    def __init__(self, **kw):
        self.__dict__.update(kw)


def fp64_roofline(ctx, samples: int, eval_ms: float):
    """Roofline of the reduction-only kernel against the DFMA peak measured in this run (tgx_probe_dfma)."""
    dfma_per_s, probe_ms = ctx.eng.probe_dfma(3)
    inst_per_s = FP64_INSTR_PER_SAMPLE * samples / (eval_ms * 1e-3)
    return {"bound": "fp64", "achieved": 2e-12 * inst_per_s, "peak": 2e-12 * dfma_per_s, "unit": "TFLOP/s",
            "frac": inst_per_s / dfma_per_s, "traffic": None,
            "kernel": "tgx::reduce_kernel<MODE, 1024>",
            "fp64_instr_per_sample": FP64_INSTR_PER_SAMPLE,
            "convention": "every FP64-pipe instruction (DFMA / DMUL / DADD / DSETP) counted as one DFMA slot = 2 FLOP; "
                          "achieved = %d instr/sample x samples / kernel time (CUDA events), peak = DFMA micro-benchmark " % FP64_INSTR_PER_SAMPLE +
                          "of this run (tgx_probe_dfma: 8 independent chains per thread, 1024 threads per SM, "
                          "%.1f ms per launch, best of 3)" % probe_ms,
            "peak_source": "measured in this run (no FP64 figure in MEASURED_PEAKS.json)"}


def leg_cfg3(ctx, steps=5, warmup=2):
    """BASELINE configs[2] as SURVEY.md 8d defines it: the mixed batch PLUS a braking trajectory (generateStopTraj from
    k = N_i / 2, TrajectoryGenerator.cpp:514-517 -> Circle.cpp:132-169) for every tenth trajectory."""
    import torch
    from trajectory_generator_ros2_b200 import abi, workloads
    eng, dev, n, rank, world = ctx.eng, ctx.dev, ctx.n, ctx.rank, ctx.world
    params = workloads.mixed_cfg3(world * n, lo=rank * n, hi=(rank + 1) * n)
    d_params = eng.upload_params(params)
    counts, _ = eng.count(d_params)
    total = int(counts.sum(dtype=torch.int64).item())
    row = (int(counts.max().item()) + 1023) // 1024 * 1024
    chunks = 1
    while True:
        try:
            rows = (n + chunks - 1) // chunks
            out = torch.empty((rows, abi.TGX_NCHAN, row), dtype=torch.float64, device=dev)
            break
        except torch.OutOfMemoryError:
            chunks *= 2
            if chunks > 64:
                raise
    rows = (n + chunks - 1) // chunks
    chunk_params, stop_idx, stop_params, stop_k = [], [], [], []
    for c in range(chunks):
        dp = d_params[c * rows: min(n, (c + 1) * rows)]
        idx = torch.arange(0, dp.shape[0], 10, device=dev)                 # every tenth trajectory brakes
        chunk_params.append(dp)
        stop_idx.append(idx)
        stop_params.append(dp[idx].contiguous())
        stop_k.append((counts[c * rows: c * rows + dp.shape[0]][idx] // 2).long())
    # size the braking rows from an untimed pass
    stop_cap, stop_total = 4, 0
    for c in range(chunks):
        eng.plan(chunk_params[c], want_outputs=False)
        eng.eval(out[: chunk_params[c].shape[0]])
        d_from = out[stop_idx[c], :, stop_k[c]].contiguous()
        sp = eng.plan_stop(stop_params[c], d_from)
        stop_cap = max(stop_cap, (int(sp.counts.max().item()) + 3) // 4 * 4)
        stop_total += sp.total_samples
    stop_out = torch.empty((int(stop_idx[0].shape[0]), abi.TGX_NCHAN, stop_cap), dtype=torch.float64, device=dev)
    # Braking runs on a second engine and a side stream: the setpoints are gathered on the main stream right after the
    # chunk's samples exist (before the next chunk overwrites the buffer), and tgx_plan_stop + tgx_eval of chunk c are
    # issued after tgx_generate of chunk c+1, so they run under its evaluation kernels.
    eng2 = ctx.eng2
    side = torch.cuda.Stream(device=dev)
    pending = []
    ev = {"eval": [], "stop": []}

    def brake(record):
        while pending:
            c, d_from, ready = pending.pop(0)
            with torch.cuda.stream(side):
                side.wait_event(ready)
                b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                b0.record()
                eng2.plan_stop(stop_params[c], d_from)
                eng2.eval(stop_out[: stop_idx[c].shape[0]])
                b1.record()
                d_from.record_stream(side)
                if record:
                    ev["stop"].append((b0, b1))

    def step(record):
        main = torch.cuda.current_stream()
        for c in range(chunks):
            dp = chunk_params[c]
            if pipelined:
                eng.generate(dp, out[: dp.shape[0]])
            else:
                eng.plan(dp, want_outputs=False)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.eval(out[: dp.shape[0]])
                e1.record()
                if record:
                    ev["eval"].append((e0, e1))
            brake(record)                     # the previous chunk's braking, under this chunk's evaluation
            # END pressed at k = N_i / 2: brake from that setpoint (the caller's gather of goals[pub_index] is plumbing)
            d_from = out[stop_idx[c], :, stop_k[c]].contiguous()
            ready = torch.cuda.Event()
            ready.record()
            pending.append((c, d_from, ready))
            if not pipelined:
                brake(record)

    def drain():
        brake(True)
        done = torch.cuda.Event()
        done.record(side)
        torch.cuda.current_stream().wait_event(done)

    pipelined = not ctx.args.no_pipeline
    for _ in range(warmup):
        step(False)
    drain()
    ctx.barrier()
    eng.set_generate_profiling(pipelined)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step(True)
    drain()
    t1.record()
    ctx.barrier()
    ms = ctx.allmax(t0.elapsed_time(t1) / steps)
    if pipelined:
        prof_ms, prof_n = eng.generate_profile()
        eng.set_generate_profiling(False)
        eval_ms = ctx.allmax(prof_ms / steps)
    else:
        eval_ms = ctx.allmax(float(np.mean([a.elapsed_time(b) for a, b in ev["eval"]])) * chunks)
    stop_ms = ctx.allmax(float(np.mean([a.elapsed_time(b) for a, b in ev["stop"]])) * chunks) if ev["stop"] else 0.0
    job = ctx.allsum(total + stop_total)
    peak, _ = measured_peak()
    achieved = BYTES_PER_SAMPLE * total / (eval_ms * 1e-3) / 1e9
    del out, stop_out
    torch.cuda.empty_cache()
    return {"workload": WORKLOAD_DESC["mixed_cfg3"].format(n=n) + "; every tenth trajectory additionally brakes from "
                        "k = N_i/2 (tgx_plan_stop + tgx_eval, generateStopTraj)",
            "value": job / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "samples_per_gpu": total, "braking_trajectories_per_gpu": int(sum(int(i.shape[0]) for i in stop_idx)),
            "braking_samples_per_gpu": stop_total, "chunks": chunks, "row_stride": row,
            "eval_ms_per_step": eval_ms, "braking_ms_per_step": stop_ms,
            "not_hidden_ms_per_step": ms - eval_ms,
            "step": ("per chunk tgx_generate (planning pipelined under the evaluation on two internal streams); braking "
                     "(tgx_plan_stop + tgx_eval on a second engine and stream) issued after the next chunk's "
                     "tgx_generate so that it runs under its evaluation kernels; braking_ms is its own stream time, "
                     "not_hidden_ms what the step takes beyond the generateTraj evaluation kernels") if pipelined else
                    "per chunk tgx_plan + tgx_eval, then tgx_plan_stop + tgx_eval, all on one stream",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "kernel": "tgx::eval_kernel (generateTraj samples only)"}}


def leg_cfg4(ctx, steps=3, warmup=3):
    """BASELINE configs[3]: 10^7 wide-range circles per GPU, max-|v| / max-|a| feasibility reduction only.
    (Three warm-up steps: the engine arrives from config 3's phase plans, sees that only the reduction kernel consumes
    this batch, and learns the table slices of the new batch shape — the third plan is the steady-state one.)"""
    import torch
    from trajectory_generator_ros2_b200 import abi, workloads
    eng, dev, rank, world = ctx.eng, ctx.dev, ctx.rank, ctx.world
    n = ctx.args.cfg4_n
    params = workloads.montecarlo_cfg4(world * n, lo=rank * n, hi=(rank + 1) * n)
    d_params = eng.upload_params(params)
    del params
    lim = abi.make_limits(**workloads.MONTECARLO_LIMITS)
    flags = torch.empty(n, dtype=torch.uint8, device=dev)
    mv = torch.empty(n, dtype=torch.float64, device=dev)
    ma = torch.empty(n, dtype=torch.float64, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev)
    ev = []
    total = 0

    pipelined = not ctx.args.no_pipeline

    def step(record):
        nonlocal total
        if pipelined:
            total = eng.generate_feasibility(d_params, lim, flags=flags, max_v=mv, max_a=ma, status=st)[4]
            return
        total = eng.plan(d_params, limits=lim, want_outputs=False).total_samples
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.feasibility(lim, n, flags=flags, max_v=mv, max_a=ma, status=st)
        b.record()
        if record:
            ev.append((a, b))

    for _ in range(warmup):
        step(False)
    ctx.barrier()
    eng.set_generate_profiling(pipelined)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step(True)
    t1.record()
    ctx.barrier()
    ms = ctx.allmax(t0.elapsed_time(t1) / steps)
    if pipelined:
        prof_ms, _ = eng.generate_profile()
        eng.set_generate_profiling(False)
        eval_ms = ctx.allmax(prof_ms / steps)
    else:
        eval_ms = ctx.allmax(float(np.mean([a.elapsed_time(b) for a, b in ev])))
    job = ctx.allsum(total)
    res = {"workload": WORKLOAD_DESC["montecarlo_cfg4"].format(n=n), "value": job / (ms * 1e-3), "unit": UNIT,
           "ms_per_step": ms, "steps": steps, "warmup": warmup, "samples_per_gpu": total,
           "eval_ms_per_step": eval_ms, "not_hidden_ms_per_step": ms - eval_ms,
           "feasible_fraction": float(flags.float().mean()),
           "step": ("tgx_generate_feasibility (1 Mi-trajectory chunks, planning pipelined under the reduction kernel on "
                    "two internal streams)" if pipelined else "tgx_plan + tgx_feasibility") + ", parameters resident in HBM",
           "roofline": fp64_roofline(ctx, total, eval_ms)}
    del d_params, flags, mv, ma, st
    torch.cuda.empty_cache()
    return res


def leg_cfg5(ctx, steps=2, warmup=1):
    """BASELINE configs[4]: the 10^8-trajectory sweep sharded over the GPUs of the box (strong scaling: the job is
    fixed), every shard drawn on its own device (tgx_fill_montecarlo), then the NCCL all-gather of the 1-byte flags
    issued through the C-ABI (tgx_gather_flags)."""
    import torch
    from trajectory_generator_ros2_b200 import abi, workloads
    from trajectory_generator_ros2_b200.engine import Comm, shard_range
    eng, dev, rank, world = ctx.eng, ctx.dev, ctx.rank, ctx.world
    n_total = ctx.args.cfg5_total
    lo, hi = shard_range(n_total, rank, world)
    m = hi - lo
    lim = abi.make_limits(**workloads.MONTECARLO_LIMITS)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d_params = torch.empty((m, 128), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    a.record()
    eng.fill_montecarlo(m, seed=1237, first_index=lo, out=d_params)
    b.record()
    torch.cuda.synchronize()
    fill_ms = ctx.allmax(a.elapsed_time(b))
    pipelined = not ctx.args.no_pipeline
    chunk = m if pipelined else 1 << 24               # trajectories per call (tgx_generate_feasibility chunks by itself)
    flags = torch.empty(m, dtype=torch.uint8, device=dev)
    full = torch.empty(n_total, dtype=torch.uint8, device=dev)
    mv = torch.empty(min(m, chunk), dtype=torch.float64, device=dev)
    ma = torch.empty(min(m, chunk), dtype=torch.float64, device=dev)
    st = torch.empty(min(m, chunk), dtype=torch.int32, device=dev)
    comm = Comm.from_torch_distributed(ctx.local_rank) if world > 1 else None
    total = 0
    ev_eval, ev_gather = [], []

    def step(record):
        nonlocal total
        total = 0
        for s in range(0, m, chunk):
            k = min(chunk, m - s)
            if pipelined:
                total += eng.generate_feasibility(d_params[s:s + k], lim, flags=flags[s:s + k], max_v=mv[:k],
                                                  max_a=ma[:k], status=st[:k])[4]
                continue
            total += eng.plan(d_params[s:s + k], limits=lim, want_outputs=False).total_samples
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.feasibility(lim, k, flags=flags[s:s + k], max_v=mv[:k], max_a=ma[:k], status=st[:k])
            e1.record()
            if record:
                ev_eval.append((e0, e1))
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        if comm is not None:
            comm.gather_flags(flags, n_total, out=full)
        else:
            full.copy_(flags)
        g1.record()
        if record:
            ev_gather.append((g0, g1))

    for _ in range(warmup):
        step(False)
    ctx.barrier()
    eng.set_generate_profiling(pipelined)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step(True)
    t1.record()
    ctx.barrier()
    ms = ctx.allmax(t0.elapsed_time(t1) / steps)
    nchunks = (m + chunk - 1) // chunk
    if pipelined:
        prof_ms, _ = eng.generate_profile()
        eng.set_generate_profiling(False)
        eval_ms = ctx.allmax(prof_ms / steps)
    else:
        eval_ms = ctx.allmax(float(np.mean([x.elapsed_time(y) for x, y in ev_eval])) * nchunks)
    gather_ms = ctx.allmax(float(np.mean([x.elapsed_time(y) for x, y in ev_gather])))
    job = ctx.allsum(total)
    ok = bool(torch.equal(full[lo:hi], flags))
    feasible = int(full.sum(dtype=torch.int64).item())
    # every rank must hold the same gathered vector: compare a checksum of checksums
    chk = int((full.view(torch.int64)[: n_total // 8].sum().item()) & 0x7fffffffffff) if n_total >= 8 else feasible
    same = ctx.allmax(float(chk)) == float(chk) and -ctx.allmax(-float(chk)) == float(chk)
    res = {"workload": f"BASELINE configs[4]: {n_total} config-4 circles sharded over {world} GPU(s) "
                       f"(tgx_shard_range), each shard drawn on its device (tgx_fill_montecarlo, Philox4x32-10, "
                       f"seed 1237), feasibility only",
           "value": job / (ms * 1e-3), "unit": UNIT, "scaling": "strong", "ms_per_step": ms, "steps": steps,
           "warmup": warmup, "trajectories": n_total, "trajectories_per_gpu": m, "samples": job,
           "eval_ms_per_step": eval_ms, "not_hidden_ms_per_step": ms - eval_ms - gather_ms,
           "flags_allgather_ms": gather_ms, "device_fill_ms": fill_ms,
           "gather": ("tgx_gather_flags -> ncclAllGather of %d B per rank over NCCL %d, on the evaluation stream, "
                      "inside the timed step" % (m, Comm.nccl_version())) if comm is not None else
                     "one shard: device-to-device copy",
           "gathered_equals_local_shard": ok, "gathered_identical_on_all_ranks": bool(same),
           "feasible": feasible,
           "step": ("tgx_generate_feasibility over the shard (1 Mi-trajectory chunks, planning pipelined under the "
                    "reduction kernel)" if pipelined else "per 2^24-trajectory chunk tgx_plan + tgx_feasibility")
                   + ", then the flag all-gather"}
    if comm is not None:
        comm.close()
    del d_params, flags, full, mv, ma, st
    torch.cuda.empty_cache()
    return res


def leg_cfg1(ctx):
    """BASELINE configs[0] / BASELINE.md 4: ms per generateTraj of the default.yaml circle — the C++ drop-in class
    (tests/cpp/bin/dropin_latency: count, plan, evaluate, D2H, repack into std::vector<Goal>) beside the reference class
    on one host core (oracle/_ref; the reference's own timing hook is Circle.cpp:92).  Rank 0 only."""
    from trajectory_generator_ros2_b200 import workloads
    res = {"workload": "BASELINE configs[0]: Circle::generateTraj of config/default.yaml (25 001 samples), median of 20"}
    exe = os.path.join(ROOT, "tests", "cpp", "bin", "dropin_latency")
    if os.path.exists(exe):
        env = dict(os.environ, TGX_DEVICE=str(ctx.local_rank))
        r = subprocess.run([exe, "20"], capture_output=True, text=True, timeout=300, env=env)
        try:
            res["dropin"] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            res["dropin"] = {"error": (r.stdout + r.stderr)[-300:]}
    else:
        res["dropin"] = {"unavailable": "tests/cpp/bin/dropin_latency not built (needs the reference headers at build time)"}
    try:
        from oracle_lib import Oracle, Reference
        p = workloads.default_circle()
        impl, kind = (Reference(), "reference") if Reference.available() else (Oracle(), "port")
        for _ in range(3):
            impl.time_batch(p, 1)
        ts = []
        for _ in range(20):
            t0 = time.perf_counter()
            impl.time_batch(p, 1)
            ts.append(1e3 * (time.perf_counter() - t0))
        res["cpu"] = {"kind": kind, "median_ms": float(np.median(ts)), "min_ms": float(min(ts)), "cores": 1,
                      "call": "Circle::generateTraj into std::vector<Goal>, -O2 (the reference's flag-less build is slower)"}
    except Exception as exc:     # the checker is optional here
        res["cpu"] = {"error": str(exc)[:200]}
    return res


def run_extras(ctx):
    extra = {}
    t0 = time.perf_counter()
    extra["cfg3"] = leg_cfg3(ctx)
    extra["cfg4"] = leg_cfg4(ctx)
    extra["cfg5"] = leg_cfg5(ctx)
    if ctx.rank == 0:
        extra["cfg1_latency"] = leg_cfg1(ctx)
    extra["wall_s"] = time.perf_counter() - t0
    return extra


# ---- our arm ---------------------------------------------------------------------------------------------------
def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, read+write bytes)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback 6650 GB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def ncu_traffic(workload: str):
    """dram bytes per eval launch from the committed ncu capture of the same command, if there is one."""
    path = os.path.join(ROOT, "profiles", "eval_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return d.get(workload)
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from trajectory_generator_ros2_b200 import abi, workloads
    from trajectory_generator_ros2_b200.engine import Engine, PinnedArray

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    eng = Engine(local_rank)
    eng.set_store_path(args.store_path == "tma")
    if args.tile_shift or args.spt:
        eng.set_tuning(args.tile_shift or 10, args.spt or 4)

    n = args.n_per_gpu
    lo, hi = rank * n, (rank + 1) * n                      # weak scaling: every GPU owns n trajectories
    params = make_params(args.workload, world * n, lo, hi)
    d_params = eng.upload_params(params)
    feas_only = args.workload == "montecarlo_cfg4"
    poly = args.workload in POLYLINE_WORKLOADS
    lim = abi.make_limits(**workloads.MONTECARLO_LIMITS) if feas_only else None

    # size the output from a first (untimed) count
    counts, _ = eng.count(d_params)
    max_count = int(counts.max().item())
    total_samples = int(counts.sum(dtype=torch.int64).item())
    row = max(1024, (max_count + 1023) // 1024 * 1024) if not feas_only else 0
    chunks = 1
    out = None
    if not feas_only:
        while True:
            try:
                rows = (n + chunks - 1) // chunks
                shape = (abi.TGX_NCHAN, rows, row) if args.plane_major else (rows, abi.TGX_NCHAN, row)
                out = torch.empty(shape, dtype=torch.float64, device=dev)
                break
            except torch.OutOfMemoryError:
                chunks *= 2
                if chunks > 64:
                    raise
    if args.records:
        # consumer side: the samples are packed into clamped 128-byte records; planes + records of the
        # whole batch do not fit 180 GB, so the batch runs in 8 chunks through one plane buffer and one record buffer
        del out
        torch.cuda.empty_cache()
        chunks = max(chunks, 8)
        rows = (n + chunks - 1) // chunks
        out = torch.empty((rows, abi.TGX_NCHAN, row), dtype=torch.float64, device=dev)
        rec = torch.empty((rows, row, 128), dtype=torch.uint8, device=dev)
        d_counts = counts.to(torch.int32).contiguous()
        box_lim = abi.make_limits(box=(-4.0, 4.0, -4.0, 4.0, 0.0, 2.0))
    rows = (n + chunks - 1) // chunks
    chunk_params = [d_params[c * rows: min(n, (c + 1) * rows)] for c in range(chunks)]
    pack_pairs, fused_pairs = [], []

    ev_pairs = []

    pipelined = not args.no_pipeline and not poly and not args.records and not args.plane_major

    def step(record: bool):
        for c in range(chunks):
            dp = chunk_params[c]
            if pipelined:
                if feas_only:
                    eng.generate_feasibility(dp, lim, flags=flags[c], max_v=mv[c], max_a=ma[c], status=st[c])
                else:
                    eng.generate(dp, out[: dp.shape[0]], chunk=args.gen_chunk)
                continue
            if poly:
                eng.plan_polyline(dp, want_outputs=False)
            else:
                eng.plan(dp, limits=lim, want_outputs=False)
            if record:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
            if feas_only:
                eng.feasibility(lim, int(dp.shape[0]), flags=flags[c], max_v=mv[c], max_a=ma[c], status=st[c])
            else:
                view = out[:, : dp.shape[0]] if args.plane_major else out[: dp.shape[0]]
                if view.is_contiguous():
                    eng.eval(view, plane_major=args.plane_major)
                else:   # plane-major partial chunk: describe the full buffer
                    lay = abi.Layout()
                    lay.d_base, lay.traj_stride, lay.chan_stride, lay.capacity = out.data_ptr(), row, rows * row, row
                    eng.eval_layout(lay)
            if record:
                b.record()
                ev_pairs.append((a, b))
            if args.records:
                m = int(dp.shape[0])
                eng.pack_goals(out[:m], d_counts[c * rows: c * rows + m], box_lim, records=rec, rec_capacity=row)
                if record:
                    e2 = torch.cuda.Event(enable_timing=True)
                    e2.record()
                    pack_pairs.append((b, e2))
                # the fused form of the same work: records straight from the evaluation kernel (same plan)
                eng.eval_records(m, row, box_lim, records=rec)
                if record:
                    e3 = torch.cuda.Event(enable_timing=True)
                    e3.record()
                    fused_pairs.append((e2, e3))

    if feas_only:
        flags = [torch.empty(int(p.shape[0]), dtype=torch.uint8, device=dev) for p in chunk_params]
        mv = [torch.empty(int(p.shape[0]), dtype=torch.float64, device=dev) for p in chunk_params]
        ma = [torch.empty(int(p.shape[0]), dtype=torch.float64, device=dev) for p in chunk_params]
        st = [torch.empty(int(p.shape[0]), dtype=torch.int32, device=dev) for p in chunk_params]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x: float) -> float:
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def allsum(x: int) -> int:
        if world == 1:
            return int(x)
        t = torch.tensor([x], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t[0])

    for _ in range(args.warmup):
        step(False)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    if sampler:
        sampler.mark_start()
    launches0 = eng.launch_count
    eng.set_generate_profiling(pipelined)
    t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        step(True)
    t_stop.record()
    barrier()
    if sampler:
        sampler.mark_stop()
    elapsed_ms = t_start.elapsed_time(t_stop)
    launches = eng.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    eval_launches = chunks
    if pipelined:
        # the library brackets every evaluation launch with an event pair on the stream it is launched on
        prof_ms, prof_n = eng.generate_profile()
        eng.set_generate_profiling(False)
        eval_ms = prof_ms / args.steps
        eval_launches = prof_n // args.steps
    else:
        eval_ms = float(np.mean([a.elapsed_time(b) for a, b in ev_pairs])) * chunks   # per step
    if world > 1:
        t = torch.tensor([elapsed_ms, eval_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, eval_ms = float(t[0]), float(t[1])
        tot = torch.tensor([total_samples], dtype=torch.int64, device=dev)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        job_samples = int(tot[0])
    else:
        job_samples = total_samples
    ms_per_step = elapsed_ms / args.steps
    value = job_samples / (ms_per_step * 1e-3)

    # ---- end-to-end through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed region) ----
    e2e = None
    if not feas_only and not args.no_e2e and not args.records:
        has_bounce = bool((params["type"] == abi.TGX_BOUNCE).any())     # Bounce moves along z: nothing is constant
        fmt = "all" if has_bounce else args.e2e_format
        eng.set_host_fill(fmt != "all")
        eng.set_host_layout(not args.e2e_traj_major)
        call_n = min(n, args.e2e_call)                       # trajectories per tgx_generate_host call
        hrow = (max_count + 3) // 4 * 4                      # host rows: the caller's capacity, no padding shipped
        hplanes = abi.TGX_NCHAN_VARYING if fmt == "compact" else abi.TGX_NCHAN
        pin_out = PinnedArray((call_n, hplanes, hrow) if args.e2e_traj_major else (hplanes, call_n, hrow), engine=eng)
        # the step's inputs live in pinned host memory (copied there once, byte-wise: numpy copies this union dtype field
        # by field); every call uploads its own slice inside the timed region
        pin_params = PinnedArray((n,), dtype=abi.PARAMS_DTYPE, engine=eng)
        pin_params.array.view(np.uint8)[:] = np.ascontiguousarray(params).view(np.uint8)
        gen = eng.generate_host_compact if fmt == "compact" else eng.generate_host

        def e2e_step():
            tot = 0
            for s in range(0, n, call_n):
                m = min(call_n, n - s)
                dst = pin_out.array[:m] if args.e2e_traj_major else pin_out.array.reshape(-1)[:hplanes * m * hrow].reshape(hplanes, m, hrow)
                c = gen(pin_params.array[s:s + m], hrow, out=dst)[1]
                tot += int(c.sum())
            return tot

        del out
        torch.cuda.empty_cache()
        e2e_step()                                           # warm-up (first touch of the pinned pages, plan learning)
        step_s = []
        h_counts_total = 0
        for _ in range(max(3, args.e2e_steps)):
            barrier()
            t0 = time.perf_counter()
            h_counts_total = e2e_step()
            barrier()
            step_s.append(allmax(time.perf_counter() - t0))
        assert h_counts_total == total_samples
        e2e_s = float(np.median(step_s))
        planes = 14 if fmt == "all" else 10
        d2h_bytes = int(n * (planes * hrow * 8 + 8))
        # the ceiling of this box, measured in this run: a plain device->host copy of the same byte count into the
        # same pinned buffer on every rank CONCURRENTLY (what PCIe and the host's memory absorb from N GPUs at once)
        copy_bytes = min(pin_out.nbytes, call_n * planes * hrow * 8)
        reps = max(1, int(round(n * planes * hrow * 8 / copy_bytes)))
        barrier()
        ceil_s = allmax(eng.probe_d2h(pin_out, copy_bytes, reps))
        barrier()
        ceiling_gbs = world * copy_bytes * reps / ceil_s / 1e9
        achieved_gbs = world * d2h_bytes / e2e_s / 1e9
        e2e = {"value": job_samples / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": int(n * 128), "d2h_bytes_per_step": d2h_bytes, "host_row_capacity": hrow,
               "wire_format": {"all": "all 14 planes over PCIe",
                               "fill": "10 varying planes over PCIe; the 4 constant planes (p.z = alt, v.z = a.z = j.z = 0) "
                                       "are written into the host buffer by host threads (112 B of host-memory writes "
                                       "per sample)",
                               "compact": "compact10 (tgx_generate_host_compact): the host buffer holds the 10 varying "
                                          "planes; p.z = params.alt and v.z = a.z = j.z = 0 are the reference's literal "
                                          "constants (Circle.cpp:109-121) and are not materialised (80 B of PCIe traffic "
                                          "and of host-memory writes per sample)"}[fmt],
               "steps": len(step_s), "ms_per_step": 1e3 * e2e_s, "ms_per_step_all": [1e3 * x for x in step_s],
               "statistic": "median of the timed steps, each the max over ranks",
               "achieved_gbs": achieved_gbs, "d2h_ceiling_gbs": ceiling_gbs, "frac": achieved_gbs / ceiling_gbs,
               "ceiling": f"plain pinned D2H copy, {reps} x {copy_bytes / 1e9:.2f} GB per rank, all {world} rank(s) "
                          f"concurrently, CUDA events, max over ranks (tgx_probe_d2h), aggregate GB/s",
               "host": eng.host_info(),
               "call": f"{'tgx_generate_host_compact' if fmt == 'compact' else 'tgx_generate_host'}, {call_n} "
                       f"trajectories per call into one reused pinned host buffer, "
                       + ("trajectory-major [n][planes][row]" if args.e2e_traj_major else
                          "plane-major [planes][n][row] (tgx_set_host_layout): contiguous 1-D copies per plane")}
        eng.set_host_layout(False)
        eng.set_host_fill(True)
        pin_out.free()
        pin_params.free()
    elif feas_only and not args.no_e2e:
        # feasibility: H2D of the parameter records, D2H of flags + maxima
        h_flags = np.empty(n, dtype=np.uint8)
        pin = torch.from_numpy(np.ascontiguousarray(params).view(np.uint8).reshape(n, 128)).pin_memory()

        def e2e_step():
            dp = pin.to(dev, non_blocking=True)
            eng.plan(dp, limits=lim, want_outputs=False)
            f, v_, a_, s_ = eng.feasibility(lim, n)
            return f.cpu(), v_.cpu(), a_.cpu()

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        e2e = {"value": job_samples / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(n * 128),
               "d2h_bytes_per_step": int(n * 17), "steps": args.e2e_steps, "ms_per_step": 1e3 * e2e_s,
               "call": "H2D params, tgx_plan + tgx_feasibility, D2H flags + max_v + max_a"}

    gather_ms = None
    if feas_only and world > 1:
        # BASELINE configs[4]: the only exchange of the path, an all-gather of the 1-byte feasibility flags (NCCL)
        # issued by libtgx itself: tgx_gather_flags -> ncclAllGather (torch.distributed only couriers the unique id)
        from trajectory_generator_ros2_b200.engine import Comm
        comm = Comm.from_torch_distributed(local_rank)
        local = torch.cat(flags)
        full = torch.empty(world * n, dtype=torch.uint8, device=dev)
        for _ in range(2):
            comm.gather_flags(local, world * n, out=full)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        comm.gather_flags(local, world * n, out=full)
        b.record()
        barrier()
        gather_ms = allmax(a.elapsed_time(b))
        assert torch.equal(full[rank * n:(rank + 1) * n], local)
        comm.close()
    extra = None
    if args.workload == "circles_cfg2" and not args.no_extras and not args.records:
        try:
            del out
        except NameError:
            pass
        del d_params, chunk_params
        torch.cuda.empty_cache()
        eng2 = Engine(local_rank)
        extra = run_extras(Ctx(eng=eng, eng2=eng2, dev=dev, n=n, rank=rank, world=world, local_rank=local_rank,
                               args=args, barrier=barrier, allmax=allmax, allsum=allsum))
        eng2.close()
    feas_roof = fp64_roofline(Ctx(eng=eng), total_samples, eval_ms) if (feas_only and rank == 0) else None
    if rank == 0:
        peak, peak_src = measured_peak()
        eval_bytes = BYTES_PER_SAMPLE * total_samples if not feas_only else 17 * n
        achieved = eval_bytes / (eval_ms * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu:
            threads = host_threads()
            s, dt, kind = cpu_time_sample(params[: args.cpu_sample], threads)
            cpu = {"value": s / dt, "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": f"first {args.cpu_sample} trajectories of the workload ({s} samples), "
                             f"generateTraj into std::vector<Goal>, {dt:.2f} s wall"}
        roof_extra = {}
        traffic = ncu_traffic(args.workload)
        if args.records:
            pack_ms = float(np.mean([a.elapsed_time(b) for a, b in pack_pairs])) * chunks
            pack_bytes = (BYTES_PER_SAMPLE + 128) * total_samples
            roof_extra = {"pack_kernel": {"kernel": "tgx::pack_goals_kernel", "bound": "hbm", "ms_per_step": pack_ms,
                                          "bytes_per_sample": BYTES_PER_SAMPLE + 128,
                                          "achieved": pack_bytes / (pack_ms * 1e-3) / 1e9, "unit": "GB/s",
                                          "frac": pack_bytes / (pack_ms * 1e-3) / 1e9 / peak,
                                          "note": "112 B of planes read + 128 B of records written per sample"}}
            fused_ms = float(np.mean([a.elapsed_time(b) for a, b in fused_pairs])) * chunks
            roof_extra["fused_records_kernel"] = {
                "kernel": "tgx::eval_kernel<..., RECORDS>", "bound": "hbm", "ms_per_step": fused_ms,
                "bytes_per_sample": 128, "achieved": 128.0 * total_samples / (fused_ms * 1e-3) / 1e9, "unit": "GB/s",
                "frac": 128.0 * total_samples / (fused_ms * 1e-3) / 1e9 / peak,
                "samples_per_s": total_samples / (fused_ms * 1e-3),
                "note": "tgx_eval_records: the same records without the plane round trip "
                        "(vs eval %.1f ms + pack %.1f ms)" % (eval_ms, pack_ms)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload].format(n=n)},
            "detail": {
                "trajectories_per_gpu": n, "samples_per_gpu": total_samples, "row_stride": row,
                "layout": "plane-major [14][n][row]" if args.plane_major else "trajectory-major [n][14][row]",
                "store_path": ("TMA: 32-sample x 14-channel boxes staged in shared memory (polyline plans and irregular "
                               "layouts: 256-bit vector stores)") if args.store_path == "tma" else "256-bit vector stores",
                "chunks": chunks,
                "step": ("tgx_plan_polyline (hold-length table + waypoints / per-leg step counts in strict IEEE "
                         "arithmetic, one thread per trajectory) + tgx_eval, parameters resident in HBM") if poly else
                        (("tgx_generate_feasibility" if feas_only else "tgx_generate") + ": the batch in %d chunks, each "
                         "planned (hold-length table + one strict-IEEE replay per trajectory) on one of two internal "
                         "streams while the previous chunk's evaluation kernel runs on the other; parameters resident "
                         "in HBM" % eval_launches) if pipelined else
                        "tgx_plan (hold-length table + one strict-IEEE replay per trajectory into fixed slices; the "
                        "first plan of an engine measures the slice sizes with a count + scan + fill pass) + "
                        + ("tgx_feasibility" if feas_only else "tgx_eval") + ", parameters resident in HBM",
                "plan_paths": dict(zip(("single_replay", "two_replay"), eng.plan_path_counts())),
                "not_hidden_ms_per_step": ms_per_step - eval_ms,
                "l2": "each step writes %.1f GB >> 126 MB L2, no flush needed" % (eval_bytes / 1e9)
                      if not feas_only else "reduction only",
                "parallelism": f"{world} independent shards, no data-path collective",
                "eval_ms_per_step": eval_ms,
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "kernel": "tgx::eval_poly_kernel" if poly else "tgx::eval_kernel",
                         "bytes_per_launch": eval_bytes / eval_launches, "launches_per_step": eval_launches,
                         "ms_per_launch": eval_ms / eval_launches,
                         "peak_source": peak_src + " — of measured" if "MEASURED" in peak_src else peak_src},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        if extra is not None:
            line["extra"] = extra
        if traffic is not None:
            line["roofline"]["traffic_source"] = ("profiles/eval_traffic.json: dram__bytes_read.sum + "
                                                  "dram__bytes_write.sum of this kernel from the committed ncu --set full "
                                                  "capture, per sample, scaled to this launch (ncu cannot run inside a "
                                                  "timed bench)")
        line["roofline"].update(roof_extra)
        if args.records:
            line["detail"]["step"] += " + tgx_pack_goals (clamp to the room box, pack to 128-byte records)"
        if feas_only:
            line["detail"]["feasible_fraction"] = float(torch.cat(flags).float().mean())
            line["detail"]["flags_allgather_ms"] = gather_ms
            # reduction-only path: 17 B written per trajectory; bound by the FP64 pipe / instruction issue, not by HBM
            feas_roof["hbm_for_information"] = {"achieved_gbs": achieved, "frac": achieved / peak}
            line["roofline"] = feas_roof
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    eng.close()


def run_transitions(args):
    """the node's own take-off / go-to / landing recurrences for a fleet, one thread per vehicle, one
    128-byte record per tick.  Single GPU; rank 0 prints one JSON line (same keys, the roofline counts record bytes)."""
    import torch
    from trajectory_generator_ros2_b200 import abi, workloads
    from trajectory_generator_ros2_b200.engine import Engine
    if int(os.environ.get("RANK", "0")) != 0:
        return
    eng = Engine(0)
    dev = torch.device("cuda", 0)
    n = args.n_per_gpu if args.n_per_gpu != (1 << 20) else (1 << 17)
    t = workloads.fleet_transitions(n)
    d_t = torch.from_numpy(t.view(np.uint8).reshape(n, 128)).to(dev)
    lim = abi.make_limits(box=(-5.0, 5.0, -5.0, 5.0, 0.0, 5.0))
    _, counts, status = eng.transitions(d_t, 0, lim)
    torch.cuda.synchronize()
    cap = int(counts.max().item())
    cap = (cap + 3) // 4 * 4
    total = int(counts.sum(dtype=torch.int64).item())
    rec = torch.empty((n, cap, 128), dtype=torch.uint8, device=dev)
    for _ in range(args.warmup):
        eng.transitions(d_t, cap, lim, records=rec)
    sampler = ClockSampler(0)
    torch.cuda.synchronize()
    sampler.mark_start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launch_count
    a.record()
    for _ in range(args.steps):
        eng.transitions(d_t, cap, lim, records=rec)
    b.record()
    torch.cuda.synchronize()
    sampler.mark_stop()
    ms = a.elapsed_time(b) / args.steps
    launches = eng.launch_count - l0
    clocks = sampler.stop()
    peak, peak_src = measured_peak()
    achieved = 128.0 * total / (ms * 1e-3) / 1e9
    # CPU leg: the oracle restatement (pinned to the unmodified node by tests/test_node_oracle.py), one host thread
    from oracle_lib import Oracle
    orc = Oracle()
    m = min(n, 512)
    t0 = time.perf_counter()
    ticks = sum(len(orc.transition(t[i:i + 1], box=(-5.0, 5.0, -5.0, 5.0, 0.0, 5.0))[0]) for i in range(m))
    cpu_s = time.perf_counter() - t0
    emit({
        "metric": "transition setpoints/s (take-off / go-to / landing recurrences)", "value": total / (ms * 1e-3),
        "unit": "samples/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"SURVEY 8(f4): {n} vehicles, one take-off / go-to / landing phase each (rng 1240), "
                               f"{total} ticks, record rows of {cap}", "l2": "%.1f GB of records per step" % (128e-9 * total)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "tgx::transition_kernel", "bytes_per_launch": 128.0 * total,
                     "peak_source": peak_src},
        "cpu_baseline": {"value": ticks / cpu_s, "unit": "samples/s", "cores": 1, "kind": "port",
                         "sample": f"first {m} vehicles ({ticks} ticks) through orc_transition incl. the Python call overhead"},
        "e2e": None, "gpu_launches": launches, "clocks": clocks})
    eng.close()


_JSON_FD = None


def _claim_stdout():
    """Keep stdout for the ONE JSON line: everything else a library prints there (NCCL's version banner under torchrun,
    for instance) is sent to stderr, and the JSON line is written to the original descriptor."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="circles_cfg2", choices=sorted(WORKLOAD_DESC) + ["transitions"])
    ap.add_argument("--n-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--plane-major", action="store_true")
    ap.add_argument("--tile-shift", type=int, default=0)
    ap.add_argument("--spt", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=1 << 17, help="trajectories timed on the CPU legs")
    ap.add_argument("--e2e-steps", type=int, default=3, help="timed end-to-end steps (at least 3; the median is reported)")
    ap.add_argument("--e2e-format", default="compact", choices=["compact", "fill", "all"],
                    help="host wire format of the e2e leg: compact10 (tgx_generate_host_compact, default), 10 planes + "
                         "host-filled constants (tgx_generate_host), or all 14 planes")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg1 / cfg3 / cfg4 / cfg5 legs of the default run")
    ap.add_argument("--cfg4-n", type=int, default=10_000_000, help="trajectories per GPU of the config-4 leg")
    ap.add_argument("--cfg5-total", type=int, default=100_000_000, help="trajectories of the config-5 sweep (all GPUs)")
    ap.add_argument("--e2e-call", type=int, default=1 << 16)
    ap.add_argument("--e2e-traj-major", action="store_true",
                    help="e2e leg with the trajectory-major host layout [n][14][row] (2-D copies) instead of plane-major")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--records", action="store_true",
                    help="also run the consumer-side kernel: clamp + pack every sample into a 128-byte record")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--gen-chunk", type=int, default=0, help="trajectories per chunk of tgx_generate (0: library default)")
    ap.add_argument("--pipeline", dest="no_pipeline", action="store_false", default=True,
                    help="time the pipelined tgx_generate (planning of chunk c+1 under the evaluation of chunk c) instead "
                         "of tgx_plan + tgx_eval back to back on one stream; measured: no faster (DESIGN.md 12)")
    ap.add_argument("--store-path", default="tma", choices=["tma", "stg"],
                    help="tgx_eval's planes through TMA (default) or through vector stores (tgx_set_store_path)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    _claim_stdout()
    if args.workload == "transitions":
        run_transitions(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
