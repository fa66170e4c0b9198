#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the committed summaries under profiles/.

    python tools/summarize_profiles.py r01

reads gpurun_out/<tag>_launches.csv (ncu --metrics gpu__time_duration.sum launch list of `bench.py`) and
gpurun_out/<tag>_prof.ncu-rep (ncu --set full capture of eval_kernel / plan_fill_kernel) and writes
profiles/<tag>_launches_summary.txt, profiles/<tag>_kernels_ncu.txt and profiles/eval_traffic.json.
"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)

# ---- launch list ---------------------------------------------------------------------------------------------
rows = list(csv.reader(l for l in open(os.path.join(src, f"{tag}_launches.csv")) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(list)
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    unit = r[ui]
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    agg[r[ki].split("(")[0][:100]].append(ns)
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(dst, f"{tag}_launches_summary.txt"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  (launch list of `python bench.py --no-e2e --no-cpu "
            f"--steps 3 --warmup 2`; cold-cache, serialised: compare SHARES)\n")
    f.write(f"# {sum(len(v) for v in agg.values())} launches, {tot / 1e6:.3f} ms total\n")
    f.write(f"{'total ms':>10} {'n':>4} {'avg us':>10} {'share':>7}  kernel\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{sum(v) / 1e6:10.3f} {len(v):4d} {sum(v) / len(v) / 1e3:10.1f} {100 * sum(v) / tot:6.1f}%  {k}\n")
print(open(os.path.join(dst, f"{tag}_launches_summary.txt")).read())

# ---- full capture ----------------------------------------------------------------------------------------------
rep = os.path.join(src, f"{tag}_prof.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_write.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "sm__sass_inst_executed_op_global_st.sum"]
traffic = {}
with open(os.path.join(dst, f"{tag}_kernels_ncu.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on, `python bench.py --no-e2e --no-cpu --steps 2 --warmup 2 "
            "--n-per-gpu 65536` (65 536 config-2 circles = 65 568 623 samples = 7.3437 GB algorithmic per eval launch)\n")
    for r in rows[2:]:
        f.write("\n")
        vals = {}
        for i, name in enumerate(h):
            if name in want:
                f.write(f"{name:90s} {u[i]:16s} {r[i][:110]}\n")
                vals[name] = (r[i], u[i])
        if "eval_kernel" in vals.get("Kernel Name", ("", ""))[0]:
            def gb(key):
                v, unit = vals[key]
                return float(v.replace(",", "")) * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}[unit]
            rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
            algo = 65568623 * 112 / 1e9
            f.write(f"# dram traffic {rd + wr:.4f} GB (read {rd:.4f} + write {wr:.4f}) vs algorithmic {algo:.4f} GB: ratio "
                    f"{(rd + wr) / algo:.4f}\n")
            traffic["ratio"] = (rd + wr) / algo
print(open(os.path.join(dst, f"{tag}_kernels_ncu.txt")).read()[:6000])
if traffic:
    bench_bytes = 1049100173 * 112
    with open(os.path.join(dst, "eval_traffic.json"), "w") as f:
        json.dump({"circles_cfg2": traffic["ratio"] * bench_bytes,
                   "_how": f"dram__bytes_read.sum + dram__bytes_write.sum of tgx::eval_kernel from profiles/{tag}_kernels_ncu.txt "
                           f"(ncu --set full at 65 536 trajectories) divided by that launch's algorithmic bytes = "
                           f"{traffic['ratio']:.4f}, times the bench launch's algorithmic bytes ({bench_bytes})"}, f, indent=1)

# ---- the other captures of tools/refresh_profiles.sh: records (f3), reduction-only (configs 4-5), plan_fill ----------------
extra = [("prof_rec_tma.ncu-rep", "ncu --set full -k regex:eval_kernel -s 3 -c 1 of `python bench.py --records --steps 1 --warmup 1 "
          "--no-cpu --no-e2e`: the RECORDS instantiation through TMA (tgx_eval_records) on one chunk of 131 072 circles; "
          "algorithmic bytes 128 B/sample"),
         ("prof_feas.ncu-rep", "ncu --set full -k regex:reduce_kernel -s 3 -c 1 of `python bench.py --workload montecarlo_cfg4 "
          "--steps 2 --warmup 3 --no-cpu --no-e2e --n-per-gpu 2000000`: tgx_feasibility's kernel on 2 000 000 config-4 circles "
          "= 2.157e9 samples; writes 17 B per trajectory, FP64- / issue-bound"),
         ("prof_plan_fill.ncu-rep", "ncu --set full -k regex:plan_fill_kernel -s 3 -c 1 of `python bench.py --workload "
          "montecarlo_cfg4 --steps 2 --warmup 3 --no-cpu --no-e2e --n-per-gpu 1000000`: the replay that writes the segment "
          "tables, 1 000 000 config-4 circles"),
         ("prof_cfg3.ncu-rep", "ncu --set full --kernel-name-base demangled -k regex:tgx::eval_kernel|tgx::plan_phase_kernel -s 2 -c 2 "
          "of `python bench.py --workload mixed_cfg3 --steps 1 --warmup 2 --no-cpu --no-e2e --no-extras --n-per-gpu 262144`: "
          "config 3's mixed circle / line / figure-eight batch as phase records, one CTA per trajectory (97 .. 1889 samples), "
          "and the phase planner on the class-sorted batch; algorithmic bytes 112 B/sample")]
want2 = want + ["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
                "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
                "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
with open(os.path.join(dst, f"{tag}_f3_cfg4_plan_kernels_ncu.txt"), "w") as f:
    for name, title in extra:
        rep = os.path.join(src, name)
        if not os.path.exists(rep):
            continue
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        h, u = rows[0], rows[1]
        f.write("# " + title + "\n")
        for r in rows[2:]:
            for i, nm in enumerate(h):
                if nm in want2:
                    f.write(f"{nm:90s} {u[i]:16s} {r[i][:110]}\n")
            f.write("\n")
print(open(os.path.join(dst, f"{tag}_f3_cfg4_plan_kernels_ncu.txt")).read()[:3000])
