#!/usr/bin/env python
"""Print the figures of a bench.py JSON line that matter when comparing runs."""
import json, sys
for path in sys.argv[1:]:
    d = json.load(open(path))
    print(path)
    print("  value %.2f G/s  ms/step %.3f  eval_ms %.3f  not_hidden %.3f  roofline.frac %.4f  launches %s" % (
        d["value"] / 1e9, d["ms_per_step"], d.get("detail", {}).get("eval_ms_per_step", 0),
        d.get("detail", {}).get("not_hidden_ms_per_step", 0),
        d["roofline"].get("frac", 0), d.get("gpu_launches")))
    e = d.get("e2e")
    if e:
        print("  e2e %.1f M/s  frac %.3f  (%.1f of %.1f GB/s)" % (e["value"] / 1e6, e.get("frac", 0), e.get("achieved_gbs", 0), e.get("d2h_ceiling_gbs", 0)))
    c = d.get("cpu_baseline")
    if c:
        print("  cpu %.1f M/s on %d cores" % (c["value"] / 1e6, c["cores"]))
    for k, v in (d.get("extra") or {}).items():
        if isinstance(v, dict) and "value" in v:
            r = v.get("roofline", {})
            print("  %s: %.2f G/s  ms/step %.3f  eval %.3f  not_hidden %.3f  braking %s  gather %s  frac %.3f" % (
                k, v["value"] / 1e9, v["ms_per_step"], v.get("eval_ms_per_step", 0), v.get("not_hidden_ms_per_step", v.get("plan_ms_per_step", 0)),
                v.get("braking_ms_per_step"), v.get("flags_allgather_ms"), r.get("frac", 0)))
        elif k == "cfg1_latency":
            print("  cfg1:", json.dumps(v)[:300])
