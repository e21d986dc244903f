#!/usr/bin/env python
"""Prints the headline numbers of a bench.py JSON line (a reading aid; the judged numbers are the line itself)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ratio", round(d["e2e"]["value"] / d["value"], 3), "ms/step", round(d["ms_per_step"], 4), "n_gpus", d["n_gpus"])
r = d["roofline"]
print("roofline frac", round(r["frac"], 3), "ms", round(r["ms"], 4), r.get("clocks"))
for b in r.get("by_batch", []):
    print("   B", b["positions_per_launch"], "ms", round(b["ms"], 4), "TF", round(b["achieved"], 1), "frac", round(b["frac"], 3))
print("stages", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r["stages_ms"].items() if k != "clocks"})
if r.get("sustained"):
    s = r["sustained"]
    print("sustained", s["device_batches"], "batches", round(s["seconds"], 2), "s", round(s["whole_net_tflops"]), "TF frac", round(s["frac"], 3), s["clocks"])
print("sweep", [(x["batch"], round(x["ms"], 4)) for x in d["batch_sweep"]])
if d.get("leaf_latency"):
    print("leaf median us", round(d["leaf_latency"]["median_us"], 1))
for leg in [d["selfplay"]] + d["selfplay_others"] if d.get("selfplay") else []:
    h = leg.get("host_driver", {})
    print(leg["workload"][:34], "| device", round(leg["value"]), "sims/s games", leg["games"], "sec", round(leg["seconds"], 2), "ms/wave", round(leg["ms_per_wave"], 4),
          "mean_batch", round(leg["mean_batch"]), "| host driver", round(h.get("value", 0)), "| cpu", round(leg.get("cpu_baseline", {}).get("value", 0)))
print("clocks", d["clocks"])
print("gpu_launches", d["gpu_launches"], "cpu_baseline", d["cpu_baseline"] and round(d["cpu_baseline"]["value"]))
