"""One-line summary of a bench.py JSON line read from stdin."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d.get("kernels", {})
print(d["config"]["workload"][:24], "| %.2f samples/s %.2f ms | e2e %.2f (%.2f ms) | launches %s | conv %.1f TF frac %.3f | %s" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("ms_per_step", 0), d.get("gpu_launches"),
    d["roofline"]["achieved"], d["roofline"]["frac"],
    {n: (round(v["ms_per_step"], 3), round(v.get("frac") or 0, 3)) for n, v in k.items()}))
