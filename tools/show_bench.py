"""Readable digest of a bench.py JSON line.  Usage: show_bench.py file.json"""
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.2f G vd/s, %.3f ms/step, e2e %.2f G vd/s, vd/step %d, launches %d" % (
    d["value"] / 1e9, d["ms_per_step"], d["e2e"]["value"] / 1e9, d.get("vertex_dabs_per_step", 0), d.get("gpu_launches", 0)))
r = d["roofline"]
print("dominant", r["kernel"], r["achieved"], "GB/s frac", r["frac"], "| whole path", r["whole_path"])
for k, v in r["stages"].items():
    print("  %-12s %s" % (k, v))
for s in r.get("radius_sweep", []):
    print("  ", s)
if "cpu_baseline" in d:
    print("cpu", d["cpu_baseline"])
print("clocks", d.get("clocks"))
