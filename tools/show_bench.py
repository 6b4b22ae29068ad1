"""Prints the key numbers of a bench.py JSON line: python tools/show_bench.py gpurun_out/x.json"""
import json
import sys


def one(name, c):
    if "error" in c:
        print(name, "ERROR", c["error"])
        return
    rl = c.get("roofline") or {}
    par = c.get("parity_fullsize") or {}
    print("%-3s value %8.3f G vd/s  %8.3f ms/step  e2e %8.3f G (%.2f ms)  dom %-12s frac %.4f  whole %.4f  8d %s  launches %d  parity ok=%s bit_exact=%s  cpu %.1f M/s" % (
        name, c["value"] / 1e9, c["ms_per_step"], c["e2e"]["value"] / 1e9, c["e2e"]["ms_per_step"], rl.get("kernel"), rl.get("frac") or 0,
        (rl.get("whole_path") or {}).get("frac") or 0, ((rl.get("whole_path") or {}).get("headline_8d") or {}).get("frac"),
        c["gpu_launches"], par.get("ok"), par.get("bit_exact"), (c.get("cpu_baseline") or {}).get("value", 0) / 1e6))
    if rl:
        print("    kernels/step: %s" % rl.get("kernel_time_in_step"))
        for k, v in rl["stages"].items():
            if v["ms"]:
                print("    %-12s %8.3f ms  %5d launches  %8.1f GB/s" % (k, v["ms"], v["launches"], v["gbs"] or 0))
        for g in rl["groups"]:
            print("    [%s r=%5.2f%% x%d] %8.2f us/dab  %6.2f G vd/s  f=%.3f  dom %7.2f us frac %s  whole %s" % (
                g["stroke"], g["radius_pct_diag"], g["dabs"], g["us_per_dab"], g["gvd_per_s"] or 0, g["moved_frac"], g["dominant_kernel_us_per_dab"],
                g["dominant_kernel_frac"], g["whole_path_frac"]))


d = json.load(open(sys.argv[1]))
one(d["config"]["workload"][:2].lower(), d)
for c in d.get("configs", []):
    one(c["name"], c)
