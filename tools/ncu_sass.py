"""SASS listing with warp-level execution counts.  Usage: ncu_sass.py report.ncu-rep lo hi [kernel-instance]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; lo = int(sys.argv[2]); hi = int(sys.argv[3]); inst = int(sys.argv[4]) if len(sys.argv) > 4 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blk = txt.split('"Kernel Name"')[1:][inst]
lines = blk.split("\n")
rows = [r for r in csv.DictReader(io.StringIO("\n".join(lines[1:]))) if r.get("Address")]
for i in range(lo, min(hi, len(rows))):
    r = rows[i]
    print("%5d %9s %5s  %s" % (i, r["Instructions Executed"], r["# Samples"], r["Source"].strip()[:90]))
