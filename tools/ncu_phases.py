"""Warp-level instructions executed between barriers of a kernel in an ncu report (SASS page).
Usage: ncu_phases.py report.ncu-rep [kernel-instance]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; inst = int(sys.argv[2]) if len(sys.argv) > 2 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blk = txt.split('"Kernel Name"')[1:][inst]
lines = blk.split("\n")
rd = csv.DictReader(io.StringIO("\n".join(lines[1:])))
rows = [r for r in rd if r.get("Address")]
tot = sum(int(r["Instructions Executed"] or 0) for r in rows)
samples = sum(int(r["# Samples"]) for r in rows)
seg_start = 0; acc = 0; sacc = 0; kinds = collections.Counter()
print("total warp instructions", tot, "samples", samples)
for i, r in enumerate(rows):
    ex = int(r["Instructions Executed"] or 0)
    acc += ex; sacc += int(r["# Samples"])
    op = r["Source"].strip().split()
    name = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "")
    kinds[name.split(".")[0]] += ex
    if "BAR.SYNC" in r["Source"] or "EXIT" in r["Source"] or i == len(rows) - 1:
        print("sass %5d..%5d  %6.2f%% of instructions  %6.2f%% of samples   ends with %s" % (seg_start, i, 100.0 * acc / max(tot, 1), 100.0 * sacc / max(samples, 1), r["Source"].strip()[:40]))
        seg_start = i + 1; acc = 0; sacc = 0
print("by opcode:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in kinds.most_common(25)))
