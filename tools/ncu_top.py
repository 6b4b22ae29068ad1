"""Top stall sites of an ncu report's source page (SASS view).  Usage: ncu_top.py report.ncu-rep [kernel-instance] [N]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; inst = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = txt.split('"Kernel Name"')[1:]
blk = blocks[inst]
lines = blk.split("\n")
print("kernel:", lines[0][:100])
rd = csv.DictReader(io.StringIO("\n".join(lines[1:])))
rows = [r for r in rd if r.get("Address")]
tot = sum(int(r["# Samples"]) for r in rows)
reasons = [k for k in rows[0].keys() if k.startswith("stall_") and "Not Issued" not in k]
agg = collections.Counter()
for r in rows:
    for k in reasons:
        agg[k] += int(r[k] or 0)
print("total samples", tot, {k: v for k, v in agg.most_common(8)})
rows_s = sorted(enumerate(rows), key=lambda ir: -int(ir[1]["# Samples"]))[:top]
for i, r in sorted(rows_s):
    rs = {k[6:]: int(r[k]) for k in reasons if int(r[k] or 0) > 0}
    rs = dict(sorted(rs.items(), key=lambda kv: -kv[1])[:3])
    print("%5d %6.2f%% %-58s exec=%-8s %s" % (i, 100.0 * int(r["# Samples"]) / max(tot, 1), r["Source"].strip()[:58], r["Instructions Executed"], rs))
