"""Key metrics of every kernel in an ncu report, as text.  Usage: ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv, io, subprocess, sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__cycles_active.avg",
]
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    print("== %s: %d kernel instance(s)" % (rep, len(rows) - 2))
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        vals = {}
        for w in WANT:
            if w in hdr:
                vals[w] = r[hdr.index(w)]
                print("  %-62s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
        try:
            t = float(vals["gpu__time_duration.sum"].replace(",", ""))
            tu = units[hdr.index("gpu__time_duration.sum")]
            t_s = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(tu, 1e-6)
            def b(k):
                u = units[hdr.index(k)]
                return float(vals[k].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            tot = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
            print("  -> DRAM traffic %.1f MB, %.0f GB/s over the launch" % (tot / 1e6, tot / t_s / 1e9))
        except Exception as e:  # noqa: BLE001
            print("  (no traffic summary: %s)" % e)
