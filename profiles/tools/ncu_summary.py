#!/usr/bin/env python
"""Key raw metrics + stall mix of the first kernel in an ncu --set full report.
usage: ncu_summary.py report.ncu-rep [title]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sm__icc_request_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-72s %s %s" % (w, vals[i][:90], units[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
stall = {c: h.index(c) for c in h if c.startswith("stall_") and "Not Issued" not in c}
isamp = h.index("# Samples")
tot, mix = 0, collections.Counter()
for r in rows[2:]:
    if len(r) < len(h) or not r[isamp].isdigit():
        continue
    tot += int(r[isamp])
    for c, i in stall.items():
        if r[i] and r[i].isdigit():
            mix[c] += int(r[i])
print("stall samples %d: %s" % (tot, ", ".join("%s %.1f%%" % (c.replace("stall_", ""), 100.0 * v / max(tot, 1)) for c, v in mix.most_common(9))))
