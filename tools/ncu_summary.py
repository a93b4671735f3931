"""Summarise an .ncu-rep (read on the CPU box): python tools/ncu_summary.py rep.ncu-rep [regex-of-metric-names]"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|sm__throughput.avg.pct|"
                 r"sm__inst_executed.avg.per_cycle_elapsed|smsp__inst_executed.sum$|issue_active.avg.pct|warps_active.avg.pct|"
                 r"registers_per_thread|occupancy_limit|waves_per|wavefronts_mem_shared.sum$|bank_conflicts.*shared.sum$|"
                 r"issue_stalled_.*_per_warp_active|pipe_tensor.*pct|l1tex__throughput.avg.pct|lts__throughput.avg.pct|"
                 r"smsp__thread_inst_executed_per_inst_executed.ratio|sm__pipe_.*cycles_active.avg.pct_of_peak_sustained_active$")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==== %s  grid=%s block=%s" % (r[hdr.index("Kernel Name")][:100], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
    for i, h in enumerate(hdr):
        if pat.search(h):
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if "issue_stalled" in h and v < 0.05:
                continue
            print("  %-95s %14s %s" % (h, r[i], units[i]))
