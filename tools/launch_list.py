"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X.csv of `python tools/profile_step.py 2`):
keeps the launches of the LAST forward (second half), groups by kernel, prints total time, share, count.
usage: python tools/launch_list.py gpurun_out/launches.csv "header comment" [forwards in capture] > profiles/rN_launches_step.txt"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
nfwd = int(sys.argv[3]) if len(sys.argv) > 3 else 2          # identical forwards in the capture: keep the last one
half = rows[len(rows) - len(rows) // nfwd:]
agg = collections.OrderedDict()
for k, ns in half:
    k = re.sub(r"\(.*", "", k)[:110]
    d = agg.setdefault(k, [0.0, 0])
    d[0] += ns
    d[1] += 1
tot = sum(d[0] for d in agg.values())
print("# " + (sys.argv[2] if len(sys.argv) > 2 else ""))
print("# per-launch times are cold-cache and serialised: compare SHARES. launches in step: %d, sum %.2f ms" % (len(half), tot / 1e6))
for k, (ns, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%9.3f ms %5.1f%% %5d  %s" % (ns / 1e6, 100 * ns / tot, n, k))
