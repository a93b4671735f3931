"""Per-kernel SASS opcode histogram of libdtlr_b200.so (static check before spending GPU time)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "dtlr_b200/libdtlr_b200.so"
pat = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None
hist = {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()[:110]
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", line)
    if m and cur:
        ins = m.group(1).split()
        op = ins[1] if ins[0].startswith("@") else ins[0]
        hist[cur][op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "LDG", "STG", "STS", "SHFL", "UTC", "LDTM", "UTMA", "UBLKCP")) and "." in op else "")] += 1
for k, h in hist.items():
    if pat in k:
        print("==", k, "total", sum(h.values()))
        print("   ", ", ".join("%s:%d" % kv for kv in h.most_common(top)))
