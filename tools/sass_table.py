"""Per-kernel table of the SASS mnemonics that prove which hardware paths a kernel uses (B200_PROFILING.md): tcgen05 MMA (UTCHMMA /
UTCQMMA ...), TMEM loads / stores (LDTM / STTM), TMA (UTMALDG / UTMASTG / UBLKCP), legacy tensor cores (HMMA), ldmatrix (LDSM),
cp.async (LDGSTS), vector reductions (RED / REDG), plus the instruction count.
usage: python tools/sass_table.py [lib.so] > profiles/rN_sass_stats.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "dtlr_b200/libdtlr_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
COLS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "LDSM", "LDGSTS", "RED", "MUFU", "BAR", "SYNCS"]
cur, hist = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", line)
    if m and cur:
        ins = m.group(1).split()
        op = (ins[1] if ins[0].startswith("@") else ins[0]).split(".")[0]
        hist[cur]["_n"] += 1
        for c in COLS:
            if op.startswith(c):
                hist[cur][c] += 1
names = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
print("# %s: SASS mnemonic counts per kernel (cuobjdump -sass); tcgen05 = UTCHMMA, TMEM = LDTM/STTM, TMA = UTMALDG/UTMASTG" % lib)
print("%-7s " % "instr" + " ".join("%7s" % c for c in COLS) + "  kernel")
tot = collections.Counter()
for (k, h), n in sorted(zip(hist.items(), names), key=lambda t: t[1]):
    n = re.sub(r"\(.*", "", n).replace("void dtlr::", "").replace("dtlr::", "")
    print("%-7d " % h["_n"] + " ".join("%7s" % (h[c] or ".") for c in COLS) + "  " + n[:100])
    tot.update(h)
print("%-7d " % tot["_n"] + " ".join("%7d" % tot[c] for c in COLS) + "  TOTAL")
