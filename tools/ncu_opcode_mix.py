"""Dynamic opcode mix of one kernel from an ncu report's source page:

    ncu -i REPORT.ncu-rep --page source --csv --kernel-name regex:k_leaf_hash > src.csv
    python tools/ncu_opcode_mix.py src.csv [permutations]
"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
si, ei, st = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
cnt, smp, tot = collections.Counter(), collections.Counter(), 0
for r in rows[rows.index(hdr) + 1:]:
    if len(r) <= ei:
        continue
    src = re.sub(r"^@!?U?P\d+\s+", "", r[si].strip())
    op = src.split()[0].rstrip(";") if src else "?"
    n = int(r[ei] or 0)
    cnt[op] += n
    tot += n
    smp[op] += int(r[st] or 0)
print("kernel:", rows[0][1][:120])
print("warp instructions executed: %d  (thread instructions %d)" % (tot, tot * 32))
if len(sys.argv) > 2:
    print("per permutation (%s permutations): %.0f thread instructions" % (sys.argv[2], tot * 32 / float(sys.argv[2])))
for op, n in cnt.most_common(24):
    print("%-24s %6.2f %%   stall samples %5.1f %%" % (op, 100 * n / tot, 100 * smp[op] / max(1, sum(smp.values()))))
