"""Aggregate an ncu `--page source --csv` dump by opcode: instructions per warp and stall samples.
usage: ncu -i rep.ncu-rep --page source --csv --kernel-id ::regex:step_kernel:1 | python tools/ncu_opmix.py [warps]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samp, tot = collections.Counter(), collections.Counter(), 0
for r in rows[hi + 1:]:
    if len(r) <= iE or not r[iE].isdigit():
        continue
    src, n, s = r[iS].strip(), int(r[iE]), int(r[iSamp] or 0)
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2).split(".")[0] if m else src
    ops[op] += n
    samp[op] += s
    tot += n
warps = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
print(f"total warp-instructions {tot}  per warp {tot / warps:.1f}")
for op, n in ops.most_common(45):
    print(f"{op:12s} {n / warps:8.1f} /warp   stall samples {samp[op]}")
