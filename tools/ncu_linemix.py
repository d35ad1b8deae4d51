"""Join ncu per-SASS-instruction execution counts with nvdisasm line info -> warp-instructions per source line.
usage: python tools/ncu_linemix.py <rep.ncu-rep> <lib.so> <mangled kernel name> [warps] [n_launches_in_rep]"""
import collections
import csv
import io
import re
import subprocess
import sys
import tempfile
import os

rep, lib, kname = sys.argv[1:4]
warps = int(sys.argv[4]) if len(sys.argv) > 4 else 32768
nl = int(sys.argv[5]) if len(sys.argv) > 5 else 1
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
dis = ""
for f in sorted(os.listdir(tmp)):  # one cubin per translation unit: take the one that holds the kernel
    if f.endswith(".cubin"):
        d = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], stdout=subprocess.PIPE, text=True).stdout
        if "\t.section\t.text." + kname + "," in d:
            dis = d
            break
lines_by_off = {}
cur, on = None, False
for ln in dis.splitlines():
    if ln.startswith("\t.section\t.text."):
        on = ln.startswith("\t.section\t.text." + kname + ",")
        continue
    if not on:
        continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines_by_off[int(m.group(1), 16)] = (cur, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", "::regex:" + (sys.argv[6] if len(sys.argv) > 6 else "step_kernel") + ":1"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
iA, iE, iSamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [r for r in rows[hi + 1:] if len(r) > iE and r[iE].isdigit()]
base = int(data[0][iA], 16)
per_line, samp_line, tot = collections.Counter(), collections.Counter(), 0
for r in data:
    off = int(r[iA], 16) - base
    key = lines_by_off.get(off, (None, "?"))[0]
    n = int(r[iE]) / nl
    per_line[key] += n
    samp_line[key] += int(r[iSamp] or 0)
    tot += n
src = {}
print(f"total {tot / warps:.1f} warp-instr per warp")
for key, n in per_line.most_common(60):
    text = ""
    if key:
        fn = [p for p in (os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", key[0]),) if os.path.exists(p)]
        if fn:
            if fn[0] not in src:
                src[fn[0]] = open(fn[0]).read().splitlines()
            text = src[fn[0]][key[1] - 1].strip()[:110]
    print(f"{n / warps:7.1f}  samp {samp_line[key]:5d}  {key}  {text}")
