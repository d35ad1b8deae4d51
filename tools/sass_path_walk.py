"""Dynamic instruction count of ONE path through a kernel's loop body, offline (no GPU): annotate the SASS with source lines (nvdisasm -g),
then walk from a start label along the fall-through / taken decisions given on the command line until the back-branch to the start label.

    python tools/sass_path_walk.py list  <obj> <mangled-kernel-substring> > listing.txt        # address, source line, instruction
    python tools/sass_path_walk.py walk  listing.txt [.L_x_123] ["addr:1,addr:0,..."]          # 1 = branch at addr taken; default: not taken

`walk` prints every conditional branch it meets with its source line (so the decisions can be filled in: e.g. "skip the reset path",
"no new sub-episode", "walk block cached / not cached"), the instruction count of the path, the opcode mix and the busiest source lines.
This is how the round-2 instruction work on env_step was steered between GPU sessions (profiles/r02b_summary.md): the PMSM rollout path went
457 -> 380 instructions by this count, 498 -> 400 by ncu's smsp__inst_executed (terminations and new sub-episodes add ~5 %)."""
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile


def listing(obj, pat):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
    out = "".join(subprocess.run(["nvdisasm", "-g", "-c", c], stdout=subprocess.PIPE, text=True).stdout for c in glob.glob(tmp + "/*.cubin"))
    on, line = False, None
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            on = pat in m.group(1)
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", ln)
        if m:
            print(f"{int(m.group(1), 16):05x} {line[0][:14] if line else '?'}:{line[1] if line else 0:<5d} {m.group(2)}")
        elif re.match(r"\s*\.L_x_\d+:", ln):
            print(ln.strip())


def walk(path, start=None, decisions=""):
    dec = dict((int(a.split(":")[0], 16), a.split(":")[1] == "1") for a in decisions.split(",") if a)
    ins, labels = [], {}
    for ln in open(path):
        ln = ln.rstrip("\n")
        m = re.match(r"(\.L_x_\d+):", ln)
        if m:
            labels[m.group(1)] = len(ins)
            continue
        m = re.match(r"([0-9a-f]{5}) (\S+):(\d+)\s+(.*)", ln)
        if m:
            ins.append((int(m.group(1), 16), int(m.group(3)), m.group(4).strip()))
    start = start or min(labels, key=labels.get)  # the first label of a rollout kernel is its loop head
    pc, n, per, ops = labels[start], 0, collections.Counter(), collections.Counter()
    for _ in range(20000):
        a, ln, t = ins[pc]
        n += 1
        per[ln] += 1
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        ops[op.split(".")[0]] += 1
        m = re.search(r"(BRA(?:\.U)?)\s+(.*?)`\((\.L_x_\d+)\)", t)
        if m:
            cond = t.startswith("@") or m.group(2).strip() != ""
            tgt = m.group(3)
            if tgt == start and (not cond or dec.get(a, True)):
                break
            if not cond:
                pc = labels[tgt]
                continue
            taken = dec.get(a)
            if taken is None:
                print(f"  branch {a:05x} line {ln}: {t[:70]} -> default not taken")
                taken = False
            if taken:
                pc = labels[tgt]
                continue
        if t.startswith("EXIT"):
            break
        pc += 1
    print("dynamic instructions on path:", n)
    print(sorted(ops.items(), key=lambda x: -x[1])[:25])
    print(sorted(per.items(), key=lambda x: -x[1])[:40])


if __name__ == "__main__":
    if sys.argv[1] == "list":
        listing(sys.argv[2], sys.argv[3])
    else:
        walk(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] else None, sys.argv[4] if len(sys.argv) > 4 else "")
