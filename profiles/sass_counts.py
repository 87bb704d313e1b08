"""Per-kernel counts of the SASS instructions that show what the kernels are built from (TMA bulk copies / reductions,
mbarriers, peer-capable loads and stores, atomics), from `cuobjdump -sass` of the in-tree library.
    python profiles/sass_counts.py > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cn-rma_b200", "libcnrma_b200.so")
KEYS = ["UBLKCP", "UBLKRED", "UBLKPF", "SYNCS", "UTMACMDFLUSH", "UTMALDG", "UTMASTG", "UTC", "HMMA", "LDG", "STG", "LDS", "STS",
        "ATOM", "RED", "MUFU", "BAR"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
demangle = lambda names: subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
counts, order, cur = {}, [], None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m:
        op = m.group(1)
        counts[cur]["_all"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k in ("UTC", "SYNCS", "ATOM", "RED") and op.startswith(k)):
                counts[cur][k] += 1
names = demangle(order)
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: cubins for {', '.join(arch)}; {len(order)} kernels")
print("# UBLKCP = cp.async.bulk (TMA bulk copy), UBLKRED = cp.reduce.async.bulk, SYNCS = mbarrier ops, UTMACMDFLUSH = bulk-group")
print("# commit/wait; no UTC*MMA / HMMA anywhere: the path is gather / reduce, nothing is a contraction.\n")
tot = collections.Counter()
w = max(len(re.sub(r"\(.*", "", n)) for n in names)
print(f"{'kernel':{w}s} {'instr':>7s} " + " ".join(f"{k:>7s}" for k in KEYS))
for fn, name in sorted(zip(order, names), key=lambda t: t[1]):
    c = counts[fn]
    tot.update(c)
    print(f"{re.sub(r'\(.*', '', name):{w}s} {c['_all']:7d} " + " ".join(f"{c[k]:7d}" for k in KEYS))
print(f"{'TOTAL':{w}s} {tot['_all']:7d} " + " ".join(f"{tot[k]:7d}" for k in KEYS))
