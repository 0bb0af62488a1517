"""SASS evidence for profiles/: per kernel of librpforest.so the counts of the mnemonics that matter for the parity and
B200 claims (DFMA must be 0 in the exact-arithmetic folds; UBLKCP = cp.async.bulk (TMA), LDGSTS = cp.async, UTMAPF /
CCTL-type prefetch, DMMA = FP64 tensor cores), plus the first lines carrying each B200-specific mnemonic.
Usage: python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "rp-tree_b200", "librpforest.so")
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
filt = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
names = dict(zip(re.findall(r"Function : (\S+)", out), filt))
WATCH = ["DFMA", "DMUL", "DADD", "DMMA", "UBLKCP", "UBLKPF", "LDGSTS", "SYNCS", "SHFL", "ATOMS", "BAR"]
cur, counts, first = None, collections.defaultdict(collections.Counter), collections.defaultdict(dict)
arch = re.findall(r"arch = (sm_\w+)", out)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = names.get(m.group(1), m.group(1))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(2)
        for w in WATCH:
            if op.startswith(w):
                counts[cur][w] += 1
                first[cur].setdefault(w, line.strip()[:110])
print("librpforest.so: cubins for", sorted(set(arch)))
print("%-110s %s" % ("kernel", " ".join("%6s" % w for w in WATCH)))
for kname in sorted(counts):
    short = re.sub(r"\(.*", "", kname)
    print("%-110s %s" % (short[:110], " ".join("%6d" % counts[kname][w] for w in WATCH)))
print()
print("first occurrence of the B200-specific / evidence mnemonics:")
for kname in sorted(first):
    for w in ("UBLKCP", "UBLKPF", "LDGSTS", "DMMA", "SYNCS"):
        if w in first[kname]:
            print("  %-70s %s" % (re.sub(r"\(.*", "", kname)[:70], first[kname][w]))
print()
exact = [k for k in counts if re.match(r"void k_project", k)]
print("exact-arithmetic check: DFMA in k_project* instances =", sum(counts[k]["DFMA"] for k in exact), "(must be 0: innerSD uses separate mul / add roundings)")
