#!/usr/bin/env python3
"""SASS instruction histogram per kernel of the built library (cuobjdump -sass): tools/sass_histogram.py > profiles/rN_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "jtokkit_b200", "libjtokkit_b200.so")], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
out, tot = [], collections.Counter()
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    ops = collections.Counter()
    for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", f, re.M):
        ops[m.group(1).split(".")[0]] += 1
        tot[m.group(1).split(".")[0]] += 1
    short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "")
    out.append((short, sum(ops.values()), ops))
print("SASS instruction histogram per kernel of jtokkit_b200/libjtokkit_b200.so (cuobjdump -sass, sm_100a); top 14 mnemonics each")
print("data-path evidence: UBLKCP = cp.async.bulk (bulk asynchronous copy global -> shared, the TMA engine), SYNCS = mbarrier operations\n")
for name, n, ops in sorted(out, key=lambda x: -x[1]):
    print("%s: %d instructions" % (name[:140], n))
    print("   " + " ".join("%s:%d" % kv for kv in ops.most_common(14)))
    extra = {k: ops[k] for k in ("UBLKCP", "SYNCS", "LDGSTS", "REDUX", "ATOMS", "ATOMG", "LDG", "STG", "LDS", "STS", "BAR", "SHFL", "VOTE") if ops[k]}
    print("   data path: %s" % extra)
print("\nwhole library: UBLKCP %d, SYNCS %d, LDGSTS %d, REDUX %d, ATOMS %d, ATOMG %d" % (tot["UBLKCP"], tot["SYNCS"], tot["LDGSTS"], tot["REDUX"], tot["ATOMS"], tot["ATOMG"]))
