#!/usr/bin/env python3
"""Top source lines (stall samples / instructions) of one launch of an .ncu-rep: tools/ncu_lines.py rep launch_index [top]"""
import collections, csv, io, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--launch-skip", str(idx), "--launch-count", "1", "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, agg, name = None, collections.defaultdict(lambda: [0, 0, 0, ""]), ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        name = r[1]
    elif r[0].isdigit() and len(r) > 8:
        try:
            k = (cur, int(r[0]))
            agg[k][0] += int(r[6]); agg[k][1] += int(r[7]); agg[k][2] += int(r[8]); agg[k][3] = r[1]
        except ValueError:
            pass
ts = sum(v[0] for v in agg.values()) or 1
ti = sum(v[1] for v in agg.values()) or 1
print(name, "| samples", ts, "| warp instructions", ti)
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-16s %5d smp %5.1f%% ins %5.1f%% thr/ins %4.1f | %s" % (f, l, 100 * v[0] / ts, 100 * v[1] / ti, v[2] / max(1, v[1]), v[3].strip()[:110]))
