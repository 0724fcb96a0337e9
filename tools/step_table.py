#!/usr/bin/env python3
"""Per-kernel time and DRAM traffic of ONE step (from one jtk_tile_first_doc_kernel launch to the next) out of an ncu launch list
captured with --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:jtk_ --csv.
Usage: tools/step_table.py launches.csv [json-out]"""
import collections, csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, mi, vi, ui, idi = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
launch = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].split("(")[0].split("::")[-1]
    d = launch.setdefault(r[idi], {"name": name})
    v, u = float(r[vi].replace(",", "")), r[ui]
    if r[mi].startswith("gpu__time"):
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    else:
        v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    d[r[mi]] = v
L = list(launch.values())
starts = [i for i, x in enumerate(L) if x["name"].startswith("jtk_tile_first_doc")]
a = starts[-1] if len(starts) == 1 else starts[-2]
b = len(L) if len(starts) == 1 else starts[-1]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for x in L[a:b]:
    g = agg[x["name"]]
    g[0] += 1
    g[1] += x.get("gpu__time_duration.sum", 0)
    g[2] += x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0)
tot, totb = sum(v[1] for v in agg.values()), sum(v[2] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-34s launches %3d  %9.1f us %5.1f%%  dram %8.1f MB" % (k, v[0], v[1], 100 * v[1] / tot, v[2] / 1e6))
print("one step (serialised, cold cache under ncu): %.1f us in %d launches, dram %.1f MB" % (tot, b - a, totb / 1e6))
if len(sys.argv) > 2:
    json.dump({"dram_bytes_per_step": totb, "launches_per_step": b - a, "serialised_us": tot,
               "per_kernel": {k: {"launches": v[0], "us": v[1], "dram_bytes": v[2]} for k, v in agg.items()}}, open(sys.argv[2], "w"), indent=1)
