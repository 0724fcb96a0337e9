#!/usr/bin/env python3
"""Summarise an .ncu-rep of the tile kernel: headline metrics, stall reasons, per-function instruction / sample shares.
Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [nbytes [kernel-regex]]"""
import collections, csv, io, re, subprocess, sys, os

rep = sys.argv[1]
KFILTER = ["-k", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def run(args):
    return subprocess.run(["ncu", "-i", rep] + KFILTER + args, capture_output=True, text=True).stdout

raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
hdr, vals = raw[0], raw[2] if len(raw) > 2 else raw[1]
m = dict(zip(hdr, vals))
def g(k):
    return m.get(k, "?")
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for k in keys:
    print("%-70s %s" % (k, g(k)))
print("--- stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active / pcsamp)")
st = [(k, v) for k, v in m.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
for k, v in sorted(st, key=lambda kv: -float(kv[1].replace(",", "") or 0))[:10]:
    print("  %-80s %s" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))

rows = list(csv.reader(io.StringIO(run(["--page", "source", "--print-source", "cuda,sass", "--csv"]))))
cur_file = None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] and r[0].isdigit() and len(r) > 8:
        key = (cur_file, int(r[0]))
        try:
            agg[key][0] += int(r[6]); agg[key][1] += int(r[7]); agg[key][2] += int(r[8])
        except ValueError:
            pass
        agg[key][3] = r[1]
tot_i = sum(v[1] for v in agg.values()) or 1
tot_s = sum(v[0] for v in agg.values()) or 1
def funcs(path):
    out, cur = {}, None
    for i, l in enumerate(open(path).read().split("\n"), 1):
        mm = re.match(r"^(?:template.*)?(?:JTK_HD|__device__|__global__|static)\s+[\w:<>\*& ]*?\s*\**(\w+)\(", l)
        if mm:
            cur = mm.group(1)
        mm = re.search(r"/\* ---- (P\d)", l)
        if mm:
            cur = "tile:" + mm.group(1)
        out[i] = cur
    return out
fmap = {"jtk_device.cuh": funcs(os.path.join(ROOT, "jtokkit_b200/csrc/jtk_device.cuh")), "jtk_kernels.cu": funcs(os.path.join(ROOT, "jtokkit_b200/csrc/jtk_kernels.cu"))}
groups = collections.defaultdict(lambda: [0, 0, 0])
for (f, l), v in agg.items():
    gname = (f.split(".")[0][4:] + ":" + str(fmap[f].get(l))) if f in fmap else f
    for i in range(3):
        groups[gname][i] += v[i]
print("--- by function: instr share, stall-sample share, active threads per instruction")
for gname, v in sorted(groups.items(), key=lambda kv: -kv[1][1])[:32]:
    print("  %-36s ins %5.1f%%  smp %5.1f%%  thr/ins %4.1f" % (gname, 100 * v[1] / tot_i, 100 * v[0] / tot_s, v[2] / max(v[1], 1)))
if len(sys.argv) > 2:
    print("warp instructions per input byte: %.2f (per 7 072-byte tile: %.0f)" % (tot_i / int(sys.argv[2]), tot_i / (int(sys.argv[2]) / 7072)))
print("--- top lines by stall samples")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
    print("  %-16s %4d smp %5.1f%% ins %5.1f%% | %s" % (k[0], k[1], 100 * v[0] / tot_s, 100 * v[1] / tot_i, v[3].strip()[:100]))
