#!/usr/bin/env python3
"""gpurun_out/r1_* (tools/capture_profiles.sh) -> tracked summaries under profiles/."""
import csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
shutil.copy(os.path.join(G, "r1_launches.csv"), os.path.join(P, "r1_launches.csv"))
table = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_table.py"), os.path.join(G, "r1_launches.csv")], capture_output=True, text=True).stdout
open(os.path.join(P, "r1_launches_summary.txt"), "w").write(
    "ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jtk_ (two timed steps of: python bench.py --steps 2 --warmup 3 --no-cpu-baseline)\n"
    "per-launch times are cold-cache and serialised (in production the four merge kernels of a sub-batch overlap on forked streams): compare shares, not absolutes\n\n" + table)
raw = subprocess.run(["ncu", "-i", os.path.join(G, "r1_split_lookup.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out, traffic = [], []
for r in rows[2:]:
    m, u = dict(zip(h, r)), dict(zip(h, units))
    out.append("launch %s" % m.get("ID"))
    for k in keys:
        if k in m:
            out.append("  %-70s %s %s" % (k, m[k], u.get(k, "")))
    st = [(k, float(v.replace(",", ""))) for k, v in m.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v]
    out.append("  stall reasons (warps stalled per issue-active cycle):")
    for k, v in sorted(st, key=lambda kv: -kv[1])[:8]:
        out.append("    %-40s %.3f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    traffic.append(sum(float(m[k].replace(",", "")) * scale.get(u[k], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")))
open(os.path.join(P, "r1_split_lookup_ncu_full.txt"), "w").write(
    "ncu --set full --clock-control none --import-source on -k regex:jtk_split_lookup -s 17 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n"
    "(selected raw metrics; the launch covers the 256 MiB sub-batch = 32768 tiles of the first timed step over the 1 GiB multilingual corpus;\n"
    " a step runs five sub-batches of 16, 64, 256, 512 and 176 MiB)\n\n" + "\n".join(out) + "\n")
summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, "r1_split_lookup.ncu-rep"), "268435456"], capture_output=True, text=True).stdout
open(os.path.join(P, "r1_split_lookup_source_hotspots.txt"), "w").write("per-function / per-line shares from the ncu source page (first captured launch)\n\n" + summ)
# the dominant kernel launches once per sub-batch; the captured launch covers 256 MiB of the step's 1 GiB: scale by 4
per_launch = sum(traffic) / len(traffic)
json.dump({"kernel": "jtk_split_lookup_kernel", "dram_bytes_per_launch": per_launch, "launch": "the 256 MiB sub-batch (32768 tiles) of a step",
           "dram_bytes_per_step": per_launch * 4, "source": "profiles/r1_split_lookup_ncu_full.txt"}, open(os.path.join(P, "tile_kernel_traffic.json"), "w"), indent=1)
mg = os.path.join(G, "r1_merge_gather.ncu-rep")
if os.path.exists(mg):
    raw = subprocess.run(["ncu", "-i", mg, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    out = ["JTK_SIDE_STREAMS=0 ncu --set full --clock-control none -k regex:jtk_merge_short|jtk_merge_medium|jtk_gather -s 85 -c 5 python bench.py --steps 2 --warmup 3 --no-cpu-baseline",
           "(the 256 MiB sub-batch of the first timed step; in production the four merge kernels run side by side on forked streams)", ""]
    for r in rows[2:]:
        m, u = dict(zip(h, r)), dict(zip(h, units))
        out.append(m.get("Kernel Name", "?"))
        for k in keys[1:]:
            if k in m:
                out.append("  %-70s %s %s" % (k, m[k], u.get(k, "")))
        st = [(k, float(v.replace(",", ""))) for k, v in m.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v]
        out.append("  stall reasons (warps stalled per issue-active cycle): " + ", ".join("%s %.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for k, v in sorted(st, key=lambda kv: -kv[1])[:5]))
    open(os.path.join(P, "r1_merge_gather_ncu_full.txt"), "w").write("\n".join(out) + "\n")
for f in ("r1_pcie.txt", "r1_per_language.txt"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f))
for f in ("bench_r1.json", "bench_r1_reference.json"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f))
print(table)
