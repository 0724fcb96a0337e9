#!/usr/bin/env python3
"""gpurun_out/r2_* (tools/capture_r2.sh) -> tracked summaries under profiles/r2_*."""
import csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def raw_rows(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    return rows[0], rows[1], rows[2:]


def describe(rep, head):
    h, units, rows = raw_rows(rep)
    out, traffic = list(head) + [""], []
    for r in rows:
        m, u = dict(zip(h, r)), dict(zip(h, units))
        out.append("launch %s  %s" % (m.get("ID"), m.get("Kernel Name", "?")))
        for k in KEYS[1:]:
            if k in m:
                out.append("  %-70s %s %s" % (k, m[k], u.get(k, "")))
        st = [(k, float(v.replace(",", ""))) for k, v in m.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v]
        out.append("  stall reasons (warps stalled per issue-active cycle):")
        for k, v in sorted(st, key=lambda kv: -kv[1])[:8]:
            out.append("    %-40s %.3f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
        traffic.append((m.get("Kernel Name", "?"), sum(float(m[k].replace(",", "")) * SCALE.get(u[k], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in m)))
    return "\n".join(out) + "\n", traffic


def cp(src, dst=None):
    s = os.path.join(G, src)
    if os.path.exists(s):
        shutil.copy(s, os.path.join(P, dst or src))


os.makedirs(P, exist_ok=True)
L = os.path.join(G, TAG + "_launches.csv")
if os.path.exists(L):
    cp(TAG + "_launches.csv")
    table = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "step_table.py"), L, os.path.join(P, TAG + "_step_traffic.json")], capture_output=True, text=True).stdout
    open(os.path.join(P, TAG + "_launches_summary.txt"), "w").write(
        "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:jtk_ -s 3*47 -c 2*47 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n"
        "(the two timed device-resident steps of the bench command over its 1 GiB multilingual corpus; the table is the first of them.\n"
        " per-launch times are cold-cache and serialised - in production the merge kernels of a sub-batch overlap on forked streams: compare shares, not absolutes)\n\n" + table)
    print(table)
rep = os.path.join(G, TAG + "_split.ncu-rep")
if os.path.exists(rep):
    txt, traffic = describe(rep, ["ncu --set full --clock-control none --import-source on -k regex:jtk_split_lookup -s 17 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline",
                                  "(selected raw metrics; the launch is the 256 MiB sub-batch of the first timed step over the 1 GiB multilingual corpus;",
                                  " a step runs five sub-batches of 16, 64, 256, 512 MiB and the rest)"])
    open(os.path.join(P, TAG + "_split_lookup_ncu_full.txt"), "w").write(txt)
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "268435456"], capture_output=True, text=True).stdout
    open(os.path.join(P, TAG + "_split_lookup_source_hotspots.txt"), "w").write("per-function / per-line shares from the ncu source page\n\n" + summ)
rep = os.path.join(G, TAG + "_merge_gather.ncu-rep")
if os.path.exists(rep):
    txt, _ = describe(rep, ["JTK_SIDE_STREAMS=0 ncu --set full --clock-control none --import-source on -k 'regex:jtk_merge_short|jtk_merge_medium|jtk_gather' -s 85 -c 5 python bench.py --steps 2 --warmup 3 --no-cpu-baseline",
                            "(the 256 MiB sub-batch of the first timed step)",
                            "(in production the merge kernels of a sub-batch run side by side on forked streams)"])
    open(os.path.join(P, TAG + "_merge_gather_ncu_full.txt"), "w").write(txt)
rep = os.path.join(G, TAG + "_decode.ncu-rep")
if os.path.exists(rep):
    txt, traffic = describe(rep, ["ncu --set full --clock-control none --import-source on -k regex:jtk_decode_fused -s 2 -c 1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline",
                                  "(the bench's decode leg: the 404 M ids of the step back to the 1 GiB corpus, device-resident; first timed launch)"])
    open(os.path.join(P, TAG + "_decode_ncu_full.txt"), "w").write(txt)
for f in ("per_language.txt", "decode_probe.txt", "general.txt", "configs_8gpu.txt", "pcie_8gpu.txt", "tests_gpu.log", "tests_multi_2gpu.log",
          "bench_n1.json", "bench_n2.json", "bench_n4.json", "bench_n8.json", "bench_reference.json"):
    cp(TAG + "_" + f)
