#!/bin/bash
# Run under gpurun on a B200: the bench, then the ncu launch list and one `--set full` capture of the dominant kernel for
# the very same command (B200_PROFILING.md recipe).  Outputs land in gpurun_out/; tools/summarise_profiles.py turns them into
# the tracked files under profiles/.
set -u
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_reference.json 2> gpurun_out/bench_r1_reference.err; echo "reference rc=$?"
$CMD > gpurun_out/r1_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
# launches of the two timed steps (skip: 3 warm-up steps x launches per step; the bench prints launches per step)
LPS=$(python -c "import json;print(json.load(open('gpurun_out/bench_r1.json'))['roofline']['launches_per_step'])")
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jtk_ -s $((3 * LPS)) -c $((2 * LPS)) --csv --log-file gpurun_out/r1_launches.csv $CMD > gpurun_out/r1_ncu1.log 2>&1; echo "ncu list rc=$?"
# the sub-batches of a step cover 16, 64, 256, 512 and 176 MiB: capture the 256 MiB one of the first timed step (3 warm-up steps x 5 launches + 2)
ncu --set full --clock-control none --import-source on -k regex:jtk_split_lookup -s 17 -c 1 -o gpurun_out/r1_split_lookup -f $CMD > gpurun_out/r1_ncu2.log 2>&1; echo "ncu full rc=$?"
# the same 256 MiB sub-batch: one launch of each merge kernel and of the gather kernel (side streams off so that ncu sees them one by one)
JTK_SIDE_STREAMS=0 ncu --set full --clock-control none --import-source on -k 'regex:jtk_merge_short|jtk_merge_medium|jtk_gather' -s 85 -c 5 -o gpurun_out/r1_merge_gather -f $CMD > gpurun_out/r1_ncu3.log 2>&1; echo "ncu merge/gather rc=$?"
python tools/pcie_probe.py > gpurun_out/r1_pcie.txt 2>&1
python tools/gpu_probe.py 512 > gpurun_out/r1_per_language.txt 2>&1
