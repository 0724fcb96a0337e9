set -u
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r1_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:jtk_split_lookup -s 17 -c 1 -o gpurun_out/r1_split_lookup -f $CMD > gpurun_out/r1_ncu2.log 2>&1; echo "ncu full rc=$?"
JTK_SIDE_STREAMS=0 ncu --set full --clock-control none --import-source on -k 'regex:jtk_merge_short|jtk_merge_medium|jtk_gather' -s 85 -c 5 -o gpurun_out/r1_merge_gather -f $CMD > gpurun_out/r1_ncu3.log 2>&1; echo "ncu merge/gather rc=$?"
