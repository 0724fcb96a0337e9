#!/bin/bash
# Round-2 final evidence on one B200 (run under gpurun): tests, bench (both arms), ncu launch list with DRAM bytes and --set full captures
# taken from the SAME command as the bench line (python bench.py ...), probes.  Everything lands in gpurun_out/r2_*;
# tools/summarise_r2.py turns it into profiles/r2_*.
set -u
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --timeout=300 > $O/r2_tests_gpu.log 2>&1; tail -2 $O/r2_tests_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -c 300 $O/r2_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err
timeout 200 python tools/gpu_probe.py 512 > $O/r2_per_language.txt 2>&1
timeout 100 python tools/decode_probe.py 1024 > $O/r2_decode_probe.txt 2>&1
timeout 300 python tools/general_probe.py 256 > $O/r2_general.txt 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > $O/r2_plain.log 2>&1 || { echo "plain failed"; tail -5 $O/r2_plain.log; exit 1; }
LPS=$(python -c "import json;print(json.loads(open('$O/r2_bench_n1.json').read().strip().splitlines()[-1])['gpu_launches']//5)")
echo "launches per step: $LPS"
# the two timed device-resident steps of the bench command (skip the three warm-up steps)
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:jtk_ -s $((3 * LPS)) -c $((2 * LPS)) --csv --log-file $O/r2_launches.csv $CMD > $O/r2_ncu1.log 2>&1; echo "ncu list rc=$?"
# split+lookup: the 256 MiB sub-batch of the first timed step (3 warm-up steps x 5 launches + 2)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:jtk_split_lookup -s 17 -c 1 -o $O/r2_split -f $CMD > $O/r2_split.log 2>&1; echo "ncu split rc=$?"
# gather + merge kernels of that sub-batch (side streams off so that ncu sees them one by one); decode kernel of the same command
JTK_SIDE_STREAMS=0 timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:jtk_merge_short|jtk_merge_medium|jtk_gather' -s 85 -c 5 -o $O/r2_merge_gather -f $CMD > $O/r2_merge_gather.log 2>&1; echo "ncu merge/gather rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:jtk_decode_fused -s 2 -c 1 -o $O/r2_decode -f $CMD > $O/r2_decode.log 2>&1; echo "ncu decode rc=$?"
echo done
