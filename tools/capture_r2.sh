#!/bin/bash
# Round-2 evidence on one B200 (run under gpurun): tests, bench (both arms), per-language probe, ncu launch list with DRAM bytes,
# ncu --set full captures of the split+lookup kernel, of the gather / merge kernels and of the decode kernel.
# Everything lands in gpurun_out/; tools/summarise_r2.sh turns it into profiles/r2_*.
set -u
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --timeout=300 > $O/r2_tests_gpu.log 2>&1; tail -2 $O/r2_tests_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -c 200 $O/r2_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference.json 2> $O/r2_bench_reference.err
timeout 200 python tools/gpu_probe.py 512 > $O/r2_per_language.txt 2>&1
timeout 100 python tools/decode_probe.py 1024 > $O/r2_decode_probe.txt 2>&1
CMD="python tools/gpu_one.py mix 1024 1"
timeout 120 $CMD > $O/one_plain.log 2>&1 || { echo "plain failed"; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:jtk_ -c 400 --csv --log-file $O/r2_launches.csv $CMD > $O/ncu_plain.log 2>&1
CMD2="python tools/gpu_one.py mix 256 2"
timeout 120 $CMD2 > $O/one_plain2.log 2>&1 || { echo "plain2 failed"; exit 1; }
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jtk_split_lookup -s 4 -c 1 -o $O/r2_split -f $CMD2 > $O/r2_split.log 2>&1
JTK_SIDE_STREAMS=0 timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:jtk_merge_short|jtk_merge_medium|jtk_gather' -s 20 -c 5 -o $O/r2_merge_gather -f $CMD2 > $O/r2_merge_gather.log 2>&1
timeout 100 python tools/decode_probe.py 256 > $O/dec_plain.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:jtk_decode_fused -s 2 -c 1 -o $O/r2_decode -f python tools/decode_probe.py 256 > $O/r2_decode.log 2>&1
echo done
