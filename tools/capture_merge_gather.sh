set -u
python tools/gpu_probe.py 256 > gpurun_out/probe_r1b.txt 2>&1; echo "probe rc=$?"
CMD="python tools/gpu_one.py mix 256 2"
$CMD > gpurun_out/one_plain.log 2>&1 || { echo "plain failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k 'regex:jtk_merge_short|jtk_merge_medium|jtk_gather' -s 15 -c 10 -o gpurun_out/r1b_merge_gather $CMD > gpurun_out/r1b_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r1b_ncu.log
cat gpurun_out/probe_r1b.txt
