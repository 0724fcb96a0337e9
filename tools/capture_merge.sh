#!/bin/bash
# ncu --set full capture of the merge kernels of one full-size sub-batch on the multilingual mix (run under gpurun)
set -u
CMD="python tools/gpu_one.py mix 256 2"
$CMD > gpurun_out/one_plain.log 2>&1 || { echo "plain failed"; exit 1; }
JTK_SIDE_STREAMS=0 ncu --set full --clock-control none --import-source on -k 'regex:jtk_merge_short|jtk_merge_medium' -s 16 -c 4 -o gpurun_out/merge_now -f $CMD > gpurun_out/merge_now.log 2>&1; echo "ncu rc=$?"
