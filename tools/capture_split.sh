#!/bin/bash
# ncu --set full capture of one full-size split+lookup launch on the multilingual mix (run under gpurun)
set -u
CMD="python tools/gpu_one.py mix 256 2"
$CMD > gpurun_out/one_plain.log 2>&1 || { echo "plain failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:jtk_split_lookup -s 4 -c 1 -o gpurun_out/split_now -f $CMD > gpurun_out/split_now.log 2>&1; echo "ncu rc=$?"
