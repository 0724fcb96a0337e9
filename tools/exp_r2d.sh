#!/bin/bash
# A/B runs on one B200: single-read decode variants; ncu full capture of the general-pattern slice kernel (DFA)
set -u
O=gpurun_out
L=jtokkit_b200/libjtokkit_b200
for v in _old _dec6 "" _s83 _s45 _s46 _s48; do echo "decode variant '$v'"; JTK_LIB=$L$v.so timeout 120 python tools/decode_probe.py 1024 2>&1 | tail -1; done > $O/r2d_decode.txt 2>&1
cat $O/r2d_decode.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jtk_general_slice -c 1 -o $O/r2d_slice -f python tools/general_one.py 128 1 > $O/r2d_slice.log 2>&1
tail -3 $O/r2d_slice.log
