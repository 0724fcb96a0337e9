#!/bin/bash
set -u
O=gpurun_out
L=jtokkit_b200/libjtokkit_b200
timeout 700 python -m pytest tests -m gpu -x -q --timeout=300 > $O/r2t_tests.log 2>&1; tail -3 $O/r2t_tests.log
for v in _dec6 ""; do echo "general variant '$v'"; JTK_LIB=$L$v.so timeout 120 python tools/general_one.py 256 3 2>&1 | tail -1; done > $O/r2f_general.txt 2>&1
cat $O/r2f_general.txt
timeout 120 python tools/decode_probe.py 1024 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jtk_general_slice -c 1 -o $O/r2f_slice -f python tools/general_one.py 128 1 > $O/r2f_slice.log 2>&1
tail -2 $O/r2f_slice.log
