#!/bin/bash
# A/B runs on one B200: decode keep/occupancy variants; general path flat DFA loop (main) vs nested loops (_dec6 build), tests of the general path
set -u
O=gpurun_out
L=jtokkit_b200/libjtokkit_b200
for v in _old _dec6 "" _k4c6 _k4c8 _k8c5 _k8c6 _k8c8; do echo "decode variant '$v'"; JTK_LIB=$L$v.so timeout 120 python tools/decode_probe.py 1024 2>&1 | tail -1; done > $O/r2e_decode.txt 2>&1
cat $O/r2e_decode.txt
for v in _dec6 ""; do echo "general variant '$v'"; JTK_LIB=$L$v.so timeout 120 python tools/general_one.py 256 3 2>&1 | tail -1; done > $O/r2e_general.txt 2>&1
cat $O/r2e_general.txt
timeout 600 python -m pytest tests/test_gpu_general.py tests/test_gpu_api.py -m gpu -x -q --timeout=300 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jtk_general -c 20 --csv --log-file $O/r2e_general_launches.csv python tools/general_one.py 128 2 > $O/r2e_general_ncu.log 2>&1
grep -o '"jtk_general[a-z_]*.*' $O/r2e_general_launches.csv | awk -F'","' '{print $1, $NF}' | head
