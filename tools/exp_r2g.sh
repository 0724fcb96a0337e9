#!/bin/bash
set -u
O=gpurun_out
L=jtokkit_b200/libjtokkit_b200
timeout 700 python -m pytest tests -m gpu -x -q --timeout=300 > $O/r2u_tests.log 2>&1; tail -3 $O/r2u_tests.log
for v in _dec6 ""; do echo "encode variant '$v'"; JTK_LIB=$L$v.so timeout 120 python tools/gpu_one.py mix 1024 4 2>&1 | tail -1; done > $O/r2g_encode.txt 2>&1
cat $O/r2g_encode.txt
