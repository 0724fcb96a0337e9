#!/bin/bash
# A/B runs on one B200: decode variants, gather variants, launch list of the general-pattern path
set -u
O=gpurun_out
L=jtokkit_b200/libjtokkit_b200
for v in _old "" _dec5 _dec6 _dec8; do echo "decode variant '$v'"; JTK_LIB=$L$v.so timeout 120 python tools/decode_probe.py 1024 2>&1 | tail -1; done > $O/r2c_decode.txt 2>&1
cat $O/r2c_decode.txt
for v in _old "" _g128 _g512; do echo "encode variant '$v'"; JTK_LIB=$L$v.so timeout 120 python tools/gpu_one.py mix 1024 4 2>&1 | tail -1; done > $O/r2c_gather.txt 2>&1
cat $O/r2c_gather.txt
timeout 100 python tools/general_one.py 128 2 > $O/r2c_general_plain.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jtk_ -c 200 --csv --log-file $O/r2c_general_launches.csv python tools/general_one.py 128 2 > $O/r2c_general_ncu.log 2>&1
tail -2 $O/r2c_general_plain.log
python tools/step_table.py $O/r2c_general_launches.csv | head -20
