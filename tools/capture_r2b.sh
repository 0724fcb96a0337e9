#!/bin/bash
# general-pattern DFA: GPU tests + probe (run under gpurun)
set -u
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --timeout=300 > $O/r2s_tests.log 2>&1; tail -3 $O/r2s_tests.log
timeout 300 python tools/general_probe.py 256 > $O/r2_general.txt 2>&1; cat $O/r2_general.txt
