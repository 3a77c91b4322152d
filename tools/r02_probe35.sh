#!/bin/bash
uptime
for i in 1 2; do PARSEC_MCA_device_b200_enabled=1 timeout 120 oracle/_ref/bin/ex05_b200 -K 4096 -t 65536 -m gpu -c 32 -r 8 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print([round(t*1e3,1) for t in d['times_s']], d['errors'], d['b200']['kernel_launches'])"; done
python tools/prof_gemm.py 2>&1 | tail -1
