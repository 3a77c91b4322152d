#!/bin/bash
B=oracle/_ref/bin/ex05_b200
mkdir -p gpurun_out
for v in parsec_b200/libvariant_*.so; do
  echo "== $v" | tee -a gpurun_out/r02_p13_sweep.log
  PB2_LIB_PATH=$PWD/$v timeout 300 python tools/sweep_hbm.py 0,0,0 2>&1 | tail -1 | tee -a gpurun_out/r02_p13_sweep.log
done
echo "== e2e profile, -c 16"
PARSEC_B200_PROFILE=1 PARSEC_MCA_device_b200_enabled=1 timeout 120 $B -K 4096 -t 65536 -m gpu -c 16 -r 5 -v 2>&1 | grep -E "repeat|Mcycles" | cut -c1-300
echo "== serial completion"
PARSEC_MCA_device_b200_parallel_completion=0 PARSEC_B200_PROFILE=1 PARSEC_MCA_device_b200_enabled=1 timeout 120 $B -K 4096 -t 65536 -m gpu -c 16 -r 4 -v 2>&1 | grep -E "repeat|Mcycles" | cut -c1-300
grep MHz /proc/cpuinfo | sort | uniq -c | sort -rn | head -3; grep "model name" /proc/cpuinfo | head -1
