#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_p6_all.log 2>&1; echo "all rc=$?"; tail -4 gpurun_out/r02_p6_all.log
for v in parsec_b200/libparsec_b200.so parsec_b200/libvariant_*.so; do
  echo "== $v" | tee -a gpurun_out/r02_p6_sweep.log
  PB2_LIB_PATH=$PWD/$v timeout 300 python tools/sweep_hbm.py 0,0,0 2>&1 | tail -1 | tee -a gpurun_out/r02_p6_sweep.log
done
