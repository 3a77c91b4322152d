#!/bin/bash
B=oracle/_ref/bin/ex05_b200
mkdir -p gpurun_out
for v in parsec_b200/libparsec_b200.so parsec_b200/libvariant_*.so; do
  echo "== $v" | tee -a gpurun_out/r02_p12_sweep.log
  PB2_LIB_PATH=$PWD/$v timeout 300 python tools/sweep_hbm.py 0,0,0 2>&1 | tail -1 | tee -a gpurun_out/r02_p12_sweep.log
done
for c in 16 32; do
echo "== e2e timing, -c $c"
PARSEC_MCA_device_b200_enabled=1 timeout 120 $B -K 4096 -t 65536 -m gpu -c $c -r 8 -v 2>&1 | grep -E "repeat|times_s" | cut -c1-400
done
timeout 900 python -m pytest tests/test_mca_component.py tests/test_stream_gpu.py tests/test_engine_gpu.py -x -q -m gpu 2>&1 | tail -3
