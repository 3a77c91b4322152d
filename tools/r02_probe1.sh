#!/bin/bash
# first GPU call of round 2: parity of the refactored kernels + streaming ring, launch-bound sweep, L2 ceiling
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r02_p1_smi.txt 2>&1
timeout 900 python -m pytest tests/test_stream_gpu.py -x -q -m gpu > gpurun_out/r02_p1_stream.log 2>&1; echo "stream rc=$?" 
tail -5 gpurun_out/r02_p1_stream.log
timeout 1200 python -m pytest tests -x -q -m gpu --deselect tests/test_stream_gpu.py > gpurun_out/r02_p1_all.log 2>&1; echo "all rc=$?"
tail -5 gpurun_out/r02_p1_all.log
./tools/l2_probe > gpurun_out/r02_p1_l2.json 2>&1; cat gpurun_out/r02_p1_l2.json
for v in libparsec_b200 libvariant_mb20 libvariant_mb16; do
  echo "== $v" | tee -a gpurun_out/r02_p1_sweep.log
  PB2_LIB_PATH=$PWD/parsec_b200/$v.so timeout 300 python tools/sweep_hbm.py 0,64,0 2>&1 | tee -a gpurun_out/r02_p1_sweep.log
  PB2_LIB_PATH=$PWD/parsec_b200/$v.so timeout 300 python tools/quick_ex05.py 4096 0 64 2>&1 | grep -E "valid=False it=3|valid=True it=3|ep " | tee -a gpurun_out/r02_p1_sweep.log
done
