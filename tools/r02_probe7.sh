#!/bin/bash
mkdir -p gpurun_out
for v in parsec_b200/libvariant_*.so; do
  echo "== $v" | tee -a gpurun_out/r02_p7_sweep.log
  PB2_LIB_PATH=$PWD/$v timeout 300 python tools/sweep_hbm.py 0,0,0 2>&1 | tail -1 | tee -a gpurun_out/r02_p7_sweep.log
done
timeout 900 python bench.py --steps 50 > gpurun_out/r02_p7_bench.json 2> gpurun_out/r02_p7_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_p7_bench.err; cat gpurun_out/r02_p7_bench.json | cut -c1-3000
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-800
timeout 600 python -m pytest tests/test_mca_component.py -x -q -m gpu 2>&1 | tail -3
