#!/bin/bash
# second GPU call: the real MCA component on the GPU, vs the reference's cuda component and CPU bodies; kernel variant sweep
mkdir -p gpurun_out
B=oracle/_ref/bin/ex05_b200
echo "== b200 small"; PARSEC_MCA_device_b200_enabled=1 timeout 120 $B -K 64 -t 65536 -m gpu -c 8 -v 2>&1 | tail -4
echo "== b200 K=4096"; PARSEC_MCA_device_b200_enabled=1 timeout 300 $B -K 4096 -t 65536 -m gpu -c 16 -r 3 -v 2>&1 | tail -5
echo "== ref cuda K=4096"; PARSEC_MCA_device_cuda_enabled=1 timeout 300 $B -K 4096 -t 65536 -m gpu -c 16 -r 3 -v 2>&1 | tail -5
echo "== cpu K=1024 all cores"; timeout 300 $B -K 1024 -t 65536 -m cpu -r 3 -v 2>&1 | tail -5
for v in parsec_b200/libvariant_*.so; do
  echo "== $v"
  PB2_LIB_PATH=$PWD/$v timeout 300 python tools/sweep_hbm.py 0,0,0 2>&1 | tail -1
done
