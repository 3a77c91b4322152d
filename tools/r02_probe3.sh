#!/bin/bash
B=oracle/_ref/bin/ex05_b200
ulimit -c 0
for c in 8 16 32 64; do
echo "== b200 K=4096 c=$c"; PARSEC_MCA_device_b200_enabled=1 timeout 300 $B -K 4096 -t 65536 -m gpu -c $c -r 4 2>&1 | cut -c1-420 | tail -1
done
echo "== b200 K=4096 c=16 serial completion"; PARSEC_MCA_device_b200_parallel_completion=0 PARSEC_MCA_device_b200_enabled=1 timeout 300 $B -K 4096 -t 65536 -m gpu -c 16 -r 4 2>&1 | cut -c1-420 | tail -1
echo "== b200 K=4096 c=16 wb"; PARSEC_MCA_device_b200_enabled=1 timeout 300 $B -K 4096 -t 65536 -m gpu -c 16 -r 4 -w -v 2>&1 | cut -c1-1000 | tail -5
