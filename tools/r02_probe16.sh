#!/bin/bash
B=oracle/_ref/bin/ex05_b200
mkdir -p gpurun_out
echo "== memory pressure"
PARSEC_MCA_device_b200_enabled=1 PARSEC_MCA_device_b200_memory_number_of_blocks=96 PARSEC_B200_DEBUG=1 timeout 60 $B -K 256 -t 65536 -m gpu -c 8 2>&1 | tail -3 | cut -c1-400
for c in 16 32; do
echo "== e2e profile, -c $c"
PARSEC_B200_PROFILE=1 PARSEC_MCA_device_b200_enabled=1 timeout 120 $B -K 4096 -t 65536 -m gpu -c $c -r 6 -v 2>&1 | grep -E "repeat [2345]|Mcycles" | tail -7 | cut -c1-300
done
timeout 600 python -m pytest tests/test_mca_component.py -x -q -m gpu 2>&1 | tail -5
