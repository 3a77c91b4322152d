#!/bin/bash
B=oracle/_ref/bin/ex05_b200
for w in 8 32 128; do
echo "== real c=16 window=${w}MiB"; PARSEC_MCA_device_b200_stage_window=$((w*1048576)) PARSEC_B200_PROFILE=1 PARSEC_MCA_device_b200_enabled=1 timeout 120 $B -K 4096 -t 65536 -m gpu -c 16 -r 5 2>&1 | cut -c1-330 | tail -2
done
echo "== real c=8 window=32MiB"; PARSEC_MCA_device_b200_enabled=1 timeout 120 $B -K 4096 -t 65536 -m gpu -c 8 -r 5 2>&1 | cut -c1-330 | tail -1
echo "== real c=32 window=32MiB"; PARSEC_MCA_device_b200_enabled=1 timeout 120 $B -K 4096 -t 65536 -m gpu -c 32 -r 5 2>&1 | cut -c1-330 | tail -1
timeout 600 python -m pytest tests/test_mca_component.py -x -q -m gpu 2>&1 | tail -2
