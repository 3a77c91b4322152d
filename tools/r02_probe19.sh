#!/bin/bash
B=oracle/_ref/bin/dtd_b200
for o in "" "-o" "-g"; do
PARSEC_MCA_device_b200_enabled=1 timeout 60 $B -M 16 -n 4096 -N 8 -c 8 $o 2>&1 | tail -4 | cut -c1-600
done
PARSEC_MCA_device_cuda_enabled=1 timeout 60 $B -M 16 -n 4096 -N 8 -c 8 -o 2>&1 | tail -2 | cut -c1-600
timeout 600 python -m pytest tests/test_mca_component.py -x -q -m gpu 2>&1 | tail -5
