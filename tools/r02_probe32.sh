#!/bin/bash
timeout 900 python -m pytest tests/test_stream_gpu.py tests/test_mca_component.py tests/test_golden.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do PARSEC_MCA_device_b200_enabled=1 PARSEC_MCA_device_b200_memory_number_of_blocks=1024 timeout 60 oracle/_ref/bin/ex02_b200 -m gpu -N 999 -c 2 -r 5 2>&1 | tail -1 | cut -c150-330; done
PARSEC_MCA_device_b200_enabled=1 timeout 120 oracle/_ref/bin/ex05_b200 -K 4096 -t 65536 -m gpu -c 32 -r 6 2>&1 | tail -1 | cut -c100-330
