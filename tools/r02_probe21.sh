#!/bin/bash
B=oracle/_ref/bin/batch_b200
PARSEC_MCA_device_b200_enabled=1 timeout 60 $B -M 96 -c 8 2>&1 | tail -3 | cut -c1-600
PARSEC_MCA_device_cuda_enabled=1 timeout 60 $B -M 96 -c 8 2>&1 | tail -2 | cut -c1-600
timeout 900 python -m pytest tests/test_mca_component.py -x -q -m gpu 2>&1 | tail -5
