#!/bin/bash
timeout 900 python -m pytest tests/test_mca_component.py -x -q -m gpu 2>&1 | tail -5
PARSEC_MCA_device_b200_trace=/tmp/tr PARSEC_MCA_device_b200_enabled=1 timeout 60 oracle/_ref/bin/ex05_b200 -K 16 -t 65536 -m gpu -c 4 > /dev/null 2>&1; head -c 700 /tmp/tr.*.json
