#!/bin/bash
B=oracle/_ref/bin/ex05_b200
PARSEC_B200_PROFILE=1 PARSEC_MCA_device_b200_enabled=1 timeout 120 $B -K 4096 -t 65536 -m gpu -c 16 -r 5 -v 2>&1 | grep -E "manager Mcycles|repeat|times_s" | cut -c1-260
echo "== no flush between repeats is not possible; try sleeping workers: c=4"
PARSEC_B200_PROFILE=1 PARSEC_MCA_device_b200_enabled=1 timeout 120 $B -K 4096 -t 65536 -m gpu -c 4 -r 5 -v 2>&1 | grep -E "repeat" | cut -c1-260
