#!/bin/bash
B=oracle/_ref/bin/ex05_b200
for c in 16 8; do
echo "== e2e timing, -c $c"
PARSEC_MCA_device_b200_enabled=1 timeout 120 $B -K 4096 -t 65536 -m gpu -c $c -r 6 -v 2>&1 | grep -E "  dev |repeat" | cut -c1-260
done
