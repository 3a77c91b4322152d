#!/bin/bash
B=oracle/_ref/bin/ex05_b200
echo "== accounting"
PARSEC_MCA_device_b200_enabled=1 timeout 60 $B -K 512 -t 65536 -m gpu -c 8 -r 2 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['errors'], d['h2d_bytes'], d['required_in'], d['b200'])"
for dma in 1 0; do
echo "== e2e -c 32 stage_dma=$dma"
PARSEC_MCA_device_b200_stage_dma=$dma PARSEC_B200_PROFILE=1 PARSEC_MCA_device_b200_enabled=1 timeout 120 $B -K 4096 -t 65536 -m gpu -c 32 -r 6 -v 2>&1 | grep -E "repeat [2345]|starter" | tail -5 | cut -c1-330
done
