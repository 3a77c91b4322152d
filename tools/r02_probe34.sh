#!/bin/bash
B=oracle/_ref/bin/ex05_b200
for i in 1 2; do PARSEC_MCA_device_b200_enabled=2 PARSEC_MCA_device_b200_memory_number_of_blocks=64 timeout 120 $B -K 2048 -t 65536 -m gpu -c 8 -r 1 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('small heaps:', d['errors'], d['executed_on_gpu'], d['best_s'], {k:d['b200'][k] for k in ('evictions','forwarded','peer_pulls','check_mismatches')})"; done
timeout 600 python -m pytest tests/test_mca_component.py -x -q -m gpu 2>&1 | tail -3
