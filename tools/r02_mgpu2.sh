#!/bin/bash
mkdir -p gpurun_out
B=oracle/_ref/bin/ex05_b200
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k: d[k] for k in ('best_s','mean_s','errors','executed_on_gpu')}, {k: d['b200'][k] for k in ('check_mismatches','peer_pulls','peer_detours','bytes_h2d_kernel')})"; }
echo "== 2 devices"; PARSEC_MCA_device_b200_enabled=2 timeout 60 $B -K 4096 -t 65536 -m gpu -c 16 -r 4 2>&1 | tail -1 | show
echo "== 2 devices small"; PARSEC_MCA_device_b200_enabled=2 timeout 60 $B -K 1024 -t 1024 -m gpu -c 16 -r 2 2>&1 | tail -1 | show
PARSEC_MCA_device_b200_enabled=2 timeout 120 oracle/_ref/bin/stage_b200 -m gpu -c 4 2>&1 | tail -1
timeout 300 python -m pytest tests/test_mca_component.py -x -q -m gpu 2>&1 | tail -2
