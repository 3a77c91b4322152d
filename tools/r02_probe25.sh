#!/bin/bash
timeout 600 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -4
run() { echo "== $1"; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 3 --no-secondary --e2e-steps 0 --kp 64 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N=2 kp=64', d['ms_per_step'], d['value'], d['roofline']['achieved'])"; }
run "push" A=1
run "pull" PB2_MGPU_PUSH=0
