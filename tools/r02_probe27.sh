#!/bin/bash
N=${1:-2}
timeout 600 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -2
run() { echo "== $1"; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 3 --no-secondary --e2e-steps 0 $KP 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N=$N $KP', d['ms_per_step'], d['value'], d['roofline']['achieved'])"; }
KP="--kp 64"
run "two rings, cap 384" A=1
run "two rings, cap 128" PB2_PULL_CAP=128
run "two rings, cap 1024" PB2_PULL_CAP=1024
run "one FIFO" PB2_PULL_CAP=0
