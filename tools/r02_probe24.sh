#!/bin/bash
run() { echo "== $1"; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 3 --no-secondary --e2e-steps 0 --kp 64 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N=2 kp=64', d['ms_per_step'], d['value'])"; }
one() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --steps 60 --no-secondary --e2e-steps 0 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N=1', d['ms_per_step'], d['value'])"; }
one "depth 2 x 4K (default)" A=1
one "depth 3 x 4K" PB2_LIB_PATH=$PWD/parsec_b200/libvariant_depth3.so
one "depth 4 x 2K" PB2_LIB_PATH=$PWD/parsec_b200/libvariant_d4c2k.so
run "depth 2 x 4K (default)" A=1
run "depth 3 x 4K" PB2_LIB_PATH=$PWD/parsec_b200/libvariant_depth3.so
run "depth 4 x 2K" PB2_LIB_PATH=$PWD/parsec_b200/libvariant_d4c2k.so
