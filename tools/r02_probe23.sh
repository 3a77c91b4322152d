#!/bin/bash
run() { echo "== $1"; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 3 --no-secondary --e2e-steps 0 --kp 64 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['achieved'])"; }
run "acquire.sys poll (default)" A=1
run "gpu-scope poll (experiment)" PB2_LIB_PATH=$PWD/parsec_b200/libvariant_gpupoll.so
