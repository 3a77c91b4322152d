#!/bin/bash
mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"; grep -v "Warning\|warn\|^\*\|OMP" gpurun_out/r02_bench_n$N.err | tail -8 | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n$N.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}); print("roofline", d.get("roofline")); print("e2e", d.get("e2e")); print("secondary", json.dumps(d.get("secondary"))[:2400])
PY
