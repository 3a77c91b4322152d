#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_p18_bench.json 2> gpurun_out/r02_p18_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_p18_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_p18_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}); print("e2e", d.get("e2e")); print("cpu", d.get("cpu_baseline")); print("roofline", {k:d["roofline"][k] for k in ("frac","dram_frac","kernel_ms")})
print("sec", json.dumps(d.get("secondary"))[:1500])
PY
