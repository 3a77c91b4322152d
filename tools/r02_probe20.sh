#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 100 --no-secondary > gpurun_out/r02_p20_bench.json 2> gpurun_out/r02_p20_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_p20_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_p20_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}); print("e2e", d.get("e2e")); print("e2e_standalone", d.get("e2e_standalone")); print("cpu", d.get("cpu_baseline"))
PY
python tools/pcie_probe.py 2>&1 | tail -8
