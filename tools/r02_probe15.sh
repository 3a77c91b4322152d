#!/bin/bash
B=oracle/_ref/bin/ex05_b200
mkdir -p gpurun_out
echo "== e2e profile, -c 16"
PARSEC_B200_PROFILE=1 PARSEC_MCA_device_b200_enabled=1 timeout 120 $B -K 4096 -t 65536 -m gpu -c 16 -r 6 -v 2>&1 | grep -E "repeat|Mcycles" | cut -c1-300
echo "== -c 32"
PARSEC_MCA_device_b200_enabled=1 timeout 120 $B -K 4096 -t 65536 -m gpu -c 32 -r 6 -v 2>&1 | grep -E "repeat" | cut -c1-300
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_p15_bench.json 2> gpurun_out/r02_p15_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_p15_bench.err; cat gpurun_out/r02_p15_bench.json | cut -c1-6000
