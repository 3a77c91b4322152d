#!/bin/bash
# final HBM kernel (TMA ring depth 3): ncu launch list of the bench command (no secondary records) and --set full of one launch
mkdir -p gpurun_out
timeout 300 ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02g_bench_launches.csv \
   python bench.py --steps 20 --warmup 3 --no-secondary --e2e-steps 0 > gpurun_out/r02g_bench_under_ncu.log 2>&1
wc -l gpurun_out/r02g_bench_launches.csv
timeout 300 ncu --target-processes application-only --set full --clock-control none --import-source on -k regex:pb2_engine_hbm_kernel -s 3 -c 1 -f -o gpurun_out/r02g_ex05_full \
   python bench.py --steps 3 --warmup 3 --no-secondary --e2e-steps 0 > gpurun_out/r02g_full_under_ncu.log 2>&1
ls -la gpurun_out/r02g_ex05_full.ncu-rep
