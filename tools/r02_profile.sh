#!/bin/bash
# r02 final profiles: (1) ncu launch list of the bench command (engine + GEMM secondary in-process), (2) ncu --set full of one
# pb2_engine_hbm_kernel launch, (3) ncu --set full of one pb2_engine_gemm2_kernel launch
mkdir -p gpurun_out
echo "== ncu launch list"
timeout 900 ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_bench_launches.csv \
   python bench.py --steps 20 --warmup 3 --e2e-steps 0 > gpurun_out/r02f_bench_under_ncu.log 2>&1
tail -c 400 gpurun_out/r02f_bench_under_ncu.log; wc -l gpurun_out/r02f_bench_launches.csv
echo "== ncu full hbm"
timeout 900 ncu --target-processes application-only --set full --clock-control none --import-source on -k regex:pb2_engine_hbm_kernel -s 3 -c 1 -f -o gpurun_out/r02f_ex05_full \
   python bench.py --steps 3 --warmup 3 --no-secondary --e2e-steps 0 > gpurun_out/r02f_full_under_ncu.log 2>&1
tail -c 200 gpurun_out/r02f_full_under_ncu.log; ls -la gpurun_out/r02f_ex05_full.ncu-rep
echo "== ncu full gemm2"
timeout 900 ncu --target-processes application-only --set full --clock-control none --import-source on -k regex:pb2_engine_gemm2_kernel -s 1 -c 1 -f -o gpurun_out/r02f_gemm2_full \
   python tools/prof_gemm.py > gpurun_out/r02f_gemm_under_ncu.log 2>&1
tail -c 200 gpurun_out/r02f_gemm_under_ncu.log; ls -la gpurun_out/r02f_gemm2_full.ncu-rep
