#!/bin/bash
# r02: (1) where the e2e repeats lose time, (2) ncu launch list of bench.py, (3) ncu --set full of the Ex05 window kernel
B=oracle/_ref/bin/ex05_b200
mkdir -p gpurun_out
for c in 16; do
echo "== e2e timing, -c $c"
PARSEC_MCA_device_b200_enabled=1 timeout 120 $B -K 4096 -t 65536 -m gpu -c $c -r 6 -v 2>&1 | grep -E "  dev |repeat" | cut -c1-200
done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv \
   python bench.py --steps 20 --warmup 3 --no-secondary --e2e-steps 0 > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -c 600 gpurun_out/r02_bench_under_ncu.log; wc -l gpurun_out/r02_bench_launches.csv
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pb2_engine_hbm_kernel -s 3 -c 1 -f -o gpurun_out/r02_ex05_full \
   python bench.py --steps 3 --warmup 3 --no-secondary --e2e-steps 0 > gpurun_out/r02_full_under_ncu.log 2>&1
tail -c 300 gpurun_out/r02_full_under_ncu.log; ls -la gpurun_out/r02_ex05_full.ncu-rep
