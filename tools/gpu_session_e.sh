#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== ncu launch list of bench.py"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/e_ncu_bench.log 2>&1
wc -l gpurun_out/e_launches.csv
echo "== ncu full capture of the march decoder kernel (dense step of a 512x512 render)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 8 -c 1 -o gpurun_out/e_tc_fwd python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/e_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
