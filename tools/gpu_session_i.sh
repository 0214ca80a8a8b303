#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; cut -c1-200 gpurun_out/i_bench.json
echo "== reference arm" ; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/i_bench_ref.json 2> gpurun_out/i_bench_ref.err; cut -c1-700 gpurun_out/i_bench_ref.json
echo "== other configs" ; timeout 900 python tools/bench_configs.py > gpurun_out/i_configs.txt 2>&1; tail -22 gpurun_out/i_configs.txt
echo "== ncu launch list" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 1300 --csv --log-file gpurun_out/i_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/i_ncu_bench.log 2>&1; wc -l gpurun_out/i_launches.csv
