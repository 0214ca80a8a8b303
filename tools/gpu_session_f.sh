#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== two-tier tests" ; timeout 600 python -m pytest tests/test_gpu_two_tier.py -q -x 2>&1 | tail -30 > gpurun_out/f_two_tier.log; tail -6 gpurun_out/f_two_tier.log
echo "== microbench" ; timeout 300 python tools/tier_microbench.py 2>&1 | tee gpurun_out/f_micro.txt | tail -4
echo "== full gpu suite" ; timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/f_pytest.log; tail -6 gpurun_out/f_pytest.log
echo "== bench (screen on)" ; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; cut -c1-250 gpurun_out/f_bench.json
echo "== ncu launch list" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 1200 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/f_ncu_bench.log 2>&1; wc -l gpurun_out/f_launches.csv
