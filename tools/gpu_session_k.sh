#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== mask cache tests" ; timeout 600 python -m pytest tests/test_gpu_two_tier.py -q -s -k "mask_cache" 2>&1 | grep -v Warn | tail -6
echo "== full gpu suite" ; timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -5
echo "== bench" ; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; cut -c1-200 gpurun_out/k_bench.json
echo "== bench (mask cache off)" ; DIST_MASK_CACHE=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/k_bench_nomc.json 2> gpurun_out/k_bench_nomc.err; cut -c1-200 gpurun_out/k_bench_nomc.json
echo "== other configs" ; timeout 900 python tools/bench_configs.py > gpurun_out/k_configs.txt 2>&1; head -12 gpurun_out/k_configs.txt; tail -2 gpurun_out/k_configs.txt
