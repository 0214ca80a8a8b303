#!/bin/bash
# Round-2 GPU session A: new parity tests, two-tier precision first light, bench, far-field statistics.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/a_gpu.txt 2>&1
echo "== two-tier unit tests" ; timeout 600 python -m pytest tests/test_gpu_two_tier.py -q -x 2>&1 | tail -40 > gpurun_out/a_two_tier.log; tail -5 gpurun_out/a_two_tier.log
echo "== full gpu suite" ; timeout 2400 python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > gpurun_out/a_pytest.log; tail -15 gpurun_out/a_pytest.log
echo "== far field stats" ; timeout 600 python tools/far_field_stats.py --out gpurun_out/a_far_field.txt > /dev/null 2>gpurun_out/a_far_field.err; tail -3 gpurun_out/a_far_field.txt
echo "== bench (screen on)" ; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; cut -c1-400 gpurun_out/a_bench.json
echo "== bench (screen off)" ; DIST_SCREEN=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/a_bench_noscreen.json 2> gpurun_out/a_bench_noscreen.err; cut -c1-300 gpurun_out/a_bench_noscreen.json
echo "== other configs" ; timeout 900 python tools/bench_configs.py > gpurun_out/a_configs.txt 2>&1; tail -25 gpurun_out/a_configs.txt
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
