#!/bin/bash
# One GPU session (gpurun -- 'bash tools/gpu_checks.sh'): the GPU test suite, the bench line, the other BASELINE
# configurations, the tier microbenchmark and the ncu launch list of the bench -- everything lands in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== gpu suite" ; timeout 2400 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|Error" | tail -8
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench.json
echo "== other configs" ; timeout 900 python tools/bench_configs.py > gpurun_out/configs.txt 2>&1; tail -22 gpurun_out/configs.txt
echo "== tier microbench" ; timeout 300 python tools/tier_microbench.py 2>&1 | tee gpurun_out/micro.txt | tail -3
echo "== ncu launch list" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 1300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/launches.csv
