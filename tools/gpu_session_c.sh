#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== two-tier tests" ; timeout 600 python -m pytest tests/test_gpu_two_tier.py -q -x 2>&1 | tail -30 > gpurun_out/c_two_tier.log; tail -4 gpurun_out/c_two_tier.log
echo "== full gpu suite" ; timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/c_pytest.log; tail -8 gpurun_out/c_pytest.log
echo "== microbench" ; timeout 300 python tools/tier_microbench.py 2>&1 | tee gpurun_out/c_micro.txt | tail -4
echo "== bench (screen on)" ; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; cut -c1-250 gpurun_out/c_bench.json
echo "== other configs" ; timeout 900 python tools/bench_configs.py > gpurun_out/c_configs.txt 2>&1; head -14 gpurun_out/c_configs.txt
echo "== timeline" ; DIST_EXTRA_NVCC_FLAGS=-DDIST_TC_TIMELINE python dist-renderer_b200/build.py --force > /dev/null 2>&1; DIST_TC_DEBUG=4 timeout 300 python tools/tc_timeline.py 2> gpurun_out/c_timeline2.txt; tail -18 gpurun_out/c_timeline2.txt
