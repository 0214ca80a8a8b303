#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== tests" ; timeout 900 python -m pytest tests/test_gpu_two_tier.py tests/test_gpu_baseline_sizes.py -q -s 2>&1 | grep -v Warning | tail -120 > gpurun_out/b_pytest.log; tail -8 gpurun_out/b_pytest.log
echo "== microbench" ; timeout 300 python tools/tier_microbench.py 2>&1 | tee gpurun_out/b_micro.txt | tail -5
