#!/bin/bash
# quick GPU sanity after a decoder-kernel change: microbench (hang detector), the kernel-level tests, the bench line
mkdir -p gpurun_out
echo "== micro"; timeout 120 python tools/tier_microbench.py 2>&1 | tail -3
echo "== tests (two-tier + parity)"; timeout 900 python -m pytest tests/test_gpu_two_tier.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cut -c1-220 gpurun_out/bench_quick.json
