#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== full gpu suite" ; timeout 2400 python -m pytest tests -m gpu -q -s 2>&1 | grep -v Warning | tail -70 > gpurun_out/h_pytest.log; tail -22 gpurun_out/h_pytest.log
echo "== memcheck" ; timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_target.py > gpurun_out/h_memcheck.txt 2>&1; tail -6 gpurun_out/h_memcheck.txt
