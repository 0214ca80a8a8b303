#!/bin/bash
mkdir -p gpurun_out
echo "== mesh tests"; timeout 900 python -m pytest tests/test_gpu_mesh.py -m gpu -q -x 2>&1 | tail -15
echo "== memcheck (small cases)"; timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_mesh.py -m gpu -q -k "bit_exact or nearest or surface_sample or outside_range" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" | head -10 | tee gpurun_out/mesh_memcheck.txt
echo "== bench"; timeout 600 python tools/bench_mesh.py 2>&1 | tee gpurun_out/mesh_bench.txt | tail -12
