#!/bin/bash
# round-2 GPU session B: trimmed strict kernels — A/B, GPU tests, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
GSRAST_LIB=$PWD/variants/ab/libgsrast.so timeout 900 python tools/math_ab.py C1d C5 C2 > gpurun_out/r2b_math_ab.log 2>&1
echo "math_ab rc=$?" >> gpurun_out/r2b_math_ab.log
cp gpurun_out/math_ab.jsonl gpurun_out/r2b_math_ab.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2b_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_gpu_tests.log
timeout 600 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -5 gpurun_out/r2b_gpu_tests.log
grep -v "^exp probe" gpurun_out/r2b_math_ab.log | cut -c1-400
cut -c1-1500 gpurun_out/r2b_bench.json
