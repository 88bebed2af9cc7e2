#!/bin/bash
# round-2 GPU session A: arithmetic-policy A/B, then the GPU test-suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python tools/math_ab.py C1d C5 C2 > gpurun_out/r2a_math_ab.log 2>&1
echo "math_ab rc=$?" >> gpurun_out/r2a_math_ab.log
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_gpu_tests.log
tail -5 gpurun_out/r2a_gpu_tests.log
tail -30 gpurun_out/r2a_math_ab.log
