#!/bin/bash
# ncu --set full of the two compositing kernels (default strict math), C2 step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2c}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_ -s 6 -c 2 -f -o gpurun_out/prof_render_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1
ncu -i gpurun_out/prof_render_$TAG.ncu-rep --page details > gpurun_out/ncu_render_details_$TAG.txt 2>&1
ls -la gpurun_out/prof_render_$TAG.ncu-rep
grep -E "Duration|Executed Ipc Active|Issue Slots Busy|Registers Per|Achieved Occupancy|L1/TEX Cache Throughput|Executed Instructions  " gpurun_out/ncu_render_details_$TAG.txt
