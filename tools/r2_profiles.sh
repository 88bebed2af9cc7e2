#!/bin/bash
# final single-GPU evidence set of round 2: bench lines, ncu launch list (+ DRAM bytes), ncu --set full of the compositing
# kernels (+ raw metrics with the L2 RED / atomic counters), GPU test log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r2
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${T}_gpu_tests.log 2>&1; tail -3 gpurun_out/${T}_gpu_tests.log
python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; tail -c 300 gpurun_out/${T}_bench_n1.json
python bench.py --impl reference --steps 5 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err; tail -c 300 gpurun_out/${T}_bench_reference_arm.json
python bench.py --math reference --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_n1_refmath.json 2> /dev/null
python bench.py --math fast --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_n1_fastmath.json 2> /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_${T}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_ -s 6 -c 2 -f -o gpurun_out/prof_render_${T} python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_${T}.log 2>&1
ncu -i gpurun_out/prof_render_${T}.ncu-rep --page details > gpurun_out/ncu_render_details_${T}.txt 2>&1
ncu -i gpurun_out/prof_render_${T}.ncu-rep --page raw --csv > gpurun_out/ncu_render_raw_${T}.csv 2>&1
ls -la gpurun_out/*${T}*
