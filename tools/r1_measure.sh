set -x
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r1_gpu_tests.log 2>&1; tail -4 gpurun_out/r1_gpu_tests.log
python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err; tail -c 600 gpurun_out/r1_bench_n1.json
python bench.py --math reference --no-cpu-baseline --no-e2e > gpurun_out/r1_bench_n1_refmath.json 2> /dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference_arm.json 2> gpurun_out/r1_bench_reference_arm.err; tail -c 400 gpurun_out/r1_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 39 -c 26 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_ -s 6 -c 2 -o gpurun_out/prof_render_r1_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*final*
