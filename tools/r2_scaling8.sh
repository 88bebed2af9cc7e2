#!/bin/bash
# 8-GPU session: weak-scaling series of the headline config, C3 view-batch strong scaling, C4 / C5, oracle parity of the
# fused exchange at 8 ranks, NVLink byte counters around one run.   TAG = output prefix
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2k}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
# oracle parity of the fused exchange, 8 ranks (one and two views per rank), odd N
timeout 600 $TR --nproc-per-node 8 --master-port 29531 tools/peers_check.py > gpurun_out/${TAG}_peers_w8.log 2>&1; tail -1 gpurun_out/${TAG}_peers_w8.log
timeout 600 $TR --nproc-per-node 8 --master-port 29532 tools/peers_check.py odd > gpurun_out/${TAG}_peers_w8_odd.log 2>&1; tail -1 gpurun_out/${TAG}_peers_w8_odd.log
timeout 600 $TR --nproc-per-node 4 --master-port 29533 tools/peers_check.py > gpurun_out/${TAG}_peers_w4.log 2>&1; tail -1 gpurun_out/${TAG}_peers_w4.log
# headline config, N = 1, 2, 4, 8
timeout 600 python bench.py --steps 30 > gpurun_out/${TAG}_bench_C2_n1.json 2> gpurun_out/${TAG}_bench_C2_n1.err
for n in 2 4; do
  timeout 600 $TR --nproc-per-node $n --master-port 2954$n bench.py --gpus $n --steps 30 > gpurun_out/${TAG}_bench_C2_n$n.json 2> gpurun_out/${TAG}_bench_C2_n$n.err
done
nvidia-smi nvlink -gt d -i 0 > gpurun_out/${TAG}_nvlink_before.txt 2>&1
timeout 600 $TR --nproc-per-node 8 --master-port 29548 bench.py --gpus 8 --steps 30 --no-e2e --no-parity-check > gpurun_out/${TAG}_bench_C2_n8_noe2e.json 2> gpurun_out/${TAG}_bench_C2_n8_noe2e.err
nvidia-smi nvlink -gt d -i 0 > gpurun_out/${TAG}_nvlink_after.txt 2>&1
timeout 600 $TR --nproc-per-node 8 --master-port 29549 bench.py --gpus 8 --steps 30 > gpurun_out/${TAG}_bench_C2_n8.json 2> gpurun_out/${TAG}_bench_C2_n8.err
# C3: 8 views per step over 2 / 4 / 8 ranks (N=1 is in the single-GPU set)
for n in 2 4 8; do
  timeout 600 $TR --nproc-per-node $n --master-port 2955$n bench.py --gpus $n --config C3 --steps 20 > gpurun_out/${TAG}_bench_C3_n$n.json 2> gpurun_out/${TAG}_bench_C3_n$n.err
done
timeout 600 python bench.py --config C3 --steps 20 --no-cpu-baseline > gpurun_out/${TAG}_bench_C3_n1.json 2> gpurun_out/${TAG}_bench_C3_n1.err
# C4 / C5 at 8 ranks
timeout 600 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --config C4 --steps 20 > gpurun_out/${TAG}_bench_C4_n8.json 2> gpurun_out/${TAG}_bench_C4_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port 29562 bench.py --gpus 8 --config C5 --steps 30 > gpurun_out/${TAG}_bench_C5_n8.json 2> gpurun_out/${TAG}_bench_C5_n8.err
# reference arm under torchrun (must use all host cores)
timeout 600 $TR --nproc-per-node 8 --master-port 29563 bench.py --impl reference --gpus 8 --steps 3 > gpurun_out/${TAG}_bench_reference_n8.json 2> gpurun_out/${TAG}_bench_reference_n8.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        j = json.load(open(f))
        e = j.get("e2e") or {}
        pc = j.get("parity_check") or {}
        print(f.split("/")[-1], round(j["value"], 2), j["unit"], round(j["ms_per_step"], 3), "ms | e2e", e.get("value") and round(e["value"], 1),
              "| parity", pc.get("max_rel"), pc.get("ok"), "| plain", j.get("config", {}).get("value_through_the_plain_path"), "| cores", (j.get("cpu_baseline") or {}).get("cores"))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
