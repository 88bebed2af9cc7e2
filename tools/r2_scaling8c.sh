#!/bin/bash
# final 8-GPU confirmation with the round's last kernels: oracle parity of the fused exchange at 8 ranks, C2 at N = 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29631 tools/peers_check.py > gpurun_out/r2v_peers_w8.log 2>&1; tail -1 gpurun_out/r2v_peers_w8.log
timeout 600 $TR --nproc-per-node 8 --master-port 29648 bench.py --gpus 8 --steps 30 > gpurun_out/r2v_bench_C2_n8.json 2> gpurun_out/r2v_bench_C2_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port 29649 bench.py --gpus 8 --config C3 --steps 20 --no-e2e > gpurun_out/r2v_bench_C3_n8.json 2> gpurun_out/r2v_bench_C3_n8.err
python - <<PY
import json
for f in ("r2v_bench_C2_n8", "r2v_bench_C3_n8"):
    j = json.load(open("gpurun_out/%s.json" % f))
    print(f, round(j["value"], 1), round(j["ms_per_step"], 3), "e2e", (j.get("e2e") or {}).get("value"), "scatter-only", j["config"]["value_with_reduce_scatter_only"], "plain", j["config"]["value_through_the_plain_path"], "parity", j["parity_check"]["max_rel"], j["parity_check"]["ok"])
PY
