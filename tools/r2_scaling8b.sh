#!/bin/bash
# second 8-GPU session: concurrent PCIe floor at 1 / 2 / 4 / 8 GPUs, then the headline config at N = 8 (+ reduce-scatter figure)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2n}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python tools/pcie_probe.py > gpurun_out/${TAG}_pcie_n1.log 2>&1
for n in 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port 2957$n tools/pcie_probe.py > gpurun_out/${TAG}_pcie_n$n.log 2>&1
done
grep -h "H2D\|D2H\|both" gpurun_out/${TAG}_pcie_n*.log
for n in 2 8; do
timeout 600 $TR --nproc-per-node $n --master-port 2958$n bench.py --gpus $n --steps 30 > gpurun_out/${TAG}_bench_C2_n$n.json 2> gpurun_out/${TAG}_bench_C2_n$n.err
python - <<PY
import json
j=json.load(open("gpurun_out/${TAG}_bench_C2_n$n.json"))
print("N=$n", round(j["value"],1), round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"],1), "scatter-only", j["config"]["value_with_reduce_scatter_only"], j["config"]["reduce_scatter_slice_vs_allreduce_table"], "plain", j["config"]["value_through_the_plain_path"], j["parity_check"]["max_rel"])
PY
done
