#!/bin/bash
# bench.py on every BASELINE configuration, one GPU (TAG = output prefix)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2i}
for c in C2 C5 C3 C4; do
  timeout 900 python bench.py --config $c --steps 20 > gpurun_out/${TAG}_bench_${c}_n1.json 2> gpurun_out/${TAG}_bench_${c}_n1.err
  echo "$c rc=$?"; tail -c 300 gpurun_out/${TAG}_bench_${c}_n1.err
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/${TAG}_bench_${c}_n1.json"))
    print("$c", round(j["value"],2), j["unit"], round(j["ms_per_step"],3), "ms; e2e", j["e2e"] and round(j["e2e"]["value"],2), "cpu", j["cpu_baseline"] and round(j["cpu_baseline"]["value"],3), "parity", j.get("parity_check") and j["parity_check"]["max_rel"], "plain", j["config"]["value_through_the_plain_path"])
    print("   ", {k:v["ms"] for k,v in j["stages"].items()})
except Exception as e: print("$c failed", e)
PY
done
