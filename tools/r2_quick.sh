#!/bin/bash
# quick GPU check: parity tests + device-resident bench line (TAG = output prefix)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-q}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e ${@:2} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_gpu_tests.log
python - <<PY
import json
j=json.load(open("gpurun_out/${TAG}_bench.json"))
print(round(j["value"],1), "steps/s", round(j["ms_per_step"],4), "ms", {k:v["ms"] for k,v in j["stages"].items()})
PY
