"""stdin: one bench.py JSON line; prints value and the stage times (ms)."""
import json
import sys
j = json.loads(sys.stdin.read())
print(sys.argv[1] if len(sys.argv) > 1 else "", round(j["value"], 1), j["unit"], round(j["ms_per_step"], 4), "ms",
      {k: v["ms"] for k, v in j["stages"].items()})
