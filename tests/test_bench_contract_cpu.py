"""bench.py's reference arm (the CPU restatement timed on the host cores) runs without a GPU: the JSON line it prints must
carry the contract keys, use every host core even under a launcher's OMP_NUM_THREADS=1, and report the number of steps it
actually timed (the arm used to extrapolate from a band of tile rows and ran single-threaded under torchrun)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C5", "--steps", "1",
                        "--warmup", "1", "--gpus", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    j = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in j, k
    assert j["impl"] == "reference" and j["dtype"] == "f32" and j["vs_baseline"] is None
    assert j["steps"] == 1 and j["value"] > 0 and abs(j["ms_per_step"] * j["value"] - 1e3) < 1e-3 * 1e3
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1), "all host cores, whatever OMP_NUM_THREADS says"
    assert cb["value"] == j["value"] and "full forward+backward" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and j["config"]["workload"].startswith("C5")


def test_non_zero_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="3", WORLD_SIZE="8")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "8"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""
