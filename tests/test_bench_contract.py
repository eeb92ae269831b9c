"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the CPU restatement timed on the host
cores) prints one JSON line with the keys the driver reads; under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None, args=()):
    env = dict(os.environ, **(env_extra or {}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--qubits", "14", "--steps", "1",
                          "--warmup", "0", *args], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    return out.stdout.strip()


def test_reference_arm_json_line():
    line = _run()
    d = json.loads(line.splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "gates_per_sec" and d["unit"] == "gates/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "brickwork" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["config"]["qubits"] == 14


def test_reference_arm_other_ranks_stay_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, ("--gpus", "2")) == ""
