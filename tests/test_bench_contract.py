"""bench.py contract checks that need no GPU: the `--impl reference` arm (the CPU CG+AMG port timed on
the host cores) prints one JSON line with the keys the driver reads, and ranks other than 0 stay silent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ["--impl", "reference", "--rows", "90", "--cols", "90", "--pairs", "4", "--steps", "1", "--warmup", "0",
        "--cpu-sample", "2", "--ref-budget-s", "60"]


def _run(extra_env=None):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + ARGS, capture_output=True, text=True,
                          env=env, cwd=ROOT, timeout=600)


def test_reference_arm_prints_the_contract_line():
    p = _run()
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    assert j["impl"] == "reference"
    assert j["metric"] == "pair_solves_per_sec" and j["unit"] == "pair-solves/s"
    assert j["higher_is_better"] is True and j["n_gpus"] == 1 and j["vs_baseline"] is None
    assert j["dtype"] == "f64" and j["data"] == "synthetic"
    for k in ("value", "steps", "warmup", "ms_per_step", "scaling"):
        assert k in j, k
    assert j["value"] > 0 and j["ms_per_step"] > 0
    assert "workload" in j["config"] and "90x90" in j["config"]["workload"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == j["value"]
    e = j["e2e"]
    assert e["value"] == j["value"] and e["unit"] == j["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    p = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout.strip() == ""
