"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm (`--impl reference`) prints
exactly one JSON line on stdout with the agreed keys, and non-zero ranks of a multi-rank launch print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = run()
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mtris/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["config"]["workload"].startswith("C3/M1")
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and 1 <= d["cpu_baseline"]["cores"] <= len(os.sched_getaffinity(0)) and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mtris/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    assert run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}).strip() == ""
