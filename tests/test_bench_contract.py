"""bench.py's reference arm runs on the CPU (it is the oracle's port of the reference's retrieve_one): check, without a
GPU, that it prints exactly one JSON line with the contract's keys, and that ranks other than 0 print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--gallery", "20000", "--queries", "64"], capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "knn_queries_per_s" and d["unit"] == "queries/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    # "reference" where /root/reference is mounted (the unmodified utils.retrieve_one), "port" on the GPU box
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_port_when_the_reference_is_not_mounted():
    lines = _run({"MMSIM_BENCH_FORCE_PORT": "1"})
    d = json.loads(lines[0])
    assert d["cpu_baseline"]["kind"] == "port" and "port of utils.retrieve_one" in d["note"]


def test_reference_arm_other_ranks_stay_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}) == []
