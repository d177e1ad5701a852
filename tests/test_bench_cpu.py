"""bench.py's reference arm (the reference node on the host cores) on a tiny sample, and its refusal of the configuration
the node cannot run.  No GPU involved: this is the oracle leg of the measurement."""
import json
import os
import subprocess
import sys

import pytest

from helpers import ROOT


def _run(args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    from oracle.pyref import have_ref
    if not have_ref():
        pytest.skip("oracle/_ref/libimgenv_ref.so not built")
    d = _run(["--impl", "reference", "--workload", "c1", "--steps", "2", "--warmup", "1"])
    assert d["impl"] == "reference" and d["unit"] == "robot-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "robot-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("c1")


def test_reference_arm_refuses_what_the_node_cannot_run():
    d = _run(["--impl", "reference", "--workload", "c5", "--steps", "1", "--warmup", "0"])
    assert d["impl"] == "reference" and "unavailable" in d


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                       timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
