"""bench.py's reference arm (`--impl reference`) on CPU: it must run without a GPU, print ONE JSON line with the contract's keys,
and under torchrun only rank 0 may print. (The repo arm needs a B200 and is exercised by the driver.)"""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(env_extra=None, args=()):
    env = dict(os.environ)
    env.update(env_extra or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--config", "1", *args],
                       capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    return [line for line in p.stdout.splitlines() if line.startswith("{")]


def test_reference_arm_prints_the_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MS/s" and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["metric"] == "IQ MSamples/s through FFT+power+detect" and d["value"] > 0 and d["steps"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["single_thread"] > 0
    assert abs(sum(cb["stage_split_single_thread"].values()) - 1.0) < 1e-6
    assert d["e2e"] == {"value": d["value"], "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["fft_size"] == 4096 and "configs[0]" in d["config"]["workload"]


def test_reference_arm_is_silent_on_other_ranks():
    assert _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, ("--gpus", "2")) == []
