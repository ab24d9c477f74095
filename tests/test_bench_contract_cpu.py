"""CPU: the measurement contract of bench.py that can be checked without a GPU - the reference arm (`--impl
reference`: the C restatement of the reference on the host cores) prints ONE JSON line with the contract's keys, and
the product arm fails loudly (no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT, env=env)


def test_reference_arm_json_line():
    out = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--log2n", "10")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                  # exactly one line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "proofs/s" and d["higher_is_better"] is True
    assert d["metric"] == "bandersnatch_thin_vrf_batch_verified_proofs_per_sec"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["config"]["batch"] == 1 << 10 and "2^10" in d["config"]["workload"]     # the label is what was run


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    out = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline", timeout=300)
    assert out.returncode != 0                              # fails loudly, prints no result line
    assert not any(l.strip().startswith("{") and "\"value\"" in l for l in out.stdout.splitlines())
