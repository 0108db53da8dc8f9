"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm (`--impl reference`) times
the unmodified reference installed under baseline/_ref (or, when that directory is absent, the pinned CPU port) on the
host cores and prints ONE JSON line with the contract's keys; the product arm refuses to run without CUDA (no CPU
fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize("installed", [True, False])
def test_reference_arm_prints_the_contract_line(installed):
    have = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "fastgen"))
    if installed and not have:
        pytest.skip("baseline/_ref not installed here (python -c 'import __graft_entry__ as g; g.build()')")
    env = dict(os.environ)
    env.pop("FASTGEN_B200_REF_DIR", None)
    if not installed:
        env["FASTGEN_B200_REF_DIR"] = "/nonexistent"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-sample", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["unit"] == "samples/s" and d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert "workload" in d["config"] and "ImageNet-64" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == ("reference" if installed else "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        return   # on a GPU box the product arm is exercised by the bench itself
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert not any(ln.strip().startswith("{") for ln in r.stdout.splitlines())
