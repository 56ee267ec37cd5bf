"""The reference arm of bench.py (`--impl reference`) runs on the host cores alone, so its JSON contract can be
checked without a GPU: one line, the keys the driver reads, the compiled reference as the thing timed."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libquisk_rx_ref.so")

KEYS = ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"]


def run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=280)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


@pytest.mark.skipif(not os.path.exists(REF), reason="compiled reference not built (oracle/build_ref.sh)")
@pytest.mark.parametrize("workload", ["rx_chain", "channelizer"])
def test_reference_arm_contract(workload):
    d = run_bench("--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "1")
    for k in KEYS:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "complex MS/s through RX chain" and d["unit"] == "MS/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device the product arm must stop with an error, not print a number."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-cpu-baseline"],
                         capture_output=True, text=True, timeout=280)
    assert out.returncode != 0
    assert not [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
