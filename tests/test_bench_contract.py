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


@pytest.mark.gpu
def test_product_arm_contract_on_gpu():
    """The line the driver records: one JSON line with the contract's keys, the roofline and cpu_baseline objects, an e2e
    figure measured on the same channel count with host<->device copies, a launch count, and the other configurations
    under `workloads` (each a complete sub-line).  Small sizes: this checks the shape of the line, not its numbers."""
    d = run_bench("--steps", "2", "--warmup", "3", "--channels", "256", "--e2e-steps", "1")
    for k in ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline", "workloads"]:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] >= 3 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["gpu_launches"] > 0
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["channels"] == d["config"]["channels_per_gpu"]            # same channel count as `value`
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert "fused_decim" in r["kernel"] and 0.1 < r["kernel_share_of_step"] <= 1.0      # (small batches leave the step latency bound)
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] > 0
    assert set(d["workloads"]) == {"panadapter", "rxa_usb", "rxa_fm", "channelizer", "pipeline", "tx_filter"}
    for name, w in d["workloads"].items():
        assert w["value"] > 0 and w["gpu_launches"] > 0 and w["e2e"]["value"] > 0, name
        assert w["roofline"]["frac"] > 0 and w["cpu_baseline"]["value"] > 0, name
