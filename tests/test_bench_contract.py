"""bench.py contract on the CPU side: the reference arm (`--impl reference`, both workloads) prints ONE JSON line with the
keys the driver reads, and the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def run_bench(*args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True,
                       timeout=timeout)
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    return r, lines


def check_reference_line(line, metric):
    d = json.loads(line)
    assert d["impl"] == "reference"
    if "unavailable" in d:          # a box without the vendored reference: one line saying why, exit 0
        assert isinstance(d["unavailable"], str) and d["unavailable"]
        return
    assert BASE_KEYS <= set(d), sorted(BASE_KEYS - set(d))
    assert d["metric"] == metric and d["unit"] == "img/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(cb) and cb["kind"] in ("reference", "port")
    assert cb["value"] == d["value"] and cb["cores"] >= 1
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_train_line():
    r, lines = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(lines) == 1, lines          # exactly one line on stdout
    check_reference_line(lines[0], "train_step_images_per_sec_640x640_bs64_per_gpu")


def test_reference_arm_detect_line():
    r, lines = run_bench("--impl", "reference", "--workload", "detect", "--steps", "1", "--warmup", "0", "--size", "320")
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(lines) == 1, lines
    check_reference_line(lines[0], "detect_images_per_sec_1280x1280_bs128")


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a GPU")
def test_product_arm_fails_loudly_without_a_gpu():
    for args in ((), ("--workload", "detect")):
        r, lines = run_bench(*args, timeout=300)
        assert r.returncode != 0 and not lines
        assert "no CUDA device" in r.stderr
