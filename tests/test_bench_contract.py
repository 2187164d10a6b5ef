"""CPU: the bench.py contract that can be checked without a GPU -- the reference arm's JSON line, the product arm failing loudly
when there is no CUDA device (no CPU fallback), the workload table against the algorithmic bytes of SURVEY.md section 8(d), and
that both arms describe the configuration with the same dict."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def _bench_module():
    sys.path.insert(0, ROOT)
    import bench
    return bench


def test_reference_arm_prints_the_contract_line():
    res = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "1"], capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    d = json.loads(res.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "sequences/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["metric"] == "FastGRNN sequences/sec (fwd infer)" and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None and d["scaling"] == "weak"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0                       # the reference arm never touches the GPU
    bench = _bench_module()
    assert d["config"] == bench.config_of(bench.WORKLOADS["c2"], "c2", 1)       # the very dict the product arm prints


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the product arm runs")
    res = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert res.returncode != 0
    assert "no CUDA device" in (res.stdout + res.stderr) and "no CPU fallback" in (res.stdout + res.stderr)
    assert not res.stdout.strip().startswith("{")        # no metric line is printed


def test_workload_table_matches_the_survey_bytes():
    bench = _bench_module()
    w = bench.WORKLOADS
    assert bench.algorithmic_bytes_per_seq(w["c2"]) == 63360            # 4*T*I + 4*T*H, SURVEY 8(d)
    assert bench.algorithmic_bytes_per_seq(w["c4"]) == 114048
    assert bench.algorithmic_bytes_per_seq(w["c5"]) == 576000           # bf16 x, fp32 states
    assert bench.algorithmic_bytes_per_seq(w["c3"]) == 177408           # fwd + grad_h + h + x
    assert bench.flops_per_seq(w["c2"]) == 4055040 and bench.flops_per_seq(w["c4"]) == 4156416 and bench.flops_per_seq(w["c5"]) == 40960000
    assert (w["c2"]["B"], w["c3"]["B"], w["c4"]["B"], w["c5"]["B"]) == (8192, 2048, 32768, 8192)
    for name, wl in w.items():
        c = bench.config_of(wl, name, 4)
        assert c["global_batch"] == 4 * wl["B"] and c["workload"] == wl["desc"] and "model" not in c
        assert ("gradient all-reduce" in c["parallelism"]) == (wl["mode"] == "train")
