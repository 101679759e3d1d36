"""bench.py pieces that can be checked without a GPU: the host copies of the synthetic generators agree with the
oracle's (and therefore with the device generator the GPU tests pin), and the `--impl reference` arm prints one JSON
line with the contract's keys."""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np

import oracle as O
from conftest import ROOT


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_bench_generators_match_oracle_bitwise():
    b = load_bench()
    assert np.array_equal(b.synth_rows_np(0x5EED0004, 12345, 50), O.synth_rows(0x5EED0004, 12345, 50, 768))
    q1, p1 = b.planted_queries_np(0x0BADC0DE, 9, 0x5EED0002, 1_000_000)
    q2, p2 = O.planted_queries(0x0BADC0DE, 9, 0x5EED0002, 1_000_000, 768)
    assert np.array_equal(p1, p2) and np.array_equal(q1, q2)


def test_cpu_sample_rows_is_bounded_and_never_exceeds_the_corpus():
    b = load_bench()
    assert b.cpu_sample_rows(154, 1) == 154
    assert b.cpu_sample_rows(1_000_000, 64) == 1_000_000
    assert b.cpu_sample_rows(100_000_000, 64) == 2_000_000
    assert 50_000 <= b.cpu_sample_rows(10_000_000, 4096) <= 2_000_000


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "20000",
                          "--batch", "4", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the default workload is config 4 at every N (one workload across the scaling curve), steps / warm-up as asked for
    assert "config4" in line["config"]["workload"] and line["steps"] == 2
    assert set(line["config"]) == {"workload", "batch", "k", "rows", "store_dtype", "space"}


def test_both_arms_print_the_same_config_block():
    """The driver compares the `config` of the two arms: both come from one function."""
    b = load_bench()
    import argparse
    args = argparse.Namespace(batch=64, space="cosine")
    cfg = b.shared_config(args, b.WORKLOADS["c4"])
    assert cfg["rows"] == 100_000_000 and cfg["batch"] == 64 and cfg["store_dtype"] == "bf16"
