"""bench.py's driver contract, as far as it can be exercised without a GPU: the reference arm (CPU oracle port) prints
ONE JSON line with the keys the driver reads, and the GPU arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--workload", "tiny", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "vertices/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("mesh vertices/sec") and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "vertices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["data"] == "synthetic"


def test_result_line_is_alone_on_stdout_even_if_a_library_writes_to_fd_1():
    """NCCL prints its version banner to fd 1 on the GPU boxes; bench.py keeps the real stdout for the result line."""
    code = ("import os, sys; sys.argv = ['bench.py', '--impl', 'reference', '--workload', 'tiny', '--steps', '1', "
            "'--warmup', '1']; import bench; real = bench.run_reference; "
            "bench.run_reference = lambda *a: (os.write(1, b'NCCL version 0.0.0\\n'), real(*a)); bench.main()")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1 and json.loads(lines[0])["impl"] == "reference"
    assert "NCCL version 0.0.0" in r.stderr


def test_reference_arm_non_zero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--workload", "tiny", "--gpus", "2"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_gpu_arm_fails_loudly_without_a_device():
    r = subprocess.run([sys.executable, BENCH, "--workload", "tiny", "--steps", "1"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)


def test_traffic_file_is_well_formed():
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
        t = json.load(f)
    for wl, fams in t.items():
        for fam, ent in fams.items():
            assert ent["dram_bytes_per_launch"] > 0 and isinstance(ent["note"], str)


def test_path_roofline_sums_the_per_kernel_bounds():
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    summ = {"linear_fwd[8x8]": {"calls": 2, "ms": 4.0, "bytes": 1e9, "flops": 2e12},       # tensor bound: 2e12 / 1e15 = 2 ms
            "edge_message_fwd[H8]": {"calls": 2, "ms": 1.0, "bytes": 5e9, "flops": 1e9}}    # HBM bound: 5e9 / 5e12 = 1 ms
    r = bench.path_roofline(summ, 2, hbm_gbs=5000.0, bf16_tflops=1000.0, passes=3, rate=0.5)
    assert abs(r["roofline_ms_per_step"] - (2.0 + 1.0) / 2) < 1e-9
    assert abs(r["roofline_ms_per_step_mode_ceiling"] - (12.0 + 1.0) / 2) < 1e-9
    assert abs(r["measured_ms_per_step_eager"] - 2.5) < 1e-9
