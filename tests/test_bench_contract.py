"""CPU checks of bench.py's bookkeeping: the SURVEY.md section 8(d) work formulas, the per-workload step
definitions, the peak table, and the reference arm's JSON line (rank > 0 prints nothing)."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_algorithmic_work_matches_survey(bench):
    B, N, k = 32, 1024, 20                                             # config A, SURVEY.md section 8(d)
    assert bench.algorithmic_bytes("edge_fwd_C64", B, N, k) == pytest.approx(349.2e6, rel=1e-3)
    assert bench.algorithmic_bytes("edge_bwd_C128", B, N, k) == pytest.approx(693.1e6, rel=1e-3)
    assert bench.algorithmic_bytes("edge_fwd_C3", B, N, k) == pytest.approx(21.4e6, rel=5e-3)
    assert bench.algorithmic_flops("knn_C64", B, N, k) == pytest.approx(4.40e9, rel=2e-3)
    assert bench.algorithmic_flops("knn_C128", B, N, k) == pytest.approx(8.69e9, rel=2e-3)
    assert bench.algorithmic_bytes("knn_C64", B, N, k) == pytest.approx(13.6e6, rel=5e-3)
    assert bench.algorithmic_bytes("chamfer_fwd", B, N, k) is None     # no work model: never the roofline kernel


def test_workloads(bench):
    a, s = bench.WORKLOADS["A"], bench.WORKLOADS["S"]
    assert a["layers"] == (3, 3, 64, 64, 128) and a["near"] == 20 and sum(a["fps_split"]) == 1024   # PointDA-10
    assert s["layers"] == (3, 3, 64, 64) and s["near"] == 10 and sum(s["fps_split"]) == 2048        # PointSegDA
    bench.set_workload("S")
    assert bench.LAYER_CHANNELS == s["layers"] and bench.SHIFT == 10 and bench.PERGROUP == 5
    bench.set_workload("A")
    assert bench.LAYER_CHANNELS == a["layers"] and bench.RADIUS == 0.13


def test_peaks_table(bench):
    p = bench.peaks()
    assert 5000 < p["hbm_gbs"] < 8100 and 1000 < p["bf16_tflops"] < 2300 and p["source"] in ("measured", "fallback")


def test_reference_arm_line():
    """`bench.py --impl reference` prints one JSON line with the contract's keys on rank 0 and nothing on other ranks."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workload_x_reference_arm_line():
    """configs[4] (256 x 4096, k = 40, feature-space kNN): the reference arm of `--workload X` prints the contract's line
    on a bounded sample of the workload."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "X", "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "clouds/s" and line["value"] > 0
    assert line["config"]["points"] == 4096 and line["config"]["k"] == 40 and line["config"]["feature_dims"] == [64, 128]
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "ggf_fwd_C64" not in line.get("op_ms_per_step", {})


def test_fused_call_accounting(bench):
    """The DGCNN layers make the fused get_graph_feature(idx=None) call: its work model is the edge tensor's bytes."""
    B, N, k = 32, 1024, 20
    assert bench.algorithmic_bytes("ggf_fwd_C64", B, N, k) == bench.algorithmic_bytes("edge_fwd_C64", B, N, k)
    assert bench.LAUNCHES["ggf_tensor"] == 3 and bench.LAUNCHES["ggf3"] == 1


def test_workload_e_reference_arm_line(bench):
    """SURVEY 8f rank 1 (`--workload E`, the EdgeConv backbone): the reference arm runs the reference's layer composition on
    a bounded sample and prints the contract's line; the layer list is the PointDA DGCNN's conv1..conv4."""
    assert bench.EC_LAYERS == ((3, 64), (64, 64), (64, 128), (128, 256))          # PointDA/Models.py:91-94
    assert bench.SEG_LAYERS == ((3, (64, 64)), (64, (64, 64)), (64, (64,)))        # PointSegDA/Models.py:159-163
    seg = bench._ec_make_layers("cpu", seg=True)
    assert [len(m) for m in seg] == [2, 2, 1] and seg[1][0].in_channels == 128 and seg[0][0].bias is not None
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "E", "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "clouds/s" and line["value"] > 0
    assert line["config"]["workload"] == "edgeconv-E" and line["config"]["points"] == 1024 and line["config"]["k"] == 20
    assert line["cpu_baseline"]["kind"] == "port" and line["gpu_launches"] == 0
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")          # other ranks exit 0 without work
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "E", "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_stdout_is_protected_from_library_prints():
    """bench.protect_stdout(): bytes written to file descriptor 1 below Python (NCCL's version banner) land on stderr, print()
    still reaches the real stdout -- a multi-GPU run prints exactly one JSON line there."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.protect_stdout(); "
            "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'); print('{\\\"metric\\\": 1}', flush=True)") % root
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == '{"metric": 1}', r.stdout
    assert "NCCL version" in r.stderr
