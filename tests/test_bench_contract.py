"""CPU-side checks of bench.py: the reference arm runs without a GPU and prints the contract's JSON line; the
algorithmic-byte accounting matches SURVEY.md section 8(d) / BASELINE.md section 5."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_match_baseline_md():
    sys.path.insert(0, ROOT)
    import bench

    # B_iter, B_check, B_io at (32,64) and (64,128)
    assert bench.algorithmic_bytes(32, 64, 1, 0, 0, 0) == 41216
    assert bench.algorithmic_bytes(64, 128, 1, 0, 0, 0) == 164352
    assert bench.algorithmic_bytes(32, 64, 0, 1, 0, 0) == 24576
    assert bench.algorithmic_bytes(64, 128, 0, 1, 0, 0) == 98304
    assert bench.algorithmic_bytes(32, 64, 0, 0, 0, 1) == 26640
    assert bench.algorithmic_bytes(64, 128, 0, 0, 0, 1) == 102416
    assert bench.algorithmic_bytes(64, 128, 0, 0, 1, 0) == 8 * (64 * 64 + 128 * 64)


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-sample", "8"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "QP/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("QP-subproblems/sec") and line["dtype"] == "f64" and line["n_gpus"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": "QP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["value"] > 0 and line["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
