"""bench.py's JSON contract, checked on the arm that runs without a GPU (--impl reference = the CPU oracle port)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"].startswith("rgbbox 1000x1000 64spp + irreg 1000x1000 64spp")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=60, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_full_frame_known_answer_helper(oracle):
    """bench.full_frame_equals_oracle (the per-N / per-config `…sha256_equals_oracle…` flags of the JSON line): True for the
    oracle's own frame, False for a frame with one pixel changed, None for a config without a recorded answer."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    frame, _, _ = oracle.render_scene("irreg", 4000, 4000)
    assert bench.full_frame_equals_oracle(frame, "irreg_4000x4000_1spp") is True
    frame[1234, 567] ^= 1
    assert bench.full_frame_equals_oracle(frame, "irreg_4000x4000_1spp") is False
    assert bench.full_frame_equals_oracle(frame, "irreg_123x45_6spp") is None
