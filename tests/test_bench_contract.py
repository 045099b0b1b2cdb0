"""CPU test of the bench.py contract for the reference arm (the b200 arm needs a GPU and is exercised by the driver):
one JSON line with the agreed keys, `impl: reference`, a cpu_baseline describing the run and an e2e block equal to the line's value."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Msamples/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    assert "workload" in d["config"] and "256-channel fir_decimate_cc" in d["config"]["workload"]
    # the two arms must describe the SAME configuration (the driver compares key sets): both take it from base_config()
    sys.path.insert(0, str(ROOT))
    import bench
    assert d["config"] == bench.base_config(1)
    src = (ROOT / "bench.py").read_text()
    assert src.count('"config": base_config(') == 2


def test_non_zero_ranks_of_the_reference_arm_stay_silent():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, cwd=str(ROOT), env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
