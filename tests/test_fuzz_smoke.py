"""CPU tier: a few seconds of each kernel fuzzer (tests/fuzz/*.py: random geometries through the real launchers under the CUDA-thread
emulator, checked against the oracle) with a fixed seed, so that the fuzzers keep working and a regression in a launcher's edge handling
shows up here.  The long runs are manual: `python tests/fuzz/fuzz_emulated.py <seed> <seconds>`."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "host_shim"))
import emul_build  # noqa: E402


@pytest.mark.parametrize("script", ["fuzz_emulated.py", "fuzz_emulated_small.py", "fuzz_emulated_fft.py"])
def test_fuzzer_runs_clean(script):
    if not emul_build.available():
        pytest.skip("needs g++ and the CUDA toolkit headers")
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "fuzz" / script), "12345", "8"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "iterations" in r.stdout
