"""CPU tier: csdr_b200/csrc/phase_table.cuh (the phase wrap of a fixed-increment chain as a table lookup) compiled for the host and
compared bit for bit with the reference's while loops (libcsdr_gpl.c:49-50): every piece boundary, strided sweeps of whole windows, and
whole chains as the shift / DDC / fastddc banks run them."""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
CUDA_INC = Path("/usr/local/cuda/include")
PI = np.float32(3.14159265358979323846)


@pytest.fixture(scope="module")
def wrap(tmp_path_factory):
    if not shutil.which("g++") or not (CUDA_INC / "cuda_runtime.h").exists():
        pytest.skip("needs g++ and the CUDA headers")
    so = tmp_path_factory.mktemp("wrapt") / "wrap_host.so"
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", f"-I{CUDA_INC}",
                    str(ROOT / "tests" / "host_shim" / "wrap_host.cpp"), "-o", str(so)], check=True, capture_output=True)
    L = C.CDLL(str(so))
    L.table_check.argtypes = [C.c_float, C.c_long, C.POINTER(C.c_int), C.POINTER(C.c_long), C.POINTER(C.c_float)]; L.table_check.restype = C.c_long
    L.table_chain.argtypes = [C.c_float, C.c_float, C.c_long]; L.table_chain.restype = C.c_long
    return L


def _check(wrap, inc, stride):
    pieces, checked, bad = C.c_int(0), C.c_long(0), C.c_float(0)
    miss = wrap.table_check(float(inc), stride, C.byref(pieces), C.byref(checked), C.byref(bad))
    assert miss == 0, (inc, pieces.value, bad.value)
    return pieces.value, checked.value


def test_every_float_of_the_window_for_large_increments(wrap):
    """coarse grids: the whole window is a few ten thousand floats -- all of them are compared"""
    rng = np.random.default_rng(0)
    total = 0; most = 0
    for inc in np.concatenate([rng.uniform(600, 4000, 24), -rng.uniform(600, 4000, 8), [1024.0, 2048.0, 4096.0, 2047.9999, 3216.99]]).astype(np.float32):
        p, n = _check(wrap, inc, 1)
        assert p > 0, inc                                                  # these do get a table
        total += n; most = max(most, p)
    assert total > 400_000 and most <= 48


def test_strided_windows_across_all_binades(wrap):
    rng = np.random.default_rng(1)
    seen = set()
    for E in range(0, 20):
        for inc in rng.uniform(2.0 ** E, 2.0 ** (E + 1), 8).astype(np.float32):
            ulps = 9.0 / np.spacing(np.float32(inc))
            budget = max(300.0, min(25_000.0, 1.5e7 / (float(inc) / 6.28 + 1.0)))     # the reference loop costs |x| / 2 pi iterations per value
            p, n = _check(wrap, inc, max(1, int(ulps // budget)))
            seen.add((E, p > 0))
    assert (2, False) in seen and (9, True) in seen                        # tiny increments need no table, the common ones have one


def test_increments_of_real_banks(wrap):
    """inc = fl(fl(rate2*PI)*n) as the kernels compute it: CLI chunk 1024 (shift / DDC banks), 224 outputs per fastddc block"""
    rng = np.random.default_rng(2)
    for n in (1024, 224, 16384, 100):
        for rate in rng.uniform(-0.5, 0.5, 40).astype(np.float32):
            rate2 = np.float32(rate * np.float32(2))
            inc = np.float32(np.float32(rate2 * PI) * np.float32(n))
            _check(wrap, inc, max(1, int(9.0 / np.spacing(np.float32(abs(inc) + 1)) // 20_000)))


def test_whole_chains(wrap):
    rng = np.random.default_rng(3)
    for rate in rng.uniform(-0.5, 0.5, 60).astype(np.float32):
        inc = np.float32(np.float32(np.float32(rate * np.float32(2)) * PI) * np.float32(1024))
        assert wrap.table_chain(float(inc), 0.0, 20_000) == -1, inc
    # a caller's first phase may be anything: the first steps take the fallback, then the chain is inside the window
    assert wrap.table_chain(2900.0, 1.0e5, 2000) == -1
    assert wrap.table_chain(-2900.0, -7.5, 2000) == -1


def test_fp64_pipe_walk_rounds_like_fp32():
    """fastddc_phasor_kernel<true> runs the float recursion on the FP64 pipe: (float)((double)a * (double)b) and (float)((double)x +- (double)y) must be the
    float product / sum bit for bit (53 >= 2*24 + 2).  Checked on 4 M random pairs incl. tiny and cancelling operands, and on a 4096-step phasor walk."""
    rng = np.random.default_rng(11)
    n = 1 << 22
    a = (rng.standard_normal(n) * np.exp(rng.uniform(-60, 3, n))).astype(np.float32)
    b = (rng.standard_normal(n) * np.exp(rng.uniform(-60, 3, n))).astype(np.float32)
    b[: n // 8] = -a[: n // 8] * np.float32(1 + 2.0 ** -20)                                          # near cancellation
    with np.errstate(under="ignore"):
        assert np.array_equal((a.astype(np.float64) * b.astype(np.float64)).astype(np.float32).view(np.uint32), (a * b).view(np.uint32))
        assert np.array_equal((a.astype(np.float64) + b.astype(np.float64)).astype(np.float32).view(np.uint32), (a + b).view(np.uint32))
        assert np.array_equal((a.astype(np.float64) - b.astype(np.float64)).astype(np.float32).view(np.uint32), (a - b).view(np.uint32))
    rate = rng.uniform(-0.5, 0.5, 512).astype(np.float32)
    cd, sd = np.cos(np.float64(rate) * np.pi).astype(np.float32), np.sin(np.float64(rate) * np.pi).astype(np.float32)
    c32, s32 = np.ones(512, np.float32), np.zeros(512, np.float32)
    c64, s64 = c32.copy(), s32.copy()
    f = lambda v: v.astype(np.float64)
    for _ in range(4096):
        c32, s32 = c32 * cd - s32 * sd, s32 * cd + c32 * sd
        a1, a2 = (f(c64) * f(cd)).astype(np.float32), (f(s64) * f(sd)).astype(np.float32)
        a3, a4 = (f(s64) * f(cd)).astype(np.float32), (f(c64) * f(sd)).astype(np.float32)
        c64, s64 = (f(a1) - f(a2)).astype(np.float32), (f(a3) + f(a4)).astype(np.float32)
    assert np.array_equal(c32.view(np.uint32), c64.view(np.uint32)) and np.array_equal(s32.view(np.uint32), s64.view(np.uint32))
