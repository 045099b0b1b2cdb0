"""CPU tier: the SHIPPED CUDA kernels executed thread by thread on the host.

tests/host_shim/cuda_emul.h turns every CUDA thread of a CTA into a fiber and __syncthreads()/__syncwarp() into real barriers, so the
kernel sources under csdr_b200/csrc (included unmodified) run with their real index arithmetic, shared-memory traffic, barrier
structure and IEEE single-precision rounding -- in a container without a GPU.  This is test infrastructure (nothing in the product
can reach it; the product still fails loudly without a GPU); it complements, not replaces, the -m gpu parity tests.
The scheduling order of the fibers between barriers is varied (forward / reverse / random) so that a missing barrier shows up as a
wrong result in at least one order.
"""
import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
SHIM = ROOT / "tests" / "host_shim"
CUDA_INC = Path("/usr/local/cuda/include")
GOLD = np.load(Path(__file__).parent / "golden" / "hotpath_golden.npz")
vp = C.c_void_p


def _rel(a, b):
    return float(np.sqrt(np.sum(np.abs(np.asarray(a, np.complex128) - np.asarray(b, np.complex128)) ** 2) / max(np.sum(np.abs(np.asarray(b, np.complex128)) ** 2), 1e-300)))


def _build(tmp, name):
    if not shutil.which("g++") or not (CUDA_INC / "cuda_runtime.h").exists():
        pytest.skip("needs g++ and the CUDA toolkit headers")
    so = tmp / f"{name}.so"
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w", f"-I{CUDA_INC}", f"-I{SHIM}",
                    str(SHIM / f"{name}.cpp"), "-o", str(so)], check=True, capture_output=True)
    return so


@pytest.fixture(scope="module")
def fft_so(tmp_path_factory):
    return _build(tmp_path_factory.mktemp("emul"), "fft_emul")           # one g++ run (~15 s) for the module


@pytest.fixture(scope="module", params=["alternate", "reverse", "random"])
def fft(request, fft_so):
    so = fft_so.with_name(f"fft_emul_{request.param}.so")                # one copy per order: the order is read when a copy initialises
    shutil.copy(fft_so, so)
    os.environ["CUDA_EMUL_ORDER"] = request.param
    L = C.CDLL(str(so))
    L.emul_fft_c2c.argtypes = [vp, C.c_long, vp, C.c_long, C.c_int, C.c_int, C.c_int]
    L.emul_olafir.argtypes = [vp, C.c_long, vp, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_long, vp, C.c_int]
    L.emul_fastddc_fwd.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int]
    L.emul_apply_fir_fft.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int]
    L.emul_fastddc_inv_bank.argtypes = [vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_long, vp, C.c_int]
    L.emul_barriers.restype = C.c_long
    return L


def _cplx(rng, *shape):
    return (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(np.complex64)


def test_k7_fft_every_size_both_directions(fft):
    rng = np.random.default_rng(0)
    for n in (2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384):
        x = _cplx(rng, 2, n); y = np.zeros_like(x)
        for inv in (0, 1):
            assert fft.emul_fft_c2c(x.ctypes.data, n, y.ctypes.data, n, n, 2, inv) == 0
            want = np.fft.ifft(x.astype(np.complex128), axis=1) * n if inv else np.fft.fft(x.astype(np.complex128), axis=1)
            assert _rel(y, want) < 1e-6, (n, inv)                        # same bar as tests/test_gpu_parity2.py::test_fft_all_sizes_vs_float64_dft
        buf = _cplx(rng, n + 1); out = np.zeros(n + 1, np.complex64)    # a row that is only 8-byte aligned
        fft.emul_fft_c2c(buf[1:].ctypes.data, n, out[1:].ctypes.data, n, n, 1, 0)
        assert _rel(out[1:], np.fft.fft(buf[1:].astype(np.complex128))) < 1e-6
    assert fft.emul_barriers() > 100                                      # the barriers were real


def _overlap_add(x, H, N, isz):
    nb = x.size // isz; out = np.zeros(nb * isz + N - isz, np.complex128)
    for b in range(nb):
        blk = np.zeros(N, np.complex128); blk[:isz] = x[b * isz:(b + 1) * isz]
        out[b * isz:b * isz + N] += np.fft.ifft(np.fft.fft(blk) * H)
    return out[:nb * isz], out[nb * isz:]


@pytest.mark.parametrize("N,isz,nb,bpc", [(4096, 2098, 5, 2), (4096, 2098, 3, 8), (512, 300, 7, 3), (64, 40, 9, 4), (256, 178, 6, 6), (1024, 224, 12, 5),
                                          (2048, 1500, 4, 2), (16, 9, 11, 3), (128, 128, 3, 2), (32, 1, 70, 16)])
def test_k9_overlap_add_bank(fft, N, isz, nb, bpc):
    """bandpass_fir_fft_cc block loop: CTA runs of `bpc` blocks (lead-in recomputation), overlap > input_size, no overlap, streaming tails."""
    rng = np.random.default_rng(N + isz)
    ch = 2
    x = _cplx(rng, ch, nb * isz); H = _cplx(rng, ch, N)
    y = np.zeros_like(x); tail = np.zeros((ch, N), np.complex64)
    assert fft.emul_olafir(x.ctypes.data, x.shape[1], y.ctypes.data, y.shape[1], ch, N, isz, nb, H.ctypes.data, N, tail.ctypes.data, bpc) == 0
    for c in range(ch):
        want, wt = _overlap_add(x[c].astype(np.complex128), H[c].astype(np.complex128), N, isz)
        assert _rel(y[c], want) < 2e-6
        if N > isz:
            assert _rel(tail[c, :N - isz], wt) < 2e-6
    h = nb // 2                                                           # two calls carrying the tail == one call
    xa = np.ascontiguousarray(x[:, :h * isz]); xb = np.ascontiguousarray(x[:, h * isz:]); ya = np.zeros_like(xa); yb = np.zeros_like(xb)
    t = np.zeros((ch, N), np.complex64)
    fft.emul_olafir(xa.ctypes.data, xa.shape[1], ya.ctypes.data, ya.shape[1], ch, N, isz, h, H.ctypes.data, N, t.ctypes.data, bpc)
    fft.emul_olafir(xb.ctypes.data, xb.shape[1], yb.ctypes.data, yb.shape[1], ch, N, isz, nb - h, H.ctypes.data, N, t.ctypes.data, bpc)
    assert _rel(np.concatenate([ya, yb], 1), y) < 1e-6 and _rel(t, tail) < 1e-6 + (N == isz)


def test_k9_golden_and_dropin_kernel(fft, oracle):
    """the golden bandpass stream of the compiled reference through the bank kernel, and apply_fir_fft_cc's one-block kernel"""
    bw = 0.05
    T = oracle.firdes_filter_len(bw); N = 256; isz = N - T + 1
    taps = np.zeros(N, np.complex64); taps[:T] = oracle.firdes_bandpass_c(T, -0.1, 0.2)
    H = oracle.dft(taps)
    x = GOLD["bp_in"]; nb = x.size // isz
    y = np.zeros(nb * isz, np.complex64); tail = np.zeros((1, N), np.complex64)
    assert fft.emul_olafir(x.ctypes.data, x.size, y.ctypes.data, y.size, 1, N, isz, nb, H.ctypes.data, N, tail.ctypes.data, 4) == 0
    assert _rel(y, GOLD["bp_out"][:y.size]) < 5e-6                        # same bar as the GPU test
    rng = np.random.default_rng(3)
    blk = np.zeros(N, np.complex64); blk[:isz] = _cplx(rng, isz); last = _cplx(rng, T - 1); out = np.zeros(N, np.complex64)
    assert fft.emul_apply_fir_fft(blk.ctypes.data, H.ctypes.data, last.ctypes.data, T - 1, out.ctypes.data, N) == 0
    want = np.fft.ifft(np.fft.fft(blk.astype(np.complex128)) * H.astype(np.complex128)); want[:T - 1] += last
    assert _rel(out, want) < 2e-6


def test_fastddc_forward_and_inverse_against_golden(fft, oracle):
    """a12/a13: golden spectra and golden channel output of the compiled reference (bw 0.05, decimation 8, shift 0.123)."""
    bw, dec, shift = 0.05, 8, 0.123
    ddc, _ = oracle.fastddc_init(bw, dec, shift)
    x = GOLD["ddc_in"]; nb = x.size // ddc.input_size
    sp = np.zeros((nb, ddc.fft_size), np.complex64); carry = np.zeros(ddc.overlap_length, np.complex64)
    assert fft.emul_fastddc_fwd(x.ctypes.data, sp.ctypes.data, carry.ctypes.data, ddc.fft_size, ddc.input_size, nb) == 0
    assert _rel(sp, GOLD["ddc_fwd_out"]) < 1e-6
    assert np.array_equal(carry, x[nb * ddc.input_size - ddc.overlap_length:nb * ddc.input_size])
    from oracle.pyoracle import _CF, _p, WINDOWS
    tf = np.empty(ddc.fft_size, np.complex64)
    oracle.L.oracle_fastddc_make_taps_fft(C.byref(ddc), shift, dec, WINDOWS["HAMMING"], _p(tf, _CF))
    chan = np.zeros(1, np.dtype([("offsetbin", np.int32), ("sindelta", np.float32), ("cosdelta", np.float32), ("rate", np.float32)]))
    chan["offsetbin"] = ddc.offsetbin; chan["sindelta"] = ddc.dsadata.sindelta; chan["cosdelta"] = ddc.dsadata.cosdelta; chan["rate"] = ddc.dsadata.rate
    golden_sp = np.ascontiguousarray(GOLD["ddc_fwd_out"])                 # keep the array alive while its pointer is in use
    for force_simple in (0, 1):                                           # the tiled kernel and the one-CTA-per-(block, channel) kernel
        remain = np.zeros(1, np.int32); phase = np.zeros(1, np.float32); total = np.zeros(1, np.int32)
        out = np.zeros((1, nb * ddc.post_input_size), np.complex64)
        rc = fft.emul_fastddc_inv_bank(golden_sp.ctypes.data, nb, tf.ctypes.data, chan.ctypes.data, 1, ddc.fft_size, ddc.fft_inv_size, ddc.pre_decimation,
                                       ddc.scrap, ddc.post_input_size, ddc.post_decimation, remain.ctypes.data, phase.ctypes.data, out.ctypes.data, out.shape[1],
                                       total.ctypes.data, force_simple)
        assert rc == (0 if force_simple else 1)
        assert total[0] == GOLD["ddc_inv_out"].size
        assert _rel(out[0, :total[0]], GOLD["ddc_inv_out"]) < 5e-6       # same bar as the GPU test
