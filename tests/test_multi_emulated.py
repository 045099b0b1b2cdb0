"""CPU tier: the one-process multi-GPU bank (csrc/multi.cu: channel slices per device, ncclBroadcast of the wideband block, submit / collect
pipeline) on the emulated library with TWO pretend devices (CUDA_EMUL_DEVICES=2, all host memory) and a memcpy stand-in for NCCL
(tests/host_shim/fake_nccl.c through CSDRB_NCCL_LIB).  What is checked is the host logic the GPU tier cannot reach without several GPUs:
slicing, buffer rotation, ticket order, and that the sliced bank equals the unsliced one bit for bit over a stream of blocks."""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "host_shim"))
import emul_build  # noqa: E402

vp, it, lg = C.c_void_p, C.c_int, C.c_long


@pytest.fixture(scope="module")
def L(tmp_path_factory):
    if not emul_build.available():
        pytest.skip("needs g++ and the CUDA toolkit headers")
    lib, _cli = emul_build.build_full_once(tmp_path_factory)
    fake = tmp_path_factory.mktemp("fake_nccl") / "libfake_nccl.so"
    subprocess.run(["gcc", "-O1", "-fPIC", "-shared", str(ROOT / "tests" / "host_shim" / "fake_nccl.c"), "-o", str(fake)], check=True)
    os.environ["CUDA_EMUL_DEVICES"] = "2"
    os.environ["CSDRB_NCCL_LIB"] = str(fake)
    L = C.CDLL(str(lib))
    L.csdrb_last_error.restype = C.c_char_p
    L.csdrb_multi_bank_create.restype = vp
    L.csdrb_multi_bank_create.argtypes = [it, vp, it, vp, it, vp, it, it, it, it]
    L.csdrb_multi_bank_destroy.argtypes = [vp]
    L.csdrb_multi_bank_submit.argtypes = [vp, vp, it, vp, lg]
    L.csdrb_multi_bank_collect.argtypes = [vp, it]
    L.csdrb_multi_bank_process_host.argtypes = [vp, vp, it, vp, lg]
    L.csdrb_multi_bank_slice.argtypes = [vp, it, vp, vp, vp]
    L.csdrb_multi_bank_set_rate.argtypes = [vp, it, C.c_float]
    L.fake = C.CDLL(str(fake))
    yield L
    del os.environ["CUDA_EMUL_DEVICES"], os.environ["CSDRB_NCCL_LIB"]


def _aligned(shape, dtype):
    n = int(np.prod(shape)); item = np.dtype(dtype).itemsize
    raw = np.zeros(n * item + 32, np.uint8)
    off = (-raw.ctypes.data) % 16
    return raw[off:off + n * item].view(dtype).reshape(shape)


def _run(L, ndev, rates, taps, D, blocks, demod, pipelined, retune=None):
    """stream `blocks` (list of complex64 arrays, each starting where the previous one stopped consuming) through a multi bank; returns [C, total]"""
    C_ = rates.size
    maxb = max(b.size for b in blocks)
    m = L.csdrb_multi_bank_create(ndev, None, C_, rates.ctypes.data, D, taps.ctypes.data, taps.size, demod, 1024, maxb)
    assert m, L.csdrb_last_error()
    dt = np.float32 if demod else np.complex64
    outs, bufs, tickets = [], [], []
    for k, b in enumerate(blocks):
        if retune and retune[0] == k:
            assert L.csdrb_multi_bank_set_rate(m, retune[1], retune[2]) == 0
        x = _aligned(b.size, np.complex64); x[:] = b
        n_out = (b.size - taps.size) // D + 1
        o = _aligned((C_, n_out + 3), dt); o[:] = np.nan
        bufs.append((x, o, n_out))
        if pipelined:
            t = L.csdrb_multi_bank_submit(m, x.ctypes.data, b.size, o.ctypes.data, n_out + 3)
            assert t >= 0, L.csdrb_last_error()
            tickets.append(t)
            if len(tickets) == 2:                                          # two in flight: the oldest must be collected before the next submit
                assert L.csdrb_multi_bank_submit(m, x.ctypes.data, b.size, o.ctypes.data, n_out + 3) < 0 and b"in flight" in L.csdrb_last_error()
                assert L.csdrb_multi_bank_collect(m, tickets[1]) < 0       # out of order
                assert L.csdrb_multi_bank_collect(m, tickets.pop(0)) == bufs[-2][2]
        else:
            assert L.csdrb_multi_bank_process_host(m, x.ctypes.data, b.size, o.ctypes.data, n_out + 3) == n_out, L.csdrb_last_error()
    for t in tickets:
        assert L.csdrb_multi_bank_collect(m, t) >= 0
    sl = []
    for i in range(ndev):
        d, c0, nc = it(), it(), it()
        assert L.csdrb_multi_bank_slice(m, i, C.byref(d), C.byref(c0), C.byref(nc)) == 0
        sl.append((d.value, c0.value, nc.value))
    L.csdrb_multi_bank_destroy(m)
    return np.concatenate([o[:, :n] for _, o, n in bufs], axis=1), sl


@pytest.mark.parametrize("demod", [1, 0])
def test_sliced_bank_equals_the_unsliced_one(L, demod):
    from oracle.pyoracle import Oracle
    o = Oracle()
    D, bw = 50, 0.005
    T = o.firdes_filter_len(bw)
    taps = np.ascontiguousarray(o.firdes_lowpass_f(T, 0.5 / D), np.float32)
    rates = np.array([-0.41, -0.27, -0.13, 0.01, 0.15, 0.29, 0.43], np.float32)
    rng = np.random.default_rng(11)
    stream = ((rng.uniform(-1, 1, 40_000) + 1j * rng.uniform(-1, 1, 40_000)) * 0.5).astype(np.complex64)
    blocks, pos = [], 0
    for size in (9000, 7013, 12_000, 5000):                                # every block starts where the previous one stopped consuming (csdr.c:1172-1174)
        blocks.append(stream[pos:pos + size]); pos += ((size - T) // D + 1) * D
    calls0 = L.fake.fake_nccl_broadcast_calls()
    one, s1 = _run(L, 1, rates, taps, D, blocks, demod, pipelined=False)
    assert L.fake.fake_nccl_broadcast_calls() == calls0                    # a one-device bank never touches NCCL
    two, s2 = _run(L, 2, rates, taps, D, blocks, demod, pipelined=True)
    assert L.fake.fake_nccl_broadcast_calls() == calls0 + 2 * len(blocks)  # one broadcast per device and block
    assert s1 == [(0, 0, 7)] and s2 == [(0, 0, 4), (1, 4, 3)]
    assert np.array_equal(one, two) and not np.isnan(one.view(np.float32)).any()
    # against the reference chain itself, one channel of each slice
    for c in (1, 5):
        sh, _ = o.shift_addition_cc(stream[:pos + T], float(rates[c]), 0.0, 1024)
        base = o.fir_decimate_cc(sh, D, taps)[:one.shape[1]]
        want = o.fmdemod_quadri_cf(base)[0] if demod else base
        from oracle.pyoracle import rel_rms
        e = rel_rms(two[c], want)
        assert e < (1e-5 if demod else 2e-6), (c, e)
    # a retune reaches the right slice
    a, _ = _run(L, 2, rates, taps, D, blocks, demod, pipelined=True, retune=(2, 5, 0.2))
    b, _ = _run(L, 1, rates, taps, D, blocks, demod, pipelined=False, retune=(2, 5, 0.2))
    from oracle.pyoracle import rel_rms as rr
    assert np.array_equal(a, b)                                           # independent of the slicing: every slice re-chunks its NCO at the retune sample
    assert rr(a[5], two[5]) > 0.1 and rr(a[4], two[4]) < 1e-3             # channel 5 moved; its neighbours only see the re-chunking (a few 1e-5)
