"""CPU tier: argument checking and error reporting of the device-resident C ABI (csrc/capi.cu + launchers), exercised on the emulated
library (tests/host_shim/emul_build.build_full): every refusal is a negative return code with a message in csdrb_last_error(), nothing is
launched, nothing crashes.  (On the real library the same code runs; without a GPU it cannot get this far, tests/test_abi.py.)"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "host_shim"))
import emul_build  # noqa: E402

vp = C.c_void_p


@pytest.fixture(scope="module")
def L(tmp_path_factory):
    if not emul_build.available():
        pytest.skip("needs g++ and the CUDA toolkit headers")
    lib, _cli = emul_build.build_full_once(tmp_path_factory)
    L = C.CDLL(str(lib))
    L.csdrb_last_error.restype = C.c_char_p
    L.csdrb_kernel_launches.restype = C.c_long
    return L


def _aligned(n, dtype):
    raw = np.zeros(n * np.dtype(dtype).itemsize + 32, np.uint8)
    off = (-raw.ctypes.data) % 16
    return raw[off:off + n * np.dtype(dtype).itemsize].view(dtype)


def test_refusals_have_codes_and_messages(L):
    f = _aligned(4096, np.float32); g = _aligned(4096, np.float32); s = _aligned(4096, np.int16)
    before = L.csdrb_kernel_launches()
    L.csdrb_convert_f_s16.argtypes = [vp, vp, C.c_long, vp]
    assert L.csdrb_convert_f_s16(None, s.ctypes.data, 16, None) < 0 and b"null" in L.csdrb_last_error()
    assert L.csdrb_convert_f_s16(f.ctypes.data + 4, s.ctypes.data, 16, None) < 0 and b"aligned" in L.csdrb_last_error()
    L.csdrb_limit_ff.argtypes = [vp, vp, C.c_long, C.c_float, vp]
    assert L.csdrb_limit_ff(f.ctypes.data, g.ctypes.data + 8, 16, 1.0, None) < 0 and b"aligned" in L.csdrb_last_error()
    # FIR bank: nonsense geometry
    L.csdrb_fir_decimate_bank_cc.argtypes = [vp, C.c_long, vp, C.c_long, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp]
    taps = np.ones(199, np.float32)
    x = _aligned(2 * 4096, np.float32); y = _aligned(2 * 4096, np.float32)
    assert L.csdrb_fir_decimate_bank_cc(x.ctypes.data, 4096, y.ctypes.data, 512, 1, 4096, 0, taps.ctypes.data, 199, -1, None) < 0          # decimation 0
    assert L.csdrb_fir_decimate_bank_cc(x.ctypes.data, 4096, y.ctypes.data, 512, 1, 4096, 10, None, 199, -1, None) < 0                      # no taps
    assert L.csdrb_fir_decimate_bank_cc(x.ctypes.data, 4096, y.ctypes.data, 512, 1, 4096, 10, taps.ctypes.data, 199, 99, None) < 0         # unknown tiling
    # FFT: unsupported size, overlap-add with input_size > fft_size
    L.csdrb_fft_c2c_batch.argtypes = [vp, C.c_long, vp, C.c_long, C.c_int, C.c_int, C.c_int, vp]
    assert L.csdrb_fft_c2c_batch(x.ctypes.data, 100, y.ctypes.data, 100, 100, 1, 0, None) < 0 and b"power of two" in L.csdrb_last_error()
    assert L.csdrb_fft_c2c_batch(x.ctypes.data, 32768, y.ctypes.data, 32768, 32768, 1, 0, None) < 0
    L.csdrb_bandpass_fir_fft_bank_cc.argtypes = [vp, C.c_long, vp, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_long, vp, vp]
    assert L.csdrb_bandpass_fir_fft_bank_cc(x.ctypes.data, 4096, y.ctypes.data, 4096, 1, 256, 300, 2, x.ctypes.data, 0, y.ctypes.data, None) < 0
    # fused DDC bank: a geometry without a fused kernel answers -2 ("use the unfused bank calls"), a wideband pointer that is not 16-byte aligned -1
    L.csdrb_ddc_bank_scratch_bytes.restype = C.c_size_t; L.csdrb_ddc_bank_scratch_bytes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    L.csdrb_ddc_bank.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, C.c_long, vp, vp, vp, C.c_size_t, vp]
    params = np.zeros(3, np.float32); phase = np.zeros(1, np.float32)
    sb = L.csdrb_ddc_bank_scratch_bytes(1, 2000, 1024, 0); scratch = np.zeros(sb + 64, np.uint8)
    t79 = np.ones(79, np.float32)
    rc = L.csdrb_ddc_bank(x.ctypes.data, 2000, 1, params.ctypes.data, phase.ctypes.data, 1024, 0, 7, t79.ctypes.data, 79, 0, y.ctypes.data, 400, None, None,
                          scratch.ctypes.data, sb, None)
    assert rc == -2 and b"no fused kernel" in L.csdrb_last_error()
    rc = L.csdrb_ddc_bank(x.ctypes.data + 8, 2000, 1, params.ctypes.data, phase.ctypes.data, 1024, 0, 10, t79.ctypes.data, 79, 0, y.ctypes.data, 400, None, None,
                          scratch.ctypes.data, sb, None)
    assert rc == -1 and b"16-byte aligned" in L.csdrb_last_error()
    rc = L.csdrb_ddc_bank(x.ctypes.data, 2000, 1, params.ctypes.data, phase.ctypes.data, 1024, 0, 10, t79.ctypes.data, 79, 0, y.ctypes.data, 400, None, None,
                          scratch.ctypes.data, 8, None)
    assert rc < 0 and b"scratch" in L.csdrb_last_error()
    # NFM de-emphasis: unknown rate is 0 outputs, not an error (libcsdr.c:1119); too many taps for the caller-taps entry is an error
    L.csdrb_deemphasis_nfm_bank_ff.argtypes = [vp, C.c_long, vp, C.c_long, C.c_int, C.c_int, C.c_int, C.c_float, vp]
    assert L.csdrb_deemphasis_nfm_bank_ff(f.ctypes.data, 4096, g.ctypes.data, 4096, 1, 4096, 22050, 0.0, None) == 0
    L.csdrb_fir_valid_bank_ff.argtypes = [vp, C.c_long, vp, C.c_long, C.c_int, C.c_int, vp, C.c_int, C.c_float, vp]
    big = np.ones(209, np.float32)
    assert L.csdrb_fir_valid_bank_ff(f.ctypes.data, 4096, g.ctypes.data, 4096, 1, 4096, big.ctypes.data, 209, 0.0, None) < 0
    # channels ride in gridDim.y: a bank of more than 65535 channels is refused with a message, not with a launch-configuration error
    assert L.csdrb_fir_decimate_bank_cc(x.ctypes.data, 200, y.ctypes.data, 14, 70000, 200, 10, t79.ctypes.data, 79, -1, None) < 0 and b"65535" in L.csdrb_last_error()
    assert L.csdrb_kernel_launches() == before                              # none of the refused calls launched anything


def test_bank_object_lifecycle(L):
    """csdrb_ddc_bank_create / set_rate / offset / process / destroy on a small stream: bad handles and geometries are refused, a good one streams"""
    L.csdrb_ddc_bank_create.restype = vp
    L.csdrb_ddc_bank_create.argtypes = [C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int]
    L.csdrb_ddc_bank_destroy.argtypes = [vp]
    L.csdrb_ddc_bank_set_rate.argtypes = [vp, C.c_int, C.c_float]
    L.csdrb_ddc_bank_offset.argtypes = [vp]
    L.csdrb_ddc_bank_process.argtypes = [vp, vp, C.c_int, vp, C.c_long, vp]
    rates = np.array([0.1, -0.2], np.float32); t79 = np.ones(79, np.float32) / 79
    assert not L.csdrb_ddc_bank_create(2, rates.ctypes.data, 7, t79.ctypes.data, 79, 1, 1024)       # no fused kernel for decimation 7
    assert not L.csdrb_ddc_bank_create(0, rates.ctypes.data, 10, t79.ctypes.data, 79, 1, 1024)
    bank = L.csdrb_ddc_bank_create(2, rates.ctypes.data, 10, t79.ctypes.data, 79, 1, 1024)
    assert bank
    assert L.csdrb_ddc_bank_set_rate(bank, 5, 0.3) < 0 and L.csdrb_ddc_bank_set_rate(bank, 1, 0.3) >= 0
    x = _aligned(2 * 5000, np.float32); x[:] = np.random.default_rng(0).uniform(-1, 1, x.size)
    out = np.zeros((2, 600), np.float32)
    n_out = L.csdrb_ddc_bank_process(bank, x.ctypes.data, 5000, out.ctypes.data, 600, None)
    assert n_out == (5000 - 79) // 10 + 1 and L.csdrb_ddc_bank_offset(bank) == (n_out * 10) % 1024
    assert np.all(np.isfinite(out[:, :n_out]))
    L.csdrb_ddc_bank_destroy(bank)
    L.csdrb_ddc_bank_destroy(None)                                          # harmless
