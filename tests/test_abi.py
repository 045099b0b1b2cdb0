"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports every function that
include/csdr_b200.h declares; the host-side (non-CUDA) entry points agree with the oracle / golden vectors."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from oracle.pyoracle import rel_rms

ROOT = Path(__file__).resolve().parent.parent
GOLD = np.load(Path(__file__).parent / "golden" / "hotpath_golden.npz")


@pytest.fixture(scope="module")
def b200():
    import csdr_b200
    from csdr_b200.build import build
    build()
    csdr_b200.lib()
    return csdr_b200


def declared_functions():
    text = (ROOT / "include" / "csdr_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    names = set()
    for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text):
        n = m.group(1)
        if n not in {"sizeof", "defined", "if", "while", "for", "return"}:
            names.add(n)
    return sorted(names)


def test_library_exports_every_declared_symbol(b200):
    names = declared_functions()
    assert len(names) > 50, names
    lib = C.CDLL(str(b200.LIB_PATH))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/csdr_b200.h but not exported: {missing}"


def test_reference_names_are_exported_unmangled(b200):
    lib = C.CDLL(str(b200.LIB_PATH))
    for n in ("convert_u8_f", "convert_f_s16", "convert_s16_f", "shift_addition_init", "shift_addition_cc", "decimating_shift_addition_cc",
              "fir_decimate_cc", "fmdemod_quadri_cf", "fractional_decimator_ff_init", "fractional_decimator_ff", "fastagc_ff",
              "apply_fir_fft_cc", "make_fft_c2c", "fft_execute", "fft_destroy", "fastddc_init", "fastddc_inv_cc", "fft_swap_sides",
              "firdes_lowpass_f", "firdes_bandpass_c", "firdes_filter_len", "next_pow2", "log2n"):
        assert hasattr(lib, n), n


def test_host_design_functions_match_oracle_and_golden(b200, oracle):
    assert [b200.firdes_filter_len(float(b)) for b in GOLD["filter_len_bw"]] == list(GOLD["filter_len"])
    for T, c, w in ((79, 0.05, "HAMMING"), (199, 0.05, "HAMMING"), (101, 0.1, "BLACKMAN"), (801, 0.01, "HAMMING"), (33, 0.2, "BOXCAR")):
        assert np.array_equal(b200.firdes_lowpass_f(T, c, w), oracle.firdes_lowpass_f(T, c, w)), (T, c, w)
    assert rel_rms(b200.firdes_lowpass_f(199, 0.05), GOLD["lowpass_199"]) < 1e-6
    assert np.array_equal(b200.firdes_bandpass_c(79, 0.1, 0.3), oracle.firdes_bandpass_c(79, 0.1, 0.3))
    assert rel_rms(b200.firdes_bandpass_c(79, 0.1, 0.3), GOLD["bandpass_79"]) < 1e-6
    for rate in (-0.085, 0.2, 0.4999, 1e-4):
        assert b200.shift_addition_init(rate) == oracle.shift_addition_init(rate)
    keys = [str(k) for k in GOLD["ddc_keys"]]
    for case, row in zip(GOLD["ddc_cases"], GOLD["ddc_geometry"]):
        d = b200.fastddc_init(float(case[0]), int(case[1]), float(case[2]))
        for k, v in zip(keys, row):
            assert np.float32(getattr(d, k)) == np.float32(v), (case, k)
    assert b200.bandpass_geometry(0.002) == (1999, 4096, 2098, 1998)      # BASELINE config 5 geometry
    L = b200.lib()
    assert [L.next_pow2(x) for x in (0, 1, 2, 3, 2048, 2049)] == [1, 2, 4, 4, 4096, 4096]
    L.log2n.argtypes = [C.c_int]
    assert [L.log2n(x) for x in (1, 2, 3, 4096, 4097)] == [0, 1, -1, 12, -1]


def test_compute_entry_points_fail_loudly_without_gpu(b200):
    """No silent CPU fallback: without a CUDA device every compute entry point reports an error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = b200.lib()
    assert L.csdrb_device_count() <= 0
    rc = L.csdrb_convert_u8_f(C.c_void_p(4096), C.c_void_p(8192), 16, None)        # pointers are never touched on the host
    assert rc < 0 and b"CUDA error" in L.csdrb_last_error()
    taps = np.ones(199, np.float32)
    rc = L.csdrb_fir_decimate_bank_cc(C.c_void_p(4096), 4096, C.c_void_p(8192), 512, 1, 4096, 10, taps.ctypes.data_as(C.POINTER(C.c_float)), 199, -1, None)
    assert rc < 0
    with pytest.raises(b200.CsdrB200Error):
        b200._check(rc, "fir_decimate_bank_cc")
