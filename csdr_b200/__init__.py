"""csdr_b200 -- B200-native csdr block-DSP hot path.

This package is a thin Python mirror of the C ABI in ``include/csdr_b200.h`` (ctypes; no torch types cross
the boundary -- tensors are passed as raw device pointers + the current CUDA stream handle).  PyTorch is
used for device memory, streams and torch.distributed only.  There is no CPU fallback: every compute call
goes to ``libcsdr_b200.so`` and raises if that library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libcsdr_b200.so"
WINDOWS = {"BOXCAR": 0, "BLACKMAN": 1, "HAMMING": 2}


class CsdrB200Error(RuntimeError):
    pass


class _CF(C.Structure):
    _fields_ = [("i", C.c_float), ("q", C.c_float)]


class _Shift(C.Structure):
    _fields_ = [("sindelta", C.c_float), ("cosdelta", C.c_float), ("rate", C.c_float)]


class _DShiftStatus(C.Structure):
    _fields_ = [("decimation_remain", C.c_int), ("starting_phase", C.c_float), ("output_size", C.c_int)]


class _FracDec(C.Structure):            # include/csdr_b200.h fractional_decimator_ff_t (= libcsdr.h:151-168)
    _fields_ = [("where", C.c_float), ("input_processed", C.c_int), ("output_size", C.c_int), ("num_poly_points", C.c_int),
                ("poly_precalc_denomiator", C.c_void_p), ("coeffs_buf", C.c_void_p), ("filtered_buf", C.c_void_p),
                ("xifirst", C.c_int), ("xilast", C.c_int), ("rate", C.c_float), ("taps", C.c_void_p), ("taps_length", C.c_int)]


class _FastAgc(C.Structure):            # fastagc_ff_t (= libcsdr.h:118-128)
    _fields_ = [("buffer_1", C.c_void_p), ("buffer_2", C.c_void_p), ("buffer_input", C.c_void_p), ("peak_1", C.c_float),
                ("peak_2", C.c_float), ("input_size", C.c_int), ("reference", C.c_float), ("last_gain", C.c_float)]


class _Unroll(C.Structure):             # shift_unroll_data_t (= libcsdr.h:199-205)
    _fields_ = [("dsin", C.POINTER(C.c_float)), ("dcos", C.POINTER(C.c_float)), ("phase_increment", C.c_float), ("size", C.c_int)]


class _Table(C.Structure):              # shift_table_data_t (= libcsdr.h:180-184)
    _fields_ = [("table", C.POINTER(C.c_float)), ("table_size", C.c_int)]


class _Ima(C.Structure):                # ima_adpcm_state_t (= ima_adpcm.h:35-38)
    _fields_ = [("index", C.c_int), ("previousValue", C.c_int)]


class _AddFast(C.Structure):            # shift_addfast_data_t (= libcsdr.h:189-194)
    _fields_ = [("dsin", C.c_float * 4), ("dcos", C.c_float * 4), ("phase_increment", C.c_float)]


class _Plan(C.Structure):               # struct fft_plan_s (= fft_fftw.h:14-20)
    _fields_ = [("size", C.c_int), ("input", C.c_void_p), ("output", C.c_void_p), ("plan", C.c_void_p)]


class FastDDC(C.Structure):             # fastddc_t (= fastddc.h:5-24)
    _fields_ = [(n, C.c_int) for n in ("pre_decimation", "post_decimation", "taps_length", "taps_min_length", "overlap_length",
                                       "fft_size", "fft_inv_size", "input_size", "post_input_size")] + \
               [("pre_shift", C.c_float), ("startbin", C.c_int), ("v", C.c_int), ("offsetbin", C.c_int), ("post_shift", C.c_float),
                ("output_scrape", C.c_int), ("scrap", C.c_int), ("dsadata", _Shift)]


_lib = None


def build(force: bool = False):
    from .build import build as _b
    return _b(force=force)


def lib() -> C.CDLL:
    """Load libcsdr_b200.so (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise CsdrB200Error(f"{LIB_PATH} not built: run `python -m csdr_b200.build` (needs nvcc); there is no CPU fallback")
    L = C.CDLL(str(LIB_PATH))
    L.csdrb_last_error.restype = C.c_char_p
    L.csdrb_version.restype = C.c_char_p
    L.csdrb_kernel_launches.restype = C.c_long
    vp, lg, it = C.c_void_p, C.c_long, C.c_int
    L.csdrb_convert_u8_f.argtypes = [vp, vp, lg, vp]
    L.csdrb_convert_s16_f.argtypes = [vp, vp, lg, vp]
    L.csdrb_convert_f_s16.argtypes = [vp, vp, lg, vp]
    L.csdrb_fir_decimate_bank_cc.argtypes = [vp, lg, vp, lg, it, it, it, C.POINTER(C.c_float), it, it, vp]
    L.csdrb_fmdemod_quadri_bank_cf.argtypes = [vp, lg, vp, lg, it, it, vp, vp, vp]
    L.csdrb_stream_synchronize.argtypes = [vp]
    L.csdrb_fir_decimate_bank_cc_host.argtypes = [vp, lg, vp, lg, it, it, it, C.POINTER(C.c_float), it, it]
    L.csdrb_fir_decimate_bank_u8_cc.argtypes = [vp, lg, vp, lg, it, it, it, C.POINTER(C.c_float), it, vp]
    L.csdrb_fir_decimate_bank_u8_host.argtypes = [vp, lg, vp, lg, it, it, it, C.POINTER(C.c_float), it, it]
    L.csdrb_host_alloc.argtypes = [C.c_size_t]; L.csdrb_host_alloc.restype = vp
    L.csdrb_host_free.argtypes = [vp]
    sz = C.c_size_t
    L.csdrb_shift_addition_bank_scratch_bytes.argtypes = [it, it, it]; L.csdrb_shift_addition_bank_scratch_bytes.restype = sz
    L.csdrb_shift_addition_bank_cc.argtypes = [vp, lg, vp, lg, it, it, vp, vp, it, vp, sz, vp]
    L.csdrb_decimating_shift_addition_bank_cc.argtypes = [vp, lg, vp, lg, it, it, vp, it, vp, vp, vp, vp]
    L.csdrb_fractional_decimator_bank_scratch_bytes.argtypes = [it, it, C.c_float]; L.csdrb_fractional_decimator_bank_scratch_bytes.restype = sz
    L.csdrb_fractional_decimator_bank_ff.argtypes = [vp, lg, vp, lg, it, it, C.c_float, it, vp, it, vp, vp, sz, vp]
    L.csdrb_fastagc_bank_scratch_bytes.argtypes = [it, it]; L.csdrb_fastagc_bank_scratch_bytes.restype = sz
    L.csdrb_fastagc_bank_ff.argtypes = [vp, lg, vp, lg, it, it, it, C.c_float, vp, vp, vp, sz, vp]
    L.csdrb_fastagc_bank_f_s16.argtypes = [vp, lg, vp, lg, it, it, it, C.c_float, vp, vp, vp, sz, vp]
    L.csdrb_fft_c2c_batch.argtypes = [vp, lg, vp, lg, it, it, it, vp]
    L.csdrb_bandpass_fir_fft_bank_cc.argtypes = [vp, lg, vp, lg, it, it, it, it, vp, lg, vp, vp]
    L.csdrb_ddc_bank_scratch_bytes.argtypes = [it, it, it, it]; L.csdrb_ddc_bank_scratch_bytes.restype = sz
    L.csdrb_ddc_bank.argtypes = [vp, it, it, vp, vp, it, it, it, C.POINTER(C.c_float), it, it, vp, lg, vp, vp, vp, sz, vp]
    L.csdrb_ddc_bank_create.argtypes = [it, C.POINTER(C.c_float), it, C.POINTER(C.c_float), it, it, it]; L.csdrb_ddc_bank_create.restype = vp
    L.csdrb_ddc_bank_destroy.argtypes = [vp]
    L.csdrb_ddc_bank_set_rate.argtypes = [vp, it, C.c_float]
    L.csdrb_ddc_bank_offset.argtypes = [vp]
    L.csdrb_ddc_bank_process.argtypes = [vp, vp, it, vp, lg, vp]
    L.csdrb_ddc_bank_rechunk.argtypes = [vp]
    L.csdrb_fastddc_inv_plan_create.argtypes = [vp, it, vp, it]; L.csdrb_fastddc_inv_plan_create.restype = vp
    L.csdrb_fastddc_inv_plan_run.argtypes = [vp, vp, vp, vp, lg, vp, vp]
    L.csdrb_fastddc_inv_plan_set_channel.argtypes = [vp, it, vp]
    L.csdrb_fastddc_inv_plan_get_state.argtypes = [vp, vp, vp]
    L.csdrb_fastddc_inv_plan_set_state.argtypes = [vp, vp, vp]
    L.csdrb_fastddc_inv_plan_destroy.argtypes = [vp]
    L.csdrb_multi_bank_create.argtypes = [it, C.POINTER(C.c_int), it, C.POINTER(C.c_float), it, C.POINTER(C.c_float), it, it, it, it]; L.csdrb_multi_bank_create.restype = vp
    L.csdrb_multi_bank_destroy.argtypes = [vp]
    L.csdrb_multi_bank_devices.argtypes = [vp]
    L.csdrb_multi_bank_slice.argtypes = [vp, it, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.csdrb_multi_bank_set_rate.argtypes = [vp, it, C.c_float]
    L.csdrb_multi_bank_submit.argtypes = [vp, vp, it, vp, lg]
    L.csdrb_multi_bank_collect.argtypes = [vp, it]
    L.csdrb_multi_bank_process_host.argtypes = [vp, vp, it, vp, lg]
    L.csdrb_apply_window_rows_c.argtypes = [vp, vp, vp, it, lg, vp]
    L.csdrb_logpower_cf.argtypes = [vp, vp, lg, C.c_float, vp]
    L.csdrb_accumulate_power_cf.argtypes = [vp, vp, lg, vp]
    L.csdrb_log_ff.argtypes = [vp, vp, lg, C.c_float, vp]
    L.csdrb_shift_unroll_bank_cc.argtypes = [vp, lg, vp, lg, it, it, vp, vp, vp, lg, it, vp, vp, sz, vp]
    L.precalculate_window.argtypes = [it, it]; L.precalculate_window.restype = C.POINTER(C.c_float)
    L.apply_precalculated_window_c.argtypes = [vp, vp, it, vp]
    L.apply_window_c.argtypes = [vp, vp, it, it]
    L.logpower_cf.argtypes = [vp, vp, it, C.c_float]
    L.accumulate_power_cf.argtypes = [vp, vp, it]
    L.log_ff.argtypes = [vp, vp, it, C.c_float]
    L.shift_unroll_init.argtypes = [C.c_float, it]; L.shift_unroll_init.restype = _Unroll
    L.shift_unroll_cc.argtypes = [vp, vp, it, C.POINTER(_Unroll), C.c_float]; L.shift_unroll_cc.restype = C.c_float
    L.shift_math_cc.argtypes = [vp, vp, it, C.c_float, C.c_float]; L.shift_math_cc.restype = C.c_float
    L.shift_table_init.argtypes = [it]; L.shift_table_init.restype = _Table
    L.shift_table_deinit.argtypes = [_Table]
    L.shift_table_cc.argtypes = [vp, vp, it, C.c_float, _Table, C.c_float]; L.shift_table_cc.restype = C.c_float
    L.csdrb_shift_table_bank_cc.argtypes = [vp, lg, vp, lg, it, it, vp, vp, vp, it, vp, sz, vp]
    L.encode_ima_adpcm_i16_u8.argtypes = [vp, vp, it, _Ima]; L.encode_ima_adpcm_i16_u8.restype = _Ima
    L.csdrb_encode_ima_adpcm_rows_i16_u8.argtypes = [vp, lg, vp, lg, it, it, vp, vp]
    L.csdrb_compress_fft_adpcm_rows_f_u8.argtypes = [vp, lg, vp, lg, it, it, vp]
    L.csdrb_shift_math_bank_scratch_bytes.argtypes = [it, it]; L.csdrb_shift_math_bank_scratch_bytes.restype = sz
    L.csdrb_shift_math_bank_cc.argtypes = [vp, lg, vp, lg, it, it, vp, vp, vp, sz, vp]
    L.shift_addfast_init.argtypes = [C.c_float]; L.shift_addfast_init.restype = _AddFast
    L.shift_addfast_cc.argtypes = [vp, vp, it, C.POINTER(_AddFast), C.c_float]; L.shift_addfast_cc.restype = C.c_float
    L.csdrb_shift_addfast_bank_cc.argtypes = [vp, lg, vp, lg, it, it, vp, vp, it, vp, sz, vp]
    L.csdrb_limit_ff.argtypes = [vp, vp, lg, C.c_float, vp]
    L.csdrb_deemphasis_wfm_bank_ff.argtypes = [vp, lg, vp, lg, it, it, C.c_float, it, vp, vp]
    L.limit_ff.argtypes = [vp, vp, it, C.c_float]
    L.deemphasis_nfm_ff.argtypes = [vp, vp, it, it]
    L.csdrb_deemphasis_nfm_bank_ff.argtypes = [vp, lg, vp, lg, it, it, it, C.c_float, vp]
    L.csdrb_deemphasis_nfm_taps.argtypes = [it, C.POINTER(it)]; L.csdrb_deemphasis_nfm_taps.restype = C.POINTER(C.c_float)
    L.csdrb_fir_valid_bank_ff.argtypes = [vp, lg, vp, lg, it, it, C.POINTER(C.c_float), it, C.c_float, vp]
    L.deemphasis_wfm_ff.argtypes = [vp, vp, it, C.c_float, it, C.c_float]; L.deemphasis_wfm_ff.restype = C.c_float
    L.csdrb_fastddc_fwd_cc.argtypes = [vp, vp, vp, it, it, it, vp]
    L.csdrb_fastddc_inv_bank_scratch_bytes.argtypes = [it, it]; L.csdrb_fastddc_inv_bank_scratch_bytes.restype = sz
    L.csdrb_fastddc_inv_bank_cc.argtypes = [vp, it, vp, vp, it, C.POINTER(FastDDC), vp, vp, vp, lg, vp, vp, sz, vp]
    L.fastddc_init.argtypes = [C.POINTER(FastDDC), C.c_float, it, C.c_float]
    L.decimating_shift_addition_init.argtypes = [C.c_float, it]; L.decimating_shift_addition_init.restype = _Shift
    L.shift_addition_cc.argtypes = [vp, vp, it, _Shift, C.c_float]; L.shift_addition_cc.restype = C.c_float
    L.decimating_shift_addition_cc.argtypes = [vp, vp, it, _Shift, it, _DShiftStatus]; L.decimating_shift_addition_cc.restype = _DShiftStatus
    L.fractional_decimator_ff_init.argtypes = [C.c_float, it, vp, it]; L.fractional_decimator_ff_init.restype = _FracDec
    L.fractional_decimator_ff.argtypes = [vp, vp, it, C.POINTER(_FracDec)]
    L.fastagc_ff.argtypes = [C.POINTER(_FastAgc), vp]
    L.make_fft_c2c.argtypes = [it, vp, vp, it, it]; L.make_fft_c2c.restype = C.POINTER(_Plan)
    L.fft_execute.argtypes = [C.POINTER(_Plan)]
    L.fft_destroy.argtypes = [C.POINTER(_Plan)]
    L.apply_fir_fft_cc.argtypes = [C.POINTER(_Plan), C.POINTER(_Plan), vp, vp, it]
    L.fastddc_inv_cc.argtypes = [vp, vp, C.POINTER(FastDDC), C.POINTER(_Plan), vp, _DShiftStatus]; L.fastddc_inv_cc.restype = _DShiftStatus
    L.fft_swap_sides.argtypes = [vp, it]
    # host-side design helpers (Part A)
    L.firdes_filter_len.argtypes = [C.c_float]
    L.firdes_lowpass_f.argtypes = [C.POINTER(C.c_float), it, C.c_float, it]
    L.firdes_bandpass_c.argtypes = [C.POINTER(_CF), it, C.c_float, C.c_float, it]
    L.shift_addition_init.argtypes = [C.c_float]; L.shift_addition_init.restype = _Shift
    L.next_pow2.argtypes = [it]
    # Part A compute wrappers on host pointers
    L.convert_u8_f.argtypes = [vp, vp, it]
    L.convert_s16_f.argtypes = [vp, vp, it]
    L.convert_f_s16.argtypes = [vp, vp, it]
    L.fir_decimate_cc.argtypes = [vp, vp, it, it, C.POINTER(C.c_float), it]
    L.fmdemod_quadri_cf.argtypes = [vp, vp, it, vp, _CF]; L.fmdemod_quadri_cf.restype = _CF
    _lib = L
    return L


def _check(rc: int, what: str) -> int:
    if rc < 0:
        raise CsdrB200Error(f"{what}: {lib().csdrb_last_error().decode()} (rc={rc})")
    return rc


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def kernel_launches() -> int:
    return int(lib().csdrb_kernel_launches())


# --------------------------------------------------------------------------------------------------
# host-side design helpers (run on the CPU exactly where the reference runs them)
# --------------------------------------------------------------------------------------------------
def firdes_filter_len(transition_bw: float) -> int:
    return int(lib().firdes_filter_len(transition_bw))


def firdes_lowpass_f(length: int, cutoff_rate: float, window: str = "HAMMING") -> np.ndarray:
    t = np.empty(length, np.float32)
    lib().firdes_lowpass_f(_fp(t), length, cutoff_rate, WINDOWS[window])
    return t


def firdes_bandpass_c(length: int, lowcut: float, highcut: float, window: str = "HAMMING") -> np.ndarray:
    t = np.empty(length, np.complex64)
    lib().firdes_bandpass_c(t.ctypes.data_as(C.POINTER(_CF)), length, lowcut, highcut, WINDOWS[window])
    return t


# --------------------------------------------------------------------------------------------------
# device-resident bank API (torch CUDA tensors in, torch CUDA tensors out)
# --------------------------------------------------------------------------------------------------
def _as_cf32_rows(x):
    """Accept [C, N] complex64 or [C, N, 2] float32 CUDA tensors; return (tensor, data_ptr, row stride in samples, C, N)."""
    import torch
    if x.dtype == torch.complex64:
        xr = torch.view_as_real(x)
    elif x.dtype == torch.float32 and x.shape[-1] == 2:
        xr = x
    else:
        raise TypeError("expected complex64 [C,N] or float32 [C,N,2]")
    if xr.dim() == 2:
        xr = xr.unsqueeze(0)
    if not xr.is_cuda:
        raise CsdrB200Error("bank API needs CUDA tensors (no CPU fallback)")
    if xr.stride(2) != 1 or xr.stride(1) != 2:
        raise ValueError("samples must be contiguous within a channel row")
    return xr, xr.data_ptr(), xr.stride(0) // 2, xr.shape[0], xr.shape[1]


def convert_u8_f(x, out=None):
    import torch
    assert x.dtype == torch.uint8 and x.is_cuda and x.is_contiguous()
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device) if out is None else out
    _check(lib().csdrb_convert_u8_f(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "convert_u8_f")
    return out


def convert_s16_f(x, out=None):
    import torch
    assert x.dtype == torch.int16 and x.is_cuda and x.is_contiguous()
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device) if out is None else out
    _check(lib().csdrb_convert_s16_f(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "convert_s16_f")
    return out


def convert_f_s16(x, out=None):
    import torch
    assert x.dtype == torch.float32 and x.is_cuda and x.is_contiguous()
    out = torch.empty(x.shape, dtype=torch.int16, device=x.device) if out is None else out
    _check(lib().csdrb_convert_f_s16(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "convert_f_s16")
    return out


def fir_out_len(input_size: int, decimation: int, taps_length: int) -> int:
    """Outputs of one fir_decimate_cc call (reference libcsdr.c:537-547)."""
    return (input_size - taps_length) // decimation + 1 if input_size >= taps_length else 0


def fir_decimate_bank_cc(x, decimation: int, taps: np.ndarray, out=None, variant: int = -1):
    """C independent cf32 streams [C, N] -> [C, n_out]; all channels share ``taps`` (host float32 array)."""
    import torch
    xr, ptr, stride, ch, n = _as_cf32_rows(x)
    taps = np.ascontiguousarray(taps, np.float32)
    n_out = fir_out_len(n, decimation, taps.size)
    if out is None:
        out = torch.empty((ch, n_out + (n_out & 1)), dtype=torch.complex64, device=xr.device)
    orr, optr, ostride, och, on = _as_cf32_rows(out)
    assert och == ch and on >= n_out
    rc = _check(lib().csdrb_fir_decimate_bank_cc(ptr, stride, optr, ostride, ch, n, decimation, _fp(taps), taps.size, variant, _stream()),
                "fir_decimate_bank_cc")
    assert rc == n_out, (rc, n_out)
    return out[:, :n_out]


class PinnedArray:
    """A page-locked host buffer from csdrb_host_alloc exposed as a numpy array (freed on close/del)."""

    def __init__(self, shape, dtype):
        self.shape = tuple(shape); self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self.ptr = lib().csdrb_host_alloc(nbytes)
        if not self.ptr:
            raise CsdrB200Error(f"csdrb_host_alloc({nbytes}): {lib().csdrb_last_error().decode()}")
        buf = (C.c_char * nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def close(self):
        if getattr(self, "ptr", None):
            self.array = None
            lib().csdrb_host_free(self.ptr); self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fir_decimate_bank_cc_host(x: np.ndarray, decimation: int, taps: np.ndarray, out: np.ndarray | None = None, chunk_channels: int = 0):
    """End-to-end call on HOST arrays: x [C, N] complex64 -> out [C, n_out] complex64 (H2D, kernel, D2H inside)."""
    assert x.dtype == np.complex64 and x.ndim == 2 and x.strides[1] == 8
    taps = np.ascontiguousarray(taps, np.float32)
    ch, n = x.shape
    n_out = fir_out_len(n, decimation, taps.size)
    if out is None:
        out = np.empty((ch, n_out), np.complex64)
    assert out.dtype == np.complex64 and out.shape[0] == ch and out.shape[1] >= n_out and out.strides[1] == 8
    rc = _check(lib().csdrb_fir_decimate_bank_cc_host(x.ctypes.data, x.strides[0] // 8, out.ctypes.data, out.strides[0] // 8, ch, n,
                                                      decimation, _fp(taps), taps.size, chunk_channels), "fir_decimate_bank_cc_host")
    assert rc == n_out
    return out[:, :n_out]


def fir_decimate_bank_u8_cc(x, decimation: int, taps: np.ndarray, out=None):
    """convert_u8_f | fir_decimate_cc fused: x [C, N, 2] uint8 CUDA tensor (interleaved I,Q; rows N samples apart or padded) -> [C, n_out] cf32."""
    import torch
    assert x.dtype == torch.uint8 and x.dim() == 3 and x.shape[2] == 2 and x.stride(2) == 1 and x.stride(1) == 2 and x.stride(0) % 2 == 0
    ch, n = x.shape[0], x.shape[1]
    taps = np.ascontiguousarray(taps, np.float32)
    n_out = fir_out_len(n, decimation, taps.size)
    if out is None:
        out = torch.empty((ch, n_out + (n_out & 1)), dtype=torch.complex64, device=x.device)
    orr, optr, ostride, och, on = _as_cf32_rows(out)
    assert och == ch and on >= n_out
    rc = _check(lib().csdrb_fir_decimate_bank_u8_cc(x.data_ptr(), x.stride(0) // 2, optr, ostride, ch, n, decimation, _fp(taps), taps.size, _stream()),
                "fir_decimate_bank_u8_cc")
    assert rc == n_out, (rc, n_out)
    return out[:, :n_out]


def fir_decimate_bank_u8_host(x: np.ndarray, decimation: int, taps: np.ndarray, out: np.ndarray | None = None, chunk_channels: int = 0):
    """End-to-end call on HOST arrays with rtl_sdr-style input: x [C, N, 2] uint8 -> out [C, n_out] complex64 (2 bytes per sample over PCIe)."""
    assert x.dtype == np.uint8 and x.ndim == 3 and x.shape[2] == 2 and x.strides[2] == 1 and x.strides[1] == 2
    taps = np.ascontiguousarray(taps, np.float32)
    ch, n = x.shape[0], x.shape[1]
    n_out = fir_out_len(n, decimation, taps.size)
    if out is None:
        out = np.empty((ch, n_out), np.complex64)
    assert out.dtype == np.complex64 and out.shape[0] == ch and out.shape[1] >= n_out and out.strides[1] == 8
    rc = _check(lib().csdrb_fir_decimate_bank_u8_host(x.ctypes.data, x.strides[0] // 2, out.ctypes.data, out.strides[0] // 8, ch, n,
                                                      decimation, _fp(taps), taps.size, chunk_channels), "fir_decimate_bank_u8_host")
    assert rc == n_out
    return out[:, :n_out]


def fmdemod_quadri_bank_cf(x, last=None, out=None, return_last: bool = False):
    """[C, N] cf32 -> [C, N] f32.  ``last`` is a [C] complex64 CUDA tensor (sample before each block) or None."""
    import torch
    xr, ptr, stride, ch, n = _as_cf32_rows(x)
    if out is None:
        out = torch.empty((ch, n + (n & 1)), dtype=torch.float32, device=xr.device)
    assert out.stride(1) == 1 and out.shape[0] == ch
    last_out = torch.empty(ch, dtype=torch.complex64, device=xr.device) if return_last else None
    _check(lib().csdrb_fmdemod_quadri_bank_cf(ptr, stride, out.data_ptr(), out.stride(0), ch, n,
                                              last.data_ptr() if last is not None else None,
                                              last_out.data_ptr() if last_out is not None else None, _stream()),
           "fmdemod_quadri_bank_cf")
    res = out[:, :n]
    return (res, last_out) if return_last else res


# --------------------------------------------------------------------------------------------------
# libcsdr drop-in calls on HOST arrays (Part A of the C ABI) -- what the reference's callers bind
# --------------------------------------------------------------------------------------------------
class libcsdr:
    """numpy-in / numpy-out front-end over the libcsdr-named host-pointer entry points."""

    @staticmethod
    def convert_u8_f(x):
        x = np.ascontiguousarray(x, np.uint8); y = np.empty(x.size, np.float32)
        lib().convert_u8_f(x.ctypes.data, y.ctypes.data, x.size); return y

    @staticmethod
    def convert_s16_f(x):
        x = np.ascontiguousarray(x, np.int16); y = np.empty(x.size, np.float32)
        lib().convert_s16_f(x.ctypes.data, y.ctypes.data, x.size); return y

    @staticmethod
    def convert_f_s16(x):
        x = np.ascontiguousarray(x, np.float32); y = np.empty(x.size, np.int16)
        lib().convert_f_s16(x.ctypes.data, y.ctypes.data, x.size); return y

    @staticmethod
    def fir_decimate_cc(x, decimation, taps):
        x = np.ascontiguousarray(x, np.complex64); taps = np.ascontiguousarray(taps, np.float32)
        y = np.empty(max(x.size // decimation + 1, 1), np.complex64)
        n = lib().fir_decimate_cc(x.ctypes.data, y.ctypes.data, x.size, decimation, _fp(taps), taps.size)
        return y[:n].copy()

    @staticmethod
    def shift_addition_cc(x, rate, phase=0.0, chunk=None):
        x = np.ascontiguousarray(x, np.complex64); y = np.empty_like(x); d = lib().shift_addition_init(rate)
        chunk = chunk or max(x.size, 1)
        for s in range(0, x.size, chunk):
            n = min(chunk, x.size - s)
            phase = lib().shift_addition_cc(x[s:].ctypes.data, y[s:].ctypes.data, n, d, phase)
        return y, float(np.float32(phase))

    @staticmethod
    def decimating_shift_addition_cc(x, rate, decimation, remain=0, phase=0.0):
        x = np.ascontiguousarray(x, np.complex64); y = np.empty(x.size // decimation + 2, np.complex64)
        d = lib().decimating_shift_addition_init(rate, decimation)
        st = lib().decimating_shift_addition_cc(x.ctypes.data, y.ctypes.data, x.size, d, decimation, _DShiftStatus(remain, phase, 0))
        return y[:st.output_size].copy(), (st.decimation_remain, st.starting_phase)

    @staticmethod
    def fractional_decimator_ff(x, rate, num_poly_points=12, taps=None, block=None):
        """block=None: one call; else the CLI's block loop with tail re-feeding (csdr.c:1510-1522)."""
        x = np.ascontiguousarray(x, np.float32)
        tp = np.ascontiguousarray(taps, np.float32) if taps is not None else None
        d = lib().fractional_decimator_ff_init(rate, num_poly_points, tp.ctypes.data if tp is not None else None, tp.size if tp is not None else 0)
        if block is None:
            out = np.empty(int(x.size / max(rate, 1.0)) + 16, np.float32)
            lib().fractional_decimator_ff(x.ctypes.data, out.ctypes.data, x.size, C.byref(d))
            return out[:d.output_size].copy()
        buf = np.zeros(block, np.float32); out = np.empty(block, np.float32); outs = []; pos = 0
        while True:
            if d.input_processed == 0:
                need, keep = block, 0
            else:
                need = d.input_processed; keep = block - need
                buf[:keep] = buf[need:].copy()
            if pos + need > x.size:
                break
            buf[keep:] = x[pos:pos + need]; pos += need
            if d.input_processed == 0:
                d.input_processed = block
            lib().fractional_decimator_ff(buf.ctypes.data, out.ctypes.data, block, C.byref(d))
            outs.append(out[:d.output_size].copy())
        return np.concatenate(outs) if outs else np.zeros(0, np.float32)

    @staticmethod
    def fastagc_ff(x, block=1024, reference=1.0):
        x = np.ascontiguousarray(x, np.float32); nblk = x.size // block
        bufs = [np.zeros(block, np.float32) for _ in range(3)]
        byaddr = {b.ctypes.data: b for b in bufs}
        st = _FastAgc(bufs[0].ctypes.data, bufs[1].ctypes.data, bufs[2].ctypes.data, 0, 0, block, reference, 0)
        y = np.empty(nblk * block, np.float32)
        for b in range(nblk):
            byaddr[st.buffer_input][:] = x[b * block:(b + 1) * block]
            lib().fastagc_ff(C.byref(st), y[b * block:].ctypes.data)
        return y

    @staticmethod
    def deemphasis_wfm_ff(x, tau, sample_rate, last=0.0, block=None):
        x = np.ascontiguousarray(x, np.float32); y = np.empty_like(x); block = block or max(x.size, 1)
        for s0 in range(0, x.size, block):
            n = min(block, x.size - s0)
            last = lib().deemphasis_wfm_ff(x[s0:].ctypes.data, y[s0:].ctypes.data, n, tau, sample_rate, last)
        return y, float(np.float32(last))

    @staticmethod
    def limit_ff(x, max_amplitude=1.0):
        x = np.ascontiguousarray(x, np.float32); y = np.empty_like(x)
        lib().limit_ff(x.ctypes.data, y.ctypes.data, x.size, max_amplitude); return y

    @staticmethod
    def deemphasis_nfm_ff(x, sample_rate):
        x = np.ascontiguousarray(x, np.float32); y = np.zeros_like(x)
        n = lib().deemphasis_nfm_ff(x.ctypes.data, y.ctypes.data, x.size, sample_rate)
        return y[:n].copy()

    @staticmethod
    def precalculate_window(size, window="HAMMING"):
        p = lib().precalculate_window(size, WINDOWS[window]); return np.ctypeslib.as_array(p, shape=(size,)).copy()

    @staticmethod
    def apply_window_c(x, window="HAMMING"):
        x = np.ascontiguousarray(x, np.complex64); y = np.empty_like(x)
        lib().apply_window_c(x.ctypes.data, y.ctypes.data, x.size, WINDOWS[window]); return y

    @staticmethod
    def logpower_cf(x, add_db=0.0):
        x = np.ascontiguousarray(x, np.complex64); y = np.empty(x.size, np.float32)
        lib().logpower_cf(x.ctypes.data, y.ctypes.data, x.size, add_db); return y

    @staticmethod
    def logaveragepower_cf(x, add_db, fft_size, avgnumber):
        x = np.ascontiguousarray(x, np.complex64); out = []
        adj = np.float32(np.float32(add_db) - np.float32(10.0 * np.log10(avgnumber)))
        for b in range(x.size // (fft_size * avgnumber)):
            acc = np.zeros(fft_size, np.float32)
            for n in range(avgnumber):
                seg = x[(b * avgnumber + n) * fft_size:(b * avgnumber + n + 1) * fft_size]
                lib().accumulate_power_cf(seg.ctypes.data, acc.ctypes.data, fft_size)
            y = np.empty(fft_size, np.float32); lib().log_ff(acc.ctypes.data, y.ctypes.data, fft_size, float(adj)); out.append(y)
        return np.concatenate(out) if out else np.zeros(0, np.float32)

    @staticmethod
    def shift_unroll_cc(x, rate, phase=0.0, size=1024):
        x = np.ascontiguousarray(x, np.complex64); y = np.empty_like(x)
        d = lib().shift_unroll_init(rate, size)
        for s0 in range(0, x.size, size):
            n = min(size, x.size - s0)
            phase = lib().shift_unroll_cc(x[s0:].ctypes.data, y[s0:].ctypes.data, n, C.byref(d), phase)
        return y, float(np.float32(phase))

    @staticmethod
    def encode_ima_adpcm_i16_u8(x, index=0, previous=0):
        x = np.ascontiguousarray(x, np.int16); y = np.empty(x.size // 2, np.uint8)
        st = lib().encode_ima_adpcm_i16_u8(x.ctypes.data, y.ctypes.data, x.size, _Ima(index, previous))
        return y, (st.index, st.previousValue)

    @staticmethod
    def shift_table_init(size=65536):
        d = lib().shift_table_init(size); t = np.ctypeslib.as_array(d.table, shape=(size,)).copy(); lib().shift_table_deinit(d); return t

    @staticmethod
    def shift_table_cc(x, rate, table, phase=0.0, chunk=None):
        x = np.ascontiguousarray(x, np.complex64); y = np.empty_like(x); table = np.ascontiguousarray(table, np.float32); chunk = chunk or max(x.size, 1)
        d = _Table(table.ctypes.data_as(C.POINTER(C.c_float)), table.size)
        for s0 in range(0, x.size, chunk):
            n = min(chunk, x.size - s0)
            phase = lib().shift_table_cc(x[s0:].ctypes.data, y[s0:].ctypes.data, n, rate, d, phase)
        return y, float(np.float32(phase))

    @staticmethod
    def shift_math_cc(x, rate, phase=0.0, chunk=None):
        x = np.ascontiguousarray(x, np.complex64); y = np.empty_like(x); chunk = chunk or max(x.size, 1)
        for s0 in range(0, x.size, chunk):
            n = min(chunk, x.size - s0)
            phase = lib().shift_math_cc(x[s0:].ctypes.data, y[s0:].ctypes.data, n, rate, phase)
        return y, float(np.float32(phase))

    @staticmethod
    def shift_addfast_cc(x, rate=None, phase=0.0, chunk=1024, steps=None):
        """calls of <= chunk samples like csdr.c:781-791; `steps` (9 floats: dsin[4], dcos[4], increment) overrides shift_addfast_init(rate);
        samples a call leaves untouched (n % 4) come back as 0"""
        x = np.ascontiguousarray(x, np.complex64); y = np.zeros_like(x); chunk = chunk or max(x.size, 1)
        d = lib().shift_addfast_init(rate) if steps is None else _AddFast((C.c_float * 4)(*steps[0:4]), (C.c_float * 4)(*steps[4:8]), float(steps[8]))
        for s0 in range(0, x.size, chunk):
            n = min(chunk, x.size - s0)
            phase = lib().shift_addfast_cc(x[s0:].ctypes.data, y[s0:].ctypes.data, n, C.byref(d), phase)
        return y, float(np.float32(phase))

    @staticmethod
    def dft(x, forward=True):
        x = np.ascontiguousarray(x, np.complex64).copy(); y = np.empty_like(x)
        pl = lib().make_fft_c2c(x.size, x.ctypes.data, y.ctypes.data, 1 if forward else 0, 0)
        lib().fft_execute(pl); lib().fft_destroy(pl); return y

    @staticmethod
    def bandpass_fir_fft_cc(x, lo, hi, bw, window="HAMMING"):
        """The CLI block loop of csdr.c:1833-1883 over the libcsdr-named entry points (complete blocks only)."""
        x = np.ascontiguousarray(x, np.complex64)
        T, N, isz, ov = bandpass_geometry(bw)
        taps = np.zeros(N, np.complex64); taps[:T] = firdes_bandpass_c(T, lo, hi, window)
        taps_fft = libcsdr.dft(taps)
        inp = np.zeros(N, np.complex64); spec = np.empty(N, np.complex64); ospec = np.empty(N, np.complex64)
        o = [np.zeros(N, np.complex64), np.zeros(N, np.complex64)]
        pf = lib().make_fft_c2c(N, inp.ctypes.data, spec.ctypes.data, 1, 0)
        pi = [lib().make_fft_c2c(N, ospec.ctypes.data, o[k].ctypes.data, 0, 0) for k in range(2)]
        out = []
        for b in range(x.size // isz):
            inp[:isz] = x[b * isz:(b + 1) * isz]
            cur, prev = (1, 0) if b & 1 else (0, 1)
            tail = o[prev][isz:]
            lib().apply_fir_fft_cc(pf, pi[cur], taps_fft.ctypes.data, tail.ctypes.data, ov)
            out.append(o[cur][:isz].copy())
        lib().fft_destroy(pf); [lib().fft_destroy(p) for p in pi]
        return np.concatenate(out) if out else np.zeros(0, np.complex64)

    @staticmethod
    def fastddc_inv(spectra, bw, decimation, shift, window="HAMMING"):
        """csdr.c:2335-2371 over the libcsdr-named entry points."""
        ddc = fastddc_init(bw, decimation, shift)
        hb = np.float32(0.5 / decimation); sh = np.float32(shift)
        taps = np.zeros(ddc.fft_size, np.complex64)
        taps[:ddc.taps_length] = firdes_bandpass_c(ddc.taps_length, float(-sh - hb), float(-sh + hb), window)
        tf = libcsdr.dft(taps); lib().fft_swap_sides(tf.ctypes.data, ddc.fft_size)
        st = _DShiftStatus(0, 0.0, 0); out = []
        for sp in spectra:
            sp = np.ascontiguousarray(sp, np.complex64).copy(); y = np.empty(ddc.post_input_size, np.complex64)
            st = lib().fastddc_inv_cc(sp.ctypes.data, y.ctypes.data, C.byref(ddc), None, tf.ctypes.data, st)
            out.append(y[:st.output_size].copy())
        return np.concatenate(out) if out else np.zeros(0, np.complex64)

    @staticmethod
    def fmdemod_quadri_cf(x, last=0j):
        x = np.ascontiguousarray(x, np.complex64); y = np.empty(x.size, np.float32)
        r = lib().fmdemod_quadri_cf(x.ctypes.data, y.ctypes.data, x.size, None, _CF(np.float32(last.real), np.float32(last.imag)))
        return y, complex(r.i, r.q)


# --------------------------------------------------------------------------------------------------
# K2 / K5 / K6 / K7 / K8 / K9 bank calls (torch CUDA tensors)
# --------------------------------------------------------------------------------------------------
def shift_addition_init(rate: float):
    d = lib().shift_addition_init(rate)
    return (d.sindelta, d.cosdelta, d.rate)


def fastddc_init(transition_bw: float, decimation: int, shift_rate: float) -> FastDDC:
    d = FastDDC()
    if lib().fastddc_init(C.byref(d), transition_bw, decimation, shift_rate):
        raise CsdrB200Error("fastddc_init: fft_size <= 2")
    return d


def _scratch(nbytes: int, device):
    import torch
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def shift_addition_bank_cc(x, rates, phases=None, chunk: int = 1024, out=None):
    """x: [N] (one shared wideband stream) or [C, N] complex64; rates: C floats.  Returns (y [C,N], new phases [C] tensor)."""
    import torch
    rates = np.atleast_1d(np.asarray(rates, np.float32))
    ch = rates.size
    shared = (x.dim() == 1) if x.dtype == torch.complex64 else (x.dim() == 2)
    xr, ptr, stride, xc, n = _as_cf32_rows(x)
    if shared:
        stride = 0
    else:
        assert xc == ch
    params = np.array([shift_addition_init(float(r)) for r in rates], np.float32)           # host: bit-exact deltas
    d_params = torch.from_numpy(params).to(xr.device)
    d_phase = torch.zeros(ch, dtype=torch.float32, device=xr.device) if phases is None else phases.clone()
    if out is None:
        out = torch.empty((ch, n), dtype=torch.complex64, device=xr.device)
    sb = lib().csdrb_shift_addition_bank_scratch_bytes(ch, n, chunk)
    scratch = _scratch(sb, xr.device)
    _check(lib().csdrb_shift_addition_bank_cc(ptr, stride, out.data_ptr(), out.stride(0), ch, n, d_params.data_ptr(), d_phase.data_ptr(), chunk,
                                              scratch.data_ptr(), scratch.numel(), _stream()), "shift_addition_bank_cc")
    return out, d_phase


def encode_ima_adpcm_rows_i16_u8(x, state=None):
    """x [R, N] int16 -> (bytes [R, N/2] uint8, state [R, 2] int32 = index, previousValue carried per row)"""
    import torch
    assert x.dtype == torch.int16 and x.is_cuda and x.dim() == 2 and x.stride(1) == 1
    rows, n = x.shape
    out = torch.empty((rows, n // 2), dtype=torch.uint8, device=x.device)
    state = torch.zeros((rows, 2), dtype=torch.int32, device=x.device) if state is None else state
    _check(lib().csdrb_encode_ima_adpcm_rows_i16_u8(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), rows, n, state.data_ptr(), _stream()), "encode_ima_adpcm_rows")
    return out, state


def compress_fft_adpcm_rows_f_u8(db):
    """db [R, fft_size] float32 (dB) -> [R, (fft_size + 10) / 2] uint8: one fresh ADPCM encoder per waterfall line (csdr.c:1745-1767)"""
    import torch
    assert db.dtype == torch.float32 and db.is_cuda and db.dim() == 2 and db.stride(1) == 1
    rows, n = db.shape
    out = torch.empty((rows, (n + 10) // 2), dtype=torch.uint8, device=db.device)
    _check(lib().csdrb_compress_fft_adpcm_rows_f_u8(db.data_ptr(), db.stride(0), out.data_ptr(), out.stride(0), rows, n, _stream()), "compress_fft_adpcm_rows")
    return out


def shift_math_bank_cc(x, rates, phases=None, out=None):
    """x: [N] (one shared wideband stream) or [C, N] complex64; returns (y [C, N], new phases [C])."""
    import torch
    rates = np.atleast_1d(np.asarray(rates, np.float32)); ch = rates.size
    shared = (x.dim() == 1) if x.dtype == torch.complex64 else (x.dim() == 2)
    xr, ptr, stride, xc, n = _as_cf32_rows(x)
    if shared:
        stride = 0
    else:
        assert xc == ch
    d_rates = torch.from_numpy(rates).to(xr.device)
    d_phase = torch.zeros(ch, dtype=torch.float32, device=xr.device) if phases is None else phases.clone()
    if out is None:
        out = torch.empty((ch, n), dtype=torch.complex64, device=xr.device)
    scratch = _scratch(lib().csdrb_shift_math_bank_scratch_bytes(ch, n), xr.device)
    _check(lib().csdrb_shift_math_bank_cc(ptr, stride, out.data_ptr(), out.stride(0), ch, n, d_rates.data_ptr(), d_phase.data_ptr(),
                                          scratch.data_ptr(), scratch.numel(), _stream()), "shift_math_bank_cc")
    return out, d_phase


def shift_table_bank_cc(x, rates, table, phases=None, out=None):
    """x: [N] (one shared wideband stream) or [C, N] complex64; table: quarter-wave sine table (numpy or CUDA tensor); returns (y [C, N], new phases [C])."""
    import torch
    rates = np.atleast_1d(np.asarray(rates, np.float32)); ch = rates.size
    shared = (x.dim() == 1) if x.dtype == torch.complex64 else (x.dim() == 2)
    xr, ptr, stride, xc, n = _as_cf32_rows(x)
    if shared:
        stride = 0
    else:
        assert xc == ch
    d_table = table if isinstance(table, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(table, np.float32)).to(xr.device)
    d_rates = torch.from_numpy(rates).to(xr.device)
    d_phase = torch.zeros(ch, dtype=torch.float32, device=xr.device) if phases is None else phases.clone()
    if out is None:
        out = torch.empty((ch, n), dtype=torch.complex64, device=xr.device)
    scratch = _scratch(lib().csdrb_shift_math_bank_scratch_bytes(ch, n), xr.device)
    _check(lib().csdrb_shift_table_bank_cc(ptr, stride, out.data_ptr(), out.stride(0), ch, n, d_rates.data_ptr(), d_phase.data_ptr(), d_table.data_ptr(),
                                           d_table.numel(), scratch.data_ptr(), scratch.numel(), _stream()), "shift_table_bank_cc")
    return out, d_phase


def shift_addfast_init(rate: float) -> np.ndarray:
    """9 floats: dsin[4], dcos[4], phase_increment (host; shift_addfast_data_t member order)"""
    d = lib().shift_addfast_init(rate)
    return np.array(list(d.dsin) + list(d.dcos) + [d.phase_increment], np.float32)


def shift_addfast_bank_cc(x, rates=None, phases=None, chunk: int = 1024, out=None, steps=None):
    """x: [N] (one shared wideband stream) or [C, N] complex64; one reference call per `chunk` samples.  `steps` [C, 9] overrides the
    per-channel shift_addfast_init(rate).  Samples a call does not touch (its n % 4 tail) are 0.  Returns (y [C, N], new phases [C])."""
    import torch
    params = np.stack([shift_addfast_init(float(r)) for r in np.atleast_1d(np.asarray(rates, np.float32))]) if steps is None \
        else np.ascontiguousarray(steps, np.float32).reshape(-1, 9)
    ch = params.shape[0]
    shared = (x.dim() == 1) if x.dtype == torch.complex64 else (x.dim() == 2)
    xr, ptr, stride, xc, n = _as_cf32_rows(x)
    if shared:
        stride = 0
    else:
        assert xc == ch
    d_params = torch.from_numpy(params).to(xr.device)
    d_phase = torch.zeros(ch, dtype=torch.float32, device=xr.device) if phases is None else phases.clone()
    if out is None:
        out = torch.zeros((ch, n), dtype=torch.complex64, device=xr.device)
    scratch = _scratch(lib().csdrb_shift_addition_bank_scratch_bytes(ch, n, chunk), xr.device)
    _check(lib().csdrb_shift_addfast_bank_cc(ptr, stride, out.data_ptr(), out.stride(0), ch, n, d_params.data_ptr(), d_phase.data_ptr(), chunk,
                                             scratch.data_ptr(), scratch.numel(), _stream()), "shift_addfast_bank_cc")
    return out, d_phase


def fractional_decimator_bank_ff(x, rate: float, num_poly_points: int = 12, taps=None, where=None):
    """x [C, N] float32 -> (y [C, cap] float32, state tensor [C,3] int32-view (where bits, input_processed, output_size))."""
    import torch
    assert x.dtype == torch.float32 and x.is_cuda and x.dim() == 2 and x.stride(1) == 1
    ch, n = x.shape
    cap = int(n / max(rate, 1.0)) + 8
    out = torch.empty((ch, cap), dtype=torch.float32, device=x.device)
    state = torch.zeros((ch, 3), dtype=torch.int32, device=x.device)
    w0 = np.float32(num_poly_points // 2 - 1 if where is None else 0)            # where = -xifirst at init (libcsdr.c:736)
    if where is None:
        state[:, 0] = int(np.array([w0], np.float32).view(np.int32)[0])
    else:
        state[:, 0] = torch.as_tensor(np.asarray(where, np.float32).view(np.int32), device=x.device)
    d_taps = torch.from_numpy(np.ascontiguousarray(taps, np.float32)).to(x.device) if taps is not None else None
    sb = lib().csdrb_fractional_decimator_bank_scratch_bytes(ch, n, rate)
    scratch = _scratch(sb, x.device)
    _check(lib().csdrb_fractional_decimator_bank_ff(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), ch, n, rate, num_poly_points,
                                                    d_taps.data_ptr() if d_taps is not None else None, d_taps.numel() if d_taps is not None else 0,
                                                    state.data_ptr(), scratch.data_ptr(), scratch.numel(), _stream()), "fractional_decimator_bank_ff")
    return out, state


def fastagc_bank_ff(x, block: int = 1024, reference: float = 1.0, state=None, hist=None):
    """x [C, nblocks*block] float32 -> y same shape (two blocks of latency); returns (y, state [C,3] f32, hist [C,2,block])."""
    import torch
    assert x.dtype == torch.float32 and x.is_cuda and x.dim() == 2 and x.stride(1) == 1
    ch, n = x.shape
    nblocks = n // block
    out = torch.empty((ch, nblocks * block), dtype=torch.float32, device=x.device)
    state = torch.zeros((ch, 3), dtype=torch.float32, device=x.device) if state is None else state
    hist = torch.zeros((ch, 2, block), dtype=torch.float32, device=x.device) if hist is None else hist
    scratch = _scratch(lib().csdrb_fastagc_bank_scratch_bytes(ch, nblocks), x.device)
    _check(lib().csdrb_fastagc_bank_ff(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), ch, block, nblocks, reference,
                                       state.data_ptr(), hist.data_ptr(), scratch.data_ptr(), scratch.numel(), _stream()), "fastagc_bank_ff")
    return out, state, hist


def fastagc_bank_f_s16(x, block: int = 1024, reference: float = 1.0, state=None, hist=None):
    """fastagc_ff | convert_f_s16 in one pass: x [C, nblocks*block] float32 -> int16 of the same shape; returns (y, state, hist)."""
    import torch
    assert x.dtype == torch.float32 and x.is_cuda and x.dim() == 2 and x.stride(1) == 1
    ch, n = x.shape
    nblocks = n // block
    out = torch.empty((ch, nblocks * block), dtype=torch.int16, device=x.device)
    state = torch.zeros((ch, 3), dtype=torch.float32, device=x.device) if state is None else state
    hist = torch.zeros((ch, 2, block), dtype=torch.float32, device=x.device) if hist is None else hist
    scratch = _scratch(lib().csdrb_fastagc_bank_scratch_bytes(ch, nblocks), x.device)
    _check(lib().csdrb_fastagc_bank_f_s16(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), ch, block, nblocks, reference,
                                          state.data_ptr(), hist.data_ptr(), scratch.data_ptr(), scratch.numel(), _stream()), "fastagc_bank_f_s16")
    return out, state, hist


def fft_c2c(x, inverse: bool = False):
    """Batched unnormalised DFT along the last axis of a [B, N] (or [N]) complex64 CUDA tensor."""
    import torch
    xr, ptr, stride, b, n = _as_cf32_rows(x)
    out = torch.empty((b, n), dtype=torch.complex64, device=xr.device)
    _check(lib().csdrb_fft_c2c_batch(ptr, stride, out.data_ptr(), out.stride(0), n, b, 1 if inverse else 0, _stream()), "fft_c2c")
    return out if x.dim() > 1 or x.dtype != torch.complex64 else out[0]


def bandpass_geometry(transition_bw: float):
    """(taps_length, fft_size, input_size, overlap) exactly as csdr.c:1833-1838 sizes them."""
    T = firdes_filter_len(transition_bw)
    N = int(lib().next_pow2(T))
    if N - T < 200:
        N <<= 1
    return T, N, N - T + 1, T - 1


def bandpass_taps_fft(low_cut: float, high_cut: float, transition_bw: float, window: str = "HAMMING", device="cuda"):
    import torch
    T, N, _, _ = bandpass_geometry(transition_bw)
    taps = np.zeros(N, np.complex64)
    taps[:T] = firdes_bandpass_c(T, low_cut, high_cut, window)
    return fft_c2c(torch.from_numpy(taps).to(device))


def bandpass_fir_fft_bank_cc(x, taps_fft, input_size: int, tail=None):
    """x [C, nblocks*input_size] complex64 -> y same shape; taps_fft [N] (shared) or [C, N]; tail [C, N] carried state."""
    import torch
    xr, ptr, stride, ch, n = _as_cf32_rows(x)
    N = taps_fft.shape[-1]
    nblocks = n // input_size
    out = torch.empty((ch, nblocks * input_size), dtype=torch.complex64, device=xr.device)
    tail = torch.zeros((ch, N), dtype=torch.complex64, device=xr.device) if tail is None else tail
    tstride = 0 if taps_fft.dim() == 1 else taps_fft.stride(0)
    _check(lib().csdrb_bandpass_fir_fft_bank_cc(ptr, stride, out.data_ptr(), out.stride(0), ch, N, input_size, nblocks, taps_fft.data_ptr(), tstride,
                                                tail.data_ptr(), _stream()), "bandpass_fir_fft_bank_cc")
    return out, tail


def fastddc_fwd_cc(x, ddc: FastDDC, overlap=None):
    """x [nblocks*input_size] complex64 -> spectra [nblocks, fft_size]; overlap [fft_size-input_size] carried state."""
    import torch
    assert x.dtype == torch.complex64 and x.is_cuda and x.dim() == 1
    nblocks = x.numel() // ddc.input_size
    spectra = torch.empty((nblocks, ddc.fft_size), dtype=torch.complex64, device=x.device)
    overlap = torch.zeros(ddc.overlap_length, dtype=torch.complex64, device=x.device) if overlap is None else overlap
    _check(lib().csdrb_fastddc_fwd_cc(x.data_ptr(), spectra.data_ptr(), overlap.data_ptr(), ddc.fft_size, ddc.input_size, nblocks, _stream()), "fastddc_fwd_cc")
    return spectra, overlap


def fastddc_make_taps_fft(ddc: FastDDC, shift_rate: float, decimation: int, window: str = "HAMMING", device="cuda"):
    """csdr.c:2342-2351: bandpass taps at -shift -+ 0.5/decimation, zero pad, forward FFT, swap sides."""
    import torch
    hb = np.float32(0.5 / decimation); sh = np.float32(shift_rate)
    taps = np.zeros(ddc.fft_size, np.complex64)
    taps[:ddc.taps_length] = firdes_bandpass_c(ddc.taps_length, float(-sh - hb), float(-sh + hb), window)
    tf = fft_c2c(torch.from_numpy(taps).to(device))
    return torch.roll(tf, ddc.fft_size // 2)


def fastddc_inv_bank_cc(spectra, shifts, decimation: int, transition_bw: float, window: str = "HAMMING", state=None):
    """All channels (one per entry of ``shifts``) consume the same spectra [nblocks, fft_size].
    Returns (out [C, nblocks*post_input_size/post_decimation + 1], counts [C] int32, state dict)."""
    import torch
    dev = spectra.device
    nblocks = spectra.shape[0]
    ch = len(shifts)
    if state is not None:
        g = state["geometry"]                                         # per-channel host work happens once, at bank creation
    else:
        ddcs = [fastddc_init(transition_bw, decimation, float(s)) for s in shifts]
        g = ddcs[0]
    if state is None:
        taps_fft = torch.stack([fastddc_make_taps_fft(d, float(s), decimation, window, dev) for d, s in zip(ddcs, shifts)]).contiguous()
        chan = np.zeros((ch, 4), np.float32)
        chan.view(np.int32)[:, 0] = [d.offsetbin for d in ddcs]
        chan[:, 1] = [d.dsadata.sindelta for d in ddcs]; chan[:, 2] = [d.dsadata.cosdelta for d in ddcs]; chan[:, 3] = [d.dsadata.rate for d in ddcs]
        state = {"taps_fft": taps_fft, "chan": torch.from_numpy(chan).to(dev), "remain": torch.zeros(ch, dtype=torch.int32, device=dev),
                 "phase": torch.zeros(ch, dtype=torch.float32, device=dev), "geometry": g}
    per_block = g.post_input_size // g.post_decimation + 1
    key = ("buffers", nblocks)
    if key not in state:
        state[key] = (torch.empty((ch, nblocks * per_block + 2), dtype=torch.complex64, device=dev), torch.zeros(ch, dtype=torch.int32, device=dev),
                      _scratch(lib().csdrb_fastddc_inv_bank_scratch_bytes(ch, nblocks), dev))
    out, counts, scratch = state[key]
    _check(lib().csdrb_fastddc_inv_bank_cc(spectra.data_ptr(), nblocks, state["taps_fft"].data_ptr(), state["chan"].data_ptr(), ch, C.byref(g),
                                           state["remain"].data_ptr(), state["phase"].data_ptr(), out.data_ptr(), out.stride(0), counts.data_ptr(),
                                           scratch.data_ptr(), scratch.numel(), _stream()), "fastddc_inv_bank_cc")
    return out, counts, state


def _fastddc_chan_rows(ddcs) -> np.ndarray:
    chan = np.zeros((len(ddcs), 4), np.float32)                          # csdrb_fastddc_chan_t {int offsetbin; float sindelta, cosdelta, rate}
    chan.view(np.int32)[:, 0] = [d.offsetbin for d in ddcs]
    chan[:, 1] = [d.dsadata.sindelta for d in ddcs]; chan[:, 2] = [d.dsadata.cosdelta for d in ddcs]; chan[:, 3] = [d.dsadata.rate for d in ddcs]
    return chan


class FastddcInvPlan:
    """csdrb_fastddc_inv_plan_*: the fastddc inverse bank with the carried post-shift state inside and a fixed number of blocks per run; the state chain and
    the phasors of run k+1 are prepared while run k executes.  Same outputs as fastddc_inv_bank_cc, bit for bit."""

    def __init__(self, shifts, decimation: int, transition_bw: float, nblocks: int, window: str = "HAMMING", device="cuda"):
        import torch
        self.shifts = [float(s) for s in shifts]
        self.decimation, self.transition_bw, self.window, self.nblocks, self.device = decimation, transition_bw, window, nblocks, device
        ddcs = [fastddc_init(transition_bw, decimation, s) for s in self.shifts]
        self.geometry = ddcs[0]
        self.channels = len(ddcs)
        self.taps_fft = torch.stack([fastddc_make_taps_fft(d, s, decimation, window, device) for d, s in zip(ddcs, self.shifts)]).contiguous()
        chan = _fastddc_chan_rows(ddcs)
        self.h = lib().csdrb_fastddc_inv_plan_create(chan.ctypes.data, self.channels, C.addressof(self.geometry), nblocks)
        if not self.h:
            raise CsdrB200Error(f"csdrb_fastddc_inv_plan_create: {lib().csdrb_last_error().decode()}")
        per_block = self.geometry.post_input_size // self.geometry.post_decimation + 1
        self.out = torch.empty((self.channels, nblocks * per_block + 2), dtype=torch.complex64, device=device)
        self.counts = torch.zeros(self.channels, dtype=torch.int32, device=device)

    def run(self, spectra, out=None):
        """spectra [nblocks, fft_size] complex64 (fastddc_fwd_cc) -> (out [C, ...], counts [C] int32 on the device)"""
        assert spectra.shape == (self.nblocks, self.geometry.fft_size) and spectra.is_contiguous()
        out = self.out if out is None else out
        _check(lib().csdrb_fastddc_inv_plan_run(self.h, spectra.data_ptr(), self.taps_fft.data_ptr(), out.data_ptr(), out.stride(0), self.counts.data_ptr(), _stream()),
               "fastddc_inv_plan_run")
        return out, self.counts

    def set_shift(self, channel: int, shift: float):
        """retune one channel from the next run on (csdr.c:2342-2351 redone for that channel)"""
        d = fastddc_init(self.transition_bw, self.decimation, float(shift))
        self.taps_fft[channel].copy_(fastddc_make_taps_fft(d, float(shift), self.decimation, self.window, self.device))
        row = _fastddc_chan_rows([d])
        _check(lib().csdrb_fastddc_inv_plan_set_channel(self.h, channel, row.ctypes.data), "fastddc_inv_plan_set_channel")
        self.shifts[channel] = float(shift)

    def state(self):
        remain = np.zeros(self.channels, np.int32); phase = np.zeros(self.channels, np.float32)
        _check(lib().csdrb_fastddc_inv_plan_get_state(self.h, remain.ctypes.data, phase.ctypes.data), "fastddc_inv_plan_get_state")
        return remain, phase

    def set_state(self, remain, phase):
        remain = np.ascontiguousarray(remain, np.int32); phase = np.ascontiguousarray(phase, np.float32)
        _check(lib().csdrb_fastddc_inv_plan_set_state(self.h, remain.ctypes.data, phase.ctypes.data), "fastddc_inv_plan_set_state")

    def close(self):
        if getattr(self, "h", None):
            lib().csdrb_fastddc_inv_plan_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ddc_bank(wide, rates, decimation: int, taps: np.ndarray, demod: bool = True, chunk: int = 1024, offset: int = 0,
             phases=None, last=None, out=None):
    """Fused shared-input bank: one wideband block [N] complex64 -> per channel shift | fir_decimate | (fmdemod).
    Returns (out [C, n_out], new chunk-start phases [C], last baseband samples [C] or None)."""
    import torch
    assert wide.dtype == torch.complex64 and wide.is_cuda and wide.dim() == 1
    rates = np.atleast_1d(np.asarray(rates, np.float32)); ch = rates.size
    taps = np.ascontiguousarray(taps, np.float32)
    n = wide.numel()
    n_out = fir_out_len(n, decimation, taps.size)
    dev = wide.device
    params = torch.from_numpy(np.array([shift_addition_init(float(r)) for r in rates], np.float32)).to(dev)
    d_phase = torch.zeros(ch, dtype=torch.float32, device=dev) if phases is None else phases.clone()
    stride = n_out + (n_out & 1)
    if out is None:
        out = torch.empty((ch, stride), dtype=torch.float32 if demod else torch.complex64, device=dev)
    last_out = torch.empty(ch, dtype=torch.complex64, device=dev) if demod else None
    scratch = _scratch(lib().csdrb_ddc_bank_scratch_bytes(ch, n, chunk, offset), dev)
    rc = _check(lib().csdrb_ddc_bank(wide.data_ptr(), n, ch, params.data_ptr(), d_phase.data_ptr(), chunk, offset, decimation, _fp(taps), taps.size,
                                     1 if demod else 0, out.data_ptr(), out.stride(0), last.data_ptr() if last is not None else None,
                                     last_out.data_ptr() if last_out is not None else None, scratch.data_ptr(), scratch.numel(), _stream()), "ddc_bank")
    assert rc == n_out
    return out[:, :n_out], d_phase, last_out


def limit_ff(x, max_amplitude: float = 1.0, out=None):
    import torch
    assert x.dtype == torch.float32 and x.is_cuda and x.is_contiguous()
    out = torch.empty_like(x) if out is None else out
    _check(lib().csdrb_limit_ff(x.data_ptr(), out.data_ptr(), x.numel(), max_amplitude, _stream()), "limit_ff")
    return out


def deemphasis_wfm_bank_ff(x, tau: float, sample_rate: int, last=None, out=None):
    """x [C, N] float32 -> (y [C, N], carried last outputs [C])."""
    import torch
    assert x.dtype == torch.float32 and x.is_cuda and x.dim() == 2 and x.stride(1) == 1
    ch, n = x.shape
    out = torch.empty((ch, n), dtype=torch.float32, device=x.device) if out is None else out
    last = torch.zeros(ch, dtype=torch.float32, device=x.device) if last is None else last.clone()
    _check(lib().csdrb_deemphasis_wfm_bank_ff(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), ch, n, tau, sample_rate, last.data_ptr(), _stream()),
           "deemphasis_wfm_bank_ff")
    return out, last


def deemphasis_nfm_taps(sample_rate: int):
    """the fixed FIR deemphasis_nfm_ff uses at this rate (host table of the library), or None"""
    n = C.c_int(0)
    p = lib().csdrb_deemphasis_nfm_taps(sample_rate, C.byref(n))
    return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else None


def deemphasis_nfm_bank_ff(x, sample_rate: int, limit_max: float = 0.0, out=None):
    """x [C, N] float32 -> y [C, N - taps] (empty when the rate has no table); limit_max > 0 clamps the input first (fused limit_ff)."""
    import torch
    assert x.dtype == torch.float32 and x.is_cuda and x.dim() == 2 and x.stride(1) == 1
    ch, n = x.shape
    out = torch.empty((ch, n), dtype=torch.float32, device=x.device) if out is None else out
    rc = _check(lib().csdrb_deemphasis_nfm_bank_ff(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), ch, n, sample_rate, limit_max, _stream()),
                "deemphasis_nfm_bank_ff")
    return out[:, :rc]


def fir_valid_bank_ff(x, taps, limit_max: float = 0.0, out=None):
    """x [C, N] float32, taps (host, <= 208) -> y [C, N - taps]: the de-emphasis FIR kernel with caller-supplied taps."""
    import torch
    assert x.dtype == torch.float32 and x.is_cuda and x.dim() == 2 and x.stride(1) == 1
    taps = np.ascontiguousarray(taps, np.float32)
    ch, n = x.shape
    out = torch.empty((ch, n), dtype=torch.float32, device=x.device) if out is None else out
    rc = _check(lib().csdrb_fir_valid_bank_ff(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), ch, n, _fp(taps), taps.size, limit_max, _stream()),
                "fir_valid_bank_ff")
    return out[:, :rc]


class DdcBank:
    """Streaming shared-input DDC/NFM bank (csdrb_ddc_bank_*): shift | fir_decimate | [fmdemod] for C channels of one wideband stream,
    one call per block, all per-channel state inside; the phase-chain pre-pass of the next block overlaps the current block."""

    def __init__(self, rates, decimation: int, taps: np.ndarray, demod: bool = True, chunk: int = 1024):
        self.rates = np.ascontiguousarray(np.atleast_1d(rates), np.float32)
        self.taps = np.ascontiguousarray(taps, np.float32)
        self.decimation, self.demod, self.channels = decimation, demod, self.rates.size
        self.h = lib().csdrb_ddc_bank_create(self.channels, _fp(self.rates), decimation, _fp(self.taps), self.taps.size, 1 if demod else 0, chunk)
        if not self.h:
            raise CsdrB200Error(f"csdrb_ddc_bank_create: {lib().csdrb_last_error().decode()}")

    def process(self, wide, out=None):
        """wide: [N] complex64 CUDA tensor starting where the previous call stopped consuming (n_out*decimation samples in)."""
        import torch
        assert wide.dtype == torch.complex64 and wide.is_cuda and wide.dim() == 1
        n_out = fir_out_len(wide.numel(), self.decimation, self.taps.size)
        if out is None:
            out = torch.empty((self.channels, n_out + (n_out & 1)), dtype=torch.float32 if self.demod else torch.complex64, device=wide.device)
        rc = _check(lib().csdrb_ddc_bank_process(self.h, wide.data_ptr(), wide.numel(), out.data_ptr(), out.stride(0), _stream()), "ddc_bank_process")
        return out[:, :rc]

    def set_rate(self, channel: int, rate: float):
        _check(lib().csdrb_ddc_bank_set_rate(self.h, channel, rate), "ddc_bank_set_rate")

    @property
    def offset(self) -> int:
        return int(lib().csdrb_ddc_bank_offset(self.h))

    def close(self):
        if getattr(self, "h", None):
            lib().csdrb_ddc_bank_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiBank:
    """One process, several GPUs (csdrb_multi_bank_*): the shared-input DDC/NFM bank in contiguous channel slices over `devices`, the wideband block
    broadcast from devices[0] with NCCL.  Host (ideally pinned) numpy buffers in and out; submit()/collect() pipeline two blocks."""

    def __init__(self, devices, rates, decimation: int, taps: np.ndarray, demod: bool = True, chunk: int = 1024, max_block: int = 1 << 21):
        self.rates = np.ascontiguousarray(np.atleast_1d(rates), np.float32)
        self.taps = np.ascontiguousarray(taps, np.float32)
        self.decimation, self.demod, self.channels = decimation, demod, self.rates.size
        devs = (C.c_int * len(devices))(*devices)
        self.h = lib().csdrb_multi_bank_create(len(devices), devs, self.channels, _fp(self.rates), decimation, _fp(self.taps), self.taps.size, 1 if demod else 0,
                                               chunk, max_block)
        if not self.h:
            raise CsdrB200Error(f"csdrb_multi_bank_create: {lib().csdrb_last_error().decode()}")

    def slices(self):
        out = []
        for i in range(lib().csdrb_multi_bank_devices(self.h)):
            d, c0, n = C.c_int(), C.c_int(), C.c_int()
            _check(lib().csdrb_multi_bank_slice(self.h, i, C.byref(d), C.byref(c0), C.byref(n)), "multi_bank_slice")
            out.append((d.value, c0.value, n.value))
        return out

    def submit(self, wide: np.ndarray, out: np.ndarray) -> int:
        assert wide.dtype == np.complex64 and wide.ndim == 1 and out.ndim == 2 and out.shape[0] == self.channels and out.strides[1] == out.itemsize
        assert out.dtype == (np.float32 if self.demod else np.complex64)
        return _check(lib().csdrb_multi_bank_submit(self.h, wide.ctypes.data, wide.size, out.ctypes.data, out.strides[0] // out.itemsize), "multi_bank_submit")

    def collect(self, ticket: int) -> int:
        return _check(lib().csdrb_multi_bank_collect(self.h, ticket), "multi_bank_collect")

    def process(self, wide: np.ndarray, out: np.ndarray) -> int:
        return self.collect(self.submit(wide, out))

    def set_rate(self, channel: int, rate: float):
        _check(lib().csdrb_multi_bank_set_rate(self.h, channel, rate), "multi_bank_set_rate")

    def close(self):
        if getattr(self, "h", None):
            lib().csdrb_multi_bank_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shift_unroll_bank_cc(x, rates, phases=None, table_size: int = 1024, out=None):
    """shift_unroll_cc for C channels; x [N] (shared wideband input) or [C, N].  Returns (y [C, N], carried phases [C])."""
    import torch
    rates = np.atleast_1d(np.asarray(rates, np.float32)); ch = rates.size
    shared = (x.dim() == 1) if x.dtype == torch.complex64 else (x.dim() == 2)
    xr, ptr, stride, xc, n = _as_cf32_rows(x)
    if shared:
        stride = 0
    dev = xr.device
    params = torch.from_numpy(np.array([shift_addition_init(float(r)) for r in rates], np.float32)).to(dev)
    ds = np.empty((ch, table_size), np.float32); dc = np.empty((ch, table_size), np.float32)
    for c, r in enumerate(rates):
        d = lib().shift_unroll_init(float(r), table_size)
        ds[c] = np.ctypeslib.as_array(d.dsin, shape=(table_size,)); dc[c] = np.ctypeslib.as_array(d.dcos, shape=(table_size,))
    d_ds, d_dc = torch.from_numpy(ds).to(dev), torch.from_numpy(dc).to(dev)
    d_phase = torch.zeros(ch, dtype=torch.float32, device=dev) if phases is None else phases.clone()
    out = torch.empty((ch, n), dtype=torch.complex64, device=dev) if out is None else out
    scratch = _scratch(lib().csdrb_shift_addition_bank_scratch_bytes(ch, n, min(table_size, n)) + 16, dev)
    _check(lib().csdrb_shift_unroll_bank_cc(ptr, stride, out.data_ptr(), out.stride(0), ch, n, params.data_ptr(), d_ds.data_ptr(), d_dc.data_ptr(), table_size,
                                            table_size, d_phase.data_ptr(), scratch.data_ptr(), scratch.numel(), _stream()), "shift_unroll_bank_cc")
    return out, d_phase


def spectrum_logpower(x, fft_size: int, window: str = "HAMMING", add_db: float = 0.0):
    """fft_cc | logpower_cf for a whole stream on the device: frames of fft_size samples -> [frames, fft_size] dB values."""
    import torch
    assert x.dtype == torch.complex64 and x.is_cuda and x.dim() == 1
    frames = x.numel() // fft_size
    xs = x[:frames * fft_size].contiguous().view(frames, fft_size)
    w = torch.from_numpy(libcsdr.precalculate_window(fft_size, window)).to(x.device)
    xw = torch.empty_like(xs)
    _check(lib().csdrb_apply_window_rows_c(xs.data_ptr(), xw.data_ptr(), w.data_ptr(), fft_size, frames, _stream()), "apply_window_rows_c")
    spec = fft_c2c(xw)
    out = torch.empty((frames, fft_size), dtype=torch.float32, device=x.device)
    _check(lib().csdrb_logpower_cf(spec.data_ptr(), out.data_ptr(), frames * fft_size, add_db, _stream()), "logpower_cf")
    return out
