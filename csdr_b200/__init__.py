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
    L.csdrb_host_alloc.argtypes = [C.c_size_t]; L.csdrb_host_alloc.restype = vp
    L.csdrb_host_free.argtypes = [vp]
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
    def fmdemod_quadri_cf(x, last=0j):
        x = np.ascontiguousarray(x, np.complex64); y = np.empty(x.size, np.float32)
        r = lib().fmdemod_quadri_cf(x.ctypes.data, y.ctypes.data, x.size, None, _CF(np.float32(last.real), np.float32(last.imag)))
        return y, complex(r.i, r.q)
