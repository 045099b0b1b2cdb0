"""CPU tier, property-based: the oracle restatement against the compiled, unmodified reference (oracle/_ref) on randomly drawn
shapes and parameters -- the same pinning as tests/test_oracle.py, but over the parameter space instead of a few hand-picked cases.
Skipped when oracle/_ref is not built (it needs /root/reference at build time)."""
import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402

from oracle.pyoracle import rel_rms  # noqa: E402

import os  # noqa: E402

# deterministic by default (the driver's CPU tier must not flake); CSDRB_HYP_EXAMPLES / CSDRB_HYP_RANDOM=1 widen the hunt by hand
COMMON = dict(deadline=None, max_examples=int(os.environ.get("CSDRB_HYP_EXAMPLES", "40")), derandomize=not os.environ.get("CSDRB_HYP_RANDOM"),
              suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


def _cplx(seed, n):
    rng = np.random.default_rng(seed)
    return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(0, 5000))
def test_conversions_any_bytes(oracle, ref, seed, n):
    rng = np.random.default_rng(seed)
    u8 = rng.integers(0, 256, n, dtype=np.uint8)
    assert np.array_equal(oracle.convert_u8_f(u8), ref.convert_u8_f(u8))
    s16 = rng.integers(-32768, 32768, n).astype(np.int16)
    assert np.array_equal(oracle.convert_s16_f(s16), ref.convert_s16_f(s16))
    f = (rng.standard_normal(n) * rng.choice([1e-3, 0.5, 1.0, 3.0])).astype(np.float32)      # 3.0: beyond +-1, the wrap-around cases
    assert np.array_equal(oracle.convert_f_s16(f), ref.convert_f_s16(f))


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(1, 6000), D=st.integers(1, 64), T=st.integers(1, 400))
def test_fir_decimate_any_geometry(oracle, ref, seed, n, D, T):
    x = _cplx(seed, n)
    taps = np.random.default_rng(seed + 1).uniform(-1, 1, T).astype(np.float32)
    a, b = oracle.fir_decimate_cc(x, D, taps), ref.fir_decimate_cc(x, D, taps)
    assert a.size == b.size                                                       # output count is the reference's, whatever it is
    if a.size:
        scale = np.abs(taps).sum()                                                # sums of <= 400 terms with cancellation: bound by the terms
        assert np.abs(a - b).max() <= 4e-6 * scale


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(1, 20000), rate=st.floats(-0.5, 0.5, width=32), phase=st.floats(-3.0, 3.0, width=32),
       chunk=st.sampled_from([None, 1, 37, 1000, 1024, 4096]))
def test_shift_addition_any_rate_and_chunking(oracle, ref, seed, n, rate, phase, chunk):
    x = _cplx(seed, n)
    a, pa = oracle.shift_addition_cc(x, rate, phase, chunk)
    b, pb = ref.shift_addition_cc(x, rate, phase, chunk)
    assert np.float32(pa) == np.float32(pb)                                       # the carried float phase: same rounding sequence
    # deltas are bit-equal (sincosf, like the shipped build); a call's SEED is (float)cos((double)phase) here and sincosf(phase) in the
    # -ffast-math build: one ulp apart for ~3 % of phases, after which the two recursions round differently and drift apart like
    # sqrt(steps) * 4e-8 -- 1.5e-6 after a 1400-sample call, far inside the 1e-5 bar
    assert rel_rms(a, b) < 5e-6


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(1, 20000), rate=st.floats(-0.5, 0.5, width=32), phase=st.floats(-50.0, 50.0, width=32))
def test_shift_math_any_rate(oracle, ref, seed, n, rate, phase):
    x = _cplx(seed, n)
    (a, pa), (b, pb) = oracle.shift_math_cc(x, rate, phase), ref.shift_math_cc(x, rate, phase)
    assert np.float32(pa) == np.float32(pb) and rel_rms(a, b) < 1e-6               # n rounded additions and wraps: the same float sequence


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(1, 20000), rate=st.floats(-0.5, 0.5, width=32), phase=st.floats(0.0, 6.28125, width=32),
       size=st.sampled_from([64, 1000, 65536]))
def test_shift_table_any_rate(oracle, ref, seed, n, rate, phase, size):
    x = _cplx(seed, n); table = ref.shift_table_init(size)
    (a, pa, bad), (b, pb) = oracle.shift_table_cc(x, rate, table, phase), ref.shift_table_cc(x, rate, table, phase)
    assert np.float32(pa) == np.float32(pb)
    if bad == 0:                                                                    # an index outside the table is undefined behaviour in the reference
        assert np.array_equal(a, b)


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(1, 20000), rate=st.floats(-0.5, 0.5, width=32), dec=st.integers(1, 40))
def test_decimating_shift_any_rate(oracle, ref, seed, n, rate, dec):
    x = _cplx(seed, n)
    (ya, sa), (yb, sb) = oracle.decimating_shift_addition_cc(x, rate, dec, 0, 0.0), ref.decimating_shift_addition_cc(x, rate, dec, 0, 0.0)
    assert ya.size == yb.size and sa[0] == sb[0] and np.float32(sa[1]) == np.float32(sb[1])      # output count, decimation_remain, carried phase
    assert ya.size == 0 or rel_rms(ya, yb) < 1e-6


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(2, 8000))
def test_fmdemod_any_length(oracle, ref, seed, n):
    x = _cplx(seed, n)
    (a, la), (b, lb) = oracle.fmdemod_quadri_cf(x), ref.fmdemod_quadri_cf(x)
    assert np.array_equal(a, b) and la == lb


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(200, 20000), rate=st.floats(1.0078125, 12.0, width=32), points=st.sampled_from([2, 4, 8, 12, 16]))
def test_fractional_decimator_any_rate(oracle, ref, seed, n, rate, points):
    x = np.random.default_rng(seed).uniform(-1, 1, n).astype(np.float32)
    ya, yb = oracle.fractional_decimator_ff(x, rate, points), ref.fractional_decimator_ff(x, rate, points)
    assert ya.size == yb.size                                                     # the float position chain picks the same indices
    assert ya.size == 0 or rel_rms(ya, yb) < 1e-6


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31), blocks=st.integers(3, 12), block=st.sampled_from([16, 256, 1000, 1024]), reference=st.floats(0.125, 2.0, width=32),
       scale=st.sampled_from([0.0, 1e-4, 0.3, 1.0, 50.0]))
def test_fastagc_any_block(oracle, ref, seed, blocks, block, reference, scale):
    x = (np.random.default_rng(seed).uniform(-1, 1, blocks * block) * scale).astype(np.float32)
    a, b = oracle.fastagc_ff(x, block, reference), ref.fastagc_ff(x, block, reference)
    if block & (block - 1) == 0:
        assert np.array_equal(a, b, equal_nan=True)                               # power-of-two blocks (the CLI default is 1024): bit-exact
    else:
        # the -ffast-math build turns the gain ramp's division by input_size into a multiplication by its reciprocal: exact only for
        # powers of two, one or two ulps otherwise
        assert np.all(np.isfinite(a) == np.isfinite(b)) and np.abs(a - b).max() <= 3e-7 * max(np.abs(b).max(), 1e-30)


@settings(**COMMON)
@given(length=st.integers(1, 600).map(lambda v: v | 1), cutoff=st.floats(0.0078125, 0.5, width=32), window=st.sampled_from(["BOXCAR", "BLACKMAN", "HAMMING"]))
def test_firdes_lowpass_any_length(oracle, ref, length, cutoff, window):
    a, b = oracle.firdes_lowpass_f(length, cutoff, window), ref.firdes_lowpass_f(length, cutoff, window)
    assert np.abs(a - b).max() <= 2e-6 * np.abs(b).max()                          # the -ffast-math build re-associates the normalising sum


@settings(**COMMON)
@given(length=st.integers(1, 400).map(lambda v: v | 1), lo=st.floats(-0.5, 0.25, width=32), width=st.floats(0.015625, 0.25, width=32))
def test_firdes_bandpass_any_band(oracle, ref, length, lo, width):
    a, b = oracle.firdes_bandpass_c(length, lo, lo + width), ref.firdes_bandpass_c(length, lo, lo + width)
    assert np.abs(a - b).max() <= 4e-6 * np.abs(b).max()


@settings(**COMMON)
@given(bw=st.floats(0.001953125, 0.25, width=32), dec=st.integers(1, 200), shift=st.floats(-0.5, 0.5, width=32))
def test_fastddc_geometry_any_parameters(oracle, ref, bw, dec, shift):
    a, b = oracle.fastddc_geometry(bw, dec, shift), ref.fastddc_geometry(bw, dec, shift)
    assert list(a) == list(b)
    for k in a:
        assert np.float32(a[k]) == np.float32(b[k]), (k, a[k], b[k])


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(1, 20000), tau=st.sampled_from([50e-6, 75e-6, 1e-3]), fs=st.sampled_from([8000, 44100, 48000, 240000]),
       last=st.floats(-1.0, 1.0, width=32), limit=st.floats(0.0625, 4.0, width=32))
def test_audio_tail_any_parameters(oracle, ref, seed, n, tau, fs, last, limit):
    x = np.random.default_rng(seed).uniform(-2, 2, n).astype(np.float32)
    (a, la), (b, lb) = oracle.deemphasis_wfm_ff(x, tau, fs, last), ref.deemphasis_wfm_ff(x, tau, fs, last)
    assert np.array_equal(a, b) and np.float32(la) == np.float32(lb)
    assert np.array_equal(oracle.limit_ff(x, limit), ref.limit_ff(x, limit))


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(1, 9000), rate=st.floats(-0.5, 0.5, width=32), phase=st.floats(-3.0, 3.0, width=32),
       size=st.sampled_from([64, 1000, 1024]))
def test_shift_unroll_any_rate(oracle, ref, seed, n, rate, phase, size):
    x = _cplx(seed, n)
    (a, pa), (b, pb) = oracle.shift_unroll_cc(x, rate, phase, size), ref.shift_unroll_cc(x, rate, phase, size)
    assert np.float32(pa) == np.float32(pb)
    assert rel_rms(a, b) < 1e-6                                                    # no recursion: table and seed differ by at most an ulp each


@settings(**COMMON)
@given(seed=st.integers(0, 2 ** 31), frames=st.integers(1, 6), size=st.sampled_from([16, 256, 1000]), add_db=st.floats(-100.0, 20.0, width=32),
       window=st.sampled_from(["BOXCAR", "BLACKMAN", "HAMMING"]))
def test_spectrum_side_path(oracle, ref, seed, frames, size, add_db, window):
    x = _cplx(seed, frames * size)
    wa, wb = oracle.precalculate_window(size, window), ref.precalculate_window(size, window)
    assert np.abs(wa - wb).max() <= 1e-6                                          # the build evaluates cos through libmvec (a few ulp)
    ya = np.concatenate([oracle.apply_precalculated_window_c(x[f * size:(f + 1) * size], wb) for f in range(frames)])
    yb = np.concatenate([ref.apply_precalculated_window_c(x[f * size:(f + 1) * size], wb) for f in range(frames)])
    assert np.array_equal(ya, yb)
    assert np.abs(oracle.logpower_cf(x, add_db) - ref.logpower_cf(x, add_db)).max() <= 2e-5      # dB; log10f vs log10 of the build
