"""Fuzz the remaining SHIPPED kernels on the CPU (tests/host_shim/cuda_emul.h): conversions, discriminator, limiter, windows/power, AGC, both
de-emphasis filters, shift_unroll / shift_addfast / decimating shift -- random sizes, strides and states through the real launchers against the
oracle, bit-exact where the GPU tests claim it.  usage: python tests/fuzz/fuzz_emulated_small.py [seed] [seconds]"""
import sys, time, tempfile, ctypes as C, numpy as np
from pathlib import Path
_ROOT = str(Path(__file__).resolve().parents[2])
sys.path.insert(0, _ROOT); sys.path.insert(0, _ROOT + '/tests/host_shim')
import emul_build as eb
from oracle.pyoracle import Oracle, rel_rms
o = Oracle()
tmp = Path(tempfile.mkdtemp(prefix='fuzz_small_'))
el, _ = eb.build_file(tmp, 'elementwise.cu'); sh, _ = eb.build_file(tmp, 'shift.cu'); au, _ = eb.build_file(tmp, 'audio.cu', host_c=("csdr_b200/host/firdes.c",))
GOLD = np.load(_ROOT + '/tests/golden/hotpath_golden.npz')
NFM = {r: np.ascontiguousarray(GOLD[f"nfm_taps_{r}"]) for r in (48000, 44100, 11025, 8000)}
P = lambda a: a.ctypes.data
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
t_end = time.time() + (float(sys.argv[2]) if len(sys.argv) > 2 else 120)
def aligned(shape, dtype):
    n = int(np.prod(shape)); it = np.dtype(dtype).itemsize; raw = np.zeros(n * it + 32, np.uint8); off = (-raw.ctypes.data) % 16
    return raw[off:off + n * it].view(dtype).reshape(shape)
def cplx(*s): return (rng.uniform(-1, 1, s) + 1j * rng.uniform(-1, 1, s)).astype(np.complex64)
it = 0; counts = {}
while time.time() < t_end:
    it += 1; kind = int(rng.integers(0, 8)); counts[kind] = counts.get(kind, 0) + 1
    if kind == 0:      # conversions
        n = int(rng.integers(1, 70000))
        u8 = aligned(n, np.uint8); u8[:] = rng.integers(0, 256, n); f = aligned(n, np.float32)
        assert el.emul_launch_convert_u8_f(P(u8), P(f), n) >= 0 and np.array_equal(f, o.convert_u8_f(u8)), ('u8', n)
        s16 = aligned(n, np.int16); s16[:] = rng.integers(-32768, 32768, n)
        assert el.emul_launch_convert_s16_f(P(s16), P(f), n) >= 0 and np.array_equal(f, o.convert_s16_f(s16)), ('s16', n)
        f[:] = rng.standard_normal(n) * rng.choice([1e-3, 0.5, 1.0, 3.0, 1e6]); f[rng.integers(0, n)] = rng.choice([np.nan, np.inf, -np.inf, 1.0, -1.0])
        q = aligned(n, np.int16)
        assert el.emul_launch_convert_f_s16(P(f), P(q), n) >= 0 and np.array_equal(q, o.convert_f_s16(f)), ('f_s16', n)
        lim = float(np.float32(rng.uniform(0.1, 3))); g = aligned(n, np.float32)
        assert el.emul_launch_limit_ff(P(f), P(g), n, lim) >= 0 and np.array_equal(g, o.limit_ff(f, lim), equal_nan=True), ('limit', n)
    elif kind == 1:    # discriminator bank
        ch = int(rng.integers(1, 6)); n = int(rng.integers(1, 20000)); stride = n + int(rng.integers(0, 9))
        x = cplx(ch, stride); last = cplx(ch); lo = np.zeros(ch, np.complex64); ostr = (n + int(rng.integers(0, 5)) + 1) & ~1; y = np.zeros((ch, ostr), np.float32)      # the launcher wants an even output stride
        assert el.emul_launch_fmdemod_quadri_bank(P(x), stride, P(y), ostr, ch, n, P(last), P(lo)) >= 0
        for c in range(ch):
            w, wl = o.fmdemod_quadri_cf(np.ascontiguousarray(x[c, :n]), complex(last[c]))
            assert np.abs(y[c, :n] - w).max() <= 3e-7 and lo[c] == np.complex64(wl), ('fmdemod', ch, n)
    elif kind == 2:    # window rows + power
        size = int(rng.integers(2, 3000)); rows = int(rng.integers(1, 6)); x = cplx(rows * size); w = o.precalculate_window(size, ["BOXCAR", "BLACKMAN", "HAMMING"][rng.integers(0, 3)])
        y = np.zeros_like(x)
        assert el.emul_launch_apply_window_rows(P(x), P(y), P(w), size, rows) >= 0
        assert np.array_equal(y, np.concatenate([o.apply_precalculated_window_c(x[r * size:(r + 1) * size], w) for r in range(rows)])), ('window', size, rows)
        p = np.zeros(x.size, np.float32); add = float(np.float32(rng.uniform(-100, 20)))
        assert el.emul_launch_power(P(x), None, P(p), x.size, add, 0) >= 0 and np.abs(p - o.logpower_cf(x, add)).max() <= 3e-5, ('logpower', size)
    elif kind == 3:    # fastagc
        ch = int(rng.integers(1, 5)); block = int([16, 100, 1000, 1024, 4096][rng.integers(0, 5)]); nb = int(rng.integers(1, 12))
        x = (rng.uniform(-1, 1, (ch, nb * block)) * rng.choice([0.0, 1e-4, 0.3, 1.0, 50.0], (ch, 1))).astype(np.float32); ref = float(np.float32(rng.uniform(0.1, 2)))
        y = np.zeros_like(x); st = np.zeros((ch, 3), np.float32); hist = np.zeros((ch, 2, block), np.float32)
        sb = au.emul_fastagc_scratch_bytes(ch, nb); scr = np.zeros(sb + 16, np.uint8)
        # two calls (streaming) == the oracle's one stream
        h = int(rng.integers(0, nb + 1))
        if h: assert au.emul_launch_fastagc_bank(P(x), x.shape[1], P(y), y.shape[1], ch, block, h, ref, P(st), P(hist), P(scr), sb) >= 0
        if nb - h:
            xs = x[:, h * block:]; ys = y[:, h * block:]
            assert au.emul_launch_fastagc_bank(P(xs), x.shape[1], P(ys), y.shape[1], ch, block, nb - h, ref, P(st), P(hist), P(scr), sb) >= 0
        for c in range(ch): assert np.array_equal(y[c], o.fastagc_ff(x[c], block, ref), equal_nan=True), ('fastagc', block, nb, h, c)
    elif kind == 4:    # de-emphasis filters
        ch = int(rng.integers(1, 40)); n = int(rng.integers(1, 6000)); x = rng.uniform(-1, 1, (ch, n)).astype(np.float32); last = rng.uniform(-1, 1, ch).astype(np.float32); l0 = last.copy()
        tau = float(rng.choice([50e-6, 75e-6, 1e-3])); fs = int(rng.choice([8000, 44100, 48000, 240000])); y = np.zeros_like(x)
        assert au.emul_launch_deemphasis_wfm_bank(P(x), n, P(y), n, ch, n, tau, fs, P(last)) >= 0
        for c in range(0, ch, max(1, ch // 4)):
            w, wl = o.deemphasis_wfm_ff(x[c], tau, fs, float(l0[c])); assert np.array_equal(y[c], w) and np.float32(wl) == last[c], ('wfm', ch, n)
        rate = int(rng.choice([48000, 44100, 11025, 8000])); T = NFM[rate].size; n = int(rng.integers(1, 8000)); ch = int(rng.integers(1, 4))
        x = rng.uniform(-2, 2, (ch, n)).astype(np.float32); lim = float(rng.choice([0.0, 1.0, 0.3])); y = np.full((ch, max(n, 1)), np.nan, np.float32)
        rc = au.emul_launch_deemphasis_nfm_bank(P(x), n, P(y), y.shape[1], ch, n, rate, lim); assert rc == max(n - T, 0), ('nfm count', n, rate, rc)
        for c in range(ch):
            xin = o.limit_ff(x[c], lim) if lim > 0 else x[c]
            if rc: assert np.abs(y[c, :rc] - o.deemphasis_nfm_ff(xin, NFM[rate])).max() <= 1e-6 * np.abs(NFM[rate]).sum(), ('nfm', n, rate)
    elif kind == 5:    # shift_addfast bank
        n = int(rng.integers(1, 20000)); chunk = int([0, 4, 37, 1000, 1024, 4096][rng.integers(0, 6)]); ch = int(rng.integers(1, 5))
        rates = rng.uniform(-0.5, 0.5, ch).astype(np.float32); x = cplx(n); ph0 = rng.uniform(-30, 30, ch).astype(np.float32); ph = ph0.copy()
        steps = np.stack([o.shift_addfast_init(float(r)) for r in rates]); y = np.zeros((ch, n), np.complex64)
        sb = sh.emul_shift_bank_scratch_bytes(ch, n, chunk); scr = np.zeros(sb + 16, np.uint8)
        assert sh.emul_launch_shift_addfast_bank(P(x), 0, P(y), n, ch, n, P(steps), P(ph), chunk, P(scr), sb) >= 0
        for c in range(ch):
            w, wp = o.shift_addfast_cc(x, float(rates[c]), float(ph0[c]), chunk or None)
            assert np.float32(wp) == ph[c] and rel_rms(y[c], w) < 2e-7 and np.array_equal(y[c] == 0, w == 0), ('addfast', n, chunk, c)
    elif kind == 6:    # shift_unroll bank
        size = int([64, 1000, 1024][rng.integers(0, 3)]); n = int(rng.integers(1, 9000)); ch = int(rng.integers(1, 4)); rates = rng.uniform(-0.5, 0.5, ch).astype(np.float32)
        x = cplx(n); tabs = [np.empty(size, np.float32) for _ in range(2 * ch)]
        for c, r in enumerate(rates): o.L.oracle_shift_unroll_init(float(r), size, tabs[2 * c].ctypes.data_as(C.POINTER(C.c_float)), tabs[2 * c + 1].ctypes.data_as(C.POINTER(C.c_float)))
        dsin = np.stack(tabs[0::2]); dcos = np.stack(tabs[1::2]); params = np.array([o.shift_addition_init(float(r)) for r in rates], np.float32)
        ph0 = rng.uniform(-3, 3, ch).astype(np.float32); ph = ph0.copy(); y = np.zeros((ch, n), np.complex64)
        sb = sh.emul_shift_bank_scratch_bytes(ch, n, size); scr = np.zeros(sb + 16, np.uint8)
        assert sh.emul_launch_shift_unroll_bank(P(x), 0, P(y), n, ch, n, P(params), P(dsin), P(dcos), size, size, P(ph), P(scr), sb) >= 0
        for c, r in enumerate(rates):
            w, wp = o.shift_unroll_cc(x, float(r), float(ph0[c]), size); assert np.float32(wp) == ph[c] and rel_rms(y[c], w) < 2e-7, ('unroll', n, size, c)
    else:              # decimating shift bank
        n = int(rng.integers(1, 12000)); dec = int(rng.integers(1, 30)); ch = int(rng.integers(1, 5)); rates = rng.uniform(-0.5, 0.5, ch).astype(np.float32)
        xs = cplx(ch, n); params = np.array([o.shift_addition_init(float(np.float32(r) * dec)) for r in rates], np.float32)
        remain = rng.integers(0, dec, ch).astype(np.int32); ph = rng.uniform(-3, 3, ch).astype(np.float32); r0, p0 = remain.copy(), ph.copy(); outsz = np.zeros(ch, np.int32)
        y = np.zeros((ch, n // dec + 2), np.complex64)
        assert sh.emul_launch_decimating_shift_bank(P(xs), n, P(y), y.shape[1], ch, n, P(params), dec, P(remain), P(ph), P(outsz)) >= 0
        for c, r in enumerate(rates):
            w, (wr, wp) = o.decimating_shift_addition_cc(xs[c], float(r), dec, int(r0[c]), float(p0[c]))
            assert outsz[c] == w.size and remain[c] == wr and ph[c] == np.float32(wp) and (w.size == 0 or rel_rms(y[c, :w.size], w) < 2e-7), ('dshift', n, dec, c)
print("iterations", it, "per kind", dict(sorted(counts.items())))
