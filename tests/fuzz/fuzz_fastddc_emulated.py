"""Fuzz the fastddc inverse bank on the CPU: random geometries (transition bandwidth, decimation incl. odd post-decimations, shifts), ragged channel and block counts
(beyond 96 blocks the state chain runs on its wrap tables), several runs with carried state -- the stateless bank call against the oracle, and the plan object
(look-ahead) against the stateless call bit for bit.  usage: python tests/fuzz/fuzz_fastddc_emulated.py [seed] [seconds]   -- test infrastructure."""
import sys, time, ctypes as C, numpy as np, tempfile
from pathlib import Path
_ROOT = str(Path(__file__).resolve().parents[2])
sys.path.insert(0, _ROOT); sys.path.insert(0, _ROOT + '/tests/host_shim')
import emul_build as eb
from oracle.pyoracle import Oracle, rel_rms, _CF, _p, WINDOWS
o = Oracle()
fft, _ = eb.build_file(Path(tempfile.mkdtemp(prefix='fuzz_fastddc_')), 'fft.cu')
P = lambda a: a.ctypes.data
chan_dt = np.dtype([("offsetbin", np.int32), ("sindelta", np.float32), ("cosdelta", np.float32), ("rate", np.float32)])


def aligned(shape, dtype):
    n = int(np.prod(shape)); it = np.dtype(dtype).itemsize; raw = np.zeros(n * it + 32, np.uint8); off = (-raw.ctypes.data) % 16
    return raw[off:off + n * it].view(dtype).reshape(shape)


rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
t_end = time.time() + (float(sys.argv[2]) if len(sys.argv) > 2 else 120)
it = 0; worst = 0.0; paths = {}
while time.time() < t_end:
    it += 1
    bw = float(rng.choice([0.05, 0.03, 0.02, 0.1]))
    dec = int(rng.choice([2, 4, 6, 8, 10, 12, 16, 24, 32, 3, 5]))
    nch = int(rng.integers(1, 7)); runs = int(rng.integers(1, 4))
    nb = int(rng.choice([1, 2, 3, 5, 17, 33, 97, 110])) if bw >= 0.05 else int(rng.choice([1, 2, 3, 5, 17]))
    shifts = [float(s) for s in rng.uniform(-0.45, 0.45, nch)]
    try:
        gs = [o.fastddc_init(bw, dec, s)[0] for s in shifts]
    except Exception:
        continue
    g = gs[0]
    if g.fft_size > 4096 or g.fft_inv_size < 2:
        continue
    taps = aligned((nch, g.fft_size), np.complex64); chan = aligned(nch, chan_dt)
    for c, (gc, s) in enumerate(zip(gs, shifts)):
        tf = np.empty(g.fft_size, np.complex64)
        o.L.oracle_fastddc_make_taps_fft(C.byref(gc), s, dec, WINDOWS["HAMMING"], _p(tf, _CF))
        taps[c] = tf
        chan[c]["offsetbin"] = gc.offsetbin; chan[c]["sindelta"] = gc.dsadata.sindelta; chan[c]["cosdelta"] = gc.dsadata.cosdelta; chan[c]["rate"] = gc.dsadata.rate
    x = ((rng.uniform(-1, 1, runs * nb * g.input_size) + 1j * rng.uniform(-1, 1, runs * nb * g.input_size)) * 0.5).astype(np.complex64)
    spectra = np.stack(o.fastddc_fwd(x, g)).astype(np.complex64)
    width = nb * (g.post_input_size // g.post_decimation + 1) + 2
    remain = aligned(nch, np.int32); phase = aligned(nch, np.float32)
    sb = fft.emul_fastddc_inv_scratch_bytes(nch, nb); scratch = aligned(sb + 16, np.uint8)
    plan = C.c_void_p()
    have_plan = fft.emul_fastddc_inv_plan_create(C.addressof(plan), P(chan), nch, nb, g.fft_size, g.fft_inv_size, g.pre_decimation, g.scrap, g.post_input_size, g.post_decimation) == 0
    streams = [[] for _ in range(nch)]
    for r in range(runs):
        sp = aligned((nb, g.fft_size), np.complex64); sp[:] = spectra[r * nb:(r + 1) * nb]
        want = aligned((nch, width), np.complex64); wt = aligned(nch, np.int32)
        rc = fft.emul_launch_fastddc_inv_bank(P(sp), nb, P(taps), P(chan), nch, g.fft_size, g.fft_inv_size, g.pre_decimation, g.scrap, g.post_input_size, g.post_decimation,
                                              P(remain), P(phase), P(want), width, P(wt), P(scratch), sb)
        assert rc >= 0, (bw, dec, nb, nch, fft.emul_last_error())
        paths[rc] = paths.get(rc, 0) + 1
        assert have_plan == (rc == 4), (rc, have_plan)
        if have_plan:
            got = aligned((nch, width), np.complex64); gt = aligned(nch, np.int32)
            assert fft.emul_fastddc_inv_plan_run(plan, P(sp), P(taps), P(got), width, P(gt)) == nb, fft.emul_last_error()
            assert np.array_equal(gt, wt), (bw, dec, nb, nch, r)
            for c in range(nch):
                assert np.array_equal(got[c, :gt[c]].view(np.uint32), want[c, :wt[c]].view(np.uint32)), ("plan != bank", bw, dec, nb, nch, r, c)
        for c in range(nch):
            streams[c].append(want[c, :wt[c]].copy())
    if have_plan:
        fft.emul_fastddc_inv_plan_destroy(plan)
    for c in range(min(nch, 2)):
        ref = o.fastddc_inv(list(spectra), bw, dec, shifts[c])
        got = np.concatenate(streams[c])
        assert got.size == ref.size, (bw, dec, nb, nch, c, got.size, ref.size)
        e = rel_rms(got, ref) if ref.size else 0.0
        worst = max(worst, e)
        assert e < 5e-6, ("bank != oracle", bw, dec, nb, nch, runs, c, e, g.fft_size, g.fft_inv_size, g.pre_decimation, g.post_decimation)
print(f"iterations {it}, worst rel-RMS vs oracle {worst:.2e}, launcher return codes (4 = fold path) {paths}")
