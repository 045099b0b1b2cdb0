"""GPU parity at BASELINE.json's full configuration sizes (-m gpu).  The whole bank runs on the GPU; the CPU oracle then replays a few
complete channels (it is much slower, so not all of them) plus size-independent properties over everything:
  configs[2]  fastddc overlap-save: 16384-pt FFT, 64 output channels from one wideband stream (1.05 M samples = 73 blocks here)
  configs[3]  NFM bank slice of one GPU: 128 channels x shift + fir_decimate_cc 50 (801 taps) + fmdemod on 1 Mi wideband samples
  configs[4]  bandpass_fir_fft_cc 4096-pt overlap-add, 512 channels x 256 k samples
(configs[1], the headline 256 x 2.4 M FIR bank, is in test_gpu_parity.py::test_fir_bank_full_size_properties.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
TOL = 1e-5


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    import csdr_b200
    csdr_b200.lib()
    return csdr_b200


def _rel(y, ref):
    from oracle.pyoracle import rel_rms
    return rel_rms(y, ref)


def test_config3_fastddc_64_channels(gpu, oracle):
    bw, dec, C = 0.002, 64, 64
    ddc = gpu.fastddc_init(bw, dec, 0.0)
    nblocks = 73
    n = nblocks * ddc.input_size                                       # 1 046 528 wideband samples
    shifts = [float(s) for s in np.linspace(-0.45, 0.45, C)]
    rng = np.random.default_rng(33)
    t = np.arange(n)
    x = np.zeros(n, np.complex128)
    for k in (0, 17, 40, 63):                                          # tones inside four of the channels (channel k passes the band around -shift_k)
        x += 0.2 * np.exp(2j * np.pi * (-shifts[k] + 0.002) * t)
    x = (x + 0.05 * (rng.normal(size=n) + 1j * rng.normal(size=n))).astype(np.complex64)
    sp, _ = gpu.fastddc_fwd_cc(torch.from_numpy(x).cuda(), ddc)
    out, counts, _ = gpu.fastddc_inv_bank_cc(sp, shifts, dec, bw)
    counts = counts.cpu().numpy()
    assert (counts == nblocks * 224).all()
    o_ddc, _ = oracle.fastddc_init(bw, dec, 0.0)
    want_sp = oracle.fastddc_fwd(x, o_ddc)
    assert _rel(sp.cpu().numpy(), np.stack(want_sp)) < 1e-6
    for c in (0, 17, 63):                                              # three complete channels against the oracle
        want = oracle.fastddc_inv(want_sp, bw, dec, shifts[c])
        assert _rel(out[c, :want.size].cpu().numpy(), want) < TOL / 2, c
    # property over all 64 channels: channels holding a tone carry far more power than their empty neighbours
    p = (out[:, :nblocks * 224].abs() ** 2).mean(1).cpu().numpy()
    for k in (17, 40):
        assert p[k] > 20 * p[k + 2] and p[k] > 20 * p[k - 2]


def test_config4_nfm_bank_slice_128_channels(gpu, oracle):
    C, N, D, bw, chunk = 128, 1 << 20, 50, 0.005, 1024
    T = gpu.firdes_filter_len(bw)
    assert T == 801
    taps = gpu.firdes_lowpass_f(T, 0.5 / D)
    rates = np.linspace(-0.45, 0.45, C).astype(np.float32)
    rng = np.random.default_rng(44)
    t = np.arange(N)
    wide = np.zeros(N, np.complex128)
    for c in (3, 64, 125):                                             # FM carriers at three channel centres (shift by rate r moves -r to 0)
        wide += 0.3 * np.exp(1j * (2 * np.pi * (-float(rates[c])) * t + np.cumsum(0.003 * np.sin(2 * np.pi * t / (4000.0 + c)))))
    wide = (wide + 0.003 * (rng.normal(size=N) + 1j * rng.normal(size=N))).astype(np.complex64)
    bank = gpu.DdcBank(rates, D, taps, demod=True, chunk=chunk)
    y = bank.process(torch.from_numpy(wide).cuda()).cpu().numpy()
    n_out = (N - T) // D + 1
    assert y.shape == (C, n_out) and np.isfinite(y).all()
    for c in (3, 64, 125):
        sh, _ = oracle.shift_addition_cc(wide, float(rates[c]), 0.0, chunk)
        want = oracle.fmdemod_quadri_cf(oracle.fir_decimate_cc(sh, D, taps))[0]
        assert _rel(y[c], want) < TOL, (c, _rel(y[c], want))
    # the same block through the one-shot call and through the unfused bank calls gives the same baseband (checksum of checksums)
    base_f, _, _ = gpu.ddc_bank(torch.from_numpy(wide).cuda(), rates, D, taps, demod=False, chunk=chunk)
    sh_all, _ = gpu.shift_addition_bank_cc(torch.from_numpy(wide).cuda(), rates, chunk=chunk)
    base_u = gpu.fir_decimate_bank_cc(sh_all, D, taps)
    assert _rel(base_f.cpu().numpy(), base_u.cpu().numpy()) < 2e-6
    bank.close()


def test_config5_bandpass_bank_512_channels(gpu, oracle):
    T, N, isz, ov = gpu.bandpass_geometry(0.002)
    assert (T, N, isz) == (1999, 4096, 2098)
    C, nblocks = 512, 125                                              # 262 250 samples per channel
    tf = gpu.bandpass_taps_fft(-0.05, 0.05, 0.002)
    g = torch.Generator(device="cuda").manual_seed(55)
    x = torch.view_as_complex(torch.rand((C, nblocks * isz, 2), generator=g, device="cuda") * 2 - 1)
    y, tail = gpu.bandpass_fir_fft_bank_cc(x, tf, isz)
    for c in (0, 255, 511):
        want = oracle.bandpass_fir_fft_cc(x[c].cpu().numpy(), -0.05, 0.05, 0.002)
        assert _rel(y[c].cpu().numpy(), want) < TOL / 2, c
    # linearity over the whole bank: F(a*x1 + x2) = a*F(x1) + F(x2)
    x2 = torch.roll(x, 1, 0)
    y2, _ = gpu.bandpass_fir_fft_bank_cc(x2, tf, isz)
    y3, _ = gpu.bandpass_fir_fft_bank_cc((0.5 * x + x2).contiguous(), tf, isz)
    err = (y3 - (0.5 * y + y2)).abs().max().item()
    assert err < 2e-5, err
    # streaming invariance: two calls with the tail carried equal one call
    cut = 60 * isz
    ya, ta = gpu.bandpass_fir_fft_bank_cc(x[:, :cut].contiguous(), tf, isz)
    yb, _ = gpu.bandpass_fir_fft_bank_cc(x[:, cut:].contiguous(), tf, isz, tail=ta)
    assert torch.equal(torch.cat([ya, yb], 1), y)
