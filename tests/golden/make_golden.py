"""Generate tests/golden/hotpath_golden.npz from the COMPILED, UNMODIFIED reference.

Run in the build container only (needs oracle/_ref/libcsdr_ref.so, i.e. /root/reference):
    make -C oracle ref && python tests/golden/make_golden.py
The reference has no golden vectors of its own for this path, so these are
outputs of the reference itself on small seeded inputs.  Inputs are stored next to the outputs so the
fixtures do not depend on numpy's generator staying stable.
Reference build flags: see oracle/Makefile (REFFLAGS) -- what the reference Makefile:32,38,39 selects.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.pyoracle import Ref  # noqa: E402


def cplx(rng, n, amp=1.0):
    return ((rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)) * amp).astype(np.complex64)


def fm_signal(n, dev=0.3, period=200.0):
    t = np.arange(n)
    return np.exp(1j * np.cumsum(dev * np.sin(2 * np.pi * t / period))).astype(np.complex64)


def main():
    r = Ref()
    rng = np.random.default_rng(20260923)
    g = {}
    # --- conversions (a1, a2)
    g["u8_in"] = np.arange(256, dtype=np.uint8)
    g["u8_out"] = r.convert_u8_f(g["u8_in"])
    g["s16_in"] = np.concatenate([np.array([-32768, -32767, -1, 0, 1, 32766, 32767], np.int16),
                                  rng.integers(-32768, 32768, 2041).astype(np.int16)])
    g["s16_out"] = r.convert_s16_f(g["s16_in"])
    g["f_in"] = np.concatenate([np.array([-1, -0.5, 0, 0.25, 0.5, 1, 1.5, 2.0, -1.5, -2.0], np.float32),
                                rng.uniform(-1, 1, 2038).astype(np.float32)])
    g["f_s16_out"] = r.convert_f_s16(g["f_in"])
    # --- filter design (a5)
    g["filter_len_bw"] = np.array([0.05, 0.0201, 0.02, 0.005, 0.002, 0.03], np.float32)
    g["filter_len"] = np.array([r.firdes_filter_len(float(b)) for b in g["filter_len_bw"]], np.int32)
    g["lowpass_79"] = r.firdes_lowpass_f(79, 0.05, "HAMMING")
    g["lowpass_199"] = r.firdes_lowpass_f(199, 0.05, "HAMMING")
    g["lowpass_101_blackman"] = r.firdes_lowpass_f(101, 0.1, "BLACKMAN")
    g["bandpass_79"] = r.firdes_bandpass_c(79, 0.1, 0.3, "HAMMING")
    # --- shift (a3, a4)
    g["shift_in"] = cplx(rng, 4096)
    y, ph = r.shift_addition_cc(g["shift_in"], -0.085, 0.0, 1024)
    g["shift_out_chunk1024"], g["shift_phase_chunk1024"] = y, np.float32(ph)
    y, ph = r.shift_addition_cc(g["shift_in"], 0.2, 0.3, None)
    g["shift_out_whole"], g["shift_phase_whole"] = y, np.float32(ph)
    y, st = r.decimating_shift_addition_cc(g["shift_in"][:448], 0.01, 2, 1, 0.5)
    g["dshift_out"], g["dshift_state"] = y, np.array(st, np.float64)
    # --- FIR (a6)
    g["fir_in"] = cplx(rng, 4096)
    g["fir_out_79_d10"] = r.fir_decimate_cc(g["fir_in"], 10, g["lowpass_79"])
    g["fir_out_199_d10"] = r.fir_decimate_cc(g["fir_in"], 10, g["lowpass_199"])
    g["fir_out_79_d7"] = r.fir_decimate_cc(g["fir_in"][:1000], 7, g["lowpass_79"])
    # --- fmdemod (a7)
    g["fm_in"] = fm_signal(2048)
    y, last = r.fmdemod_quadri_cf(g["fm_in"], 0.25 - 0.5j)
    g["fm_out"], g["fm_last"] = y, np.complex64(last)
    # --- fractional decimator (a8)
    g["fd_in"] = rng.uniform(-1, 1, 8192).astype(np.float32)
    g["fd_out_r5_blk1024"] = r.fractional_decimator_ff(g["fd_in"], 5.0, 12, None, 1024)
    g["fd_out_r2p7_whole"] = r.fractional_decimator_ff(g["fd_in"], 2.7183, 12, None, None)
    g["fd_out_r3_pts4_prefilter"] = r.fractional_decimator_ff(g["fd_in"][:3000], 3.0, 4, r.firdes_lowpass_f(31, 0.15), None)
    # --- fastagc (a9)
    env = np.repeat(np.array([0.02, 0.02, 0.5, 0.1, 0.001, 0.3, 0.3, 0.0, 0.05, 0.9], np.float32), 256)
    g["agc_in"] = (rng.uniform(-1, 1, env.size).astype(np.float32) * env)
    g["agc_out_b256"] = r.fastagc_ff(g["agc_in"], 256, 1.0)
    g["agc_out_b512_ref0p5"] = r.fastagc_ff(g["agc_in"], 512, 0.5)
    # --- audio tail (8f rank 1): 1-pole de-emphasis and limiter
    g["deemph_in"] = rng.uniform(-1.5, 1.5, 4096).astype(np.float32)
    y, last = r.deemphasis_wfm_ff(g["deemph_in"], 50e-6, 48000, 0.0, 1024)
    g["deemph_out_50us_48k"], g["deemph_last"] = y, np.float32(last)
    g["limit_out"] = r.limit_ff(g["deemph_in"], 1.0)
    # --- spectrum side path + shift_unroll (8f ranks 3, 4)
    g["win_hamming_1024"] = r.precalculate_window(1024, "HAMMING")
    g["spec_in"] = cplx(rng, 4096, 0.5)
    g["logpower_out"] = r.logpower_cf(g["spec_in"], -70.0)
    g["logavg_out"] = r.logaveragepower_cf(g["spec_in"], -70.0, 512, 4)
    y, ph = r.shift_unroll_cc(g["shift_in"], -0.085, 0.0, 1024)
    g["unroll_out"], g["unroll_phase"] = y, np.float32(ph)
    # --- overlap-add FFT FIR (a10) : bw 0.05 -> 79 taps, fft 256, 178 samples/block (csdr.c:1833-1838)
    g["bp_in"] = cplx(rng, 434 * 5)
    g["bp_out"] = r.bandpass_fir_fft_cc(g["bp_in"], -0.1, 0.2, 0.05, "HAMMING")
    # --- fastddc (a11-a13) : geometry printouts + a small end-to-end (bw 0.05, decimation 8)
    cases = [(0.002, 64, 0.1), (0.05, 10, -0.3), (0.01, 6, 0.25), (0.005, 50, 0.4), (0.05, 8, 0.123)]
    keys = None
    rows = []
    for bw, dec, sh in cases:
        geo = r.fastddc_geometry(bw, dec, sh)
        keys = keys or list(geo)
        rows.append([float(geo[k]) for k in keys])
    g["ddc_cases"] = np.array(cases, np.float64)
    g["ddc_keys"] = np.array(keys)
    g["ddc_geometry"] = np.array(rows, np.float64)
    ddc, _ = r.fastddc_init(0.05, 8, 0.123)
    n = ddc.input_size * 4
    t = np.arange(n)
    g["ddc_in"] = (np.exp(2j * np.pi * 0.125 * t) * 0.5).astype(np.complex64) + cplx(rng, n, 0.05)
    spectra = r.fastddc_fwd(g["ddc_in"], ddc)
    g["ddc_fwd_out"] = np.stack(spectra)
    g["ddc_inv_out"] = r.fastddc_inv(spectra, 0.05, 8, 0.123, "HAMMING")
    # --- NFM audio tail (8f rank 1): the four fixed de-emphasis FIRs (tables as the compiled reference exports them) on one input;
    #     own generator so that the arrays above keep their values
    rng2 = np.random.default_rng(20260923)
    g["nfm_in"] = rng2.uniform(-1.0, 1.0, 3000).astype(np.float32)
    for sr in r.NFM_RATES:
        g[f"nfm_taps_{sr}"] = r.deemphasis_nfm_taps(sr)
        g[f"nfm_out_{sr}"] = r.deemphasis_nfm_ff(g["nfm_in"], sr)
    # --- shift_addfast_cc (8f rank 3): the reference build's own step table (its init goes through libmvec sinf/cosf under
    #     -ffast-math, up to 1 ulp off the correctly rounded value) and the stream it produces from it in 1024-sample calls
    g["addfast_steps"] = r.shift_addfast_init(-0.085)
    y, ph = r.shift_addfast_cc(g["shift_in"], -0.085, 0.0, 1024)
    g["addfast_out"], g["addfast_phase"] = y, np.float32(ph)
    # --- shift_math_cc (8f rank 3): one call per 1024 samples like the CLI, starting phase outside [0, 2*PI]
    y, ph = r.shift_math_cc(g["shift_in"], -0.085, -7.5, 1024)
    g["math_out"], g["math_phase"] = y, np.float32(ph)
    out = Path(__file__).with_name("hotpath_golden.npz")
    np.savez_compressed(out, **g)
    print(f"wrote {out} ({out.stat().st_size} bytes, {len(g)} arrays)")


if __name__ == "__main__":
    main()
