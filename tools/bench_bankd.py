"""csdr-bankd end to end (u8 IQ in on stdin -> one s16 audio stream per channel out) against the reference's process chain of README.md:87.

    python tools/bench_bankd.py [channels=128] [seconds_of_signal=20] [--devices 0,1]

bankd: ONE process reads the 2.4 Msps u8 stream once, every channel runs through the fused bank kernel + the NFM audio tail on the GPU, sinks are /dev/null;
wall clock from the first byte to exit.  Reference: `csdr convert_u8_f | csdr shift_addition_cc R | csdr fir_decimate_cc 50 0.005 HAMMING | csdr fmdemod_quadri_cf |
csdr limit_ff | csdr deemphasis_nfm_ff 48000 | csdr fastagc_ff | csdr convert_f_s16` (eight processes per channel, the unmodified reference binary oracle/_ref/csdr_ref)
on the same bytes; a few chains run side by side to show that they scale with cores until the cores are used up.  Test infrastructure / measurement only."""
import os, subprocess, sys, time, tempfile
from pathlib import Path
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
BANKD = ROOT / "csdr_b200" / "csdr-bankd"
REF = ROOT / "oracle" / "_ref" / "csdr_ref"
args = [a for a in sys.argv[1:] if not a.startswith("--")]
channels = int(args[0]) if args else 128
seconds = float(args[1]) if len(args) > 1 else 20.0
devices = next((sys.argv[i + 1] for i, a in enumerate(sys.argv) if a == "--devices"), None)
FS = 2_400_000
n = int(FS * seconds) & ~1
rng = np.random.default_rng(1)
tmp = Path(tempfile.mkdtemp(prefix="bankd_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None))
iq = tmp / "iq.u8"
rng.integers(0, 256, 2 * n, dtype=np.uint8).tofile(iq)
rates = np.linspace(-0.45, 0.45, channels)


def run_bankd():
    cmd = [str(BANKD), "--u8", "--decimation", "50", "--bw", "0.005"]
    if devices:
        cmd += ["--devices", devices, "--tail", "none"]
    cmd += [f"{r:.6f}:/dev/null" for r in rates]
    best = None
    for _ in range(2):                                       # first run pays CUDA context + module load
        t0 = time.perf_counter()
        with open(iq, "rb") as f:
            p = subprocess.run(cmd, stdin=f, stderr=subprocess.PIPE)
        dt = time.perf_counter() - t0
        if p.returncode != 0:
            raise SystemExit(f"csdr-bankd failed: {p.stderr.decode()[-500:]}")
        best = dt if best is None else min(best, dt)
    return best


def run_reference(k):
    """k reference chains side by side (k different shifts) on the same file; wall clock until the last one is done"""
    env = dict(os.environ, LD_LIBRARY_PATH=f"{REF.parent}:{os.environ.get('LD_LIBRARY_PATH', '')}")
    chain = ("{c} convert_u8_f < {f} | {c} shift_addition_cc {r:.6f} | {c} fir_decimate_cc 50 0.005 HAMMING | {c} fmdemod_quadri_cf | {c} limit_ff | "
             "{c} deemphasis_nfm_ff 48000 | {c} fastagc_ff | {c} convert_f_s16 > /dev/null")
    t0 = time.perf_counter()
    procs = [subprocess.Popen(["bash", "-c", chain.format(c=REF, r=rates[i % channels], f=iq)], env=env, stderr=subprocess.DEVNULL) for i in range(k)]
    for p in procs:
        p.wait()
    return time.perf_counter() - t0


t_b = run_bankd()
print(f"csdr-bankd{' --devices ' + devices if devices else ''}: {channels} channels, {n / 1e6:.1f} M wideband samples in {t_b:.2f} s = {n / t_b / 1e6:.1f} Msps wideband "
      f"= {n / t_b / FS:.1f} x real time at 2.4 Msps ({channels * n / t_b / 1e6:.0f} M channel-samples/s), process start and CUDA context included", flush=True)
# the same with a quarter of the signal: the difference is the streaming rate without the start-up
iq_full = iq
iq = tmp / "iq_quarter.u8"
n_q = (n // 4) & ~1
with open(iq_full, "rb") as f, open(iq, "wb") as g:
    g.write(f.read(2 * n_q))
t_q = run_bankd()
iq = iq_full
rate = (n - n_q) / max(t_b - t_q, 1e-9)
print(f"csdr-bankd streaming rate (start-up taken out: {n / 1e6:.0f} M vs {n_q / 1e6:.0f} M samples, {t_b:.2f} s vs {t_q:.2f} s): {rate / 1e6:.0f} Msps wideband = {rate / FS:.0f} x real time, "
      f"{channels * rate / 1e6:.0f} M channel-samples/s", flush=True)
if REF.exists():
    cores = os.cpu_count() or 1
    for k in (1, 4, min(16, channels)):
        t_r = run_reference(k)
        print(f"reference chain x {k} (8 processes each, {cores} host threads): {t_r:.2f} s = {n / t_r / 1e6:.2f} Msps wideband each, {k * n / t_r / 1e6:.1f} M channel-samples/s in total; "
              f"{channels} channels this way ~ {channels / k * t_r:.0f} s of wall clock for the same signal (bankd: {t_b:.2f} s)", flush=True)
else:
    print("oracle/_ref/csdr_ref not built: reference chain skipped")
for f in tmp.iterdir():
    f.unlink()
tmp.rmdir()
