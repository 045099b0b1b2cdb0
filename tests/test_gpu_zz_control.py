"""GPU tests (-m gpu) of the CLI's run-time control channel (csdr.c:252-343: --fd <inherited descriptor> / --fifo <path>, one value per line):
the value that is waiting when the process starts selects the initial tuning, exactly as if it had been given on the command line -- for our CLI and
for the reference CLI alike.  (Re-tuning in mid-stream is timing dependent in both programs and is not compared.)  The same bodies run in the CPU
tier on the emulated library (tests/test_cli_emulated.py).  File name sorts last: not yet run on hardware."""
import os
import subprocess
import sys
import threading
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
from test_gpu_cli import clis, rel, run_graph  # noqa: E402,F401

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _with_fd(cli, args_before, args_after, line, data, timeout=120):
    """run `cli <args_before> --fd N <args_after>` with `line` already waiting in the pipe behind descriptor N"""
    r, w = os.pipe()
    os.write(w, line.encode())
    try:
        p = subprocess.run([cli] + args_before + ["--fd", str(r)] + args_after, input=data, stdout=subprocess.PIPE, stderr=subprocess.PIPE, pass_fds=(r,), timeout=timeout)
    finally:
        os.close(r); os.close(w)
    assert p.returncode == 0, p.stderr[-1000:]
    return p.stdout


def _with_fifo(cli, args_before, args_after, line, data, tmp_path, timeout=120):
    """the same through a named FIFO: the CLI blocks in open() until a writer shows up"""
    path = tmp_path / f"ctl_{abs(hash((cli, tuple(args_before), line))) % 10**8}"
    os.mkfifo(path)
    keep = []

    def writer():
        fd = os.open(path, os.O_WRONLY)
        os.write(fd, line.encode())
        keep.append(fd)                                                  # stay open until the CLI is done (a closed FIFO still reads as "no data")

    t = threading.Thread(target=writer, daemon=True); t.start()
    try:
        p = subprocess.run([cli] + args_before + ["--fifo", str(path)] + args_after, input=data, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
    finally:
        t.join(10)
        for fd in keep:
            os.close(fd)
    assert p.returncode == 0, p.stderr[-1000:]
    return p.stdout


def test_initial_tuning_through_the_control_channel(clis, tmp_path):
    ours, ref = clis
    rng = np.random.default_rng(12)
    z = (rng.uniform(-1, 1, 70_000) + 1j * rng.uniform(-1, 1, 70_000)).astype(np.complex64).tobytes()
    for cli in (ours, ref):
        plain = run_graph(cli, ["shift_addition_cc -0.085"], z)
        assert _with_fd(cli, ["shift_addition_cc"], [], "-0.085\n", z) == plain
        assert _with_fifo(cli, ["shift_addition_cc"], [], "-0.085\n", z, tmp_path) == plain
        assert _with_fd(cli, ["shift_unroll_cc"], [], "0.2\n", z) == run_graph(cli, ["shift_unroll_cc 0.2"], z)
        assert _with_fd(cli, ["bandpass_fir_fft_cc"], ["0.05"], "-0.1 0.2\n", z) == run_graph(cli, ["bandpass_fir_fft_cc -0.1 0.2 0.05"], z)
    # and the two programs agree with each other on the controlled runs
    a = np.frombuffer(_with_fd(ours, ["shift_addition_cc"], [], "0.3\n", z), np.complex64)
    b = np.frombuffer(_with_fd(ref, ["shift_addition_cc"], [], "0.3\n", z), np.complex64)
    assert a.size == b.size and rel(a, b) < 1e-5
    spectra_ours = run_graph(ours, ["fastddc_fwd_cc 8"], z); spectra_ref = run_graph(ref, ["fastddc_fwd_cc 8"], z)
    a = np.frombuffer(_with_fd(ours, ["fastddc_inv_cc"], ["8"], "-0.21\n", spectra_ours), np.complex64)
    b = np.frombuffer(_with_fd(ref, ["fastddc_inv_cc"], ["8"], "-0.21\n", spectra_ref), np.complex64)
    assert a.size == b.size and a.size > 0 and rel(a, b) < 1e-5
