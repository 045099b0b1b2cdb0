"""GPU parity tests (-m gpu) for SURVEY 8(f) rank 3: shift_addfast_cc (libcsdr.c:307-317, 396-433) as the libcsdr drop-in and as a bank.
The kernel replays the reference's separately rounded float recursion, so given the same step table it reproduces the compiled
reference's stream up to an occasional one-ulp difference in a call's cos/sin seed."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
GOLD = np.load(Path(__file__).parent / "golden" / "hotpath_golden.npz")


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


def _cplx(rng, n):
    return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)


def test_shift_addfast_dropin_against_golden_and_oracle(gpu, oracle):
    # the reference build's own step table -> its own stream (golden), carried phase bit-exact
    y, ph = gpu.libcsdr.shift_addfast_cc(GOLD["shift_in"], phase=0.0, chunk=1024, steps=GOLD["addfast_steps"])
    assert _rel(y, GOLD["addfast_out"]) < 1e-7 and np.float32(ph) == GOLD["addfast_phase"]
    # our init (= strict oracle init, <= 1 ulp from the reference build's libmvec one): the 256-step recursion keeps that below the bar
    y, ph = gpu.libcsdr.shift_addfast_cc(GOLD["shift_in"], -0.085, 0.0, 1024)
    assert _rel(y, GOLD["addfast_out"]) < 1e-5 and np.float32(ph) == GOLD["addfast_phase"]
    rng = np.random.default_rng(5)
    for n, chunk, rate, ph0 in ((16384, 1024, 0.2, 0.0), (5000, 1000, -0.31, 1.5), (4099, 4099, 0.4999, -3.0), (1022, 1024, 1e-4, 0.3), (3, 1024, 0.1, 0.2)):
        x = _cplx(rng, n)
        y, ph = gpu.libcsdr.shift_addfast_cc(x, rate, ph0, chunk)
        want, wph = oracle.shift_addfast_cc(x, rate, ph0, chunk)
        assert np.float32(ph) == np.float32(wph), (n, chunk, rate)
        assert _rel(y, want) < 1e-7 if n >= 4 else np.all(y == 0), (n, chunk, rate)
        tail = n - (n % chunk) + ((n % chunk) & ~3) if n % chunk else n           # the last call leaves its n % 4 tail untouched
        assert np.all(y[tail:] == 0) and np.all(want[tail:] == 0)


def test_shift_addfast_bank_replays_reference_state_chain(gpu, oracle):
    rng = np.random.default_rng(11)
    rates = np.array([-0.41, -0.085, 0.0, 0.2, 0.4999, 1e-4, 0.25], np.float32)
    for n, chunk in ((16384 + 777, 1024), (50_000, 1000), (4096, 4096), (3000, 0), (1001, 37)):
        # shared wideband input
        x = _cplx(rng, n)
        ph0 = rng.uniform(-3, 3, rates.size).astype(np.float32)
        y, ph = gpu.shift_addfast_bank_cc(torch.from_numpy(x).cuda(), rates, phases=torch.from_numpy(ph0).cuda(), chunk=chunk)
        y = y.cpu().numpy(); ph = ph.cpu().numpy()
        for c, r in enumerate(rates):
            want, wph = oracle.shift_addfast_cc(x, float(r), float(ph0[c]), chunk or None)
            assert np.float32(wph) == ph[c], (n, chunk, c)
            assert _rel(y[c], want) < 1e-7, (n, chunk, c, _rel(y[c], want))
            assert np.array_equal(y[c] == 0, want == 0)                           # untouched tails in the same places
    # one row per channel
    xs = np.stack([_cplx(rng, 20_000) for _ in rates])
    y, ph = gpu.shift_addfast_bank_cc(torch.from_numpy(xs).cuda(), rates, chunk=1024)
    for c, r in enumerate(rates):
        want, wph = oracle.shift_addfast_cc(xs[c], float(r), 0.0, 1024)
        assert _rel(y[c].cpu().numpy(), want) < 1e-7 and np.float32(wph) == ph[c].item()
