"""CPU tests of the N>1 host logic with torch.distributed/gloo, world_size 2 (and pure-python checks of the slicing).
The per-slice compute is played by the CPU oracle here (tests may use it); on GPUs it is the csdr_b200 bank chain."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from csdr_b200.sharding import BankShard, channel_slice, owner_of  # noqa: E402


def test_channel_slices_are_contiguous_balanced_and_cover():
    for C in (1, 2, 7, 64, 256, 1000, 1024):
        for W in (1, 2, 3, 4, 8):
            seen = []
            for r in range(W):
                s, n = channel_slice(C, W, r)
                seen.extend(range(s, s + n))
                assert all(owner_of(c, C, W) == r for c in range(s, s + n))
            assert seen == list(range(C))
            sizes = [channel_slice(C, W, r)[1] for r in range(W)]
            assert max(sizes) - min(sizes) <= 1
    assert channel_slice(1024, 8, 3) == (384, 128)            # BASELINE config 4: 128 channels per GPU
    assert channel_slice(256, 8, 7) == (224, 32)
    with pytest.raises(ValueError):
        channel_slice(4, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    from csdr_b200.sharding import BankShard, SharedInputBank, gather_counts
    from oracle.pyoracle import Oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ora = Oracle()
        C, N, D, T, NBLK = 6, 4096, 10, 79, 5
        rates = [(-0.3 + 0.1 * c) for c in range(C)]
        shard = BankShard(C, world, rank)
        taps = ora.firdes_lowpass_f(T, 0.5 / D)
        rng = np.random.default_rng(0)                       # only the source rank's data matters
        blocks = [(rng.uniform(-1, 1, N) + 1j * rng.uniform(-1, 1, N)).astype(np.complex64) for _ in range(NBLK)]
        phases = {c: 0.0 for c in shard.channels}

        def compute(buf, sh):
            x = buf.numpy().copy()
            out = []
            for c in sh.channels:                            # the slice this rank owns, and only that
                y, phases[c] = ora.shift_addition_cc(x, rates[c], phases[c], 1024)
                out.append(ora.fir_decimate_cc(y, D, taps))
            return np.stack(out) if out else np.zeros((0, 0), np.complex64)

        bank = SharedInputBank(shard, lambda: torch.zeros(N, dtype=torch.complex64), compute, src=0)
        # only the SOURCE's iterable counts: rank 1 passes a wrong-length one on purpose (round 1 deadlocked on that), and an empty stream ends at once
        assert list(bank.run([] if rank == 0 else [None] * 3)) == []
        feed = [torch.from_numpy(b) for b in blocks] if rank == 0 else [None] * (NBLK + 2)
        results = list(bank.run(feed))
        counts = gather_counts(shard, sum(r.shape[0] for r in results))
        q.put((rank, shard.start, shard.count, [r.copy() for r in results], counts))
    finally:
        dist.destroy_process_group()


def test_shared_input_bank_two_ranks_matches_unsharded():
    import torch.multiprocessing as mp
    from oracle.pyoracle import Oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    got.sort()
    # unsharded truth
    ora = Oracle()
    C, N, D, T, NBLK = 6, 4096, 10, 79, 5
    rates = [(-0.3 + 0.1 * c) for c in range(C)]
    taps = ora.firdes_lowpass_f(T, 0.5 / D)
    rng = np.random.default_rng(0)
    blocks = [(rng.uniform(-1, 1, N) + 1j * rng.uniform(-1, 1, N)).astype(np.complex64) for _ in range(NBLK)]
    phases = [0.0] * C
    covered = []
    for rank, start, count, results, counts in got:
        assert (start, count) == channel_slice(C, 2, rank) and counts == [NBLK * 3, NBLK * 3]
        assert len(results) == NBLK
        covered.extend(range(start, start + count))
    assert covered == list(range(C))
    for k in range(NBLK):
        for c in range(C):
            y, phases[c] = ora.shift_addition_cc(blocks[k], rates[c], phases[c], 1024)
            want = ora.fir_decimate_cc(y, D, taps)
            rank = owner_of(c, C, 2)
            _, start, _, results, _ = got[rank]
            assert np.array_equal(results[k][c - start], want), (k, c)      # same oracle, same broadcast bytes -> identical
