"""Multi-GPU layout of a channel bank: one process per GPU, contiguous channel slices, and -- only
for shared-input banks -- one broadcast of the wideband IQ block per step.

Channels are independent units (each reference channel is literally a separate process chain, ddcd_old.h:51-61), so
  * independent-input banks (fir_decimate bank, bandpass_fir_fft bank): every rank owns the inputs of its slice; NO collective;
  * shared-input banks (NFM/DDC bank, fastddc): the block is broadcast from the rank that ingests it (what `nmux`
    does over TCP in the reference, nmux.cpp:246-353), double-buffered so the copy of block k+1 overlaps the
    compute of block k; outputs stay on the owning rank (1/D of the input rate).
Backend: torch.distributed -- "nccl" over NVLink/NVSwitch on GPUs, "gloo" for the CPU tests of this host logic.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Sequence


def channel_slice(n_channels: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous balanced slices: (first channel, count) of `rank`; the first n_channels % world ranks hold one more."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_channels, world)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def owner_of(channel: int, n_channels: int, world: int) -> int:
    base, extra = divmod(n_channels, world)
    edge = extra * (base + 1)
    return channel // (base + 1) if channel < edge else extra + (channel - edge) // max(base, 1)


@dataclass
class BankShard:
    """What one rank owns of a C-channel bank."""
    n_channels: int
    world: int
    rank: int

    def __post_init__(self):
        self.start, self.count = channel_slice(self.n_channels, self.world, self.rank)

    @property
    def channels(self) -> range:
        return range(self.start, self.start + self.count)

    def take(self, per_channel: Sequence):
        """Slice any per-channel parameter list (rates, shifts, taps rows ...) down to this rank."""
        return per_channel[self.start:self.start + self.count]


class SharedInputBank:
    """Shared-input bank stepping: broadcast the wideband block, run this rank's slice.

    `compute(block, shard)` is the per-slice work -- on GPUs the csdr_b200 bank chain
    (shift_addition_bank_cc -> fir_decimate_bank_cc -> fmdemod_quadri_bank_cf ...); it must leave its result on the owning
    rank.  Two block buffers alternate so that the broadcast of block k+1 (async, its own NCCL stream) overlaps
    `compute` of block k.
    """

    def __init__(self, shard: BankShard, make_buffer: Callable[[], "object"], compute: Callable, src: int = 0):
        import torch.distributed as dist
        self.dist, self.shard, self.compute, self.src = dist, shard, compute, src
        self.buffers = [make_buffer(), make_buffer()]
        self.pending = None                      # (work handle, buffer index) of the broadcast in flight
        self.turn = 0

    def _more(self, have_block: bool) -> bool:
        """The SOURCE decides whether another block follows: one int travels ahead of every block (control plane), the other ranks follow it.
        Without it a rank whose own iterable is shorter, longer or empty would leave the others waiting in a broadcast forever."""
        if self.shard.world == 1:
            return have_block
        import torch
        flag = torch.tensor([1 if have_block else 0], dtype=torch.int32, device=self.buffers[0].device)
        self.dist.broadcast(flag, src=self.src)
        return bool(flag.item())

    def _post(self, block):
        buf = self.buffers[self.turn]
        if self.shard.rank == self.src:
            buf.copy_(block)
        work = self.dist.broadcast(buf, src=self.src, async_op=True) if self.shard.world > 1 else None
        self.pending = (work, self.turn)
        self.turn ^= 1

    def run(self, blocks):
        """`blocks`: iterable of wideband blocks, read on the source rank only (the other ranks may pass anything, e.g. None: how many blocks
        there are is the source's call).  Yields this rank's result per block, in order."""
        it = iter(blocks) if self.shard.rank == self.src else iter(())
        nxt = next(it, None)
        if not self._more(nxt is not None):
            return                                   # empty stream: every rank returns
        self._post(nxt)
        while True:
            nxt = next(it, None)
            more = self._more(nxt is not None)
            work, idx = self.pending
            if work is not None:
                work.wait()
            if more:
                self._post(nxt)                      # block k+1 starts moving ...
            yield self.compute(self.buffers[idx], self.shard)   # ... while block k is processed
            if not more:
                return


def gather_counts(shard: BankShard, local_count: int) -> list[int]:
    """All-gather one integer per rank (e.g. outputs written) -- control-plane only, never on the data path."""
    import torch
    import torch.distributed as dist
    if shard.world == 1:
        return [local_count]
    t = torch.tensor([local_count], dtype=torch.int64)
    backend = dist.get_backend()
    if backend == "nccl":
        t = t.cuda()
    out = [torch.zeros_like(t) for _ in range(shard.world)]
    dist.all_gather(out, t)
    return [int(o.item()) for o in out]
