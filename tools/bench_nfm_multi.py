"""BASELINE config 4: 1024-channel shift + fir_decimate_cc(50, 801 taps) + fmdemod NFM bank, 128 channels per GPU, the wideband IQ
block broadcast from rank 0 over NCCL (double-buffered against compute).  Launch: torchrun --nproc-per-node N tools/bench_nfm_multi.py
Prints one JSON line on rank 0 (wideband Msamples/s; every rank demodulates all of its channels for every wideband sample)."""
import json, os, sys, time
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import csdr_b200 as cb
from csdr_b200.sharding import BankShard, SharedInputBank

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
CH_PER_GPU, D, BW, CHUNK = 128, 50, 0.005, 1024
C = CH_PER_GPU * world
N = int(os.environ.get("NFM_BLOCK", 1 << 20)); NBLK = int(os.environ.get("NFM_BLOCKS", 12))
T = cb.firdes_filter_len(BW); taps = cb.firdes_lowpass_f(T, 0.5 / D)
shard = BankShard(C, world, rank)
rates = np.linspace(-0.45, 0.45, C).astype(np.float32)[shard.start:shard.start + shard.count]
n_out = cb.fir_out_len(N, D, T)
out = torch.empty((shard.count, n_out + (n_out & 1)), dtype=torch.float32, device=dev)
ddc = cb.DdcBank(rates, D, taps, demod=True, chunk=CHUNK)          # owns phases/history; pre-pass of block k+1 overlaps block k

TAIL = int(os.environ.get("NFM_TAIL", "0"))                         # 1: the whole README.md:87 graph -- ... | limit_ff | deemphasis_nfm_ff 48000 | fastagc_ff | convert_f_s16
NFM_T = cb.deemphasis_nfm_taps(48000).size
deemph = torch.empty((shard.count, n_out), dtype=torch.float32, device=dev)
agc_state = torch.zeros((shard.count, 3), dtype=torch.float32, device=dev); agc_hist = torch.zeros((shard.count, 2, 1024), dtype=torch.float32, device=dev)
pcm = torch.empty((shard.count, ((n_out - NFM_T) // 1024) * 1024), dtype=torch.int16, device=dev)

def compute(buf, sh):
    # (a streaming caller would re-present the unconsumed tail; for the throughput measurement every block is processed whole)
    audio = ddc.process(buf, out=out)
    if not TAIL: return audio
    y = cb.deemphasis_nfm_bank_ff(audio, 48000, limit_max=1.0, out=deemph)         # limiter fused into the FIR's load
    y, _, _ = cb.fastagc_bank_ff(y, 1024, 1.0, state=agc_state, hist=agc_hist)     # state carried block to block
    return cb.convert_f_s16(y, out=pcm)                                             # only s16 audio would leave the GPU

bank = SharedInputBank(shard, lambda: torch.zeros(N, dtype=torch.complex64, device=dev), compute, src=0)
g = torch.Generator(device=dev).manual_seed(1)
blocks = [torch.view_as_complex(torch.rand((N, 2), generator=g, device=dev) * 2 - 1) for _ in range(3)] if rank == 0 else [None] * 3
def feed(k): return [blocks[i % 3] for i in range(k)]
for _ in bank.run(feed(3)): pass
torch.cuda.synchronize()
if world > 1: dist.barrier(device_ids=[local])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in bank.run(feed(NBLK)): pass
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
t = torch.tensor([ms], dtype=torch.float64, device=dev)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.item()); wide = N * NBLK / ms / 1e3
    print(json.dumps({"config": "cfg4 NFM bank: shift|fir_decimate 50 (801 taps)|fmdemod" + ("|limit|deemphasis_nfm 48000|fastagc|convert_f_s16" if TAIL else "") +
                                ", 128 ch/GPU, NCCL broadcast of the wideband block",
                      "n_gpus": world, "channels": C, "block_samples": N, "blocks": NBLK, "ms_per_block": ms / NBLK,
                      "wideband_msps": wide, "channel_msps_aggregate": wide * C, "realtime_factor_at_2p4Msps": wide / 2.4}), flush=True)
if world > 1: dist.destroy_process_group()
