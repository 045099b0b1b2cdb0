#!/usr/bin/env python
"""bench.py -- headline benchmark: 256-channel fir_decimate_cc bank (decim 10, 199-tap HAMMING), BASELINE config 2.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch: 256 channels x 2.4 M cf32 samples per GPU (1 s of a
2.4 Msps stream per channel).  Multi-GPU: a 256*N-channel bank, one contiguous 256-channel slice per device,
no data-path collective (independent inputs) -> weak scaling.

Prints ONE JSON line (rank 0):
  value      whole-job Msamples/s of complex input, inputs resident in HBM, device-timed (CUDA events), max over ranks
  e2e        same metric through the host-buffer C-ABI call (csdrb_fir_decimate_bank_cc_host): pinned host input,
             H2D + kernel + D2H of the outputs inside the timed region, every step
  roofline   algorithmic bytes (8 B/sample in + 8 B/output) / measured kernel time vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the compiled reference (oracle/_ref) or the oracle port on this box's host cores, bounded sample
--impl reference: the reference's own CPU implementation of the same step on all host threads (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CHANNELS, N_IN, TAPS, DECIM = 256, 2_400_000, 199, 10
WORKLOAD = "256-channel fir_decimate_cc bank, decim=10, 199-tap HAMMING, 2.4 Msps/ch synthetic cf32 (BASELINE configs[1])"
METRIC = "Msamples/s in, 256-ch fir_decimate_cc d=10"
ALGO_BYTES_PER_SAMPLE = 8.0 + 8.0 / DECIM            # 8 B read per input sample + 8 B written per output


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    p = ROOT / "profiles" / "fir_bank_ncu_summary.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def physical_gpu_index(local: int) -> int:
    cvd = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    parts = [p for p in cvd.split(",") if p.strip()]
    if parts and local < len(parts) and parts[local].strip().isdigit():
        return int(parts[local])
    return local


# ----------------------------------------------------------------------------------------------------
def cpu_reference_runner():
    """(callable, kind): fir_decimate_cc of the compiled reference if oracle/_ref is there, else the oracle port."""
    import ctypes as C
    from oracle import pyoracle
    if pyoracle.have_ref():
        L = C.CDLL(str(pyoracle.REF_SO)); fn = L.fir_decimate_cc; kind = "reference"
    else:
        pyoracle.build(ref=False)
        L = C.CDLL(str(pyoracle.ORACLE_SO)); fn = L.oracle_fir_decimate_cc; kind = "port"
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    fn.restype = C.c_int
    return fn, kind


def cpu_fir_threads(fn, taps, n_threads, channels_per_thread, reps=1):
    """Every thread filters `channels_per_thread` synthetic channels of N_IN samples; returns (seconds, samples)."""
    rng = np.random.default_rng(0)
    x = (rng.uniform(-1, 1, 2 * N_IN).astype(np.float32))          # one shared read-only input (interleaved I,Q)
    outs = [np.empty(2 * (N_IN // DECIM + 1), np.float32) for _ in range(n_threads)]
    start = threading.Barrier(n_threads + 1)

    def work(k):
        start.wait()
        for _ in range(channels_per_thread * reps):
            fn(x.ctypes.data, outs[k].ctypes.data, N_IN, DECIM, taps.ctypes.data, TAPS)

    th = [threading.Thread(target=work, args=(k,)) for k in range(n_threads)]
    [t.start() for t in th]
    start.wait(); t0 = time.perf_counter()
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    return dt, n_threads * channels_per_thread * reps * N_IN


def host_taps():
    """Taps for the CPU arms, designed by the CPU library itself (no GPU library needed)."""
    from oracle import pyoracle
    o = pyoracle.Ref() if pyoracle.have_ref() else pyoracle.Oracle()
    return np.ascontiguousarray(o.firdes_lowpass_f(TAPS, 0.5 / DECIM, "HAMMING"), np.float32)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fn, kind = cpu_reference_runner()
    taps = host_taps()
    cores = os.cpu_count() or 1
    for _ in range(max(args.warmup, 1)):
        cpu_fir_threads(fn, taps, cores, 1)
    times, samples = [], 0
    for _ in range(args.steps):
        dt, s = cpu_fir_threads(fn, taps, cores, 1)
        times.append(dt); samples += s
    total = sum(times)
    value = samples / total / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "taps": TAPS, "decimation": DECIM, "samples_per_channel": N_IN,
                       "step_sample": f"{cores} channels x {N_IN} samples per step (one channel per host thread)"},
            "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": kind,
                             "sample": f"{args.steps} steps x {cores} channels x {N_IN} samples, reference fir_decimate_cc, {cores} threads"},
            "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import csdr_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    csdr_b200.lib()

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    taps = csdr_b200.firdes_lowpass_f(TAPS, 0.5 / DECIM, "HAMMING")
    n_out = csdr_b200.fir_out_len(N_IN, DECIM, TAPS)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    x = torch.rand((CHANNELS, N_IN, 2), generator=g, device=dev, dtype=torch.float32) * 2 - 1      # 4.9 GB >> 126 MB L2
    y = torch.empty((CHANNELS, n_out + (n_out & 1)), dtype=torch.complex64, device=dev)

    # ---- device-resident leg ---------------------------------------------------------------------
    for _ in range(args.warmup):
        csdr_b200.fir_decimate_bank_cc(x, DECIM, taps, out=y)
    barrier()
    launches0 = csdr_b200.kernel_launches()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(physical_gpu_index(local)) as clk:
        barrier()
        ev[0].record()
        for k in range(args.steps):
            csdr_b200.fir_decimate_bank_cc(x, DECIM, taps, out=y)
            ev[k + 1].record()
        barrier()
        # keep the GPU loaded a little longer when the timed region is shorter than the sampler period
        t_hold = time.perf_counter()
        while not args.quick and time.perf_counter() - t_hold < 0.6:
            csdr_b200.fir_decimate_bank_cc(x, DECIM, taps, out=y)
        torch.cuda.synchronize()
    launches = csdr_b200.kernel_launches() - launches0
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[args.steps]))
    per_kernel_ms = float(np.mean([ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]))
    per_kernel_ms = max_over_ranks(per_kernel_ms)
    samples_per_step = CHANNELS * N_IN
    value = world * samples_per_step * args.steps / total_ms / 1e3                                   # Msamples/s, whole job

    # ---- end-to-end leg: pinned host buffers through the C-ABI host call -------------------------
    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "value": value, "unit": "Msamples/s", "kernel_ms": per_kernel_ms, "gpu_launches": int(launches)}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    hin = csdr_b200.PinnedArray((CHANNELS, N_IN), np.complex64)
    hout = csdr_b200.PinnedArray((CHANNELS, n_out), np.complex64)
    stage = torch.from_numpy(hin.array.view(np.float32).reshape(CHANNELS, N_IN, 2))
    stage.copy_(x)                                                                                     # same synthetic data, now on the host
    del x, y
    torch.cuda.empty_cache()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(1):
        csdr_b200.fir_decimate_bank_cc_host(hin.array, DECIM, taps, out=hout.array)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        csdr_b200.fir_decimate_bank_cc_host(hin.array, DECIM, taps, out=hout.array)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    if world > 1:
        dist.barrier(device_ids=[local])
    e2e_value = world * samples_per_step * e2e_steps / e2e_s / 1e6
    checksum = float(np.abs(hout.array[:, :: max(1, n_out // 64)]).sum())                            # the step's result is read on the host
    hin.close(); hout.close()

    # ---- CPU baseline beside it (rank 0, N=1 only, bounded) --------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fn, kind = cpu_reference_runner()
        cores = os.cpu_count() or 1
        cpu_fir_threads(fn, taps, cores, 1)
        dt, s = 0.0, 0
        reps = 0
        while dt < 8.0 and reps < 64:
            d, n = cpu_fir_threads(fn, taps, cores, 1)
            dt += d; s += n; reps += 1
        cpu = {"value": s / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": kind,
               "sample": f"{reps} x {cores} channels x {N_IN} samples of the same workload, fir_decimate_cc of "
                         f"{'oracle/_ref (unmodified reference build)' if kind == 'reference' else 'oracle port'}, {cores} host threads, {dt:.1f} s"}

    if rank == 0:
        peak, peak_src = peaks()
        algo_bytes = samples_per_step * 8.0 + CHANNELS * n_out * 8.0
        achieved = algo_bytes / (per_kernel_ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "channels_per_gpu": CHANNELS, "samples_per_channel": N_IN, "taps": TAPS, "decimation": DECIM,
                           "parallelism": f"channel slices, {CHANNELS} per GPU x {world} GPUs, no collective",
                           "l2": "inputs (4.9 GB per GPU) exceed the 126 MB L2; no flush needed", "timing": "CUDA events on the launch stream, max over ranks"},
                "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": samples_per_step * 8, "d2h_bytes_per_step": CHANNELS * n_out * 8,
                        "steps": e2e_steps, "api": "csdrb_fir_decimate_bank_cc_host (pinned host buffers, 3-stream chunked pipeline)", "checksum": checksum},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(),
                             "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": per_kernel_ms,
                             "fp32_tflops": CHANNELS * n_out * TAPS * 4 / (per_kernel_ms * 1e-3) / 1e12,
                             "note": "8.8 algorithmic B/sample at 79.6 flop/sample sits on the FP32/HBM ridge; both are reported"},
                "cpu_baseline": cpu, "clocks": clk.summary()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="device-resident leg only (for ncu runs): no clock hold loop, no e2e, no CPU baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
