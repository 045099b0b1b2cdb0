#!/bin/bash
# Round-2 GPU call 23: csdr-bankd end to end with a long signal (start-up taken out) against the reference process chain
set -u
mkdir -p gpurun_out
timeout 1200 python tools/bench_bankd.py 128 160 2>&1 | tee gpurun_out/r2_g23_bankd_128ch.txt
du -sh gpurun_out
