#!/bin/bash
# Round-2 GPU call 31 (2 GPUs): csdr-bankd --devices 0,1 with the NFM tail (and the rest of the daemon tests)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zzz_bankd.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2_g31_bankd_tests.log
du -sh gpurun_out
