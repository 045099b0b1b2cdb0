#!/bin/bash
# Round-2 GPU call 27: eight chain warps per CTA (the guests sit on 8 SMs instead of 64) -- look-ahead orders again, K2, cfg4, parity of everything that has a chain
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_fullsize.py tests/test_gpu_shift_variants.py tests/test_gpu_zz_control.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2_g27_tests.log
for o in f d; do
  CSDRB_PLAN_ORDER=$o CSDRB_INV_TRACE=1 python tools/plan_trace.py 592 4 2>&1 | grep "plan trace" | tail -1 | sed "s/^/[order $o] /" | tee -a gpurun_out/r2_g27_plan_trace.txt
  CSDRB_PLAN_ORDER=$o C3_BLOCKS=592 python tools/bench_configs.py c3 2>&1 | grep -E "plan|fwd\+inv" | sed "s/^/[order $o] /" | tee -a gpurun_out/r2_g27_c3.txt
done
python tools/bench_configs.py k 2>&1 | grep -E "K2" | tee gpurun_out/r2_g27_k2.txt
CSDRB_SHIFT_SLICES=1 python tools/bench_configs.py k 2>&1 | grep -E "K2" | sed "s/^/[one stream] /" | tee -a gpurun_out/r2_g27_k2.txt
python tools/bench_configs.py c4 2>&1 | grep -E "FUSED" | tee gpurun_out/r2_g27_c4.txt
du -sh gpurun_out
