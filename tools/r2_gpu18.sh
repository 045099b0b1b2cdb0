#!/bin/bash
# Round-2 GPU call 18: what runs next to the fold -- plan timelines and loop times per look-ahead order; K2 with the first-form main kernel in three slices
set -u
mkdir -p gpurun_out
for o in d f s i; do
  CSDRB_PLAN_ORDER=$o CSDRB_INV_TRACE=1 python tools/plan_trace.py 592 4 2>&1 | grep "plan trace" | tail -1 | sed "s/^/[order $o] /" | tee -a gpurun_out/r2_g18_plan_trace.txt
  CSDRB_PLAN_ORDER=$o C3_BLOCKS=592 python tools/bench_configs.py c3 2>&1 | grep -E "plan" | sed "s/^/[order $o] /" | tee -a gpurun_out/r2_g18_c3.txt
done
python tools/bench_configs.py k 2>&1 | grep -E "K2" | tee gpurun_out/r2_g18_k2.txt
CSDRB_SHIFT_SLICES=1 python tools/bench_configs.py k 2>&1 | grep -E "K2" | sed "s/^/[one stream] /" | tee -a gpurun_out/r2_g18_k2.txt
CSDRB_SHIFT_SLICES=2 python tools/bench_configs.py k 2>&1 | grep -E "K2" | sed "s/^/[two slices] /" | tee -a gpurun_out/r2_g18_k2.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_shift_variants.py tests/test_gpu_fullsize.py -m gpu -x -q -k "shift" 2>&1 | tail -2
du -sh gpurun_out
