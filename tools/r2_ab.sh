#!/bin/bash
# Round-2 first GPU call (one B200): validate the branch, then A/B the kernels that only ran under the CPU emulator so far.
#   gpurun --timeout 900 -- 'bash tools/r2_ab.sh'
# Everything lands in gpurun_out/r2_ab_*.  Nothing here is a bench value for bench.py; it decides what gets merged.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2_ab_tests.log
for staged in 0 1; do
  for tile in 44 22; do
    tag="staged${staged}_tile${tile}"
    env $( [ $staged = 1 ] && echo CSDRB_OLAFIR_STAGED=1 ) CSDRB_INV_TILE=$tile python tools/bench_configs.py c3 c5 2>&1 | tee gpurun_out/r2_ab_configs_${tag}.txt | tail -12
    cp gpurun_out/configs.json gpurun_out/r2_ab_configs_${tag}.json 2>/dev/null
  done
done
# config 5 with the radix-16 overlap-add kernel (4096 = 16^3: 8 barriers and 4R+4W shared accesses per block instead of 12 / 6R+6W, round 1: 20 / 11R+11W)
CSDRB_FFT_RADIX16=1 python tools/bench_configs.py c3 c5 2>&1 | tee gpurun_out/r2_ab_configs_radix16.txt | tail -10
python tools/bench_configs.py k 2>&1 | tee gpurun_out/r2_ab_kernels.txt | tail -20        # K2 (staged row loads on this branch), K7 FFT sizes, ...
CSDRB_FFT_RADIX16=1 python tools/bench_configs.py k 2>&1 | grep -i "fft\|K7" | tee gpurun_out/r2_ab_fft_radix16.txt
# compare gpurun_out/r2_ab_kernels.txt with profiles/r01_configs_latest.txt (main, round 1) for the K2 / de-emphasis / FFT lines
timeout 120 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py 2>&1 | tail -3 | tee gpurun_out/r2_ab_memcheck.log
