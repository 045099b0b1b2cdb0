#!/bin/bash
# Round-2 GPU call 13: fastddc inverse -- rows-per-CTA fused-I/O post kernel and FP64-pipe phasor walk, each against the form it replaces (timeline + totals), parity tests
set -u
mkdir -p gpurun_out
for v in "new CSDRB_X=0" "tiledpost CSDRB_INV_POST=0" "f32walk CSDRB_PHASOR_F64=0" "old CSDRB_INV_POST=0 CSDRB_PHASOR_F64=0"; do
  set -- $v; tag=$1; shift
  echo "== $tag"
  env "$@" CSDRB_INV_TRACE=1 python tools/run_c3_once.py 5 2>&1 | grep "inv trace" | tail -2 | tee -a gpurun_out/r2_g13_trace_$tag.txt
  env "$@" python tools/bench_configs.py c3 2>&1 | grep "cfg3" | tee gpurun_out/r2_g13_c3_$tag.txt
done
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_fullsize.py -m gpu -x -q -k "fastddc or fold or ddc" 2>&1 | tail -4 | tee gpurun_out/r2_g13_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_g13_c3_launches.csv python tools/run_c3_once.py 2 > /dev/null 2>&1
grep -o '"[a-z_0-9<>: ,A-Za-z]*fastddc[^"]*","[0-9]*","[0-9]*","([0-9, ]*)","([0-9, ]*)".*' gpurun_out/r2_g13_c3_launches.csv | awk -F'"' '{print $2, $(NF-1)}' | cut -c1-60,200- | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fastddc_ifft_rows_kernel -s 1 -c 1 -f -o gpurun_out/r2_g13_ifft_rows python tools/run_c3_once.py 2 > gpurun_out/r2_g13_ncu.log 2>&1
python tools/ncu_summary.py --out gpurun_out gpurun_out/r2_g13_ifft_rows.ncu-rep 2>&1 | tail -1
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
