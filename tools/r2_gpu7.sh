#!/bin/bash
# Round-2 GPU call 7: packed-fp32 FFT butterflies (A/B against call 6), fastddc launch breakdown, ncu summaries made ON the box (reports stay there).
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity2.py tests/test_gpu_fullsize.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_g7_tests.log
python tools/bench_configs.py k c3 c5 2>&1 | tee gpurun_out/r2_g7_kernels.txt | grep -i "K7\|cfg3\|cfg5"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fastddc -c 24 --csv --log-file gpurun_out/r2_g7_ddc3_launches.csv python tools/run_ddc3_once.py > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r2_g7_ddc3_launches.csv')))
h = next(r for r in rows if 'Kernel Name' in r); kn, mv = h.index('Kernel Name'), h.index('Metric Value')
for r in rows[rows.index(h) + 1:][-8:]:
    if len(r) > mv: print(r[kn].split('(')[0][:60], float(r[mv]) / 1e3)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_g7_bench_quick_launches.csv python bench.py --quick --steps 10 > /dev/null 2>&1
cap() { name=$1; regex=$2; skip=$3; script=$4
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o /tmp/r2_g7_$name python $script >> gpurun_out/r2_g7_ncu.log 2>&1
  python tools/ncu_summary.py --out gpurun_out /tmp/r2_g7_$name.ncu-rep | tee -a gpurun_out/r2_g7_ncu.log; rm -f /tmp/r2_g7_$name.ncu-rep; }
: > gpurun_out/r2_g7_ncu.log
cap fir_cf32 "fir_bank_fast_kernel.*Lb0" 2 tools/run_fir_once.py
cap fir_u8 "fir_bank_fast_kernel.*Lb1" 2 tools/run_fir_once.py
cap ddc_v2 ddc_bank_fused2 2 tools/run_ddc_once.py
cap fold fastddc_fold 1 tools/run_ddc3_once.py
cap ifft_post fastddc_ifft_post 1 tools/run_ddc3_once.py
cap olafir16 olafir_bank_fused16 2 tools/run_cfg5_once.py
cap fft16384 fft_c2c_batch16 2 tools/run_cfg5_once.py
cap shift_bank "shift_bank_kernel" 1 tools/run_shift_once.py
ls -la gpurun_out | grep r2_g7 | tail -14; du -sh gpurun_out
