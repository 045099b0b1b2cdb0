#!/bin/bash
# Round-2 GPU call 10: fastddc post kernel with flat items, u8 FIR kernel timing.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity2.py tests/test_gpu_fullsize.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2_g10_tests.log
python tools/bench_configs.py c3 k 2>&1 | tee gpurun_out/r2_g10_kernels.txt | grep -i "cfg3\|K1+K3\|K2\|K5\|K6"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fastddc -c 30 --csv --log-file gpurun_out/r2_g10_ddc3_launches.csv python tools/run_ddc3_once.py > /dev/null 2>&1
python - <<'PY' | tee -a gpurun_out/r2_g10_kernels.txt
import csv
rows = list(csv.reader(open('gpurun_out/r2_g10_ddc3_launches.csv')))
h = next(r for r in rows if 'Kernel Name' in r); kn, mv = h.index('Kernel Name'), h.index('Metric Value')
for r in rows[rows.index(h) + 1:][-5:]:
    if len(r) > mv: print(r[kn].split('(')[0][:60], float(r[mv]) / 1e3)
PY
ncu --set full --clock-control none --import-source on -k regex:fastddc_ifft_post -s 1 -c 1 -o /tmp/r2_g10_ifft_post python tools/run_ddc3_once.py > gpurun_out/r2_g10_ncu.log 2>&1
python tools/ncu_summary.py --out gpurun_out /tmp/r2_g10_ifft_post.ncu-rep | tee -a gpurun_out/r2_g10_ncu.log; rm -f /tmp/*.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:"fir_bank_fast_kernel.*Lb1" -s 2 -c 1 -o /tmp/r2_g10_fir_u8 python tools/run_fir_once.py >> gpurun_out/r2_g10_ncu.log 2>&1
python tools/ncu_summary.py --out gpurun_out /tmp/r2_g10_fir_u8.ncu-rep | tee -a gpurun_out/r2_g10_ncu.log; rm -f /tmp/*.ncu-rep
du -sh gpurun_out
