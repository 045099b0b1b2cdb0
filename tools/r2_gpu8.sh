#!/bin/bash
# Round-2 GPU call 8: fastddc with precomputed phasors, fold thread-tile A/B, launch breakdown.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity2.py tests/test_gpu_fullsize.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_g8_tests.log
for bt in 8 4; do
  CSDRB_FOLD_BT=$bt python tools/bench_configs.py c3 2>&1 | sed "s/^/[BT=$bt] /" | tee -a gpurun_out/r2_g8_c3.txt
  CSDRB_FOLD_BT=$bt ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fastddc -c 30 --csv --log-file gpurun_out/_l.csv python tools/run_ddc3_once.py > /dev/null 2>&1
  python - "$bt" <<'PY' | tee -a gpurun_out/r2_g8_c3.txt
import csv, sys
rows = list(csv.reader(open('gpurun_out/_l.csv')))
h = next(r for r in rows if 'Kernel Name' in r); kn, mv = h.index('Kernel Name'), h.index('Metric Value')
for r in rows[rows.index(h) + 1:][-5:]:
    if len(r) > mv: print('[BT=%s]' % sys.argv[1], r[kn].split('(')[0][:60], float(r[mv]) / 1e3)
PY
done
CSDRB_FOLD_BT=4 ncu --set full --clock-control none --import-source on -k regex:fastddc_fold -s 1 -c 1 -o /tmp/r2_g8_fold_bt4 python tools/run_ddc3_once.py > gpurun_out/r2_g8_ncu.log 2>&1
python tools/ncu_summary.py --out gpurun_out /tmp/r2_g8_fold_bt4.ncu-rep | tee -a gpurun_out/r2_g8_ncu.log; rm -f /tmp/*.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:fastddc_ifft_post -s 1 -c 1 -o /tmp/r2_g8_ifft_post python tools/run_ddc3_once.py >> gpurun_out/r2_g8_ncu.log 2>&1
python tools/ncu_summary.py --out gpurun_out /tmp/r2_g8_ifft_post.ncu-rep | tee -a gpurun_out/r2_g8_ncu.log; rm -f /tmp/*.ncu-rep
du -sh gpurun_out
