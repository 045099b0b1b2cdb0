#!/bin/bash
# Round-2 GPU call 17: how the fold's time per block depends on the batch (L2 residency of spectra + folded bins?), ncu of the fold at 592 blocks; K2 with half-width tiles
set -u
mkdir -p gpurun_out
for nb in 144 288 432 576 592; do
  echo "== $nb blocks"; CSDRB_INV_TRACE=1 python tools/plan_trace.py $nb 4 2>&1 | grep "plan trace" | tail -1 | sed "s/^/[$nb] /" | tee -a gpurun_out/r2_g17_plan_trace.txt
  C3_BLOCKS=$nb python tools/bench_configs.py c3 2>&1 | grep -E "one loop" | sed "s/^/[$nb] /" | tee -a gpurun_out/r2_g17_c3.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fastddc_fold_kernel -s 2 -c 1 -f -o gpurun_out/r2_g17_fold592 python tools/plan_trace.py 592 3 > gpurun_out/r2_g17_ncu.log 2>&1
python tools/ncu_summary.py --out gpurun_out gpurun_out/r2_g17_fold592.ncu-rep 2>&1 | tail -1
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_fold592_ncu_summary.json')); m = d['metrics']
print('fold @592:', d['duration_us'], 'us, dram MB', round(d['dram_bytes_per_launch'] / 1e6, 1), 'L2 hit', m['lts__t_sector_hit_rate.pct']['value'], 'fma', m['sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active']['value'], d['stall_sampling']['by_reason_pct'])
PY
rm -f gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_shift_variants.py tests/test_gpu_fullsize.py -m gpu -x -q -k "shift" 2>&1 | tail -3 | tee gpurun_out/r2_g17_shift_tests.log
python tools/bench_configs.py k 2>&1 | grep -E "K2" | tee gpurun_out/r2_g17_k2.txt
CSDRB_SHIFT_SLICES=1 python tools/bench_configs.py k 2>&1 | grep -E "K2" | sed "s/^/[one stream] /" | tee -a gpurun_out/r2_g17_k2.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/r2_g17_shift_launches.csv python tools/run_shift_once.py > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2_g17_shift_launches.csv')) if len(r) > 10 and r[0].isdigit()]
for r in rows[:8]: print(r[4][:60], r[-1])
PY
du -sh gpurun_out
