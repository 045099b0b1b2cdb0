#!/bin/bash
# Round-2 GPU call 25: straight-line wrap step in the chain kernels -- parity, chain kernel times, cfg3 / K2 / cfg4 with it
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_fullsize.py tests/test_gpu_shift_variants.py tests/test_gpu_zz_control.py tests/test_gpu_zz_shift_table.py tests/test_gpu_zz_shift_math.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2_g25_tests.log
for nb in 256 592; do
  CSDRB_INV_TRACE=1 python tools/plan_trace.py $nb 4 2>&1 | grep "plan trace" | tail -1 | sed "s/^/[$nb] /" | tee -a gpurun_out/r2_g25_plan_trace.txt
  C3_BLOCKS=$nb python tools/bench_configs.py c3 2>&1 | grep -E "cfg3" | sed "s/^/[$nb] /" | tee -a gpurun_out/r2_g25_c3.txt
done
python tools/bench_configs.py k 2>&1 | tee gpurun_out/r2_g25_kernels.txt | grep -E "K2"
CSDRB_SHIFT_SLICES=1 python tools/bench_configs.py k 2>&1 | grep -E "K2" | sed "s/^/[one stream] /" | tee -a gpurun_out/r2_g25_kernels.txt
python tools/bench_configs.py c4 2>&1 | grep -E "FUSED" | tee gpurun_out/r2_g25_c4.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/r2_g25_shift_launches.csv python tools/run_shift_once.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_g25_plan_launches.csv python tools/plan_trace.py 592 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_g25_ddc_launches.csv python tools/run_ddc_once.py > /dev/null 2>&1
python - <<'PY'
import csv
for f, sel in (('shift', slice(4, 10)), ('plan', slice(-7, None)), ('ddc', slice(-8, None))):
    rows = [r for r in csv.reader(open(f'gpurun_out/r2_g25_{f}_launches.csv')) if len(r) > 10 and r[0].isdigit()]
    for r in rows[sel]: print(f, r[4][:60], r[-1])
PY
du -sh gpurun_out
