#!/bin/bash
# Round-2 GPU call 19: four channels per chain warp (fastddc state chain, shift chains), plan look-ahead behind the fold: timelines, loop times, K2, parity
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_fullsize.py tests/test_gpu_shift_variants.py tests/test_gpu_zz_control.py -m gpu -x -q -k "fastddc or fold or ddc or shift or retune" 2>&1 | tail -3 | tee gpurun_out/r2_g19_tests.log
for nb in 256 592; do
  CSDRB_INV_TRACE=1 python tools/plan_trace.py $nb 4 2>&1 | grep "plan trace" | tail -1 | sed "s/^/[$nb] /" | tee -a gpurun_out/r2_g19_plan_trace.txt
  C3_BLOCKS=$nb python tools/bench_configs.py c3 2>&1 | grep -E "cfg3" | sed "s/^/[$nb] /" | tee -a gpurun_out/r2_g19_c3.txt
done
CSDRB_INV_TRACE=1 python tools/run_c3_once.py 3 2>&1 | grep "inv trace" | tail -1 | tee gpurun_out/r2_g19_inv_trace.txt
python tools/bench_configs.py k 2>&1 | tee gpurun_out/r2_g19_kernels.txt | grep -E "K2"
CSDRB_SHIFT_SLICES=1 python tools/bench_configs.py k 2>&1 | grep -E "K2" | sed "s/^/[one stream] /" | tee -a gpurun_out/r2_g19_kernels.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/r2_g19_shift_launches.csv python tools/run_shift_once.py > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2_g19_shift_launches.csv')) if len(r) > 10 and r[0].isdigit()]
for r in rows[4:12]: print(r[4][:60], r[-1])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_g19_plan_launches.csv python tools/plan_trace.py 592 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2_g19_plan_launches.csv')) if len(r) > 10 and r[0].isdigit()]
for r in rows[-8:]: print(r[4][:60], r[-1])
PY
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_g19_bench.json 2> gpurun_out/r2_g19_bench.err; tail -2 gpurun_out/r2_g19_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_g19_bench.json') if l.startswith('{')][-1])
print('value', round(d['value']), 'frac', round(d['roofline']['frac'], 3), 'e2e', round(d['e2e']['value']), 'e2e_u8', round(d['e2e_u8']['value']), d['clocks'])
for e in d['extra']: print('  ', e['name'], round(e['value']), 'Msps', round(e['kernel_ms'], 3), 'ms', e['roofline']['bound'], round(e['roofline']['frac'], 3), e['clocks'].get('sm_mhz'))
PY
du -sh gpurun_out
