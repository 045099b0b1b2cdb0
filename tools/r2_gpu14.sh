#!/bin/bash
# Round-2 GPU call 14: fastddc inverse plan (look-ahead) -- parity vs the stateless bank, timings, launch list; fold operand order A/B
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_fullsize.py -m gpu -x -q -k "fastddc or fold or ddc" 2>&1 | tail -4 | tee gpurun_out/r2_g14_tests.log
for v in "hfirst CSDRB_X=0" "xfirst CSDRB_FOLD_HFIRST=0"; do
  set -- $v; tag=$1; shift
  echo "== $tag"
  env "$@" CSDRB_INV_TRACE=1 python tools/run_c3_once.py 4 2>&1 | grep "inv trace" | tail -2 | tee gpurun_out/r2_g14_trace_$tag.txt
  env "$@" python tools/bench_configs.py c3 2>&1 | grep "cfg3" | tee gpurun_out/r2_g14_c3_$tag.txt
done
C3_BLOCKS=592 python tools/bench_configs.py c3 2>&1 | grep "cfg3" | sed "s/^/[592 blocks] /" | tee gpurun_out/r2_g14_c3_592.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_g14_c3_plan_launches.csv python tools/run_c3_once.py 3 plan > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2_g14_c3_plan_launches.csv')) if len(r) > 10 and r[0].isdigit()]
for r in rows[-12:]: print(r[4][:70], r[-1])
PY
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_shift_variants.py tests/test_gpu_fullsize.py -m gpu -x -q -k "shift" 2>&1 | tail -3 | tee gpurun_out/r2_g14_shift_tests.log
python tools/bench_configs.py k 2>&1 | tee gpurun_out/r2_g14_kernels.txt | grep -E "K2|K5|K6|K1\+K3"
CSDRB_SHIFT_SLICES=1 python tools/bench_configs.py k 2>&1 | grep -E "K2" | sed "s/^/[one stream] /" | tee -a gpurun_out/r2_g14_kernels.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_g14_shift_launches.csv python tools/run_shift_once.py > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r2_g14_shift_launches.csv')) if len(r) > 10 and r[0].isdigit()]
for r in rows[-18:]: print(r[4][:60], r[-1])
PY
python tools/bench_configs.py c4 2>&1 | grep -E "FUSED" | tee gpurun_out/r2_g14_c4.txt
timeout 600 python -m pytest tests/test_gpu_parity2.py tests/test_gpu_round2.py tests/test_gpu_fullsize.py tests/test_gpu_nfm_tail.py -m gpu -x -q -k "ddc or nfm or bank" 2>&1 | tail -3 | tee gpurun_out/r2_g14_ddc_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_g14_bench.json 2> gpurun_out/r2_g14_bench.err; tail -2 gpurun_out/r2_g14_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_g14_bench.json') if l.startswith('{')][-1])
print('value', round(d['value']), 'frac', round(d['roofline']['frac'], 3), 'e2e', round(d['e2e']['value']), 'e2e_u8', round(d['e2e_u8']['value']), d['clocks'])
for e in d['extra']: print('  ', e['name'], round(e['value']), 'Msps', round(e['kernel_ms'], 3), 'ms', e['roofline']['bound'], round(e['roofline']['frac'], 3), e['clocks'].get('sm_mhz'))
PY
du -sh gpurun_out
