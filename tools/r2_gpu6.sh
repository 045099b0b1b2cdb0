#!/bin/bash
# Round-2 GPU call 6: full GPU suite, memcheck over every kernel, the bench line, ncu captures of the hot kernels for profiles/.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r2_g6_tests.log
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py 2>&1 | tail -4 | tee gpurun_out/r2_g6_memcheck.log
python bench.py > gpurun_out/r2_g6_bench.json 2> gpurun_out/r2_g6_bench.err; tail -3 gpurun_out/r2_g6_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_g6_bench_ref.json 2>/dev/null
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_g6_bench.json') if l.startswith('{')][-1])
print('value', round(d['value']), 'frac', round(d['roofline']['frac'], 3), 'sustained', round(d['sustained']['value']), 'e2e', round(d['e2e']['value']), 'e2e_u8', round(d['e2e_u8']['value']), d['clocks'])
for e in d['extra']: print(e['name'], round(e['value']), 'Msps', round(e['kernel_ms'], 3), 'ms', e['roofline']['bound'], round(e['roofline']['frac'], 3), e['clocks'].get('sm_mhz'))
PY
python tools/bench_configs.py k 2>&1 | tee gpurun_out/r2_g6_kernels.txt | grep -i "K2\|K5\|K6"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_g6_bench_quick_launches.csv python bench.py --quick --steps 10 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fastddc -c 24 --csv --log-file gpurun_out/r2_g6_ddc3_launches.csv python tools/run_ddc3_once.py > /dev/null 2>&1
cap() { name=$1; regex=$2; skip=$3; script=$4; ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/r2_g6_$name python $script >> gpurun_out/r2_g6_ncu.log 2>&1; }
: > gpurun_out/r2_g6_ncu.log
cap fir_cf32 "fir_bank_fast_kernel.*Lb0" 2 tools/run_fir_once.py
cap fir_u8 "fir_bank_fast_kernel.*Lb1" 2 tools/run_fir_once.py
cap ddc_v2 ddc_bank_fused2 2 tools/run_ddc_once.py
cap fold fastddc_fold 1 tools/run_ddc3_once.py
cap ifft_post fastddc_ifft_post 1 tools/run_ddc3_once.py
cap olafir16 olafir_bank_fused16 2 tools/run_cfg5_once.py
cap fft16384 fft_c2c_batch16 2 tools/run_cfg5_once.py
ls -la gpurun_out | grep r2_g6 | tail -12
