#!/bin/bash
# Round-2 GPU call 9: smoke(), fastddc with the coalesced phasor table, bench line.
set -u
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r2_g9_smoke.log
python -m pytest tests/test_gpu_parity2.py tests/test_gpu_fullsize.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2_g9_tests.log
python tools/bench_configs.py c3 2>&1 | tee gpurun_out/r2_g9_c3.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fastddc -c 30 --csv --log-file gpurun_out/r2_g9_ddc3_launches.csv python tools/run_ddc3_once.py > /dev/null 2>&1
python - <<'PY' | tee -a gpurun_out/r2_g9_c3.txt
import csv
rows = list(csv.reader(open('gpurun_out/r2_g9_ddc3_launches.csv')))
h = next(r for r in rows if 'Kernel Name' in r); kn, mv = h.index('Kernel Name'), h.index('Metric Value')
for r in rows[rows.index(h) + 1:][-5:]:
    if len(r) > mv: print(r[kn].split('(')[0][:60], float(r[mv]) / 1e3)
PY
python bench.py > gpurun_out/r2_g9_bench.json 2> gpurun_out/r2_g9_bench.err; tail -2 gpurun_out/r2_g9_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_g9_bench.json') if l.startswith('{')][-1])
print('value', round(d['value']), 'frac', round(d['roofline']['frac'], 3), 'sustained', round(d['sustained']['value']), 'e2e', round(d['e2e']['value']), 'e2e_u8', round(d['e2e_u8']['value']), d['clocks'])
for e in d['extra']: print(e['name'], round(e['value']), 'Msps', round(e['kernel_ms'], 3), 'ms', e['roofline']['bound'], round(e['roofline']['frac'], 3), e['clocks'].get('sm_mhz'))
PY
du -sh gpurun_out
