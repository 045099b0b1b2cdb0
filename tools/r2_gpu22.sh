#!/bin/bash
# Round-2 GPU call 22: csdr-bankd end to end against the reference process chain; bench line after the roofline relabelling of the fastddc / overlap-add legs
set -u
mkdir -p gpurun_out
nproc > gpurun_out/r2_g22_nproc.txt
timeout 900 python tools/bench_bankd.py 128 20 2>&1 | tee gpurun_out/r2_g22_bankd_128ch.txt
timeout 600 python tools/bench_bankd.py 16 20 2>&1 | head -1 | tee gpurun_out/r2_g22_bankd_16ch.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_g22_bench.json 2> gpurun_out/r2_g22_bench.err; tail -2 gpurun_out/r2_g22_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_g22_bench.json') if l.startswith('{')][-1])
print('value', round(d['value']), 'frac', round(d['roofline']['frac'], 3), 'e2e', round(d['e2e']['value']), 'e2e_u8', round(d['e2e_u8']['value']), d['clocks'])
for e in d['extra']: print('  ', e['name'], round(e['value']), 'Msps', round(e['kernel_ms'], 3), 'ms', e['roofline']['bound'], round(e['roofline']['frac'], 3), 'of hbm', round(e['roofline'].get('frac_of_hbm', e['roofline']['frac']), 3))
print(d['extra'][-1]['sweep'])
PY
du -sh gpurun_out
