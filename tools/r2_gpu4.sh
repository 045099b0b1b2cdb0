#!/bin/bash
# Round-2 GPU call 4: full GPU suite (incl. tests/test_gpu_round2.py), the new bench.py line, fastddc launch breakdown.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2_g4_tests.log
python bench.py > gpurun_out/r2_g4_bench.json 2> gpurun_out/r2_g4_bench.err; tail -3 gpurun_out/r2_g4_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_g4_bench_ref.json 2>/dev/null
python tools/bench_configs.py k c3 2>&1 | tee gpurun_out/r2_g4_kernels.txt | grep -i "K2\|cfg3"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fastddc|ifft_post" -c 30 --csv --log-file gpurun_out/r2_g4_ddc3_launches.csv python tools/run_ddc3_once.py > /dev/null 2>&1
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_g4_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'sustained', d['sustained']['value'], 'e2e', d['e2e']['value'], 'e2e_u8', d['e2e_u8']['value'], d['clocks'])
for e in d['extra']: print(e['name'], round(e['value']), 'Msps', round(e['kernel_ms'], 3), 'ms', e['roofline']['bound'], round(e['roofline']['frac'], 3), e['clocks'].get('sm_mhz'), e.get('sweep'))
PY
