#!/bin/bash
# Round-2 GPU call 28: validation of the final tree after the chain-step change -- full GPU suite, smoke, bench (both arms), memcheck, kernel table, chain kernel times
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2_g28_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2_g28_smoke.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_g28_bench_reference.json 2> gpurun_out/r2_g28_bench_reference.err
python bench.py > gpurun_out/r2_g28_bench.json 2> gpurun_out/r2_g28_bench.err; tail -2 gpurun_out/r2_g28_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_g28_bench.json') if l.startswith('{')][-1])
r = json.loads([l for l in open('gpurun_out/r2_g28_bench_reference.json') if l.startswith('{')][-1])
print('value', round(d['value']), 'frac', round(d['roofline']['frac'], 3), 'sustained', round(d['sustained']['value']), 'e2e', round(d['e2e']['value']), 'e2e_u8', round(d['e2e_u8']['value']), 'launches', d['gpu_launches'], d['clocks'])
print('reference arm', round(r['value']), r['cpu_baseline']['kind'], r['cpu_baseline']['cores'])
for e in d['extra']: print('  ', e['name'], round(e['value']), 'Msps', round(e['kernel_ms'], 3), 'ms', e['roofline']['bound'], round(e['roofline']['frac'], 3))
PY
python tools/bench_configs.py k 2>&1 | tee gpurun_out/r2_g28_kernels.txt | grep -E "K2|K5|K6"
C3_BLOCKS=592 python tools/bench_configs.py c3 2>&1 | grep cfg3 | tee gpurun_out/r2_g28_c3_592.txt
python tools/bench_configs.py c4 2>&1 | grep -E "FUSED" | tee gpurun_out/r2_g28_c4.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/r2_g28_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r2_g28_memcheck.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/r2_g28_shift_launches.csv python tools/run_shift_once.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_g28_plan_launches.csv python tools/plan_trace.py 592 3 > /dev/null 2>&1
python - <<'PY'
import csv
for f, sel in (('shift', slice(4, 10)), ('plan', slice(-7, None))):
    rows = [r for r in csv.reader(open(f'gpurun_out/r2_g28_{f}_launches.csv')) if len(r) > 10 and r[0].isdigit()]
    for r in rows[sel]: print(f, r[4][:60], r[-1])
PY
du -sh gpurun_out
