#!/bin/bash
# Round-2 GPU call 20: validation of the final tree -- full GPU suite, smoke, bench (both arms), memcheck over every kernel, launch list of the bench command, kernel table
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_g20_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r2_g20_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2_g20_smoke.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_g20_bench_reference.json 2> gpurun_out/r2_g20_bench_reference.err; tail -1 gpurun_out/r2_g20_bench_reference.err
python bench.py > gpurun_out/r2_g20_bench.json 2> gpurun_out/r2_g20_bench.err; tail -2 gpurun_out/r2_g20_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_g20_bench.json') if l.startswith('{')][-1])
r = json.loads([l for l in open('gpurun_out/r2_g20_bench_reference.json') if l.startswith('{')][-1])
print('value', round(d['value']), 'frac', round(d['roofline']['frac'], 3), 'sustained', round(d['sustained']['value']), 'e2e', round(d['e2e']['value']), 'e2e_u8', round(d['e2e_u8']['value']), 'launches', d['gpu_launches'], d['clocks'])
print('reference arm', round(r['value']), r['cpu_baseline']['kind'], r['cpu_baseline']['cores'], '| cpu_baseline in own arm', round(d['cpu_baseline']['value']))
for e in d['extra']: print('  ', e['name'], round(e['value']), 'Msps', round(e['kernel_ms'], 3), 'ms', e['roofline']['bound'], round(e['roofline']['frac'], 3), e['clocks'].get('sm_mhz'))
PY
python tools/bench_configs.py k 2>&1 | tee gpurun_out/r2_g20_kernels.txt | cut -c1-150
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/r2_g20_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2_g20_memcheck.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_g20_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-extra > gpurun_out/r2_g20_bench_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2_g20_bench_launches.csv')) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    k = r[4][:70]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[-1]) / 1e3
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:8]: print(f"{v[0]:5d} x {k:70s} {v[1]:10.1f} us {100 * v[1] / tot:5.1f} %")
PY
du -sh gpurun_out
