#!/bin/bash
# Round-2 GPU call 5 (2 GPUs): full GPU suite incl. the multi-GPU bank / daemon tests, bench.py at N=2 (NCCL broadcast legs), fastddc launch breakdown.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r2_g5_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_g5_bench_n2.json 2> gpurun_out/r2_g5_bench_n2.err; tail -3 gpurun_out/r2_g5_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2_g5_bench_n2.json') if l.startswith('{')][-1])
    print('N=2 value', round(d['value']), 'frac', round(d['roofline']['frac'], 3), 'e2e', round(d['e2e']['value']), 'h2d/rank', round(d['e2e']['h2d_gbs_per_rank'], 1), 'e2e_u8', round(d['e2e_u8']['value']), d['clocks'])
    for e in d['extra']: print(e['name'], round(e['value']), 'Msps', round(e['kernel_ms'], 3), 'ms', e['roofline']['bound'], round(e['roofline']['frac'], 3), e['clocks'].get('sm_mhz'))
except Exception as ex: print('bench n2 parse failed', ex)
PY
CUDA_VISIBLE_DEVICES=0 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fastddc -c 40 --csv --log-file gpurun_out/r2_g5_ddc3_launches.csv python tools/run_ddc3_once.py > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r2_g5_ddc3_launches.csv')))
h = next(r for r in rows if 'Kernel Name' in r); kn, mv = h.index('Kernel Name'), h.index('Metric Value')
for r in rows[rows.index(h) + 1:]:
    if len(r) > mv: print(r[kn].split('(')[0][:60], float(r[mv]) / 1e3)
PY
CUDA_VISIBLE_DEVICES=0 ncu --set full --clock-control none --import-source on -k regex:fastddc_ifft_post -s 1 -c 1 -o gpurun_out/r2_g5_ifft_post python tools/run_ddc3_once.py > gpurun_out/r2_g5_ncu.log 2>&1
CUDA_VISIBLE_DEVICES=0 ncu --set full --clock-control none --import-source on -k regex:fastddc_fold -s 1 -c 1 -o gpurun_out/r2_g5_fold python tools/run_ddc3_once.py >> gpurun_out/r2_g5_ncu.log 2>&1
ls -la gpurun_out | tail -4
