#!/bin/bash
# Round-2 GPU call 21 (2 GPUs): the final bench line under torchrun (NCCL broadcast legs with the plan object and 592-block batches), both arms; the multi-GPU tests
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_g21_bench_n2.json 2> gpurun_out/r2_g21_bench_n2.err
echo "bench N=2 rc=$?"; tail -2 gpurun_out/r2_g21_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_g21_bench_reference_n2.json 2> gpurun_out/r2_g21_bench_reference_n2.err
echo "reference arm N=2 rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_g21_bench_n2.json') if l.startswith('{')][-1])
print('N', d['n_gpus'], 'value', round(d['value']), 'frac', round(d['roofline']['frac'], 3), 'sustained', round(d['sustained']['value']), 'e2e', round(d['e2e']['value']), 'e2e_u8', round(d['e2e_u8']['value']), d['clocks'])
for e in d['extra']: print('  ', e['name'], round(e['value']), 'Msps', round(e['kernel_ms'], 3), 'ms', e['roofline']['bound'], round(e['roofline']['frac'], 3), e['clocks'].get('sm_mhz'))
r = json.loads([l for l in open('gpurun_out/r2_g21_bench_reference_n2.json') if l.startswith('{')][-1])
print('reference arm at N=2:', round(r['value']), r.get('n_gpus'), r['cpu_baseline']['kind'])
PY
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_zzz_bankd.py -m gpu -x -q -k "multi or devices or bankd" 2>&1 | tail -3 | tee gpurun_out/r2_g21_multi_tests.log
du -sh gpurun_out
