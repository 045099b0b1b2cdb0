#!/bin/bash
# Round-2 GPU call 29 (8 GPUs): the final bench line under torchrun on eight GPUs (both arms)
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_g29_bench_n8.json 2> gpurun_out/r2_g29_bench_n8.err
echo "bench N=8 rc=$?"; tail -2 gpurun_out/r2_g29_bench_n8.err | cut -c1-300
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_g29_bench_n8.json') if l.startswith('{')][-1])
print('N', d['n_gpus'], 'value', round(d['value']), 'frac', round(d['roofline']['frac'], 3), 'sustained', round(d['sustained']['value']), 'e2e', round(d['e2e']['value']), 'h2d/rank', round(d['e2e']['h2d_gbs_per_rank'], 1), 'e2e_u8', round(d['e2e_u8']['value']), d['clocks'])
for e in d['extra']: print('  ', e['name'], round(e['value']), 'Msps', round(e['kernel_ms'], 3), 'ms', e['roofline']['bound'], round(e['roofline']['frac'], 3), e.get('overlap'))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2_g29_bench_reference_n8.json 2> /dev/null
echo "reference arm N=8 rc=$?"; tail -c 400 gpurun_out/r2_g29_bench_reference_n8.json
du -sh gpurun_out
