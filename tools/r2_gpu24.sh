#!/bin/bash
# Round-2 GPU call 24 (2 GPUs): bench line under torchrun after the overlap breakdown was added to the NFM leg
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_g24_bench_n2.json 2> gpurun_out/r2_g24_bench_n2.err
echo "bench N=2 rc=$?"; tail -3 gpurun_out/r2_g24_bench_n2.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_g24_bench_n2.json') if l.startswith('{')][-1])
print('N', d['n_gpus'], 'value', round(d['value']), 'frac', round(d['roofline']['frac'], 3), 'e2e', round(d['e2e']['value']), 'e2e_u8', round(d['e2e_u8']['value']))
for e in d['extra']: print('  ', e['name'], round(e['value']), 'Msps', round(e['kernel_ms'], 3), 'ms', e['roofline']['bound'], round(e['roofline']['frac'], 3), e.get('overlap'))
PY
timeout 900 python tools/bench_bankd.py 128 60 --devices 0,1 2>&1 | head -2 | tee gpurun_out/r2_g24_bankd_2gpu.txt
du -sh gpurun_out
