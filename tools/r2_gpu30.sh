#!/bin/bash
# Round-2 GPU call 30 (2 GPUs): the serial-broadcast path of the NFM leg (the default from 8 ranks on) exercised at N=2; sanity of the library after the last two edits
set -u
mkdir -p gpurun_out
CSDRB_BENCH_BCAST=serial timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_g30_bench_n2_serial.json 2> gpurun_out/r2_g30_bench_n2_serial.err
echo "bench N=2 (serial broadcast) rc=$?"; tail -2 gpurun_out/r2_g30_bench_n2_serial.err | cut -c1-300
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_g30_bench_n2_serial.json') if l.startswith('{')][-1])
print('N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
for e in d['extra']: print('  ', e['name'], round(e['value']), 'Msps', round(e['kernel_ms'], 3), 'ms', e['config'].get('collective'), e.get('overlap'))
PY
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity2.py -m gpu -x -q -k "fastddc or fold or plan" 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
du -sh gpurun_out
