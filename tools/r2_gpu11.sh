#!/bin/bash
# Round-2 GPU call 11 (8 GPUs): bench.py at N=8 and N=4 under torchrun (NCCL broadcast legs, NUMA-local pinned buffers), the multi-GPU bank test on 8 devices,
# and the fastddc post-kernel prefetch A/B on GPU 0.
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_g11_topo.txt 2>&1
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2_g11_bench_n$n.json 2> gpurun_out/r2_g11_bench_n$n.err
  echo "N=$n rc=$?"; tail -2 gpurun_out/r2_g11_bench_n$n.err
  python - "$n" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f'gpurun_out/r2_g11_bench_n{sys.argv[1]}.json') if l.startswith('{')][-1])
    print('N', d['n_gpus'], 'value', round(d['value']), 'frac', round(d['roofline']['frac'], 3), 'e2e', round(d['e2e']['value']), 'h2d/rank', round(d['e2e']['h2d_gbs_per_rank'], 1), 'e2e_u8', round(d['e2e_u8']['value']), d['clocks'])
    for e in d['extra']: print('  ', e['name'], round(e['value']), 'Msps', round(e['kernel_ms'], 3), 'ms', e['roofline']['bound'], round(e['roofline']['frac'], 3), e['clocks'].get('sm_mhz'))
except Exception as ex: print('parse failed', ex)
PY
done
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "multi_gpu" 2>&1 | tail -3 | tee gpurun_out/r2_g11_tests.log
CUDA_VISIBLE_DEVICES=0 python tools/bench_configs.py c3 2>&1 | tee gpurun_out/r2_g11_c3.txt | grep cfg3
du -sh gpurun_out
