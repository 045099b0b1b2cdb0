#!/bin/bash
# Round-2 GPU call 15: cfg3 after the rows-kernel index arithmetic and the look-ahead reorder; K2 with three chain slices; ncu summaries (fold, rows, K2, DDC v2) made on the box
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_fullsize.py tests/test_gpu_shift_variants.py -m gpu -x -q -k "fastddc or fold or ddc or shift" 2>&1 | tail -3 | tee gpurun_out/r2_g15_tests.log
python tools/bench_configs.py c3 2>&1 | grep "cfg3" | tee gpurun_out/r2_g15_c3_256.txt
C3_BLOCKS=592 python tools/bench_configs.py c3 2>&1 | grep "cfg3" | sed "s/^/[592 blocks] /" | tee gpurun_out/r2_g15_c3_592.txt
python tools/bench_configs.py k 2>&1 | tee gpurun_out/r2_g15_kernels.txt | grep -E "K2|K7"
CSDRB_SHIFT_SLICES=1 python tools/bench_configs.py k 2>&1 | grep -E "K2" | sed "s/^/[one stream] /" | tee -a gpurun_out/r2_g15_kernels.txt
for spec in "fold:fastddc_fold_kernel:tools/run_c3_once.py 2:1" "rows:fastddc_ifft_rows_kernel:tools/run_c3_once.py 2:1" "shift_bank:shift_bank_kernel:tools/run_shift_once.py:1" "ddc_v2:ddc_bank_fused2_kernel:tools/run_ddc_once.py:1"; do
  IFS=: read tag kern cmd skip <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -f -o gpurun_out/r2_g15_$tag python $cmd > gpurun_out/r2_g15_ncu_$tag.log 2>&1
  python tools/ncu_summary.py --out gpurun_out gpurun_out/r2_g15_$tag.ncu-rep 2>&1 | tail -1
done
rm -f gpurun_out/*.ncu-rep
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_g15_bench.json 2> gpurun_out/r2_g15_bench.err; tail -2 gpurun_out/r2_g15_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r2_g15_bench.json') if l.startswith('{')][-1])
print('value', round(d['value']), 'frac', round(d['roofline']['frac'], 3), 'e2e', round(d['e2e']['value']), 'e2e_u8', round(d['e2e_u8']['value']), d['clocks'])
for e in d['extra']: print('  ', e['name'], round(e['value']), 'Msps', round(e['kernel_ms'], 3), 'ms', e['roofline']['bound'], round(e['roofline']['frac'], 3), e['clocks'].get('sm_mhz'))
PY
du -sh gpurun_out
