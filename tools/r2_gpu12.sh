#!/bin/bash
# Round-2 GPU call 12: timeline of the fastddc inverse bank (CSDRB_INV_TRACE=1), ncu of the u8 FIR kernel (4th matching launch onwards = u8 template)
set -u
mkdir -p gpurun_out
CSDRB_INV_TRACE=1 python tools/run_c3_once.py 6 2>&1 | grep "inv trace" | tee gpurun_out/r2_g12_inv_trace.txt
python tools/bench_configs.py c3 2>&1 | grep cfg3 | tee gpurun_out/r2_g12_c3.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_bank_fast_kernel -s 4 -c 1 -f -o gpurun_out/r2_g12_fir_u8 python tools/run_fir_once.py > gpurun_out/r2_g12_ncu.log 2>&1
python tools/ncu_summary.py --out gpurun_out gpurun_out/r2_g12_fir_u8.ncu-rep 2>&1 | tail -2
ncu -i gpurun_out/r2_g12_fir_u8.ncu-rep --page source --csv 2>/dev/null | python - <<'PY' > gpurun_out/r2_g12_fir_u8_hot_sass.txt
import csv, sys
rows = list(csv.reader(sys.stdin))
if rows:
    h = rows[0]
    ci = {n: i for i, n in enumerate(h)}
    s = next((i for n, i in ci.items() if n.startswith("Source")), 1)
    w = next((i for n, i in ci.items() if "Warp Stall Sampling (All" in n), None)
    if w is not None:
        body = [r for r in rows[1:] if len(r) > w]
        body.sort(key=lambda r: -float(r[w] or 0))
        for r in body[:40]: print(r[w], r[s][:110])
PY
head -30 gpurun_out/r2_g12_fir_u8_hot_sass.txt
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
