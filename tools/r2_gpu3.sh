#!/bin/bash
# Round-2 GPU call 3: warp-resident phase chains, fold-based fastddc inverse, DDC v2 variants (kernel-only times from ncu launch lists).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2_g3_tests.log
python tools/bench_configs.py k c3 2>&1 | tee gpurun_out/r2_g3_kernels.txt | grep -i "K2\|cfg3"
CSDRB_INV_FOLD=0 python tools/bench_configs.py c3 2>&1 | sed 's/^/[fold off] /' | tee -a gpurun_out/r2_g3_kernels.txt | grep -i "cfg3"
python tools/bench_configs.py c4 2>&1 | grep -i "FUSED" | tee gpurun_out/r2_g3_c4.txt
one() { tag=$1; shift; env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/_l.csv python tools/run_ddc_once.py > /dev/null 2>&1
  python - "$tag" <<'PY' | tee -a gpurun_out/r2_g3_ddc_variants.txt
import csv, sys, collections
rows = list(csv.reader(open('gpurun_out/_l.csv')))
h = next(r for r in rows if 'Kernel Name' in r); kn, mv = h.index('Kernel Name'), h.index('Metric Value')
d = collections.defaultdict(list)
for r in rows[rows.index(h) + 1:]:
    if len(r) > mv and 'csdrb' in r[kn]: d[r[kn].split('(')[0][:48]].append(float(r[mv]) / 1e3)
print(sys.argv[1], {k: round(sorted(v)[len(v) // 2], 1) for k, v in d.items()})
PY
}
: > gpurun_out/r2_g3_ddc_variants.txt
one v2_cpl2_wps12_pf1 CSDRB_DDC_V=2
one v2_cpl2_wps12_pf0 CSDRB_DDC_PF=0
one v2_cpl2_wps8_pf1 CSDRB_DDC_WPS=8
one v2_cpl2_wps16_pf1 CSDRB_DDC_WPS=16
one v2_cpl2_wps24_pf1 CSDRB_DDC_WPS=24
one v2_cpl1_wps12_pf1 CSDRB_DDC_CPL=1
one v2_cpl1_wps24_pf1 CSDRB_DDC_CPL=1 CSDRB_DDC_WPS=24
one v2_cpl1_wps24_pf0 CSDRB_DDC_CPL=1 CSDRB_DDC_WPS=24 CSDRB_DDC_PF=0
one v1 CSDRB_DDC_V=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_g3_ddc3_launches.csv python tools/run_ddc3_once.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fastddc_fold -s 1 -c 1 -o gpurun_out/r2_g3_fold python tools/run_ddc3_once.py > gpurun_out/r2_g3_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ddc_bank_fused2 -s 2 -c 1 -o gpurun_out/r2_g3_ddc_v2 python tools/run_ddc_once.py >> gpurun_out/r2_g3_ncu.log 2>&1
ls -la gpurun_out | tail -4
