#!/bin/bash
# Round-2 GPU call: validate + A/B the fused DDC bank (v2 kernel: taps in shared memory, packed phasor, ramp-aware taps) and the table-driven phase chains.
#   gpurun --timeout 900 -- 'bash tools/r2_ddc_ab.sh'
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2_ddc_tests.log
run() { tag=$1; shift; env "$@" python tools/bench_configs.py c4 2>&1 | grep -i "FUSED\|shift (shared" | sed "s/^/[$tag] /" | tee -a gpurun_out/r2_ddc_ab.txt; }
: > gpurun_out/r2_ddc_ab.txt
run v1_cpl1_wps24 CSDRB_DDC_V=1
run v2_cpl2_wps16 CSDRB_DDC_V=2
run v2_cpl2_wps12 CSDRB_DDC_V=2 CSDRB_DDC_WPS=12
run v2_cpl2_wps8  CSDRB_DDC_V=2 CSDRB_DDC_WPS=8
run v2_cpl2_wps24 CSDRB_DDC_V=2 CSDRB_DDC_WPS=24
run v2_cpl1_wps16 CSDRB_DDC_V=2 CSDRB_DDC_CPL=1
run v2_cpl1_wps32 CSDRB_DDC_V=2 CSDRB_DDC_CPL=1 CSDRB_DDC_WPS=32
python tools/bench_configs.py k c3 2>&1 | tee gpurun_out/r2_ddc_kernels.txt | grep -i "K2\|cfg3"
# launch list of the bank-object path (main kernel vs phase chain vs seeds), then one full capture of the main kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_ddc_launches.csv python tools/run_ddc_once.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ddc_bank_fused2 -s 2 -c 1 -o gpurun_out/r2_ddc_v2 python tools/run_ddc_once.py > gpurun_out/r2_ddc_ncu.log 2>&1
ls -la gpurun_out | tail -5
