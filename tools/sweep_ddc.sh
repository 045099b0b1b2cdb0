#!/bin/bash
# kernel-only time of the fused DDC kernel for CPL / resident-warp targets (ncu launch list, cold-cache serialised: compare relatives)
for cpl in 1 2; do for wps in 8 12 16 24 32 48; do
  CSDRB_DDC_CPL=$cpl CSDRB_DDC_WPS=$wps ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/sw.csv -k regex:ddc_bank_fused python tools/run_fused_once.py > /dev/null 2>&1
  echo "CPL=$cpl WPS=$wps: $(grep ddc_bank_fused /tmp/sw.csv | tail -1 | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"')"
done; done
