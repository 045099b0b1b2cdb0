#!/bin/bash
for cpl in 1 2; do for wps in 10 16 24 32 48; do
  echo -n "CPL=$cpl WPS=$wps: "; CSDRB_DDC_CPL=$cpl CSDRB_DDC_WPS=$wps python tools/bench_configs.py c4 2>&1 | grep FUSED | awk '{print $9, $10}'
done; done
