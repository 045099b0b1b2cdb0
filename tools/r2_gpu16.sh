#!/bin/bash
# Round-2 GPU call 16: plan timeline at 592 blocks; A/B of the phasor-walk placement and of the 512-thread fold
set -u
mkdir -p gpurun_out
cat > /tmp/plan_trace.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import csdr_b200 as cb
bw, dec, C, nblocks = 0.002, 64, 64, 592
ddc = cb.fastddc_init(bw, dec, 0.0)
x = torch.view_as_complex(torch.rand((nblocks * ddc.input_size, 2), device="cuda") * 2 - 1)
sp, ov = cb.fastddc_fwd_cc(x, ddc)
plan = cb.FastddcInvPlan(list(np.linspace(-0.45, 0.45, C)), dec, bw, nblocks)
for _ in range(5):
    sp, ov = cb.fastddc_fwd_cc(x, ddc, overlap=ov)
    plan.run(sp)
torch.cuda.synchronize()
PY
CSDRB_INV_TRACE=1 python /tmp/plan_trace.py 2>&1 | grep "plan trace" | tail -3 | tee gpurun_out/r2_g16_plan_trace.txt
for v in "default CSDRB_X=0" "walk_after_ifft CSDRB_PLAN_WALK=i" "fold512 CSDRB_FOLD_BT=4"; do
  set -- $v; tag=$1; shift
  env "$@" C3_BLOCKS=592 python tools/bench_configs.py c3 2>&1 | grep -E "plan" | sed "s/^/[$tag] /" | tee -a gpurun_out/r2_g16_c3.txt
done
du -sh gpurun_out
