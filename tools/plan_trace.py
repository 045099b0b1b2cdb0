"""CSDRB_INV_TRACE=1 python tools/plan_trace.py [nblocks] [runs]: timeline of csdrb_fastddc_inv_plan_run at BASELINE config 3 (64 channels), forward FFT in the loop."""
import sys, numpy as np, torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import csdr_b200 as cb
nblocks = int(sys.argv[1]) if len(sys.argv) > 1 else 592
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 5
bw, dec, C = 0.002, 64, 64
ddc = cb.fastddc_init(bw, dec, 0.0)
x = torch.view_as_complex(torch.rand((nblocks * ddc.input_size, 2), device="cuda") * 2 - 1)
sp, ov = cb.fastddc_fwd_cc(x, ddc)
plan = cb.FastddcInvPlan(list(np.linspace(-0.45, 0.45, C)), dec, bw, nblocks)
for _ in range(runs):
    sp, ov = cb.fastddc_fwd_cc(x, ddc, overlap=ov)
    plan.run(sp)
torch.cuda.synchronize()
