"""A few fastddc inverse-bank calls of BASELINE config 3 (for ncu captures and CSDRB_INV_TRACE=1 timelines)."""
import sys, numpy as np, torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import csdr_b200 as cb
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
bw, dec, C, nblocks = 0.002, 64, 64, 256
ddc = cb.fastddc_init(bw, dec, 0.0)
x = torch.view_as_complex(torch.rand((nblocks * ddc.input_size, 2), device="cuda") * 2 - 1)
sp, ov = cb.fastddc_fwd_cc(x, ddc)
shifts = list(np.linspace(-0.45, 0.45, C))
out, counts, st = cb.fastddc_inv_bank_cc(sp, shifts, dec, bw)
plan = cb.FastddcInvPlan(shifts, dec, bw, nblocks) if len(sys.argv) > 2 and sys.argv[2] == "plan" else None
for _ in range(reps):
    sp, ov = cb.fastddc_fwd_cc(x, ddc, overlap=ov)
    if plan is not None: plan.run(sp)
    else: cb.fastddc_inv_bank_cc(sp, shifts, dec, bw, state=st)
torch.cuda.synchronize()
