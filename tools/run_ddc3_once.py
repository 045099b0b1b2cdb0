import sys, numpy as np, torch
sys.path.insert(0, ".")
import csdr_b200 as cb
ddc = cb.fastddc_init(0.002, 64, 0.0)
xs = torch.view_as_complex(torch.rand((256 * ddc.input_size, 2), device="cuda"))
sp, ovl = cb.fastddc_fwd_cc(xs, ddc)
shifts = list(np.linspace(-0.45, 0.45, 64))
out, counts, st = cb.fastddc_inv_bank_cc(sp, shifts, 64, 0.002)
for _ in range(2): out, counts, st = cb.fastddc_inv_bank_cc(sp, shifts, 64, 0.002, state=st)
a = torch.rand((64, 2_400_000), device="cuda")
for _ in range(2): cb.fractional_decimator_bank_ff(a, 5.0, 12)
torch.cuda.synchronize()
