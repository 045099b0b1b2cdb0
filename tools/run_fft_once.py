import sys, numpy as np, torch
sys.path.insert(0, ".")
import csdr_b200 as cb
dev = "cuda"
z = torch.view_as_complex(torch.rand((4096, 4096, 2), device=dev))
for _ in range(2): cb.fft_c2c(z)
T, NF, isz, ov = cb.bandpass_geometry(0.002)
tf = cb.bandpass_taps_fft(-0.05, 0.05, 0.002)
x = torch.view_as_complex(torch.rand((512, 31 * isz, 2), device=dev))
tail = torch.zeros((512, NF), dtype=torch.complex64, device=dev)
for _ in range(2): cb.bandpass_fir_fft_bank_cc(x, tf, isz, tail=tail)
ddc = cb.fastddc_init(0.002, 64, 0.0)
xs = torch.view_as_complex(torch.rand((64 * ddc.input_size, 2), device=dev))
sp, ovl = cb.fastddc_fwd_cc(xs, ddc)
shifts = list(np.linspace(-0.45, 0.45, 64))
out, counts, st = cb.fastddc_inv_bank_cc(sp, shifts, 64, 0.002)
out, counts, st = cb.fastddc_inv_bank_cc(sp, shifts, 64, 0.002, state=st)
torch.cuda.synchronize()
