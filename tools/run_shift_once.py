"""ncu driver: K2 shift bank (64 x 2.4 M, chunk 1024) and the audio-rate kernels (K5 fractional decimator, K6 fastagc)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import csdr_b200 as cb
x = torch.view_as_complex(torch.rand((64, 2_400_000, 2), device="cuda") * 2 - 1)
out = torch.empty_like(x)
rates = np.linspace(-0.4, 0.4, 64).astype(np.float32)
for _ in range(3): cb.shift_addition_bank_cc(x, rates, chunk=1024, out=out)
a = torch.rand((64, 2_400_000), device="cuda")
for _ in range(2): cb.fractional_decimator_bank_ff(a, 5.0, 12)
for _ in range(2): cb.fastagc_bank_ff(a[:, :2343 * 1024].contiguous(), 1024, 1.0)
torch.cuda.synchronize()
