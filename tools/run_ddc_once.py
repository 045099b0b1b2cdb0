"""ncu driver: the config-4 bank object for a few blocks (main kernel, phase chain and seed kernels in the launch list)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import csdr_b200 as cb
C, N, D = 128, 1 << 21, 50
T = cb.firdes_filter_len(0.005); taps = cb.firdes_lowpass_f(T, 0.5 / D)
x = torch.view_as_complex(torch.rand((N, 2), device="cuda") * 2 - 1)
rates = np.linspace(-0.45, 0.45, C).astype(np.float32)
n_out = cb.fir_out_len(N, D, T)
fo = torch.empty((C, n_out + (n_out & 1)), dtype=torch.float32, device="cuda")
bank = cb.DdcBank(rates, D, taps, demod=True, chunk=1024)
for _ in range(4):
    bank.process(x, out=fo)
torch.cuda.synchronize()
bank.close()
