import sys, torch
sys.path.insert(0, ".")
import csdr_b200 as cb
z = torch.view_as_complex(torch.rand((16384, 4096, 2), device="cuda"))
for _ in range(3): cb.fft_c2c(z)
torch.cuda.synchronize()
