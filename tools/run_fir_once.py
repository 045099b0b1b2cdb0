"""ncu driver: the headline FIR bank, cf32 and u8 front ends (256 x 2.4 M, 199 taps, d = 10)."""
import sys, torch
sys.path.insert(0, ".")
import csdr_b200 as cb
taps = cb.firdes_lowpass_f(199, 0.05)
x = torch.rand((256, 2_400_000, 2), device="cuda") * 2 - 1
for _ in range(3): cb.fir_decimate_bank_cc(x, 10, taps)
del x
u8 = torch.randint(0, 256, (256, 2_400_000, 2), dtype=torch.uint8, device="cuda")
for _ in range(3): cb.fir_decimate_bank_u8_cc(u8, 10, taps)
torch.cuda.synchronize()
