"""ncu driver: the config-5 overlap-add bank (512 channels, 4096-pt) on a 64 k block, and the 16384-pt batched FFT."""
import sys, torch
sys.path.insert(0, ".")
import csdr_b200 as cb
T, NF, isz, ov = cb.bandpass_geometry(0.002)
tf = cb.bandpass_taps_fft(-0.05, 0.05, 0.002)
x = torch.view_as_complex(torch.rand((512, 31 * isz, 2), device="cuda"))
tail = torch.zeros((512, NF), dtype=torch.complex64, device="cuda")
for _ in range(3): cb.bandpass_fir_fft_bank_cc(x, tf, isz, tail=tail)
z = torch.view_as_complex(torch.rand((4096, 16384, 2), device="cuda"))
for _ in range(3): cb.fft_c2c(z)
torch.cuda.synchronize()
