// kernels.h -- internal launcher prototypes (host side of each .cu).  Not part of the public C ABI.
#pragma once
#include <cuda_runtime.h>

namespace csdrb {

// K3 fir_decimate.cu
int launch_fir_decimate_bank(const float2* d_in, long in_stride, float2* d_out, long out_stride, int channels, int n_in,
                             int D, const float* h_taps, const float* d_taps, long taps_stride, int T, int variant,
                             cudaStream_t st);
int fir_bank_variant_count();

// K1/K4 elementwise.cu
int launch_convert_u8_f(const unsigned char* d_in, float* d_out, long n, cudaStream_t st);
int launch_convert_s16_f(const short* d_in, float* d_out, long n, cudaStream_t st);
int launch_convert_f_s16(const float* d_in, short* d_out, long n, cudaStream_t st);
int launch_fmdemod_quadri_bank(const float2* d_in, long in_stride, float* d_out, long out_stride, int channels, int n,
                               const float2* d_last_in, float2* d_last_out, cudaStream_t st);

}  // namespace csdrb
