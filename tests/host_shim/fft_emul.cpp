// fft_emul.cpp -- CPU-tier execution of the shipped FFT-family kernels (csdr_b200/csrc/fft_kernels.cuh) under tests/host_shim/cuda_emul.h.
// Built and driven by tests/test_kernels_emulated.py.  TEST INFRASTRUCTURE ONLY.
#include <algorithm>
using std::max;
using std::min;
#include "cuda_emul.h"
#include "../../csdr_b200/csrc/fft_kernels.cuh"

#include <vector>

using namespace csdrb;

namespace {
std::vector<float2>& twiddles(int n)
{
    static std::vector<float2> tw[32];
    int lg = 0; while ((1 << lg) < n) lg++;
    if (tw[lg].empty()) { tw[lg].resize((size_t)3 * n); fft_fill_twiddles(n, tw[lg].data()); }
    return tw[lg];
}
template <int N>
void c2c(const float2* in, long is, float2* out, long os, int batch, int inverse)
{
    const float2* tw = twiddles(N).data();
    const size_t smem = sizeof(float2) * fft_smem_elems(N);
    if (inverse) cuda_emul::launch(dim3(batch), dim3(fft_threads(N)), smem, fft_c2c_batch_kernel<N, true>, in, is, out, os, tw);
    else cuda_emul::launch(dim3(batch), dim3(fft_threads(N)), smem, fft_c2c_batch_kernel<N, false>, in, is, out, os, tw);
}
template <int N>
void olafir(const float2* in, long is, float2* out, long os, int channels, int input_size, int nblocks, const float2* H, long hs, float2* tail_io,
            int blocks_per_cta, int fused)
{
    const float2* tw = twiddles(N).data();
    const dim3 grid((nblocks + blocks_per_cta - 1) / blocks_per_cta, channels);
    if (fused) {
        const size_t smem = sizeof(float2) * ((size_t)fft_smem_elems(N) + 2 * (size_t)(N - input_size));
        cuda_emul::launch(grid, dim3(fft_threads(N)), smem, olafir_bank_fused_kernel<N>, in, is, out, os, H, hs, tail_io, input_size, nblocks, blocks_per_cta, tw);
    } else {
        const size_t smem = sizeof(float2) * ((size_t)fft_smem_elems(N) + (size_t)N);
        cuda_emul::launch(grid, dim3(fft_threads(N)), smem, olafir_bank_kernel<N>, in, is, out, os, H, hs, tail_io, input_size, nblocks, blocks_per_cta, tw);
    }
}
template <int N>
void ddc_fwd(const float2* in, float2* spectra, const float2* overlap_in, int input_size, int nblocks)
{
    const float2* tw = twiddles(N).data();
    cuda_emul::launch(dim3(nblocks), dim3(fft_threads(N)), sizeof(float2) * fft_smem_elems(N), fastddc_fwd_kernel<N>, in, spectra, overlap_in, input_size, tw);
}
}  // namespace

#define SIZES_ALL(X) X(2) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096) X(8192) X(16384)
#define SIZES_OLA(X) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096)

extern "C" {
int emul_fft_c2c(const float2* in, long in_stride, float2* out, long out_stride, int n, int batch, int inverse)
{
    switch (n) {
#define X(N) case N: c2c<N>(in, in_stride, out, out_stride, batch, inverse); return 0;
        SIZES_ALL(X)
#undef X
    }
    return -1;
}
int emul_olafir(const float2* in, long in_stride, float2* out, long out_stride, int channels, int fft_size, int input_size, int nblocks,
                const float2* taps_fft, long taps_stride, float2* tail_io, int blocks_per_cta, int fused)
{
    switch (fft_size) {
#define X(N) case N: olafir<N>(in, in_stride, out, out_stride, channels, input_size, nblocks, taps_fft, taps_stride, tail_io, blocks_per_cta, fused); return 0;
        SIZES_OLA(X)
#undef X
    }
    return -1;
}
int emul_fastddc_fwd(const float2* in, float2* spectra, const float2* overlap_in, int fft_size, int input_size, int nblocks)
{
    switch (fft_size) {
#define X(N) case N: ddc_fwd<N>(in, spectra, overlap_in, input_size, nblocks); return 0;
        SIZES_OLA(X)
#undef X
    }
    return -1;
}
long emul_barriers(void) { return cuda_emul::st().barriers; }
}
