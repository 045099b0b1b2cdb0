// fft_emul.cpp -- CPU-tier execution of the shipped FFT-family kernels (csdr_b200/csrc/fft_kernels.cuh) under tests/host_shim/cuda_emul.h:
// every CUDA thread a fiber, __syncthreads() a real barrier.  Built and driven by tests/test_kernels_emulated.py.
// TEST INFRASTRUCTURE ONLY -- nothing under csdr_b200/ can reach this; grids/blocks/shared sizes below restate the launchers of fft.cu.
#include <algorithm>
using std::max;
using std::min;
#include "cuda_emul.h"
#include "../../csdr_b200/csrc/fft_kernels.cuh"

#include <vector>

using namespace csdrb;

namespace {
const float2* twiddles(int n)
{
    static std::vector<float2> tw[32];
    int lg = 0; while ((1 << lg) < n) lg++;
    if (tw[lg].empty()) { tw[lg].resize((size_t)3 * n); fft_fill_twiddles(n, tw[lg].data()); }
    return tw[lg].data();
}
template <int N>
void c2c(const float2* in, long is, float2* out, long os, int batch, int inverse)
{
    const size_t smem = sizeof(float2) * fft_smem_elems(N);
    if (inverse) cuda_emul::launch(dim3(batch), dim3(fft_threads(N)), smem, fft_c2c_batch_kernel<N, true>, in, is, out, os, twiddles(N));
    else cuda_emul::launch(dim3(batch), dim3(fft_threads(N)), smem, fft_c2c_batch_kernel<N, false>, in, is, out, os, twiddles(N));
}
template <int N>
void olafir(const float2* in, long is, float2* out, long os, int channels, int input_size, int nblocks, const float2* H, long hs, float2* tail_io,
            int blocks_per_cta)
{
    const dim3 grid((nblocks + blocks_per_cta - 1) / blocks_per_cta, channels);
    const size_t smem = sizeof(float2) * ((size_t)fft_smem_elems(N) + (size_t)N);
    cuda_emul::launch(grid, dim3(fft_threads(N)), smem, olafir_bank_kernel<N>, in, is, out, os, H, hs, tail_io, input_size, nblocks, blocks_per_cta, twiddles(N));
}
template <int N>
void ddc_fwd(const float2* in, float2* spectra, float2* overlap_io, int input_size, int nblocks)
{
    cuda_emul::launch(dim3(nblocks), dim3(fft_threads(N)), sizeof(float2) * fft_smem_elems(N), fastddc_fwd_kernel<N>, in, spectra, (const float2*)overlap_io, input_size,
                      twiddles(N));
    if (N - input_size > 0) cuda_emul::launch(dim3(1), dim3(1024), 0, fastddc_carry_overlap_kernel, in, overlap_io, N - input_size, (long)nblocks * input_size);
}
template <int N>
void fir_fft(const float2* in, const float2* H, const float2* last_overlap, int overlap_size, float2* out)
{
    cuda_emul::launch(dim3(1), dim3(fft_threads(N)), sizeof(float2) * fft_smem_elems(N), apply_fir_fft_kernel<N>, in, H, last_overlap, overlap_size, out, twiddles(N));
}
template <int M>
void ddc_inv(const float2* spectra, int nblocks, const float2* taps_fft, const DdcChan* chan, int channels, int fft_size, int pre, int scrap, int post_input_size,
             int post_decimation, const int* blk_remain, const float* blk_phase, const int* blk_offset, float2* out, long out_stride, int tiled)
{
    if (tiled) {
        if constexpr (M >= 8 && M <= 1024) {
            constexpr int CT = 4, BT = 4;
            const dim3 grid((nblocks + BT - 1) / BT, (channels + CT - 1) / CT);
            cuda_emul::launch(grid, dim3(256), sizeof(float2) * (size_t)CT * BT * fft_smem_elems(M), fastddc_inv_tiled_kernel<M, CT, BT>, spectra, taps_fft, chan, blk_remain,
                              blk_phase, blk_offset, out, out_stride, fft_size, pre, scrap, post_input_size, post_decimation, nblocks, channels, twiddles(M));
        }
    } else {
        cuda_emul::launch(dim3(nblocks, channels), dim3(256), 0, fastddc_inv_kernel<M>, spectra, taps_fft, chan, blk_remain, blk_phase, blk_offset, out, out_stride, fft_size,
                          pre, scrap, post_input_size, post_decimation, nblocks, twiddles(M));
    }
}
}  // namespace

#define SIZES_ALL(X) X(2) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096) X(8192) X(16384)
#define SIZES_MID(X) X(4) X(8) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096)

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
                const float2* taps_fft, long taps_stride, float2* tail_io, int blocks_per_cta)
{
    switch (fft_size) {
#define X(N) case N: olafir<N>(in, in_stride, out, out_stride, channels, input_size, nblocks, taps_fft, taps_stride, tail_io, blocks_per_cta); return 0;
        SIZES_MID(X)
#undef X
    }
    return -1;
}
int emul_fastddc_fwd(const float2* in, float2* spectra, float2* overlap_io, int fft_size, int input_size, int nblocks)
{
    switch (fft_size) {
#define X(N) case N: ddc_fwd<N>(in, spectra, overlap_io, input_size, nblocks); return 0;
        SIZES_MID(X)
#undef X
    }
    return -1;
}
int emul_apply_fir_fft(const float2* in, const float2* taps_fft, const float2* last_overlap, int overlap_size, float2* out, int fft_size)
{
    switch (fft_size) {
#define X(N) case N: fir_fft<N>(in, taps_fft, last_overlap, overlap_size, out); return 0;
        SIZES_MID(X)
#undef X
    }
    return -1;
}
// chan: channels x {int offsetbin; float sindelta, cosdelta, rate}; remain_io/phase_io/out_total as in csdrb_fastddc_inv_bank_cc
int emul_fastddc_inv_bank(const float2* spectra, int nblocks, const float2* taps_fft, const void* chan, int channels, int fft_size, int fft_inv_size, int pre_decimation,
                          int scrap, int post_input_size, int post_decimation, int* remain_io, float* phase_io, float2* out, long out_stride, int* out_total, int force_simple)
{
    std::vector<int> blk_remain((size_t)channels * nblocks), blk_offset((size_t)channels * nblocks);
    std::vector<float> blk_phase((size_t)channels * nblocks);
    const DdcChan* c = static_cast<const DdcChan*>(chan);
    cuda_emul::launch(dim3((channels + 63) / 64), dim3(64), 0, fastddc_state_chain_kernel, c, remain_io, phase_io, blk_remain.data(), blk_phase.data(), blk_offset.data(),
                      out_total, channels, nblocks, post_input_size, post_decimation);
    const int tiled = !force_simple && fft_inv_size <= 1024 && fft_inv_size >= 8 && (fft_size / fft_inv_size) % 2 == 0;      // launch_fastddc_inv_bank's choice
    switch (fft_inv_size) {
#define X(M) case M: ddc_inv<M>(spectra, nblocks, taps_fft, c, channels, fft_size, pre_decimation, scrap, post_input_size, post_decimation, blk_remain.data(), \
                                 blk_phase.data(), blk_offset.data(), out, out_stride, tiled); return tiled;
        X(2) SIZES_MID(X)
#undef X
    }
    return -1;
}
long emul_barriers(void) { return cuda_emul::st().barriers; }
}
