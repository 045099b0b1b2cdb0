// audio.cu -- K5 fractional_decimator_ff and K6 fastagc_ff (the per-channel audio-rate tail of the FM chain).
//
// K5 replaces fractional_decimator_ff (libcsdr.c:751-793; state struct libcsdr.h:151-168).
//    Output positions come from a float accumulator (`where += rate`) whose ceilf() picks sample indices,
//    so one ulp of difference flips an index (hard part 3).  We therefore split the work:
//      fracdec_positions_kernel : one thread per channel replays the accumulator chain exactly and records
//                                 (index_high, xwhere) per output -- sequential by definition, tiny;
//      fracdec_interp_kernel    : one thread per output evaluates the Lagrange polynomial with the same
//                                 operation order as the reference (IEEE mul/div/add, no contraction).
// K6 replaces fastagc_ff (libcsdr.c:944-991; state struct libcsdr.h:118-128): one CTA per channel walks the
//    blocks of its stream (the gain of block b depends on block b-1), block-wide |x| max reduction, linear
//    gain ramp evaluated in double exactly as the C expression promotes it, two blocks of latency.
#include "common.cuh"
#include "kernels.h"

namespace csdrb {

// ---------------------------------------------------------------------------------------------- K5
struct FracDecState { float where; int input_processed; int output_size; };

__global__ void fracdec_positions_kernel(FracDecState* __restrict__ state, int* __restrict__ idx_high, float* __restrict__ xwhere,
                                         int channels, int n, float rate, int num_poly_points, int xifirst, int taps_length, int cap)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= channels) return;
    float where = state[c].where;
    int produced = 0, index_high;
    for (; (index_high = (int)ceilf(where)) + num_poly_points + taps_length < n; where = __fadd_rn(where, rate)) {
        if (produced < cap) {
            idx_high[(long)c * cap + produced] = index_high;
            xwhere[(long)c * cap + produced] = __fsub_rn(where, (float)(index_high - 1));
        }
        produced++;
    }
    const int processed = (index_high - 1) + xifirst;
    state[c].input_processed = processed;
    state[c].where = __fsub_rn(where, (float)processed);
    state[c].output_size = produced < cap ? produced : cap;
}

constexpr int FD_MAX_POINTS = 64;

__global__ void __launch_bounds__(128)
fracdec_interp_kernel(const float* __restrict__ in, long in_stride, float* __restrict__ out, long out_stride,
                      const FracDecState* __restrict__ state, const int* __restrict__ idx_high, const float* __restrict__ xwhere,
                      int cap, int num_poly_points, int xifirst, int xilast, const float* __restrict__ taps, int taps_length)
{
    const int c = blockIdx.y;
    const int produced = state[c].output_size;
    const float* x = in + (long)c * in_stride;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < produced; o += gridDim.x * blockDim.x) {
        const int low = idx_high[(long)c * cap + o] - 1;
        const float xw = xwhere[(long)c * cap + o];
        float acc = 0.f;
        int slot = 0;
        for (int xi = xifirst; xi <= xilast; xi++, slot++) {
            float coef = 1.f, den = 1.f;
            for (int xj = xifirst; xj <= xilast; xj++)
                if (xi != xj) { coef = __fmul_rn(coef, __fsub_rn(xw, (float)xj)); den = __fmul_rn(den, (float)(xi - xj)); }
            float pt;
            if (taps) {
                pt = 0.f;
                const float* seg = x + low + slot;
                for (int t = 0; t < taps_length; t++) pt = __fadd_rn(pt, __fmul_rn(seg[t], taps[t]));
            } else pt = x[low + slot];
            acc = __fadd_rn(acc, __fmul_rn(__fdiv_rn(coef, den), pt));
        }
        out[(long)c * out_stride + o] = acc;
    }
}

size_t fracdec_scratch_bytes(int channels, int n, float rate)
{
    const int cap = (int)((double)n / (rate > 1.f ? rate : 1.0)) + 8;
    return (size_t)channels * cap * (sizeof(int) + sizeof(float));
}

int launch_fractional_decimator_bank(const float* d_in, long in_stride, float* d_out, long out_stride, int channels, int n,
                                     float rate, int num_poly_points, const float* d_taps, int taps_length, void* d_state,
                                     void* d_scratch, size_t scratch_bytes, cudaStream_t st)
{
    if (channels <= 0 || n <= 0) return 0;
    if (!(rate > 1.0f)) { set_error("fractional_decimator: rate must be > 1.0 (reference asserts it, libcsdr.c:756)"); return -1; }
    num_poly_points &= ~1;
    if (num_poly_points < 2 || num_poly_points > FD_MAX_POINTS) { set_error("fractional_decimator: num_poly_points must be even, 2..64"); return -1; }
    const int cap = (int)((double)n / rate) + 8;
    if (!d_scratch || scratch_bytes < fracdec_scratch_bytes(channels, n, rate)) { set_error("fractional_decimator: scratch too small"); return -1; }
    const int xifirst = -(num_poly_points / 2) + 1, xilast = num_poly_points / 2;
    int* idx = static_cast<int*>(d_scratch);
    float* xw = reinterpret_cast<float*>(idx + (size_t)channels * cap);
    if (!d_taps) taps_length = 0;
    fracdec_positions_kernel<<<(channels + 63) / 64, 64, 0, st>>>(static_cast<FracDecState*>(d_state), idx, xw, channels, n, rate, num_poly_points,
                                                                  xifirst, taps_length, cap);
    CSDRB_CUDA(cudaGetLastError());
    int gx = (cap + 127) / 128; if (gx > 1024) gx = 1024;
    fracdec_interp_kernel<<<dim3(gx, channels), 128, 0, st>>>(d_in, in_stride, d_out, out_stride, static_cast<const FracDecState*>(d_state), idx, xw, cap,
                                                              num_poly_points, xifirst, xilast, d_taps, taps_length);
    CSDRB_CUDA(cudaGetLastError());
    return 2;
}

// ---------------------------------------------------------------------------------------------- K6
struct FastAgcState { float peak_1, peak_2, last_gain; };

__global__ void __launch_bounds__(256)
fastagc_bank_kernel(const float* __restrict__ in, long in_stride, float* __restrict__ out, long out_stride, int block, int nblocks,
                    float reference, FastAgcState* __restrict__ state, float* __restrict__ hist /*[C][2][block]*/)
{
    __shared__ float red[8];
    __shared__ float s_peak;
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* x = in + (long)c * in_stride;
    float* y = out + (long)c * out_stride;
    float* h1 = hist + (long)c * 2 * block;      // block that leaves next (reference buffer_1)
    float* h2 = h1 + block;                      // block after that     (reference buffer_2)
    float peak_1 = state[c].peak_1, peak_2 = state[c].peak_2, last_gain = state[c].last_gain;
    for (int b = 0; b < nblocks; b++) {
        const float* cur = x + (long)b * block;
        float m = 0.f;
        for (int i = tid; i < block; i += blockDim.x) m = fmaxf(m, fabsf(cur[i]));
        for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) red[warp] = m;
        __syncthreads();
        if (tid == 0) { float t = red[0]; for (int w = 1; w < (int)(blockDim.x >> 5); w++) t = fmaxf(t, red[w]); s_peak = t; }
        __syncthreads();
        const float peak_in = s_peak;
        float target_peak = peak_in;
        if (target_peak < peak_2) target_peak = peak_2;
        if (target_peak < peak_1) target_peak = peak_1;
        float target = __fdiv_rn(reference, target_peak);
        if (target > 50.f) target = 50.f;                               // FASTAGC_MAX_GAIN, libcsdr.c:944
        // the block leaving now entered two calls ago: history for b < 2, otherwise the input itself
        const float* leaving = b >= 2 ? x + (long)(b - 2) * block : (b == 0 ? h1 : h2);
        for (int i = tid; i < block; i += blockDim.x) {
            const float r = __fdiv_rn((float)i, (float)block);
            const float gain = (float)((double)last_gain * (1.0 - (double)r) + (double)__fmul_rn(target, r));
            y[(long)b * block + i] = __fmul_rn(leaving[i], gain);
        }
        peak_1 = peak_2; peak_2 = peak_in; last_gain = target;
        __syncthreads();
    }
    // new history = the last two input blocks (or shifted old history when fewer than two arrived)
    if (nblocks >= 2) {
        for (int i = tid; i < block; i += blockDim.x) { h1[i] = x[(long)(nblocks - 2) * block + i]; h2[i] = x[(long)(nblocks - 1) * block + i]; }
    } else if (nblocks == 1) {
        for (int i = tid; i < block; i += blockDim.x) { h1[i] = h2[i]; }
        __syncthreads();
        for (int i = tid; i < block; i += blockDim.x) { h2[i] = x[i]; }
    }
    if (tid == 0) { state[c].peak_1 = peak_1; state[c].peak_2 = peak_2; state[c].last_gain = last_gain; }
}

int launch_fastagc_bank(const float* d_in, long in_stride, float* d_out, long out_stride, int channels, int block, int nblocks,
                        float reference, void* d_state, float* d_hist, cudaStream_t st)
{
    if (channels <= 0 || nblocks <= 0) return 0;
    if (block <= 0) { set_error("fastagc: block size must be positive"); return -1; }
    if (d_out == d_in) { set_error("fastagc: in-place operation is not supported (output lags input by two blocks)"); return -1; }
    fastagc_bank_kernel<<<channels, 256, 0, st>>>(d_in, in_stride, d_out, out_stride, block, nblocks, reference, static_cast<FastAgcState*>(d_state), d_hist);
    CSDRB_CUDA(cudaGetLastError());
    return 1;
}

}  // namespace csdrb
