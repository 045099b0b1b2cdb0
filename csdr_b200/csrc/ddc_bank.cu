// ddc_bank.cu -- fused shared-input DDC / NFM bank (BASELINE config 4):
//     shift_addition_cc(rate_c)  ->  fir_decimate_cc(D, taps)  ->  [fmdemod_quadri_cf]        for C channels of ONE wideband stream
// i.e. what ddcd_old.h:51-57 runs as one `csdr shift_addition_cc --fd N | csdr fir_decimate_cc D bw` process chain per client.
//
// Mapping: LANE = CHANNEL.  All 32 lanes of a warp walk the same wideband samples in the same order, so
//   * the wideband sample is one broadcast load per warp (shared input: 8 B per sample, L1/L2 resident),
//   * the FIR tap for sample n and output o is the same for every lane -> taps live in uniform registers
//     (kernel parameter, duplicated (h,h) for FFMA2), exactly like the independent-input bank kernel,
//   * each lane keeps its own NCO phasor (the reference's float recursion, re-seeded at every chunk boundary from the
//     replayed float phase chain) and its own M = ceil(T/D) running output accumulators in registers.
// Per wideband sample and channel: 4 flop rotation + 6 flop recursion + 2*M FFMA2 lanes; nothing goes through shared memory,
// the shifted stream never exists in HBM (the unfused chain writes and re-reads 8*C bytes per wideband sample).
// A warp owns a time segment of SEG outputs (plus M-1 trailing periods to finish its last outputs); segments are independent
// because the phasor of any sample depends only on its chunk's seed and its position inside the chunk.
//
// Bound: FP32 issue: ~ (10 + 4*M) FMA-lane slots per (sample, channel).
#include "common.cuh"
#include "kernels.h"

namespace csdrb {

template <int TPAD>
struct DdcTaps { float2 hh[TPAD]; };

#define FMDEMOD_K_D 0.340447550238101026565118445432744920253753662109375

__device__ __forceinline__ float quadri_d(float2 cur, float2 prev)
{
    const float dq = __fsub_rn(cur.y, prev.y), di = __fsub_rn(cur.x, prev.x);
    const float num = __fsub_rn(__fmul_rn(cur.x, dq), __fmul_rn(cur.y, di));
    const float den = __fadd_rn(__fmul_rn(cur.x, cur.x), __fmul_rn(cur.y, cur.y));
    return den != 0.f ? (float)(FMDEMOD_K_D * (double)num / (double)den) : 0.f;
}

// seeds: (cos, sin) of every chunk's starting phase, evaluated in double like the reference does at each call
__global__ void ddc_seed_kernel(const float* __restrict__ chunk_phase, float2* __restrict__ seeds, long total)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const double ph = (double)chunk_phase[i];
    seeds[i] = make_float2((float)cos(ph), (float)sin(ph));
}

// phase chain over ABSOLUTE chunks: the block starts `offset` samples into chunk 0; phase_io holds the phase at the start of
// chunk 0 on entry and, on return, the phase at the start of the chunk that contains sample `advance` (the next block's start).
__global__ void ddc_phase_chain_kernel(const float3* __restrict__ params, float* __restrict__ phase_io, float* __restrict__ chunk_phase,
                                       int channels, int nchunks, int chunk, int next_chunk)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= channels) return;
    const float rate2 = params[c].z;
    float ph = phase_io[c], keep = ph;
    for (int k = 0; k < nchunks; k++) {
        chunk_phase[(long)c * nchunks + k] = ph;
        if (k == next_chunk) keep = ph;
        ph = wrap_phase_pm_pi(__fadd_rn(ph, __fmul_rn(__fmul_rn(rate2, 3.14159265358979323846f), (float)chunk)));
    }
    if (next_chunk >= nchunks) {                       // the next block starts beyond the chunks this block touched
        for (int k = nchunks; k < next_chunk; k++) ph = wrap_phase_pm_pi(__fadd_rn(ph, __fmul_rn(__fmul_rn(rate2, 3.14159265358979323846f), (float)chunk)));
        keep = ph;
    }
    phase_io[c] = keep;
}

template <int D, int M, bool DEMOD>
__global__ void __launch_bounds__(128)
ddc_bank_fused_kernel(const float2* __restrict__ wide, int n_in, int offset, int chunk, int nchunks,
                      const float3* __restrict__ params, const float2* __restrict__ seeds, int channels,
                      void* __restrict__ out_v, long out_stride, int n_out, int seg_outputs,
                      const float2* __restrict__ last_in, float2* __restrict__ last_out,
                      const __grid_constant__ DdcTaps<D * M> taps)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = (blockIdx.y * 4 + warp) * 32 + lane;
    const bool live = ch < channels;
    const int chs = live ? ch : channels - 1;                           // dead lanes shadow a real channel (no divergence), never store
    const int o_first = blockIdx.x * seg_outputs;                       // first output this warp emits
    if (o_first >= n_out) return;
    const int o_end = min(n_out, o_first + seg_outputs);
    // an output's window starts at its own first sample, so the walk can begin right at o_first (outputs before it are the
    // partial ones and are simply not emitted); the demodulator needs the previous baseband sample: start one output early.
    int o_start = o_first - (DEMOD ? 1 : 0);
    if (o_start < 0) o_start = 0;
    const float3 p = params[chs];
    const float sind = p.x, cosd = p.y;
    float2 acc[M];
#pragma unroll
    for (int j = 0; j < M; j++) acc[j] = make_float2(0.f, 0.f);
    float2 prev = (DEMOD && last_in) ? last_in[chs] : make_float2(0.f, 0.f);
    // sample index n runs from o_start*D; absolute chunk position = offset + n
    long n = (long)o_start * D;
    int kchunk = (int)((offset + n) / chunk);
    int into = (int)((offset + n) % chunk);                            // samples already consumed in this chunk
    float2 cs = seeds[(long)chs * nchunks + kchunk];
    float c = cs.x, s = cs.y;
    for (int t = 0; t < into; t++) {                                    // replay the recursion up to the segment start (< chunk steps, no data)
        const float cn = __fsub_rn(__fmul_rn(c, cosd), __fmul_rn(s, sind));
        const float sn = __fadd_rn(__fmul_rn(s, cosd), __fmul_rn(c, sind));
        c = cn; s = sn;
    }
    int left = chunk - into;
    // acc[j] collects output (o_cur + j) where o_cur is the output whose window STARTS at the current period:
    // at period q (samples qD .. qD+D-1) sample qD+p contributes to output q-j with tap p + jD, j = 0..M-1.
    for (int q = o_start; q < o_end + M - 1; q++) {
        const long base = (long)q * D;                                  // samples past n_in only ever meet zero-padded taps: read as zeros
#pragma unroll
        for (int pp = 0; pp < D; pp += 2) {
            // two wideband samples per 128-bit broadcast load (D is even, base is even -> 16-byte aligned)
            float4 xx;
            if (base + pp + 1 < n_in) xx = __ldg(reinterpret_cast<const float4*>(wide + base + pp));
            else { const float2 a = base + pp < n_in ? wide[base + pp] : make_float2(0.f, 0.f); xx = make_float4(a.x, a.y, 0.f, 0.f); }
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int pidx = pp + e;
                const float xi = e ? xx.z : xx.x, xq = e ? xx.w : xx.y;
                if (left == 0) {                                        // chunk boundary: re-seed the phasor (warp-uniform branch)
                    if (kchunk < nchunks - 1) kchunk++;                 // (beyond the block the data are zeros; any phasor will do)
                    cs = seeds[(long)chs * nchunks + kchunk];
                    c = cs.x; s = cs.y; left = chunk;
                }
                const float2 sh = make_float2(fmaf(c, xi, -s * xq), fmaf(s, xi, c * xq));
                const float cn = __fsub_rn(__fmul_rn(c, cosd), __fmul_rn(s, sind));
                const float sn = __fadd_rn(__fmul_rn(s, cosd), __fmul_rn(c, sind));
                c = cn; s = sn; left--;
#pragma unroll
                for (int j = 0; j < M; j++) acc[j] = ffma2(sh, taps.hh[pidx + j * D], acc[j]);
            }
        }
        // period q done: output q-(M-1) is complete (its last tap block was j = M-1)
        const int o = q - (M - 1);
        const float2 y = acc[M - 1];
#pragma unroll
        for (int j = M - 1; j > 0; j--) acc[j] = acc[j - 1];
        acc[0] = make_float2(0.f, 0.f);
        if (o >= o_start) {
            if (DEMOD) {
                if (o >= o_first && live) static_cast<float*>(out_v)[(long)ch * out_stride + o] = quadri_d(y, prev);
                prev = y;
            } else if (o >= o_first && live) {
                static_cast<float2*>(out_v)[(long)ch * out_stride + o] = y;
            }
            if (DEMOD && last_out && live && o == n_out - 1) last_out[ch] = y;
        }
    }
}

size_t ddc_bank_scratch_bytes(int channels, int input_size, int chunk, int offset)
{
    if (chunk <= 0) chunk = input_size > 0 ? input_size : 1;
    const long nchunks = ((long)offset + input_size + chunk - 1) / chunk + 1;
    return (size_t)channels * (size_t)nchunks * (sizeof(float) + sizeof(float2)) + 64;
}

template <int D, int M>
static int launch_fused(const float2* wide, int n_in, int offset, int chunk, int nchunks, const float3* params, const float2* seeds, int channels,
                        int demod, void* out, long out_stride, int n_out, const float2* last_in, float2* last_out, const float* h_taps, int T, cudaStream_t st)
{
    DdcTaps<D * M> tp;
    for (int k = 0; k < D * M; k++) { const float h = k < T ? h_taps[k] : 0.f; tp.hh[k] = make_float2(h, h); }
    const int groups = (channels + 127) / 128;
    // enough warps to fill the machine (~12 per SM) without making the warm-up (M+1 outputs) dominate
    long want_segments = (148L * 12 + groups * 4 - 1) / (groups * 4);
    int seg = (int)((n_out + want_segments - 1) / want_segments);
    if (seg < 4 * M) seg = 4 * M;
    dim3 grid((n_out + seg - 1) / seg, groups);
    if (demod) ddc_bank_fused_kernel<D, M, true><<<grid, 128, 0, st>>>(wide, n_in, offset, chunk, nchunks, params, seeds, channels, out, out_stride, n_out, seg, last_in, last_out, tp);
    else ddc_bank_fused_kernel<D, M, false><<<grid, 128, 0, st>>>(wide, n_in, offset, chunk, nchunks, params, seeds, channels, out, out_stride, n_out, seg, last_in, last_out, tp);
    CSDRB_CUDA(cudaGetLastError());
    return 0;
}

// returns outputs per channel; *launches receives the number of kernels launched
int launch_ddc_bank(const float2* d_wide, int input_size, int channels, const float* d_params, float* d_phase_io, int chunk, int offset,
                    int decimation, const float* h_taps, int taps_length, int demod, void* d_out, long out_stride,
                    const float2* d_last_in, float2* d_last_out, void* d_scratch, size_t scratch_bytes, int* launches, cudaStream_t st)
{
    *launches = 0;
    if (channels <= 0 || decimation <= 0 || taps_length <= 0) { set_error("ddc bank: bad geometry"); return -1; }
    const int n_out = input_size >= taps_length ? (input_size - taps_length) / decimation + 1 : 0;
    if (n_out == 0) return 0;
    if (chunk <= 0) chunk = input_size;
    if (offset < 0 || offset >= chunk) { set_error("ddc bank: offset must be in [0, chunk)"); return -1; }
    if (reinterpret_cast<uintptr_t>(d_wide) & 15) { set_error("ddc bank: wideband input must be 16-byte aligned"); return -1; }
    if (scratch_bytes < ddc_bank_scratch_bytes(channels, input_size, chunk, offset) || !d_scratch) { set_error("ddc bank: scratch too small"); return -1; }
    const int nchunks = (int)(((long)offset + input_size + chunk - 1) / chunk) + 1;
    float* chunk_phase = static_cast<float*>(d_scratch);
    float2* seeds = reinterpret_cast<float2*>(static_cast<char*>(d_scratch) + (((size_t)channels * nchunks * sizeof(float) + 15) & ~(size_t)15));
    const long advance = (long)n_out * decimation;                      // the next block starts here (the caller re-presents the tail)
    const int next_chunk = (int)((offset + advance) / chunk);
    ddc_phase_chain_kernel<<<(channels + 63) / 64, 64, 0, st>>>(reinterpret_cast<const float3*>(d_params), d_phase_io, chunk_phase, channels, nchunks, chunk, next_chunk);
    CSDRB_CUDA(cudaGetLastError());
    const long total = (long)channels * nchunks;
    ddc_seed_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(chunk_phase, seeds, total);
    CSDRB_CUDA(cudaGetLastError());
    int rc = -1;
    const float3* P = reinterpret_cast<const float3*>(d_params);
    if (decimation == 50 && taps_length <= 50 * 17) rc = launch_fused<50, 17>(d_wide, input_size, offset, chunk, nchunks, P, seeds, channels, demod, d_out, out_stride, n_out, d_last_in, d_last_out, h_taps, taps_length, st);
    else if (decimation == 10 && taps_length <= 10 * 20) rc = launch_fused<10, 20>(d_wide, input_size, offset, chunk, nchunks, P, seeds, channels, demod, d_out, out_stride, n_out, d_last_in, d_last_out, h_taps, taps_length, st);
    else if (decimation == 10 && taps_length <= 10 * 8) rc = launch_fused<10, 8>(d_wide, input_size, offset, chunk, nchunks, P, seeds, channels, demod, d_out, out_stride, n_out, d_last_in, d_last_out, h_taps, taps_length, st);
    else { set_error("ddc bank: no fused kernel for decimation %d / %d taps (compiled: d=50 T<=850, d=10 T<=200); run the unfused bank calls", decimation, taps_length); return -2; }
    if (rc < 0) return rc;
    *launches = 3;
    return n_out;
}

}  // namespace csdrb
