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
// The float phase chain is a separate one-thread-per-channel pre-pass.  Folding it into this kernel as a ticketed producer CTA was
// tried and measured SLOWER (2.27 ms vs 2.06 ms at 2 M samples x 128 ch): the whole grid is resident at once, so every segment needs
// its chunk phase at t = 0 and nothing overlaps, while the producer runs ~3x slower on a shared SM.  The remaining lever is to run the
// pre-pass of block k+1 on a side stream during the main kernel of block k (it only depends on the previous pre-pass).
//
// Bound: FP32 issue: ~ (10 + 4*M) FMA-lane slots per (sample, channel).
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>

namespace csdrb {

template <int TPAD>
struct alignas(16) DdcTaps { float4 hh2[TPAD / 2]; };               // rows of MP = M rounded up to even; one float4 = taps (j, j+1), each duplicated (h,h)

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

// CPL = channels per lane.  A tap pair is loaded once per warp and sample phase (LDCU.128 fetches two of them: taps are stored
// [p][j] so the M taps that one sample meets are contiguous); with CPL = 2 every loaded tap feeds two FFMA2 and every wideband
// sample load feeds two channels, which halves the non-FMA issue slots per channel-sample (ncu r01: 868 LDCU per 850 FFMA2 at CPL = 1).
template <int D, int M, int CPL, bool DEMOD>
__global__ void __launch_bounds__(128)
ddc_bank_fused_kernel(const float2* __restrict__ wide, int n_in, int offset, int chunk, int nchunks,
                      const float3* __restrict__ params, const float2* __restrict__ seeds, int channels,
                      void* __restrict__ out_v, long out_stride, int n_out, int seg_outputs,
                      const float2* __restrict__ last_in, float2* __restrict__ last_out,
                      const __grid_constant__ DdcTaps<D * ((M + 1) & ~1)> taps)
{
    constexpr int MP = (M + 1) & ~1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch0 = (blockIdx.y * 4 + warp) * (32 * CPL) + lane;        // this lane's channels: ch0 + 32*u
    if (ch0 - lane >= channels) return;                                 // whole warp beyond the bank
    const int o_first = blockIdx.x * seg_outputs;                       // first output this warp emits
    if (o_first >= n_out) return;
    const int o_end = min(n_out, o_first + seg_outputs);
    // an output's window starts at its own first sample, so the walk can begin right at o_first (outputs before it are the
    // partial ones and are simply not emitted); the demodulator needs the previous baseband sample: start one output early.
    int o_start = o_first - (DEMOD ? 1 : 0);
    if (o_start < 0) o_start = 0;
    int chs[CPL]; bool live[CPL];
    float sind[CPL], cosd[CPL], c[CPL], s[CPL];
    float2 acc[CPL][M], prev[CPL];
    long n = (long)o_start * D;                                         // absolute chunk position = offset + n
    int kchunk = (int)((offset + n) / chunk);
    const int into = (int)((offset + n) % chunk);                       // samples already consumed in this chunk
#pragma unroll
    for (int u = 0; u < CPL; u++) {
        live[u] = ch0 + 32 * u < channels;
        chs[u] = live[u] ? ch0 + 32 * u : channels - 1;                 // dead lanes shadow a real channel (no divergence), never store
        const float3 p = params[chs[u]];
        sind[u] = p.x; cosd[u] = p.y;
#pragma unroll
        for (int j = 0; j < M; j++) acc[u][j] = make_float2(0.f, 0.f);
        prev[u] = (DEMOD && last_in) ? last_in[chs[u]] : make_float2(0.f, 0.f);
        const float2 cs = seeds[(long)chs[u] * nchunks + kchunk];
        c[u] = cs.x; s[u] = cs.y;
    }
    for (int t = 0; t < into; t++) {                                    // replay the recursion up to the segment start (< chunk steps, no data)
#pragma unroll
        for (int u = 0; u < CPL; u++) {
            const float cn = __fsub_rn(__fmul_rn(c[u], cosd[u]), __fmul_rn(s[u], sind[u]));
            const float sn = __fadd_rn(__fmul_rn(s[u], cosd[u]), __fmul_rn(c[u], sind[u]));
            c[u] = cn; s[u] = sn;
        }
    }
    int left = chunk - into;
    // acc[.][j] collects output q-j while the walk is in period q (samples qD .. qD+D-1): sample qD+p meets tap p + jD.
    for (int q = o_start; q < o_end + M - 1; q++) {
        const long base = (long)q * D;
        const bool inside = base + D <= n_in;                           // whole period inside the block: no per-load checks
        const float4* src = reinterpret_cast<const float4*>(wide + base);   // D even, base even: 16-byte aligned
#pragma unroll
        for (int pp = 0; pp < D; pp += 2) {
            float4 xx;                                                  // two wideband samples per 128-bit broadcast load
            if (inside) xx = __ldg(src + pp / 2);
            else {                                                      // samples past n_in only ever meet zero-padded taps: read as zeros
                const float2 a = base + pp < n_in ? wide[base + pp] : make_float2(0.f, 0.f);
                const float2 b = base + pp + 1 < n_in ? wide[base + pp + 1] : make_float2(0.f, 0.f);
                xx = make_float4(a.x, a.y, b.x, b.y);
            }
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int pidx = pp + e;
                const float xi = e ? xx.z : xx.x, xq = e ? xx.w : xx.y;
                if (left == 0) {                                        // chunk boundary: re-seed the phasors (warp-uniform branch)
                    if (kchunk < nchunks - 1) kchunk++;                 // (beyond the block the data are zeros; any phasor will do)
#pragma unroll
                    for (int u = 0; u < CPL; u++) { const float2 cs = seeds[(long)chs[u] * nchunks + kchunk]; c[u] = cs.x; s[u] = cs.y; }
                    left = chunk;
                }
                left--;
                float2 sh[CPL];
#pragma unroll
                for (int u = 0; u < CPL; u++) {
                    sh[u] = make_float2(fmaf(c[u], xi, -s[u] * xq), fmaf(s[u], xi, c[u] * xq));
                    const float cn = __fsub_rn(__fmul_rn(c[u], cosd[u]), __fmul_rn(s[u], sind[u]));
                    const float sn = __fadd_rn(__fmul_rn(s[u], cosd[u]), __fmul_rn(c[u], sind[u]));
                    c[u] = cn; s[u] = sn;
                }
#pragma unroll
                for (int j = 0; j < M; j += 2) {
                    const float4 h2 = taps.hh2[(pidx * MP + j) / 2];    // [p][j] layout, one 128-bit uniform load = two taps
#pragma unroll
                    for (int u = 0; u < CPL; u++) {
                        acc[u][j] = ffma2(sh[u], make_float2(h2.x, h2.y), acc[u][j]);
                        if (j + 1 < M) acc[u][j + 1] = ffma2(sh[u], make_float2(h2.z, h2.w), acc[u][j + 1]);
                    }
                }
            }
        }
        // period q done: output q-(M-1) is complete (its last tap block was j = M-1)
        const int o = q - (M - 1);
#pragma unroll
        for (int u = 0; u < CPL; u++) {
            const float2 y = acc[u][M - 1];
#pragma unroll
            for (int j = M - 1; j > 0; j--) acc[u][j] = acc[u][j - 1];
            acc[u][0] = make_float2(0.f, 0.f);
            if (o >= o_start) {
                if (DEMOD) {
                    if (o >= o_first && live[u]) static_cast<float*>(out_v)[(long)chs[u] * out_stride + o] = quadri_d(y, prev[u]);
                    prev[u] = y;
                    if (last_out && live[u] && o == n_out - 1) last_out[chs[u]] = y;
                } else if (o >= o_first && live[u]) {
                    static_cast<float2*>(out_v)[(long)chs[u] * out_stride + o] = y;
                }
            }
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
    constexpr int MP = (M + 1) & ~1;
    DdcTaps<D * MP> tp;                                                 // tap k = jD + p stored at [p][j], duplicated for FFMA2
    for (int p = 0; p < D; p++)
        for (int j = 0; j < MP; j++) {
            const int k = j * D + p; const float h = (j < M && k < T) ? h_taps[k] : 0.f;
            float4& slot = tp.hh2[(p * MP + j) / 2];
            if (j & 1) { slot.z = h; slot.w = h; } else { slot.x = h; slot.y = h; }
        }
    // tuning knobs (defaults from the r01 sweep in profiles/): channels per lane and resident-warp target per SM
    static const int cpl_env = getenv("CSDRB_DDC_CPL") ? atoi(getenv("CSDRB_DDC_CPL")) : 1;
    static const int wps_env = getenv("CSDRB_DDC_WPS") ? atoi(getenv("CSDRB_DDC_WPS")) : 24;
    const int cpl = cpl_env == 2 ? 2 : 1;
    const int warps_per_seg = (channels + 32 * cpl - 1) / (32 * cpl);
    const int groups = (warps_per_seg + 3) / 4;
    // enough warps to fill the machine while keeping the M-1 trailing periods of every segment a small fraction
    long want_segments = (148L * wps_env + warps_per_seg - 1) / warps_per_seg;
    int seg = (int)((n_out + want_segments - 1) / want_segments);
    if (seg < 2 * M) seg = 2 * M;
    dim3 grid((n_out + seg - 1) / seg, groups);
#define CSDRB_DDC_LAUNCH(CPLV, DM) ddc_bank_fused_kernel<D, M, CPLV, DM><<<grid, 128, 0, st>>>(wide, n_in, offset, chunk, nchunks, params, seeds, channels, out, out_stride, n_out, seg, last_in, last_out, tp)
    if (cpl == 2) { if (demod) CSDRB_DDC_LAUNCH(2, true); else CSDRB_DDC_LAUNCH(2, false); }
    else { if (demod) CSDRB_DDC_LAUNCH(1, true); else CSDRB_DDC_LAUNCH(1, false); }
#undef CSDRB_DDC_LAUNCH
    CSDRB_CUDA(cudaGetLastError());
    return 0;
}

// ---- the two halves of a block, exposed separately so a bank object can run the pre-pass of the NEXT block on a side stream ----
// pre-pass: float phase chain over the absolute chunks the block touches + the (cos, sin) seeds; advances d_phase_io to the chunk
// that contains the next block's first sample.
int launch_ddc_prepass(int input_size, int channels, const float* d_params, float* d_phase_io, int chunk, int offset, int decimation,
                       int taps_length, void* d_scratch, size_t scratch_bytes, cudaStream_t st)
{
    const int n_out = input_size >= taps_length ? (input_size - taps_length) / decimation + 1 : 0;
    if (n_out == 0) return 0;
    if (chunk <= 0) chunk = input_size;
    if (offset < 0 || offset >= chunk) { set_error("ddc bank: offset must be in [0, chunk)"); return -1; }
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
    return 2;
}

// main kernel: needs the seeds a matching launch_ddc_prepass() left in d_scratch
int launch_ddc_main(const float2* d_wide, int input_size, int channels, const float* d_params, int chunk, int offset, int decimation,
                    const float* h_taps, int taps_length, int demod, void* d_out, long out_stride, const float2* d_last_in, float2* d_last_out,
                    const void* d_scratch, cudaStream_t st)
{
    if (channels <= 0 || decimation <= 0 || taps_length <= 0) { set_error("ddc bank: bad geometry"); return -1; }
    const int n_out = input_size >= taps_length ? (input_size - taps_length) / decimation + 1 : 0;
    if (n_out == 0) return 0;
    if (chunk <= 0) chunk = input_size;
    if (reinterpret_cast<uintptr_t>(d_wide) & 15) { set_error("ddc bank: wideband input must be 16-byte aligned"); return -1; }
    const int nchunks = (int)(((long)offset + input_size + chunk - 1) / chunk) + 1;
    const float2* seeds = reinterpret_cast<const float2*>(static_cast<const char*>(d_scratch) + (((size_t)channels * nchunks * sizeof(float) + 15) & ~(size_t)15));
    int rc = -1;
    const float3* P = reinterpret_cast<const float3*>(d_params);
    if (decimation == 50 && taps_length <= 50 * 17) rc = launch_fused<50, 17>(d_wide, input_size, offset, chunk, nchunks, P, seeds, channels, demod, d_out, out_stride, n_out, d_last_in, d_last_out, h_taps, taps_length, st);
    else if (decimation == 10 && taps_length <= 10 * 8) rc = launch_fused<10, 8>(d_wide, input_size, offset, chunk, nchunks, P, seeds, channels, demod, d_out, out_stride, n_out, d_last_in, d_last_out, h_taps, taps_length, st);
    else if (decimation == 10 && taps_length <= 10 * 20) rc = launch_fused<10, 20>(d_wide, input_size, offset, chunk, nchunks, P, seeds, channels, demod, d_out, out_stride, n_out, d_last_in, d_last_out, h_taps, taps_length, st);
    else { set_error("ddc bank: no fused kernel for decimation %d / %d taps (compiled: d=50 T<=850, d=10 T<=200); run the unfused bank calls", decimation, taps_length); return -2; }
    return rc < 0 ? rc : n_out;
}

// one-shot: pre-pass and main kernel back to back on the caller's stream
int launch_ddc_bank(const float2* d_wide, int input_size, int channels, const float* d_params, float* d_phase_io, int chunk, int offset,
                    int decimation, const float* h_taps, int taps_length, int demod, void* d_out, long out_stride,
                    const float2* d_last_in, float2* d_last_out, void* d_scratch, size_t scratch_bytes, int* launches, cudaStream_t st)
{
    *launches = 0;
    if (channels <= 0 || decimation <= 0 || taps_length <= 0) { set_error("ddc bank: bad geometry"); return -1; }
    if (reinterpret_cast<uintptr_t>(d_wide) & 15) { set_error("ddc bank: wideband input must be 16-byte aligned"); return -1; }
    int rc = launch_ddc_prepass(input_size, channels, d_params, d_phase_io, chunk, offset, decimation, taps_length, d_scratch, scratch_bytes, st);
    if (rc <= 0) return rc;
    rc = launch_ddc_main(d_wide, input_size, channels, d_params, chunk, offset, decimation, h_taps, taps_length, demod, d_out, out_stride, d_last_in, d_last_out, d_scratch, st);
    if (rc < 0) return rc;
    *launches = 3;
    return rc;
}

}  // namespace csdrb
