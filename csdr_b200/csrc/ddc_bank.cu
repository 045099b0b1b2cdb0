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

// CPL = channels per lane.  A tap pair is loaded once per warp and sample phase (LDCU.128 fetches two of them: taps are stored
// [p][j] so the M taps that one sample meets are contiguous); with CPL = 2 every loaded tap feeds two FFMA2 and every wideband
// sample load feeds two channels, which halves the non-FMA issue slots per channel-sample (ncu r01: 868 LDCU per 850 FFMA2 at CPL = 1).
// Pipelined roles inside ONE launch (r01: the serial phase-chain pre-pass cost as much as the main kernel at multi-M-sample blocks):
// every CTA draws a ticket; the first `groups` tickets become PRODUCERS (one per 128-channel group: a thread per channel walks the
// float phase chain chunk by chunk and publishes chunk_phase[c][k] + a per-chunk flag), every later ticket is a CONSUMER that maps
// to (segment, group) in time order and only waits when it reaches a chunk whose phase is not published yet.  Tickets are handed
// out in the order CTAs start running, so a producer is always resident before any consumer of its group can wait on it.
struct DdcSync { unsigned ticket; unsigned pad[3]; };

__device__ __forceinline__ void ddc_wait_flag(const volatile unsigned* flag)
{
    while (*flag == 0u) __nanosleep(64);
    __threadfence();
}

template <int D, int M, int CPL, bool DEMOD>
__global__ void __launch_bounds__(128)
ddc_bank_fused_kernel(const float2* __restrict__ wide, int n_in, int offset, int chunk, int nchunks,
                      const float3* __restrict__ params, float* __restrict__ phase_io, float* __restrict__ chunk_phase,
                      unsigned* __restrict__ flags, DdcSync* __restrict__ sync, int next_chunk, int groups, int channels,
                      void* __restrict__ out_v, long out_stride, int n_out, int seg_outputs,
                      const float2* __restrict__ last_in, float2* __restrict__ last_out,
                      const __grid_constant__ DdcTaps<D * ((M + 1) & ~1)> taps)
{
    constexpr int MP = (M + 1) & ~1;
    constexpr int CH_PER_GROUP = 128 * CPL;
    __shared__ unsigned s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(&sync->ticket, 1u);
    __syncthreads();
    const unsigned ticket = s_ticket;
    if (ticket < (unsigned)groups) {
        // ---------------- producer: the float phase chain of this group's channels, published chunk by chunk ----------------
        const int g = (int)ticket;
        float ph[CPL], keep[CPL], rate2[CPL]; int chn[CPL];
#pragma unroll
        for (int u = 0; u < CPL; u++) {
            chn[u] = g * CH_PER_GROUP + u * 128 + threadIdx.x;
            const bool ok = chn[u] < channels;
            ph[u] = ok ? phase_io[chn[u]] : 0.f; keep[u] = ph[u]; rate2[u] = ok ? params[chn[u]].z : 0.f;
        }
        for (int k = 0; k < nchunks; k++) {
#pragma unroll
            for (int u = 0; u < CPL; u++) {
                if (chn[u] < channels) chunk_phase[(long)chn[u] * nchunks + k] = ph[u];
                if (k == next_chunk) keep[u] = ph[u];
                ph[u] = wrap_phase_pm_pi(__fadd_rn(ph[u], __fmul_rn(__fmul_rn(rate2[u], 3.14159265358979323846f), (float)chunk)));
            }
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) { *reinterpret_cast<volatile unsigned*>(flags + (long)g * nchunks + k) = 1u; }
        }
        if (next_chunk >= nchunks) {
            for (int k = nchunks; k <= next_chunk; k++) {
#pragma unroll
                for (int u = 0; u < CPL; u++) {
                    if (k == next_chunk) keep[u] = ph[u];
                    ph[u] = wrap_phase_pm_pi(__fadd_rn(ph[u], __fmul_rn(__fmul_rn(rate2[u], 3.14159265358979323846f), (float)chunk)));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < CPL; u++) if (chn[u] < channels) phase_io[chn[u]] = keep[u];
        return;
    }
    // ---------------- consumer ----------------
    const unsigned work = ticket - (unsigned)groups;
    const int seg_idx = (int)(work / (unsigned)groups), grp = (int)(work % (unsigned)groups);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch0 = grp * CH_PER_GROUP + warp * (32 * CPL) + lane;      // this lane's channels: ch0 + 32*u
    if (ch0 - lane >= channels) return;                                 // whole warp beyond the bank
    const volatile unsigned* gflags = flags + (long)grp * nchunks;
    const int o_first = seg_idx * seg_outputs;                          // first output this warp emits
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
    }
    ddc_wait_flag(gflags + kchunk);
#pragma unroll
    for (int u = 0; u < CPL; u++) {                                     // seed = cos/sin of the float phase, evaluated in double like the reference
        const double ph = (double)__ldcg(chunk_phase + (long)chs[u] * nchunks + kchunk);
        c[u] = (float)cos(ph); s[u] = (float)sin(ph);
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
                    ddc_wait_flag(gflags + kchunk);
#pragma unroll
                    for (int u = 0; u < CPL; u++) {
                        const double ph = (double)__ldcg(chunk_phase + (long)chs[u] * nchunks + kchunk);
                        c[u] = (float)cos(ph); s[u] = (float)sin(ph);
                    }
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
    const long groups = (channels + 127) / 128;
    return 64 + (size_t)groups * (size_t)nchunks * sizeof(unsigned) + (size_t)channels * (size_t)nchunks * sizeof(float) + 64;
}

template <int D, int M>
static int launch_fused(const float2* wide, int n_in, int offset, int chunk, int nchunks, const float3* params, float* phase_io, void* scratch,
                        int next_chunk, int channels, int demod, void* out, long out_stride, int n_out, const float2* last_in, float2* last_out,
                        const float* h_taps, int T, cudaStream_t st)
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
    const int groups = (channels + 128 * cpl - 1) / (128 * cpl);
    const int warps_per_seg = (channels + 32 * cpl - 1) / (32 * cpl);
    // enough warps to fill the machine while keeping the M-1 trailing periods of every segment a small fraction
    long want_segments = (148L * wps_env + warps_per_seg - 1) / warps_per_seg;
    int seg = (int)((n_out + want_segments - 1) / want_segments);
    if (seg < 2 * M) seg = 2 * M;
    const int nseg = (n_out + seg - 1) / seg;
    // scratch: [sync 64 B][flags groups*nchunks][chunk_phase channels*nchunks]; sync + flags are zeroed every call
    DdcSync* sync = static_cast<DdcSync*>(scratch);
    unsigned* flags = reinterpret_cast<unsigned*>(static_cast<char*>(scratch) + 64);
    float* chunk_phase = reinterpret_cast<float*>(flags + (size_t)groups * nchunks);
    CSDRB_CUDA(cudaMemsetAsync(scratch, 0, 64 + (size_t)groups * nchunks * sizeof(unsigned), st));
    const dim3 grid((unsigned)(nseg * groups + groups));
#define CSDRB_DDC_LAUNCH(CPLV, DM) ddc_bank_fused_kernel<D, M, CPLV, DM><<<grid, 128, 0, st>>>(wide, n_in, offset, chunk, nchunks, params, phase_io, chunk_phase, \
        flags, sync, next_chunk, groups, channels, out, out_stride, n_out, seg, last_in, last_out, tp)
    if (cpl == 2) { if (demod) CSDRB_DDC_LAUNCH(2, true); else CSDRB_DDC_LAUNCH(2, false); }
    else { if (demod) CSDRB_DDC_LAUNCH(1, true); else CSDRB_DDC_LAUNCH(1, false); }
#undef CSDRB_DDC_LAUNCH
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
    if (reinterpret_cast<uintptr_t>(d_scratch) & 15) { set_error("ddc bank: scratch must be 16-byte aligned"); return -1; }
    const int nchunks = (int)(((long)offset + input_size + chunk - 1) / chunk) + 1;
    const long advance = (long)n_out * decimation;                      // the next block starts here (the caller re-presents the tail)
    const int next_chunk = (int)((offset + advance) / chunk);
    int rc = -1;
    const float3* P = reinterpret_cast<const float3*>(d_params);
    if (decimation == 50 && taps_length <= 50 * 17) rc = launch_fused<50, 17>(d_wide, input_size, offset, chunk, nchunks, P, d_phase_io, d_scratch, next_chunk, channels, demod, d_out, out_stride, n_out, d_last_in, d_last_out, h_taps, taps_length, st);
    else if (decimation == 10 && taps_length <= 10 * 8) rc = launch_fused<10, 8>(d_wide, input_size, offset, chunk, nchunks, P, d_phase_io, d_scratch, next_chunk, channels, demod, d_out, out_stride, n_out, d_last_in, d_last_out, h_taps, taps_length, st);
    else if (decimation == 10 && taps_length <= 10 * 20) rc = launch_fused<10, 20>(d_wide, input_size, offset, chunk, nchunks, P, d_phase_io, d_scratch, next_chunk, channels, demod, d_out, out_stride, n_out, d_last_in, d_last_out, h_taps, taps_length, st);
    else { set_error("ddc bank: no fused kernel for decimation %d / %d taps (compiled: d=50 T<=850, d=10 T<=200); run the unfused bank calls", decimation, taps_length); return -2; }
    if (rc < 0) return rc;
    *launches = 1;
    return n_out;
}

}  // namespace csdrb
