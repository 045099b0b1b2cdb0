// capi.cu -- the C ABI of libcsdr_b200.so (see include/csdr_b200.h).
//
// Part B (csdrb_*) entry points are thin: validate, launch on the caller's stream, count the launch.
// Part A (libcsdr names) wraps Part B for HOST buffers: grow-only device workspace, one private
// stream, H2D -> kernel(s) -> D2H, synchronous per call -- what a drop-in for a CPU library has to be.
#include "common.cuh"
#include "kernels.h"
#include "csdr_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace csdrb {

static thread_local char g_err[512] = "";
static std::atomic<long> g_launches{0};

void set_error(const char* fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e), file, line, what);
    return -(1000 + (int)e);
}
static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int counted(int rc, int n = 1) { if (rc >= 0) g_launches += n; return rc; }

// ---- host-pointer workspace for Part A -----------------------------------------------------------
struct HostCtx {
    std::mutex mu;
    cudaStream_t stream = nullptr;
    void* buf[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t cap[4] = {0, 0, 0, 0};
    bool ready = false;
    int init()
    {
        if (ready) return 0;
        int n = 0;
        CSDRB_CUDA(cudaGetDeviceCount(&n));
        if (n <= 0) { set_error("no CUDA device visible"); return -1; }
        CSDRB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        ready = true;
        return 0;
    }
    int reserve(int slot, size_t bytes)
    {
        if (bytes <= cap[slot]) return 0;
        if (buf[slot]) CSDRB_CUDA(cudaFree(buf[slot]));
        size_t want = bytes + bytes / 2 + 4096;
        CSDRB_CUDA(cudaMalloc(&buf[slot], want));
        cap[slot] = want;
        return 0;
    }
};
static HostCtx g_ctx;

[[noreturn]] static void die(const char* who)
{
    fprintf(stderr, "libcsdr_b200: %s failed: %s\n", who, g_err[0] ? g_err : "(no detail)");
    abort();
}
#define A_CHECK(expr, who) do { if ((expr) < 0) die(who); } while (0)
#define A_CUDA(call, who) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cuda_fail(e_, #call, __FILE__, __LINE__); die(who); } } while (0)

}  // namespace csdrb

using namespace csdrb;

extern "C" {

// =====================================================================================================
// Part B
// =====================================================================================================
const char* csdrb_last_error(void) { return g_err; }
const char* csdrb_version(void) { return "csdr_b200 0.1 (sm_100a)"; }
int csdrb_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
    return n;
}
int csdrb_set_device(int device) { CSDRB_CUDA(cudaSetDevice(device)); return 0; }
int csdrb_stream_synchronize(void* stream) { CSDRB_CUDA(cudaStreamSynchronize(S(stream))); return 0; }
long csdrb_kernel_launches(void) { return g_launches.load(); }

int csdrb_convert_u8_f(const unsigned char* d_in, float* d_out, long n, void* stream) { return counted(launch_convert_u8_f(d_in, d_out, n, S(stream))); }
int csdrb_convert_s16_f(const short* d_in, float* d_out, long n, void* stream) { return counted(launch_convert_s16_f(d_in, d_out, n, S(stream))); }
int csdrb_convert_f_s16(const float* d_in, short* d_out, long n, void* stream) { return counted(launch_convert_f_s16(d_in, d_out, n, S(stream))); }

int csdrb_fir_bank_variants(void) { return fir_bank_variant_count(); }

int csdrb_fir_decimate_bank_cc(const complexf* d_in, long in_stride, complexf* d_out, long out_stride, int channels,
                               int input_size, int decimation, const float* h_taps, int taps_length, int variant, void* stream)
{
    if (!d_in || !d_out || !h_taps) { set_error("fir_decimate bank: null pointer"); return -1; }
    // the generic kernel reads taps from device memory: keep a small per-process copy
    static float* d_taps = nullptr; static int d_taps_cap = 0; static std::mutex mu;
    const float* dt = nullptr;
    const bool fast = (decimation == 10 && taps_length <= 200 && (in_stride % 2 == 0) && ((reinterpret_cast<uintptr_t>(d_in) & 15) == 0));
    if (!fast) {
        std::lock_guard<std::mutex> lk(mu);
        if (taps_length > d_taps_cap) {
            if (d_taps) CSDRB_CUDA(cudaFree(d_taps));
            CSDRB_CUDA(cudaMalloc(&d_taps, sizeof(float) * (size_t)taps_length * 2));
            d_taps_cap = taps_length * 2;
        }
        CSDRB_CUDA(cudaMemcpyAsync(d_taps, h_taps, sizeof(float) * (size_t)taps_length, cudaMemcpyHostToDevice, S(stream)));
        dt = d_taps;
    }
    return counted(launch_fir_decimate_bank(reinterpret_cast<const float2*>(d_in), in_stride, reinterpret_cast<float2*>(d_out), out_stride,
                                            channels, input_size, decimation, h_taps, dt, 0, taps_length, variant, S(stream)));
}

int csdrb_fmdemod_quadri_bank_cf(const complexf* d_in, long in_stride, float* d_out, long out_stride, int channels,
                                 int input_size, const complexf* d_last_in, complexf* d_last_out, void* stream)
{
    if (!d_in || !d_out) { set_error("fmdemod_quadri bank: null pointer"); return -1; }
    return counted(launch_fmdemod_quadri_bank(reinterpret_cast<const float2*>(d_in), in_stride, d_out, out_stride, channels, input_size,
                                              reinterpret_cast<const float2*>(d_last_in), reinterpret_cast<float2*>(d_last_out), S(stream)));
}


// ---- host-buffer bank call: the e2e path -------------------------------------------------------------
// Streams a [channels][input_size] HOST bank through the device in channel chunks on three streams so
// that the H2D copy of chunk i+1, the kernel of chunk i and the D2H copy of chunk i-1 overlap (PCIe is
// full duplex).  Host buffers should be page-locked (csdrb_host_alloc) -- pageable memory works but is
// staged by the driver.  Synchronous: returns when h_out is complete.
} // extern C (reopened below)
namespace csdrb {
struct HostBank {
    static constexpr int NS = 3;
    cudaStream_t st[NS] = {nullptr, nullptr, nullptr};
    void* din[NS] = {nullptr, nullptr, nullptr};
    void* dout[NS] = {nullptr, nullptr, nullptr};
    size_t cin = 0, cout = 0;
    std::mutex mu;
    int ensure(size_t bin, size_t bout)
    {
        for (int k = 0; k < NS; k++) if (!st[k]) CSDRB_CUDA(cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking));
        if (bin > cin) {
            for (int k = 0; k < NS; k++) { if (din[k]) CSDRB_CUDA(cudaFree(din[k])); CSDRB_CUDA(cudaMalloc(&din[k], bin)); }
            cin = bin;
        }
        if (bout > cout) {
            for (int k = 0; k < NS; k++) { if (dout[k]) CSDRB_CUDA(cudaFree(dout[k])); CSDRB_CUDA(cudaMalloc(&dout[k], bout)); }
            cout = bout;
        }
        return 0;
    }
};
static HostBank g_hb;
}  // namespace csdrb
extern "C" {

void* csdrb_host_alloc(size_t bytes)
{
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { cuda_fail(e, "cudaHostAlloc", __FILE__, __LINE__); return nullptr; }
    return p;
}
void csdrb_host_free(void* p) { if (p) cudaFreeHost(p); }

int csdrb_fir_decimate_bank_cc_host(const complexf* h_in, long in_stride, complexf* h_out, long out_stride, int channels,
                                    int input_size, int decimation, const float* h_taps, int taps_length, int chunk_channels)
{
    if (!h_in || !h_out || !h_taps || channels <= 0 || decimation <= 0 || taps_length <= 0) { set_error("fir_decimate host bank: bad argument"); return -1; }
    const int n_out = input_size >= taps_length ? (input_size - taps_length) / decimation + 1 : 0;
    if (n_out == 0) return 0;
    std::lock_guard<std::mutex> lk(g_hb.mu);
    const long dstride_in = (input_size + 1) & ~1L, dstride_out = (n_out + 1) & ~1L;
    if (chunk_channels <= 0) {                                 // ~192 MiB of input per chunk keeps all three stages busy
        chunk_channels = (int)((192L << 20) / (dstride_in * 8));
        if (chunk_channels < 1) chunk_channels = 1;
    }
    if (chunk_channels > channels) chunk_channels = channels;
    if (int rc = g_hb.ensure((size_t)chunk_channels * dstride_in * 8, (size_t)chunk_channels * dstride_out * 8)) return rc;
    int slot = 0;
    for (int c0 = 0; c0 < channels; c0 += chunk_channels, slot = (slot + 1) % HostBank::NS) {
        const int nc = channels - c0 < chunk_channels ? channels - c0 : chunk_channels;
        cudaStream_t s = g_hb.st[slot];
        CSDRB_CUDA(cudaMemcpy2DAsync(g_hb.din[slot], (size_t)dstride_in * 8, h_in + (long)c0 * in_stride, (size_t)in_stride * 8,
                                     (size_t)input_size * 8, nc, cudaMemcpyHostToDevice, s));
        int rc = csdrb_fir_decimate_bank_cc((const complexf*)g_hb.din[slot], dstride_in, (complexf*)g_hb.dout[slot], dstride_out, nc,
                                            input_size, decimation, h_taps, taps_length, -1, s);
        if (rc < 0) return rc;
        CSDRB_CUDA(cudaMemcpy2DAsync(h_out + (long)c0 * out_stride, (size_t)out_stride * 8, g_hb.dout[slot], (size_t)dstride_out * 8,
                                     (size_t)n_out * 8, nc, cudaMemcpyDeviceToHost, s));
    }
    for (int k = 0; k < HostBank::NS; k++) CSDRB_CUDA(cudaStreamSynchronize(g_hb.st[k]));
    return n_out;
}

// =====================================================================================================
// Part A -- host-pointer drop-ins
// =====================================================================================================
void convert_u8_f(unsigned char* input, float* output, int input_size)
{
    if (input_size <= 0) return;
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    A_CHECK(g_ctx.init(), "convert_u8_f");
    A_CHECK(g_ctx.reserve(0, (size_t)input_size), "convert_u8_f");
    A_CHECK(g_ctx.reserve(1, (size_t)input_size * 4), "convert_u8_f");
    A_CUDA(cudaMemcpyAsync(g_ctx.buf[0], input, (size_t)input_size, cudaMemcpyHostToDevice, g_ctx.stream), "convert_u8_f");
    A_CHECK(csdrb_convert_u8_f((const unsigned char*)g_ctx.buf[0], (float*)g_ctx.buf[1], input_size, g_ctx.stream), "convert_u8_f");
    A_CUDA(cudaMemcpyAsync(output, g_ctx.buf[1], (size_t)input_size * 4, cudaMemcpyDeviceToHost, g_ctx.stream), "convert_u8_f");
    A_CUDA(cudaStreamSynchronize(g_ctx.stream), "convert_u8_f");
}

void convert_s16_f(short* input, float* output, int input_size)
{
    if (input_size <= 0) return;
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    A_CHECK(g_ctx.init(), "convert_s16_f");
    A_CHECK(g_ctx.reserve(0, (size_t)input_size * 2), "convert_s16_f");
    A_CHECK(g_ctx.reserve(1, (size_t)input_size * 4), "convert_s16_f");
    A_CUDA(cudaMemcpyAsync(g_ctx.buf[0], input, (size_t)input_size * 2, cudaMemcpyHostToDevice, g_ctx.stream), "convert_s16_f");
    A_CHECK(csdrb_convert_s16_f((const short*)g_ctx.buf[0], (float*)g_ctx.buf[1], input_size, g_ctx.stream), "convert_s16_f");
    A_CUDA(cudaMemcpyAsync(output, g_ctx.buf[1], (size_t)input_size * 4, cudaMemcpyDeviceToHost, g_ctx.stream), "convert_s16_f");
    A_CUDA(cudaStreamSynchronize(g_ctx.stream), "convert_s16_f");
}
void convert_i16_f(short* input, float* output, int input_size) { convert_s16_f(input, output, input_size); }

void convert_f_s16(float* input, short* output, int input_size)
{
    if (input_size <= 0) return;
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    A_CHECK(g_ctx.init(), "convert_f_s16");
    A_CHECK(g_ctx.reserve(0, (size_t)input_size * 4), "convert_f_s16");
    A_CHECK(g_ctx.reserve(1, (size_t)input_size * 2), "convert_f_s16");
    A_CUDA(cudaMemcpyAsync(g_ctx.buf[0], input, (size_t)input_size * 4, cudaMemcpyHostToDevice, g_ctx.stream), "convert_f_s16");
    A_CHECK(csdrb_convert_f_s16((const float*)g_ctx.buf[0], (short*)g_ctx.buf[1], input_size, g_ctx.stream), "convert_f_s16");
    A_CUDA(cudaMemcpyAsync(output, g_ctx.buf[1], (size_t)input_size * 2, cudaMemcpyDeviceToHost, g_ctx.stream), "convert_f_s16");
    A_CUDA(cudaStreamSynchronize(g_ctx.stream), "convert_f_s16");
}
void convert_f_i16(float* input, short* output, int input_size) { convert_f_s16(input, output, input_size); }

int fir_decimate_cc(complexf* input, complexf* output, int input_size, int decimation, float* taps, int taps_length)
{
    if (input_size < taps_length || input_size <= 0) return 0;
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    A_CHECK(g_ctx.init(), "fir_decimate_cc");
    const int n_out = (input_size - taps_length) / decimation + 1;
    A_CHECK(g_ctx.reserve(0, (size_t)input_size * 8 + 16), "fir_decimate_cc");
    A_CHECK(g_ctx.reserve(1, (size_t)n_out * 8 + 16), "fir_decimate_cc");
    A_CUDA(cudaMemcpyAsync(g_ctx.buf[0], input, (size_t)input_size * 8, cudaMemcpyHostToDevice, g_ctx.stream), "fir_decimate_cc");
    int rc = csdrb_fir_decimate_bank_cc((const complexf*)g_ctx.buf[0], (input_size + 1) & ~1, (complexf*)g_ctx.buf[1], (n_out + 1) & ~1, 1,
                                        input_size, decimation, taps, taps_length, -1, g_ctx.stream);
    A_CHECK(rc, "fir_decimate_cc");
    A_CUDA(cudaMemcpyAsync(output, g_ctx.buf[1], (size_t)rc * 8, cudaMemcpyDeviceToHost, g_ctx.stream), "fir_decimate_cc");
    A_CUDA(cudaStreamSynchronize(g_ctx.stream), "fir_decimate_cc");
    return rc;
}

complexf fmdemod_quadri_cf(complexf* input, float* output, int input_size, float* temp, complexf last_sample)
{
    (void)temp;
    if (input_size <= 0) return last_sample;
    std::lock_guard<std::mutex> lk(g_ctx.mu);
    A_CHECK(g_ctx.init(), "fmdemod_quadri_cf");
    A_CHECK(g_ctx.reserve(0, (size_t)input_size * 8 + 16), "fmdemod_quadri_cf");
    A_CHECK(g_ctx.reserve(1, (size_t)input_size * 4 + 16), "fmdemod_quadri_cf");
    A_CHECK(g_ctx.reserve(2, 64), "fmdemod_quadri_cf");
    A_CUDA(cudaMemcpyAsync(g_ctx.buf[0], input, (size_t)input_size * 8, cudaMemcpyHostToDevice, g_ctx.stream), "fmdemod_quadri_cf");
    A_CUDA(cudaMemcpyAsync(g_ctx.buf[2], &last_sample, 8, cudaMemcpyHostToDevice, g_ctx.stream), "fmdemod_quadri_cf");
    A_CHECK(csdrb_fmdemod_quadri_bank_cf((const complexf*)g_ctx.buf[0], (input_size + 1) & ~1, (float*)g_ctx.buf[1], (input_size + 1) & ~1, 1,
                                         input_size, (const complexf*)g_ctx.buf[2], nullptr, g_ctx.stream), "fmdemod_quadri_cf");
    A_CUDA(cudaMemcpyAsync(output, g_ctx.buf[1], (size_t)input_size * 4, cudaMemcpyDeviceToHost, g_ctx.stream), "fmdemod_quadri_cf");
    A_CUDA(cudaStreamSynchronize(g_ctx.stream), "fmdemod_quadri_cf");
    return input[input_size - 1];
}

}  // extern "C"
