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
#include <vector>
#include <sched.h>

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
// one workspace per device (a process may csdrb_set_device() between calls): the buffers, like the stream, belong to the device that was
// current when they were made
static HostCtx& host_ctx()
{
    static std::mutex mu;
    static std::vector<HostCtx*> per_dev;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    std::lock_guard<std::mutex> lk(mu);
    if ((size_t)dev >= per_dev.size()) per_dev.resize((size_t)dev + 1, nullptr);
    if (!per_dev[(size_t)dev]) per_dev[(size_t)dev] = new HostCtx();
    return *per_dev[(size_t)dev];
}
#define g_ctx (host_ctx())

// CSDRB_TRACE=1: report at exit how many kernels this process launched (lets a caller verify that a
// preloaded/linked libcsdr_b200 really did the work instead of some other libcsdr).
struct ExitReport {
    ~ExitReport()
    {
        const char* t = getenv("CSDRB_TRACE");
        if (t && *t && *t != '0') fprintf(stderr, "libcsdr_b200: %ld kernel launches in this process\n", g_launches.load());
    }
};
static ExitReport g_exit_report;

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

// channels ride in gridDim.y (limit 65535) in the bank kernels: refuse larger banks with a message instead of an "invalid configuration" launch error
static inline bool too_many_channels(int channels, const char* who)
{
    if (channels <= 65535) return false;
    set_error("%s: %d channels in one call (at most 65535; split the bank)", who, channels);
    return true;
}
static inline bool null_io(const void* in, const void* out, const char* who) { if (in && out) return false; set_error("%s: null pointer", who); return true; }
int csdrb_convert_u8_f(const unsigned char* d_in, float* d_out, long n, void* stream)
{
    return null_io(d_in, d_out, "convert_u8_f") ? -1 : counted(launch_convert_u8_f(d_in, d_out, n, S(stream)));
}
int csdrb_convert_s16_f(const short* d_in, float* d_out, long n, void* stream)
{
    return null_io(d_in, d_out, "convert_s16_f") ? -1 : counted(launch_convert_s16_f(d_in, d_out, n, S(stream)));
}
int csdrb_convert_f_s16(const float* d_in, short* d_out, long n, void* stream)
{
    return null_io(d_in, d_out, "convert_f_s16") ? -1 : counted(launch_convert_f_s16(d_in, d_out, n, S(stream)));
}

int csdrb_fir_bank_variants(void) { return fir_bank_variant_count(); }

int csdrb_fir_decimate_bank_cc(const complexf* d_in, long in_stride, complexf* d_out, long out_stride, int channels,
                               int input_size, int decimation, const float* h_taps, int taps_length, int variant, void* stream)
{
    if (too_many_channels(channels, "fir_decimate bank")) return -1;
    if (!d_in || !d_out || !h_taps) { set_error("fir_decimate bank: null pointer"); return -1; }
    // the generic kernel reads its taps from device memory: a stream-ordered allocation per call, so that two callers on different streams (or
    // devices) never share a buffer that one of them is still reading
    float* dt = nullptr;
    const bool fast = ((decimation == 10 && taps_length <= 200) || (decimation == 50 && taps_length <= 900)) && (in_stride % 2 == 0) &&
                      ((reinterpret_cast<uintptr_t>(d_in) & 15) == 0);
    if (!fast) {
        CSDRB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&dt), sizeof(float) * (size_t)taps_length, S(stream)));
        CSDRB_CUDA(cudaMemcpyAsync(dt, h_taps, sizeof(float) * (size_t)taps_length, cudaMemcpyHostToDevice, S(stream)));
        CSDRB_CUDA(cudaStreamSynchronize(S(stream)));       // h_taps is the caller's (possibly pageable) memory: it may change once we return
    }
    const int rc = counted(launch_fir_decimate_bank(reinterpret_cast<const float2*>(d_in), in_stride, reinterpret_cast<float2*>(d_out), out_stride,
                                                    channels, input_size, decimation, h_taps, dt, 0, taps_length, variant, S(stream)));
    if (dt) CSDRB_CUDA(cudaFreeAsync(dt, S(stream)));
    return rc;
}

// convert_u8_f | fir_decimate_cc in one launch for rtl_sdr-style input: d_in holds interleaved unsigned 8-bit I,Q (2 bytes per sample), in_stride counts
// SAMPLES between channel rows.  Fused for the compiled tilings (d=10 T<=200, d=50 T<=900) when rows start on 16-byte boundaries (in_stride % 8 == 0);
// any other geometry converts into a stream-ordered temporary and runs the cf32 bank.
int csdrb_fir_decimate_bank_u8_cc(const unsigned char* d_in, long in_stride, complexf* d_out, long out_stride, int channels,
                                  int input_size, int decimation, const float* h_taps, int taps_length, void* stream)
{
    if (too_many_channels(channels, "fir_decimate u8 bank")) return -1;
    if (!d_in || !d_out || !h_taps) { set_error("fir_decimate u8 bank: null pointer"); return -1; }
    int rc = launch_fir_decimate_bank_u8(d_in, in_stride, reinterpret_cast<float2*>(d_out), out_stride, channels, input_size, decimation, h_taps, taps_length, S(stream));
    if (rc != -2) return counted(rc);
    const long fstride = (input_size + 1) & ~1L;
    float* tmp = nullptr;
    CSDRB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&tmp), sizeof(float) * 2 * (size_t)fstride * channels, S(stream)));
    rc = launch_u8_rows_to_cf32(d_in, in_stride, reinterpret_cast<float2*>(tmp), fstride, channels, input_size, S(stream));
    if (rc < 0) { cudaFreeAsync(tmp, S(stream)); return rc; }
    g_launches += 1;
    rc = csdrb_fir_decimate_bank_cc(reinterpret_cast<const complexf*>(tmp), fstride, d_out, out_stride, channels, input_size, decimation, h_taps, taps_length, -1, stream);
    CSDRB_CUDA(cudaFreeAsync(tmp, S(stream)));
    return rc;
}

int csdrb_fmdemod_quadri_bank_cf(const complexf* d_in, long in_stride, float* d_out, long out_stride, int channels,
                                 int input_size, const complexf* d_last_in, complexf* d_last_out, void* stream)
{
    if (too_many_channels(channels, "fmdemod_quadri bank")) return -1;
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
static HostBank& host_bank()                                     // one per device, like the Part A workspace
{
    static std::mutex mu;
    static std::vector<HostBank*> per_dev;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    std::lock_guard<std::mutex> lk(mu);
    if ((size_t)dev >= per_dev.size()) per_dev.resize((size_t)dev + 1, nullptr);
    if (!per_dev[(size_t)dev]) per_dev[(size_t)dev] = new HostBank();
    return *per_dev[(size_t)dev];
}

// Page-locked memory on the NUMA node the current device hangs off.  cudaHostAlloc places the pages where the calling thread runs; with one
// process per GPU on a two-socket host half the ranks would otherwise stream their H2D traffic across the socket interconnect (SCALE_r01: e2e
// efficiency 0.56 at 8 GPUs).  The thread is moved onto the device's node for the duration of the allocation (sysfs: the PCI device's
// numa_node and that node's cpulist), then gets its old affinity back.  Any failure along the way just leaves the default placement.
static bool cpus_of_device_node(cpu_set_t* set)
{
    int dev = 0; char bus[32] = "";
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(bus, sizeof bus, dev) != cudaSuccess) return false;
    for (char* p = bus; *p; p++) if (*p >= 'A' && *p <= 'Z') *p = (char)(*p - 'A' + 'a');
    char path[128]; snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r"); if (!f) return false;
    int node = -1; const int got = fscanf(f, "%d", &node); fclose(f);
    if (got != 1 || node < 0) return false;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r"); if (!f) return false;
    char list[4096] = ""; const bool ok = fgets(list, sizeof list, f) != nullptr; fclose(f);
    if (!ok) return false;
    CPU_ZERO(set);
    int n = 0;
    for (char* p = list; *p && *p != '\n';) {                         // "0-31,64-95"
        char* e; long a = strtol(p, &e, 10); if (e == p) break; long b = a;
        if (*e == '-') { p = e + 1; b = strtol(p, &e, 10); }
        for (long c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET((int)c, set); n++; }
        p = (*e == ',') ? e + 1 : e;
    }
    return n > 0;
}
}  // namespace csdrb
extern "C" {

void* csdrb_host_alloc(size_t bytes)
{
    cpu_set_t old_set, node_set;
    const bool have_old = sched_getaffinity(0, sizeof old_set, &old_set) == 0;
    const bool moved = have_old && !(getenv("CSDRB_NO_NUMA") && getenv("CSDRB_NO_NUMA")[0] == '1') && cpus_of_device_node(&node_set) &&
                       sched_setaffinity(0, sizeof node_set, &node_set) == 0;
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e == cudaSuccess && moved) memset(p, 0, bytes ? bytes : 1);  // first touch from the node, in case the driver only reserved the range
    if (moved) sched_setaffinity(0, sizeof old_set, &old_set);
    if (e != cudaSuccess) { cuda_fail(e, "cudaHostAlloc", __FILE__, __LINE__); return nullptr; }
    return p;
}
void csdrb_host_free(void* p) { if (p) cudaFreeHost(p); }

// in_bytes = bytes per input sample: 8 (cf32) or 2 (u8 IQ, converted inside the FIR kernel)
static int fir_bank_host(const void* h_in, int in_bytes, long in_stride, complexf* h_out, long out_stride, int channels,
                         int input_size, int decimation, const float* h_taps, int taps_length, int chunk_channels)
{
    if (!h_in || !h_out || !h_taps || channels <= 0 || decimation <= 0 || taps_length <= 0) { set_error("fir_decimate host bank: bad argument"); return -1; }
    const int n_out = input_size >= taps_length ? (input_size - taps_length) / decimation + 1 : 0;
    if (n_out == 0) return 0;
    HostBank& hb = host_bank();
    std::lock_guard<std::mutex> lk(hb.mu);
    const long dstride_in = in_bytes == 2 ? (input_size + 7) & ~7L : (input_size + 1) & ~1L, dstride_out = (n_out + 1) & ~1L;
    if (chunk_channels <= 0) {                                 // ~192 MiB of cf32 input (48 MiB of u8) per chunk keeps all three stages busy
        chunk_channels = (int)((192L << 20) / (dstride_in * 8));
        if (chunk_channels < 1) chunk_channels = 1;
    }
    if (chunk_channels > channels) chunk_channels = channels;
    if (int rc = hb.ensure((size_t)chunk_channels * dstride_in * in_bytes, (size_t)chunk_channels * dstride_out * 8)) return rc;
    int slot = 0;
    for (int c0 = 0; c0 < channels; c0 += chunk_channels, slot = (slot + 1) % HostBank::NS) {
        const int nc = channels - c0 < chunk_channels ? channels - c0 : chunk_channels;
        cudaStream_t s = hb.st[slot];
        CSDRB_CUDA(cudaMemcpy2DAsync(hb.din[slot], (size_t)dstride_in * in_bytes, static_cast<const char*>(h_in) + (long)c0 * in_stride * in_bytes,
                                     (size_t)in_stride * in_bytes, (size_t)input_size * in_bytes, nc, cudaMemcpyHostToDevice, s));
        int rc = in_bytes == 2
            ? csdrb_fir_decimate_bank_u8_cc((const unsigned char*)hb.din[slot], dstride_in, (complexf*)hb.dout[slot], dstride_out, nc, input_size, decimation, h_taps, taps_length, s)
            : csdrb_fir_decimate_bank_cc((const complexf*)hb.din[slot], dstride_in, (complexf*)hb.dout[slot], dstride_out, nc, input_size, decimation, h_taps, taps_length, -1, s);
        if (rc < 0) return rc;
        CSDRB_CUDA(cudaMemcpy2DAsync(h_out + (long)c0 * out_stride, (size_t)out_stride * 8, hb.dout[slot], (size_t)dstride_out * 8,
                                     (size_t)n_out * 8, nc, cudaMemcpyDeviceToHost, s));
    }
    for (int k = 0; k < HostBank::NS; k++) CSDRB_CUDA(cudaStreamSynchronize(hb.st[k]));
    return n_out;
}

int csdrb_fir_decimate_bank_cc_host(const complexf* h_in, long in_stride, complexf* h_out, long out_stride, int channels,
                                    int input_size, int decimation, const float* h_taps, int taps_length, int chunk_channels)
{
    return fir_bank_host(h_in, 8, in_stride, h_out, out_stride, channels, input_size, decimation, h_taps, taps_length, chunk_channels);
}

// the same with rtl_sdr-style u8 IQ on the host side (2 bytes per sample over PCIe instead of 8): convert_u8_f | fir_decimate_cc, csdr-fm:41
int csdrb_fir_decimate_bank_u8_host(const unsigned char* h_in, long in_stride, complexf* h_out, long out_stride, int channels,
                                    int input_size, int decimation, const float* h_taps, int taps_length, int chunk_channels)
{
    return fir_bank_host(h_in, 2, in_stride, h_out, out_stride, channels, input_size, decimation, h_taps, taps_length, chunk_channels);
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

// =====================================================================================================
// Part B, continued: K2, K5-K9
// =====================================================================================================
extern "C" {

size_t csdrb_shift_addition_bank_scratch_bytes(int channels, int input_size, int chunk) { return shift_bank_scratch_bytes(channels, input_size, chunk); }

int csdrb_shift_addition_bank_cc(const complexf* d_in, long in_stride, complexf* d_out, long out_stride, int channels, int input_size,
                                 const shift_addition_data_t* d_params, float* d_phase_io, int chunk, void* d_scratch, size_t scratch_bytes, void* stream)
{
    if (too_many_channels(channels, "shift_addition bank")) return -1;
    if (!d_in || !d_out || !d_params || !d_phase_io) { set_error("shift_addition bank: null pointer"); return -1; }
    int rc = launch_shift_addition_bank(reinterpret_cast<const float2*>(d_in), in_stride, reinterpret_cast<float2*>(d_out), out_stride, channels, input_size,
                                        reinterpret_cast<const float*>(d_params), d_phase_io, chunk, d_scratch, scratch_bytes, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}

size_t csdrb_shift_math_bank_scratch_bytes(int channels, int input_size) { return shift_math_scratch_bytes(channels, input_size); }

int csdrb_shift_math_bank_cc(const complexf* d_in, long in_stride, complexf* d_out, long out_stride, int channels, int input_size,
                             const float* d_rates, float* d_phase_io, void* d_scratch, size_t scratch_bytes, void* stream)
{
    if (too_many_channels(channels, "shift_math bank")) return -1;
    if (!d_in || !d_out || !d_rates || !d_phase_io) { set_error("shift_math bank: null pointer"); return -1; }
    int rc = launch_shift_math_bank(reinterpret_cast<const float2*>(d_in), in_stride, reinterpret_cast<float2*>(d_out), out_stride, channels, input_size,
                                    d_rates, d_phase_io, d_scratch, scratch_bytes, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}

int csdrb_shift_table_bank_cc(const complexf* d_in, long in_stride, complexf* d_out, long out_stride, int channels, int input_size,
                              const float* d_rates, float* d_phase_io, const float* d_table, int table_size, void* d_scratch, size_t scratch_bytes, void* stream)
{
    if (too_many_channels(channels, "shift_table bank")) return -1;
    if (!d_in || !d_out || !d_rates || !d_phase_io || !d_table) { set_error("shift_table bank: null pointer"); return -1; }
    int rc = launch_shift_table_bank(reinterpret_cast<const float2*>(d_in), in_stride, reinterpret_cast<float2*>(d_out), out_stride, channels, input_size,
                                     d_rates, d_phase_io, d_table, table_size, d_scratch, scratch_bytes, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}

int csdrb_shift_addfast_bank_cc(const complexf* d_in, long in_stride, complexf* d_out, long out_stride, int channels, int input_size,
                                const shift_addfast_data_t* d_params, float* d_phase_io, int chunk, void* d_scratch, size_t scratch_bytes, void* stream)
{
    if (too_many_channels(channels, "shift_addfast bank")) return -1;
    if (!d_in || !d_out || !d_params || !d_phase_io) { set_error("shift_addfast bank: null pointer"); return -1; }
    int rc = launch_shift_addfast_bank(reinterpret_cast<const float2*>(d_in), in_stride, reinterpret_cast<float2*>(d_out), out_stride, channels, input_size,
                                       reinterpret_cast<const float*>(d_params), d_phase_io, chunk, d_scratch, scratch_bytes, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}

int csdrb_decimating_shift_addition_bank_cc(const complexf* d_in, long in_stride, complexf* d_out, long out_stride, int channels, int input_size,
                                            const shift_addition_data_t* d_params, int decimation, int* d_remain_io, float* d_phase_io, int* d_out_size, void* stream)
{
    if (!d_in || !d_out || !d_params || !d_remain_io || !d_phase_io) { set_error("decimating_shift_addition bank: null pointer"); return -1; }
    int rc = launch_decimating_shift_bank(reinterpret_cast<const float2*>(d_in), in_stride, reinterpret_cast<float2*>(d_out), out_stride, channels, input_size,
                                          reinterpret_cast<const float*>(d_params), decimation, d_remain_io, d_phase_io, d_out_size, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}

size_t csdrb_fractional_decimator_bank_scratch_bytes(int channels, int input_size, float rate) { return fracdec_scratch_bytes(channels, input_size, rate); }

int csdrb_fractional_decimator_bank_ff(const float* d_in, long in_stride, float* d_out, long out_stride, int channels, int input_size, float rate,
                                       int num_poly_points, const float* d_taps, int taps_length, csdrb_fracdec_state_t* d_state,
                                       void* d_scratch, size_t scratch_bytes, void* stream)
{
    if (too_many_channels(channels, "fractional_decimator bank")) return -1;
    if (!d_in || !d_out || !d_state) { set_error("fractional_decimator bank: null pointer"); return -1; }
    int rc = launch_fractional_decimator_bank(d_in, in_stride, d_out, out_stride, channels, input_size, rate, num_poly_points, d_taps, taps_length, d_state,
                                              d_scratch, scratch_bytes, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}

size_t csdrb_fastagc_bank_scratch_bytes(int channels, int nblocks) { return fastagc_scratch_bytes(channels, nblocks); }

int csdrb_fastagc_bank_ff(const float* d_in, long in_stride, float* d_out, long out_stride, int channels, int block, int nblocks, float reference,
                          csdrb_fastagc_state_t* d_state, float* d_hist, void* d_scratch, size_t scratch_bytes, void* stream)
{
    if (too_many_channels(channels, "fastagc bank")) return -1;
    if (!d_in || !d_out || !d_state || !d_hist) { set_error("fastagc bank: null pointer"); return -1; }
    int rc = launch_fastagc_bank(d_in, in_stride, d_out, out_stride, channels, block, nblocks, reference, d_state, d_hist, d_scratch, scratch_bytes, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}

// fastagc_ff | convert_f_s16 fused (the last two blocks of the NFM graph, README.md:87).  Block sizes without a fused kernel (> 1024) run the two steps.
int csdrb_fastagc_bank_f_s16(const float* d_in, long in_stride, short* d_out, long out_stride, int channels, int block, int nblocks, float reference,
                             csdrb_fastagc_state_t* d_state, float* d_hist, void* d_scratch, size_t scratch_bytes, void* stream)
{
    if (too_many_channels(channels, "fastagc bank")) return -1;
    if (!d_in || !d_out || !d_state || !d_hist) { set_error("fastagc s16 bank: null pointer"); return -1; }
    int rc = launch_fastagc_bank_s16(d_in, in_stride, d_out, out_stride, channels, block, nblocks, reference, d_state, d_hist, d_scratch, scratch_bytes, S(stream));
    if (rc != -2) return rc < 0 ? rc : counted(0, rc);
    float* tmp = nullptr;                                              // unusual block size: AGC into a stream-ordered temporary, then the conversion row by row
    const long n = (long)block * nblocks;
    CSDRB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&tmp), sizeof(float) * (size_t)n * channels, S(stream)));
    rc = launch_fastagc_bank(d_in, in_stride, tmp, n, channels, block, nblocks, reference, d_state, d_hist, d_scratch, scratch_bytes, S(stream));
    for (int c = 0; c < channels && rc >= 0; c++) { const int r2 = launch_convert_f_s16(tmp + (long)c * n, d_out + (long)c * out_stride, n, S(stream)); rc = r2 < 0 ? r2 : rc + 1; }
    CSDRB_CUDA(cudaFreeAsync(tmp, S(stream)));
    return rc < 0 ? rc : counted(0, rc);
}

int csdrb_fft_c2c_batch(const complexf* d_in, long in_stride, complexf* d_out, long out_stride, int size, int batch, int inverse, void* stream)
{
    if (!d_in || !d_out) { set_error("fft: null pointer"); return -1; }
    int rc = launch_fft_c2c_batch(reinterpret_cast<const float2*>(d_in), in_stride, reinterpret_cast<float2*>(d_out), out_stride, size, batch, inverse, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}

int csdrb_bandpass_fir_fft_bank_cc(const complexf* d_in, long in_stride, complexf* d_out, long out_stride, int channels, int fft_size, int input_size,
                                   int nblocks, const complexf* d_taps_fft, long taps_stride, complexf* d_tail_io, void* stream)
{
    if (too_many_channels(channels, "bandpass_fir_fft bank")) return -1;
    if (!d_in || !d_out || !d_taps_fft || !d_tail_io) { set_error("bandpass_fir_fft bank: null pointer"); return -1; }
    int rc = launch_olafir_bank(reinterpret_cast<const float2*>(d_in), in_stride, reinterpret_cast<float2*>(d_out), out_stride, channels, fft_size, input_size,
                                nblocks, reinterpret_cast<const float2*>(d_taps_fft), taps_stride, reinterpret_cast<float2*>(d_tail_io), 0, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}

int csdrb_fastddc_fwd_cc(const complexf* d_in, complexf* d_spectra, complexf* d_overlap_io, int fft_size, int input_size, int nblocks, void* stream)
{
    if (!d_in || !d_spectra || !d_overlap_io) { set_error("fastddc_fwd: null pointer"); return -1; }
    int rc = launch_fastddc_fwd(reinterpret_cast<const float2*>(d_in), reinterpret_cast<float2*>(d_spectra), reinterpret_cast<float2*>(d_overlap_io),
                                fft_size, input_size, nblocks, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}

size_t csdrb_fastddc_inv_bank_scratch_bytes(int channels, int nblocks) { return fastddc_inv_scratch_bytes(channels, nblocks); }

int csdrb_fastddc_inv_bank_cc(const complexf* d_spectra, int nblocks, const complexf* d_taps_fft, const csdrb_fastddc_chan_t* d_chan, int channels,
                              const fastddc_t* g, int* d_remain_io, float* d_phase_io, complexf* d_out, long out_stride, int* d_out_total,
                              void* d_scratch, size_t scratch_bytes, void* stream)
{
    if (too_many_channels(channels, "fastddc_inv bank")) return -1;
    if (!d_spectra || !d_taps_fft || !d_chan || !g || !d_remain_io || !d_phase_io || !d_out || !d_out_total) { set_error("fastddc_inv bank: null pointer"); return -1; }
    int rc = launch_fastddc_inv_bank(reinterpret_cast<const float2*>(d_spectra), nblocks, reinterpret_cast<const float2*>(d_taps_fft), d_chan, channels,
                                     g->fft_size, g->fft_inv_size, g->pre_decimation, g->scrap, g->post_input_size, g->post_decimation,
                                     d_remain_io, d_phase_io, reinterpret_cast<float2*>(d_out), out_stride, d_out_total, d_scratch, scratch_bytes, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}

struct csdrb_fastddc_inv_plan { void* impl; };

csdrb_fastddc_inv_plan_t* csdrb_fastddc_inv_plan_create(const csdrb_fastddc_chan_t* chan, int channels, const fastddc_t* g, int nblocks)
{
    if (too_many_channels(channels, "fastddc_inv plan")) return nullptr;
    if (!chan || !g) { set_error("fastddc_inv_plan_create: null pointer"); return nullptr; }
    void* impl = nullptr;
    if (fastddc_inv_plan_create(&impl, chan, channels, nblocks, g->fft_size, g->fft_inv_size, g->pre_decimation, g->scrap, g->post_input_size, g->post_decimation) < 0) return nullptr;
    auto* p = new csdrb_fastddc_inv_plan_t{impl};
    return p;
}

int csdrb_fastddc_inv_plan_run(csdrb_fastddc_inv_plan_t* plan, const complexf* d_spectra, const complexf* d_taps_fft, complexf* d_out, long out_stride,
                               int* d_out_total, void* stream)
{
    if (!plan) { set_error("fastddc_inv_plan_run: null plan"); return -1; }
    int rc = fastddc_inv_plan_run(plan->impl, reinterpret_cast<const float2*>(d_spectra), reinterpret_cast<const float2*>(d_taps_fft), reinterpret_cast<float2*>(d_out),
                                  out_stride, d_out_total, S(stream));
    return rc < 0 ? rc : counted(rc, 4);
}

int csdrb_fastddc_inv_plan_set_channel(csdrb_fastddc_inv_plan_t* plan, int channel, const csdrb_fastddc_chan_t* chan)
{
    if (!plan) { set_error("fastddc_inv_plan_set_channel: null plan"); return -1; }
    return fastddc_inv_plan_set_channel(plan->impl, channel, chan);
}

int csdrb_fastddc_inv_plan_get_state(csdrb_fastddc_inv_plan_t* plan, int* remain, float* phase)
{
    if (!plan) { set_error("fastddc_inv_plan_get_state: null plan"); return -1; }
    return fastddc_inv_plan_get_state(plan->impl, remain, phase);
}

int csdrb_fastddc_inv_plan_set_state(csdrb_fastddc_inv_plan_t* plan, const int* remain, const float* phase)
{
    if (!plan) { set_error("fastddc_inv_plan_set_state: null plan"); return -1; }
    return fastddc_inv_plan_set_state(plan->impl, remain, phase);
}

void csdrb_fastddc_inv_plan_destroy(csdrb_fastddc_inv_plan_t* plan)
{
    if (!plan) return;
    fastddc_inv_plan_destroy(plan->impl);
    delete plan;
}

int csdrb_apply_window_rows_c(const complexf* d_in, complexf* d_out, const float* d_window, int size, long rows, void* stream)
{
    if (!d_in || !d_out || !d_window) { set_error("apply_window: null pointer"); return -1; }
    int rc = launch_apply_window_rows(reinterpret_cast<const float2*>(d_in), reinterpret_cast<float2*>(d_out), d_window, size, rows, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}
int csdrb_logpower_cf(const complexf* d_in, float* d_out, long n, float add_db, void* stream)
{
    if (!d_in || !d_out) { set_error("logpower_cf: null pointer"); return -1; }
    int rc = launch_power(reinterpret_cast<const float2*>(d_in), nullptr, d_out, n, add_db, 0, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}
int csdrb_accumulate_power_cf(const complexf* d_in, float* d_acc, long n, void* stream)
{
    if (!d_in || !d_acc) { set_error("accumulate_power_cf: null pointer"); return -1; }
    int rc = launch_power(reinterpret_cast<const float2*>(d_in), nullptr, d_acc, n, 0.f, 1, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}
int csdrb_log_ff(const float* d_in, float* d_out, long n, float add_db, void* stream)
{
    if (!d_in || !d_out) { set_error("log_ff: null pointer"); return -1; }
    int rc = launch_power(nullptr, d_in, d_out, n, add_db, 2, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}
int csdrb_shift_unroll_bank_cc(const complexf* d_in, long in_stride, complexf* d_out, long out_stride, int channels, int input_size,
                               const shift_addition_data_t* d_params, const float* d_dsin, const float* d_dcos, long table_stride, int table_size,
                               float* d_phase_io, void* d_scratch, size_t scratch_bytes, void* stream)
{
    if (too_many_channels(channels, "shift_unroll bank")) return -1;
    if (!d_in || !d_out || !d_params || !d_dsin || !d_dcos || !d_phase_io) { set_error("shift_unroll bank: null pointer"); return -1; }
    int rc = launch_shift_unroll_bank(reinterpret_cast<const float2*>(d_in), in_stride, reinterpret_cast<float2*>(d_out), out_stride, channels, input_size,
                                      reinterpret_cast<const float*>(d_params), d_dsin, d_dcos, table_stride, table_size, d_phase_io, d_scratch, scratch_bytes, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}

// ---- device memory / streams for CUDA-header-free hosts ------------------------------------------------------------------
void* csdrb_device_alloc(size_t bytes)
{
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
    if (e == cudaSuccess) e = cudaMemset(p, 0, bytes ? bytes : 16);
    if (e != cudaSuccess) { cuda_fail(e, "cudaMalloc/cudaMemset", __FILE__, __LINE__); if (p) cudaFree(p); return nullptr; }
    return p;
}
void csdrb_device_free(void* d_ptr) { if (d_ptr) cudaFree(d_ptr); }
void* csdrb_stream_create(void)
{
    cudaStream_t st = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (e != cudaSuccess) { cuda_fail(e, "cudaStreamCreateWithFlags", __FILE__, __LINE__); return nullptr; }
    return st;
}
void csdrb_stream_destroy(void* stream) { if (stream) cudaStreamDestroy(S(stream)); }
int csdrb_copy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream)
{
    if (!bytes) return 0;
    if (null_io(h_src, d_dst, "copy_h2d")) return -1;
    CSDRB_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, S(stream)));
    return 0;
}
int csdrb_copy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream)
{
    if (!bytes) return 0;
    if (null_io(d_src, h_dst, "copy_d2h")) return -1;
    CSDRB_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, S(stream)));
    return 0;
}
int csdrb_copy_d2d(void* d_dst, const void* d_src, size_t bytes, void* stream)
{
    if (!bytes) return 0;
    if (null_io(d_src, d_dst, "copy_d2d")) return -1;
    CSDRB_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, S(stream)));
    return 0;
}
int csdrb_copy2d_d2d(void* d_dst, size_t dst_pitch_bytes, const void* d_src, size_t src_pitch_bytes, size_t width_bytes, size_t rows, void* stream)
{
    if (!width_bytes || !rows) return 0;
    if (null_io(d_src, d_dst, "copy2d_d2d")) return -1;
    CSDRB_CUDA(cudaMemcpy2DAsync(d_dst, dst_pitch_bytes, d_src, src_pitch_bytes, width_bytes, rows, cudaMemcpyDeviceToDevice, S(stream)));
    return 0;
}
int csdrb_copy2d_h2d(void* d_dst, size_t dst_pitch_bytes, const void* h_src, size_t src_pitch_bytes, size_t width_bytes, size_t rows, void* stream)
{
    if (!width_bytes || !rows) return 0;
    if (null_io(h_src, d_dst, "copy2d_h2d")) return -1;
    CSDRB_CUDA(cudaMemcpy2DAsync(d_dst, dst_pitch_bytes, h_src, src_pitch_bytes, width_bytes, rows, cudaMemcpyHostToDevice, S(stream)));
    return 0;
}
int csdrb_copy2d_d2h(void* h_dst, size_t dst_pitch_bytes, const void* d_src, size_t src_pitch_bytes, size_t width_bytes, size_t rows, void* stream)
{
    if (!width_bytes || !rows) return 0;
    if (null_io(d_src, h_dst, "copy2d_d2h")) return -1;
    CSDRB_CUDA(cudaMemcpy2DAsync(h_dst, dst_pitch_bytes, d_src, src_pitch_bytes, width_bytes, rows, cudaMemcpyDeviceToHost, S(stream)));
    return 0;
}

int csdrb_encode_ima_adpcm_rows_i16_u8(const short* d_in, long in_stride, unsigned char* d_out, long out_stride, int rows, int input_length,
                                       ima_adpcm_state_t* d_state_io, void* stream)
{
    if (!d_in || !d_out || !d_state_io) { set_error("encode_ima_adpcm rows: null pointer"); return -1; }
    int rc = launch_adpcm_encode_rows(d_in, in_stride, d_out, out_stride, rows, input_length, d_state_io, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}

int csdrb_compress_fft_adpcm_rows_f_u8(const float* d_in, long in_stride, unsigned char* d_out, long out_stride, int rows, int fft_size, void* stream)
{
    if (!d_in || !d_out) { set_error("compress_fft_adpcm rows: null pointer"); return -1; }
    int rc = launch_compress_fft_adpcm_rows(d_in, in_stride, d_out, out_stride, rows, fft_size, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}

int csdrb_limit_ff(const float* d_in, float* d_out, long n, float max_amplitude, void* stream)
{
    if (!d_in || !d_out) { set_error("limit_ff: null pointer"); return -1; }
    int rc = launch_limit_ff(d_in, d_out, n, max_amplitude, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}

int csdrb_deemphasis_wfm_bank_ff(const float* d_in, long in_stride, float* d_out, long out_stride, int channels, int input_size, float tau,
                                 int sample_rate, float* d_last_io, void* stream)
{
    if (!d_in || !d_out || !d_last_io) { set_error("deemphasis_wfm bank: null pointer"); return -1; }
    int rc = launch_deemphasis_wfm_bank(d_in, in_stride, d_out, out_stride, channels, input_size, tau, sample_rate, d_last_io, S(stream));
    return rc < 0 ? rc : counted(0, rc);
}

int csdrb_fir_valid_bank_ff(const float* d_in, long in_stride, float* d_out, long out_stride, int channels, int input_size, const float* taps,
                            int taps_length, float limit_max, void* stream)
{
    if (!d_in || !d_out || !taps) { set_error("fir_valid bank: null pointer"); return -1; }
    int rc = launch_fir_valid_bank(d_in, in_stride, d_out, out_stride, channels, input_size, taps, taps_length, limit_max, S(stream));
    if (rc > 0) counted(0, 1);
    return rc;
}

int csdrb_deemphasis_nfm_bank_ff(const float* d_in, long in_stride, float* d_out, long out_stride, int channels, int input_size, int sample_rate,
                                 float limit_max, void* stream)
{
    if (!d_in || !d_out) { set_error("deemphasis_nfm bank: null pointer"); return -1; }
    int rc = launch_deemphasis_nfm_bank(d_in, in_stride, d_out, out_stride, channels, input_size, sample_rate, limit_max, S(stream));
    if (rc > 0) counted(0, 1);
    return rc;
}

size_t csdrb_ddc_bank_scratch_bytes(int channels, int input_size, int chunk, int offset) { return ddc_bank_scratch_bytes(channels, input_size, chunk, offset); }

int csdrb_ddc_bank(const complexf* d_wide, int input_size, int channels, const shift_addition_data_t* d_params, float* d_phase_io, int chunk, int offset,
                   int decimation, const float* h_taps, int taps_length, int demod, void* d_out, long out_stride,
                   const complexf* d_last_in, complexf* d_last_out, void* d_scratch, size_t scratch_bytes, void* stream)
{
    if (!d_wide || !d_params || !d_phase_io || !h_taps || !d_out) { set_error("ddc bank: null pointer"); return -1; }
    int launches = 0;
    int rc = launch_ddc_bank(reinterpret_cast<const float2*>(d_wide), input_size, channels, reinterpret_cast<const float*>(d_params), d_phase_io, chunk, offset,
                             decimation, h_taps, taps_length, demod, d_out, out_stride, reinterpret_cast<const float2*>(d_last_in),
                             reinterpret_cast<float2*>(d_last_out), d_scratch, scratch_bytes, &launches, S(stream));
    if (rc >= 0) g_launches += launches;
    return rc;
}

// ---- streaming DDC/NFM bank object ---------------------------------------------------------------------
// Owns every piece of per-channel state (NCO parameters, chunk-start phases, discriminator history, position inside the current
// chunk) so that a stream is processed block by block with one call per block, and hides the serial float phase chain: while the
// caller's stream runs the main kernel of block k, a private side stream already runs the pre-pass of block k+1 (it depends only on
// the previous pre-pass).  Two scratch/phase buffers alternate; a retune or a block of a different size simply drops the look-ahead.
struct csdrb_ddc_bank_s {
    int channels = 0, decimation = 0, taps_length = 0, demod = 0, chunk = 0;
    std::vector<float> taps;
    std::vector<shift_addition_data_t> h_params;
    float* d_params = nullptr;
    float* d_phase[2] = {nullptr, nullptr};          // phase at the start of the chunk holding the next block's first sample (ping-pong)
    float2* d_last[2] = {nullptr, nullptr};          // previous baseband sample per channel (ping-pong across blocks)
    void* d_scratch[2] = {nullptr, nullptr};
    void* d_tables = nullptr;                         // phase-wrap tables (phase_table.cuh), one per channel; rebuilt when a rate changes
    size_t scratch_cap = 0;
    int cur = 0;                                      // which phase/scratch buffer holds the state for the NEXT process() call
    int last_sel = 0;
    int offset = 0;                                   // position of the next block's first sample inside its chunk
    bool ahead_valid = false; int ahead_input_size = 0; // a pre-pass for the next block (of this size) is already in flight/done
    cudaStream_t side = nullptr;
    cudaEvent_t ev_prepass = nullptr, ev_pre_inline = nullptr, ev_main[2] = {nullptr, nullptr};
    long blocks = 0;
    bool params_dirty = false;
};

csdrb_ddc_bank_t* csdrb_ddc_bank_create(int channels, const float* h_rates, int decimation, const float* h_taps, int taps_length, int demod, int chunk)
{
    if (channels <= 0 || !h_rates || !h_taps || decimation <= 0 || taps_length <= 0) { set_error("ddc bank create: bad argument"); return nullptr; }
    if (!((decimation == 50 && taps_length <= 850) || (decimation == 10 && taps_length <= 200))) {     // the geometries launch_ddc_main has fused kernels for
        set_error("ddc bank create: no fused kernel for decimation %d / %d taps (compiled: d=50 T<=850, d=10 T<=200); run the unfused bank calls", decimation, taps_length);
        return nullptr;
    }
    auto* b = new csdrb_ddc_bank_s();
    b->channels = channels; b->decimation = decimation; b->taps_length = taps_length; b->demod = demod ? 1 : 0; b->chunk = chunk > 0 ? chunk : 1024;
    b->taps.assign(h_taps, h_taps + taps_length);
    b->h_params.resize((size_t)channels);
    for (int c = 0; c < channels; c++) b->h_params[(size_t)c] = shift_addition_init(h_rates[c]);
    bool ok = cudaMalloc(&b->d_params, sizeof(shift_addition_data_t) * (size_t)channels) == cudaSuccess;
    for (int k = 0; k < 2 && ok; k++) {
        ok = ok && cudaMalloc(&b->d_phase[k], sizeof(float) * (size_t)channels) == cudaSuccess;
        ok = ok && cudaMalloc(&b->d_last[k], sizeof(float2) * (size_t)channels) == cudaSuccess;
        if (ok) { cudaMemset(b->d_phase[k], 0, sizeof(float) * (size_t)channels); cudaMemset(b->d_last[k], 0, sizeof(float2) * (size_t)channels); }
    }
    ok = ok && cudaMemcpy(b->d_params, b->h_params.data(), sizeof(shift_addition_data_t) * (size_t)channels, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMalloc(&b->d_tables, ddc_bank_tables_bytes(channels)) == cudaSuccess;
    ok = ok && launch_ddc_tables(channels, b->d_params, b->chunk, b->d_tables, nullptr) >= 0 && cudaStreamSynchronize(nullptr) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&b->side, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&b->ev_prepass, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&b->ev_pre_inline, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; k < 2; k++) ok = ok && cudaEventCreateWithFlags(&b->ev_main[k], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { set_error("ddc bank create: CUDA allocation failed (%s)", cudaGetErrorString(cudaGetLastError())); csdrb_ddc_bank_destroy(b); return nullptr; }
    return b;
}

void csdrb_ddc_bank_destroy(csdrb_ddc_bank_t* b)
{
    if (!b) return;
    if (b->side) { cudaStreamSynchronize(b->side); cudaStreamDestroy(b->side); }
    if (b->ev_prepass) cudaEventDestroy(b->ev_prepass);
    if (b->ev_pre_inline) cudaEventDestroy(b->ev_pre_inline);
    for (int k = 0; k < 2; k++) if (b->ev_main[k]) cudaEventDestroy(b->ev_main[k]);
    cudaFree(b->d_params); cudaFree(b->d_tables);
    for (int k = 0; k < 2; k++) { cudaFree(b->d_phase[k]); cudaFree(b->d_last[k]); cudaFree(b->d_scratch[k]); }
    delete b;
}

int csdrb_ddc_bank_set_rate(csdrb_ddc_bank_t* b, int channel, float rate)
{
    if (!b || channel < 0 || channel >= b->channels) { set_error("ddc bank set_rate: bad channel"); return -1; }
    b->h_params[(size_t)channel] = shift_addition_init(rate);
    b->params_dirty = true;                           // uploaded, and the look-ahead dropped, at the next process()
    return 0;
}

// close the current NCO chunk at the next block's first sample without changing a rate (what a retune does to every channel of the bank): lets
// several banks that share a stream stay chunk-aligned with each other when only one of them retunes (csrc/multi.cu)
int csdrb_ddc_bank_rechunk(csdrb_ddc_bank_t* b)
{
    if (!b) { set_error("ddc bank rechunk: null pointer"); return -1; }
    b->params_dirty = true;
    return 0;
}

int csdrb_ddc_bank_offset(const csdrb_ddc_bank_t* b) { return b ? b->offset : -1; }

int csdrb_ddc_bank_process(csdrb_ddc_bank_t* b, const complexf* d_wide, int input_size, void* d_out, long out_stride, void* stream)
{
    if (!b || !d_wide || !d_out) { set_error("ddc bank process: null pointer"); return -1; }
    cudaStream_t st = S(stream);
    const int n_out = input_size >= b->taps_length ? (input_size - b->taps_length) / b->decimation + 1 : 0;
    if (n_out == 0) return 0;
    const size_t need = csdrb_ddc_bank_scratch_bytes(b->channels, input_size, b->chunk, b->chunk - 1) + 64;
    if (need > b->scratch_cap) {                      // grow both scratch buffers (drops any look-ahead)
        CSDRB_CUDA(cudaStreamSynchronize(b->side));
        CSDRB_CUDA(cudaStreamSynchronize(st));
        for (int k = 0; k < 2; k++) { if (b->d_scratch[k]) CSDRB_CUDA(cudaFree(b->d_scratch[k])); CSDRB_CUDA(cudaMalloc(&b->d_scratch[k], need)); }
        b->scratch_cap = need; b->ahead_valid = false;
    }
    int launches = 0;
    if (b->params_dirty) {                            // retune: a pre-pass made with the old deltas is void
        CSDRB_CUDA(cudaStreamSynchronize(b->side));
        // the look-ahead already advanced phase[cur^1] from phase[cur]; phase[cur] is still the state before it: just forget it
        b->ahead_valid = false;
        // The phase stays continuous across a retune only if the new deltas start from the phase AT the retune sample.  phase[cur] is the
        // phase at the start of the current chunk, `offset` samples back: close that chunk here -- advance every channel by `offset` samples
        // at its OLD rate, exactly what the reference does when a caller hands shift_addition_cc a shorter buffer (libcsdr_gpl.c:48-50) -- and
        // start a fresh chunk at this block's first sample.
        if (b->offset > 0) {
            int rc = launch_ddc_rechunk(b->channels, b->d_params, b->d_phase[b->cur], b->offset, st);
            if (rc < 0) return rc;
            launches += rc;
            b->offset = 0;
        }
        CSDRB_CUDA(cudaMemcpyAsync(b->d_params, b->h_params.data(), sizeof(shift_addition_data_t) * (size_t)b->channels, cudaMemcpyHostToDevice, st));
        { int rc = launch_ddc_tables(b->channels, b->d_params, b->chunk, b->d_tables, st); if (rc < 0) return rc; launches += rc; }
        CSDRB_CUDA(cudaStreamSynchronize(st));        // h_params may change again as soon as we return; the side stream reads the new tables next
        b->params_dirty = false;
    }
    int sel;                                          // scratch buffer holding this block's seeds
    if (b->ahead_valid && b->ahead_input_size == input_size) {
        sel = b->cur ^ 1;                             // the side stream produced seeds[sel] and advanced phase[sel] from phase[cur]
        CSDRB_CUDA(cudaStreamWaitEvent(st, b->ev_prepass, 0));
    } else {
        if (b->ahead_valid) CSDRB_CUDA(cudaStreamSynchronize(b->side));   // a mismatching look-ahead: let it finish, then ignore it
        sel = b->cur ^ 1;
        // phase[sel] <- phase[cur], then run the pre-pass on the caller's stream (it advances phase[sel])
        CSDRB_CUDA(cudaMemcpyAsync(b->d_phase[sel], b->d_phase[b->cur], sizeof(float) * (size_t)b->channels, cudaMemcpyDeviceToDevice, st));
        int rc = launch_ddc_prepass(input_size, b->channels, b->d_params, b->d_phase[sel], b->chunk, b->offset, b->decimation, b->taps_length,
                                    b->d_scratch[sel], b->scratch_cap, b->d_tables, st);
        if (rc < 0) return rc;
        launches += rc;
        CSDRB_CUDA(cudaEventRecord(b->ev_pre_inline, st));             // the look-ahead below starts from the phases this pre-pass produced
        CSDRB_CUDA(cudaStreamWaitEvent(b->side, b->ev_pre_inline, 0));
    }
    b->ahead_valid = false;
    // main kernel of this block: discriminator history ping-pongs last[last_sel] -> last[last_sel^1]
    int rc = launch_ddc_main(reinterpret_cast<const float2*>(d_wide), input_size, b->channels, b->d_params, b->chunk, b->offset, b->decimation, b->taps.data(),
                             b->taps_length, b->demod, d_out, out_stride, b->d_last[b->last_sel], b->d_last[b->last_sel ^ 1], b->d_scratch[sel], st);
    if (rc < 0) return rc;
    launches += 1;
    b->last_sel ^= 1;
    CSDRB_CUDA(cudaEventRecord(b->ev_main[b->blocks & 1], st));
    // state for the next block now lives in phase[sel]; its first sample sits `offset` samples into that chunk
    b->cur = sel;
    b->offset = (int)(((long)b->offset + (long)n_out * b->decimation) % b->chunk);
    // look-ahead: pre-pass of the next block (assumed to have the same size) on the side stream into the other buffer, CONCURRENT
    // with the main kernel just launched.  It only has to wait for (a) this block's pre-pass (ordered above / same stream) and
    // (b) the main kernel of the PREVIOUS block, the last reader of the scratch buffer it is about to overwrite.
    {
        const int nxt = b->cur ^ 1;
        if (b->blocks > 0) CSDRB_CUDA(cudaStreamWaitEvent(b->side, b->ev_main[(b->blocks - 1) & 1], 0));
        CSDRB_CUDA(cudaMemcpyAsync(b->d_phase[nxt], b->d_phase[b->cur], sizeof(float) * (size_t)b->channels, cudaMemcpyDeviceToDevice, b->side));
        int rp = launch_ddc_prepass(input_size, b->channels, b->d_params, b->d_phase[nxt], b->chunk, b->offset, b->decimation, b->taps_length,
                                    b->d_scratch[nxt], b->scratch_cap, b->d_tables, b->side);
        if (rp < 0) return rp;
        launches += rp;
        CSDRB_CUDA(cudaEventRecord(b->ev_prepass, b->side));
        b->ahead_valid = true; b->ahead_input_size = input_size;
    }
    b->blocks++;
    g_launches += launches;
    return n_out;
}

// =====================================================================================================
// Part A, continued: host-pointer drop-ins for shift / fractional decimator / fastagc / FFT / fastddc
// =====================================================================================================
#define A_BEGIN(who) std::lock_guard<std::mutex> lk(g_ctx.mu); A_CHECK(g_ctx.init(), who)
#define A_UP(slot, ptr, bytes, who) do { A_CHECK(g_ctx.reserve(slot, (bytes) + 16), who); \
    A_CUDA(cudaMemcpyAsync(g_ctx.buf[slot], ptr, bytes, cudaMemcpyHostToDevice, g_ctx.stream), who); } while (0)
#define A_DOWN(ptr, slot, bytes, who) A_CUDA(cudaMemcpyAsync(ptr, g_ctx.buf[slot], bytes, cudaMemcpyDeviceToHost, g_ctx.stream), who)
#define A_SYNC(who) A_CUDA(cudaStreamSynchronize(g_ctx.stream), who)

float shift_addition_cc(complexf* input, complexf* output, int input_size, shift_addition_data_t d, float starting_phase)
{
    const char* who = "shift_addition_cc";
    if (input_size <= 0) return starting_phase;      // the reference still wraps the phase; with n = 0 nothing changes unless |phase| > pi
    A_BEGIN(who);
    // slot 0: input, 1: output, 2: params(12 B) + phase(4 B) at +64, 3: scratch
    A_UP(0, input, (size_t)input_size * 8, who);
    A_CHECK(g_ctx.reserve(1, (size_t)input_size * 8 + 16), who);
    struct { shift_addition_data_t p; float pad; float phase; } blob = {d, 0.f, starting_phase};
    A_UP(2, &blob, sizeof blob, who);
    const size_t sb = csdrb_shift_addition_bank_scratch_bytes(1, input_size, input_size);
    A_CHECK(g_ctx.reserve(3, sb + 16), who);
    float* d_phase = reinterpret_cast<float*>(static_cast<char*>(g_ctx.buf[2]) + offsetof(decltype(blob), phase));
    A_CHECK(csdrb_shift_addition_bank_cc((const complexf*)g_ctx.buf[0], 0, (complexf*)g_ctx.buf[1], 0, 1, input_size,
                                         (const shift_addition_data_t*)g_ctx.buf[2], d_phase, input_size, g_ctx.buf[3], g_ctx.cap[3], g_ctx.stream), who);
    A_DOWN(output, 1, (size_t)input_size * 8, who);
    float new_phase = 0.f;
    A_CUDA(cudaMemcpyAsync(&new_phase, d_phase, 4, cudaMemcpyDeviceToHost, g_ctx.stream), who);
    A_SYNC(who);
    return new_phase;
}

float shift_table_cc(complexf* input, complexf* output, int input_size, float rate, shift_table_data_t table_data, float starting_phase)
{
    const char* who = "shift_table_cc";
    if (input_size <= 0 || !table_data.table || table_data.table_size < 2) return starting_phase;
    A_BEGIN(who);
    // slot 0: input, 1: output, 2: rate at +0 and phase at +64, 3: scratch.  The table has its own device buffer and is sent with every call
    // (256 KB for the default size: a host pointer is no proof that the contents are the ones sent last time).
    static float* d_table = nullptr; static int d_table_cap = 0;
    A_UP(0, input, (size_t)input_size * 8, who);
    A_CHECK(g_ctx.reserve(1, (size_t)input_size * 8 + 16), who);
    float blob[17] = {0};
    blob[0] = rate; blob[16] = starting_phase;
    A_UP(2, blob, sizeof blob, who);
    float* d_rate = reinterpret_cast<float*>(g_ctx.buf[2]);
    float* d_phase = d_rate + 16;
    if (table_data.table_size > d_table_cap) {
        if (d_table) A_CUDA(cudaFree(d_table), who);
        d_table = nullptr; d_table_cap = 0;
        A_CUDA(cudaMalloc(&d_table, (size_t)table_data.table_size * 4), who);
        d_table_cap = table_data.table_size;
    }
    A_CUDA(cudaMemcpyAsync(d_table, table_data.table, (size_t)table_data.table_size * 4, cudaMemcpyHostToDevice, g_ctx.stream), who);
    const size_t sb = csdrb_shift_math_bank_scratch_bytes(1, input_size);
    A_CHECK(g_ctx.reserve(3, sb + 16), who);
    A_CHECK(csdrb_shift_table_bank_cc((const complexf*)g_ctx.buf[0], 0, (complexf*)g_ctx.buf[1], 0, 1, input_size, d_rate, d_phase, d_table, table_data.table_size,
                                      g_ctx.buf[3], g_ctx.cap[3], g_ctx.stream), who);
    A_DOWN(output, 1, (size_t)input_size * 8, who);
    float new_phase = 0.f;
    A_CUDA(cudaMemcpyAsync(&new_phase, d_phase, 4, cudaMemcpyDeviceToHost, g_ctx.stream), who);
    A_SYNC(who);
    return new_phase;
}

float shift_math_cc(complexf* input, complexf* output, int input_size, float rate, float starting_phase)
{
    const char* who = "shift_math_cc";
    if (input_size <= 0) return starting_phase;
    A_BEGIN(who);
    // slot 0: input, 1: output, 2: rate at +0 and phase at +64, 3: scratch
    A_UP(0, input, (size_t)input_size * 8, who);
    A_CHECK(g_ctx.reserve(1, (size_t)input_size * 8 + 16), who);
    float blob[17] = {0};
    blob[0] = rate; blob[16] = starting_phase;
    A_UP(2, blob, sizeof blob, who);
    const size_t sb = csdrb_shift_math_bank_scratch_bytes(1, input_size);
    A_CHECK(g_ctx.reserve(3, sb + 16), who);
    float* d_rate = reinterpret_cast<float*>(g_ctx.buf[2]);
    float* d_phase = d_rate + 16;
    A_CHECK(csdrb_shift_math_bank_cc((const complexf*)g_ctx.buf[0], 0, (complexf*)g_ctx.buf[1], 0, 1, input_size, d_rate, d_phase, g_ctx.buf[3], g_ctx.cap[3], g_ctx.stream), who);
    A_DOWN(output, 1, (size_t)input_size * 8, who);
    float new_phase = 0.f;
    A_CUDA(cudaMemcpyAsync(&new_phase, d_phase, 4, cudaMemcpyDeviceToHost, g_ctx.stream), who);
    A_SYNC(who);
    return new_phase;
}

float shift_addfast_cc(complexf* input, complexf* output, int input_size, shift_addfast_data_t* d, float starting_phase)
{
    const char* who = "shift_addfast_cc";
    if (input_size <= 0 || !d) return starting_phase;
    A_BEGIN(who);
    // slot 0: input, 1: output, 2: params (36 B) + phase at +64, 3: scratch
    A_UP(0, input, (size_t)input_size * 8, who);
    A_CHECK(g_ctx.reserve(1, (size_t)input_size * 8 + 16), who);
    struct { shift_addfast_data_t p; float pad[7]; float phase; } blob;
    blob.p = *d; blob.phase = starting_phase;
    static_assert(offsetof(decltype(blob), phase) == 64, "phase sits at +64");
    A_UP(2, &blob, sizeof blob, who);
    const size_t sb = csdrb_shift_addition_bank_scratch_bytes(1, input_size, input_size);
    A_CHECK(g_ctx.reserve(3, sb + 16), who);
    float* d_phase = reinterpret_cast<float*>(static_cast<char*>(g_ctx.buf[2]) + 64);
    A_CHECK(csdrb_shift_addfast_bank_cc((const complexf*)g_ctx.buf[0], 0, (complexf*)g_ctx.buf[1], 0, 1, input_size,
                                        (const shift_addfast_data_t*)g_ctx.buf[2], d_phase, input_size, g_ctx.buf[3], g_ctx.cap[3], g_ctx.stream), who);
    const int whole = input_size & ~3;                                  // the n%4 tail of `output` is left alone, like the reference
    if (whole > 0) A_DOWN(output, 1, (size_t)whole * 8, who);
    float new_phase = 0.f;
    A_CUDA(cudaMemcpyAsync(&new_phase, d_phase, 4, cudaMemcpyDeviceToHost, g_ctx.stream), who);
    A_SYNC(who);
    return new_phase;
}

decimating_shift_addition_status_t decimating_shift_addition_cc(complexf* input, complexf* output, int input_size, shift_addition_data_t d,
                                                                int decimation, decimating_shift_addition_status_t s)
{
    const char* who = "decimating_shift_addition_cc";
    A_BEGIN(who);
    if (input_size > 0) A_UP(0, input, (size_t)input_size * 8, who); else A_CHECK(g_ctx.reserve(0, 64), who);
    const int cap = input_size / (decimation > 0 ? decimation : 1) + 2;
    A_CHECK(g_ctx.reserve(1, (size_t)cap * 8 + 16), who);
    struct { shift_addition_data_t p; int remain; float phase; int outsz; } blob = {d, s.decimation_remain, s.starting_phase, 0};
    A_UP(2, &blob, sizeof blob, who);
    char* b2 = static_cast<char*>(g_ctx.buf[2]);
    A_CHECK(csdrb_decimating_shift_addition_bank_cc((const complexf*)g_ctx.buf[0], 0, (complexf*)g_ctx.buf[1], 0, 1, input_size,
                                                    (const shift_addition_data_t*)b2, decimation, (int*)(b2 + offsetof(decltype(blob), remain)),
                                                    (float*)(b2 + offsetof(decltype(blob), phase)), (int*)(b2 + offsetof(decltype(blob), outsz)), g_ctx.stream), who);
    A_CUDA(cudaMemcpyAsync(&blob, b2, sizeof blob, cudaMemcpyDeviceToHost, g_ctx.stream), who);
    A_SYNC(who);
    if (blob.outsz > 0) { A_DOWN(output, 1, (size_t)blob.outsz * 8, who); A_SYNC(who); }
    s.decimation_remain = blob.remain; s.starting_phase = blob.phase; s.output_size = blob.outsz;
    return s;
}

fractional_decimator_ff_t fractional_decimator_ff_init(float rate, int num_poly_points, float* taps, int taps_length)
{
    // libcsdr.c:715-748 -- same field values; the three scratch arrays are kept so the struct stays layout- and
    // ownership-compatible with callers that free them.
    fractional_decimator_ff_t d;
    d.num_poly_points = num_poly_points & ~1;
    d.poly_precalc_denomiator = (float*)malloc(sizeof(float) * (size_t)(d.num_poly_points > 0 ? d.num_poly_points : 1));
    d.xifirst = -(num_poly_points / 2) + 1;
    d.xilast = num_poly_points / 2;
    int slot = 0;
    for (int xi = d.xifirst; xi <= d.xilast && slot < d.num_poly_points; xi++, slot++) {
        float prod = 1;
        for (int xj = d.xifirst; xj <= d.xilast; xj++) if (xi != xj) prod *= (float)(xi - xj);
        d.poly_precalc_denomiator[slot] = prod;
    }
    d.where = (float)(-d.xifirst);
    d.coeffs_buf = (float*)malloc(sizeof(float) * (size_t)(d.num_poly_points > 0 ? d.num_poly_points : 1));
    d.filtered_buf = (float*)malloc(sizeof(float) * (size_t)(d.num_poly_points > 0 ? d.num_poly_points : 1));
    d.rate = rate; d.taps = taps; d.taps_length = taps_length; d.input_processed = 0; d.output_size = 0;
    return d;
}

void fractional_decimator_ff(float* input, float* output, int input_size, fractional_decimator_ff_t* d)
{
    const char* who = "fractional_decimator_ff";
    if (input_size <= 0) { d->output_size = 0; return; }
    A_BEGIN(who);
    A_UP(0, input, (size_t)input_size * 4, who);
    const int cap = (int)((double)input_size / (d->rate > 1.f ? d->rate : 1.0)) + 8;
    A_CHECK(g_ctx.reserve(1, (size_t)cap * 4 + 16), who);
    const int tl = d->taps ? d->taps_length : 0;
    struct Blob { csdrb_fracdec_state_t st; int pad; } blob = {{d->where, 0, 0}, 0};
    // slot 2: [state | taps]
    const size_t tap_off = 64;
    A_CHECK(g_ctx.reserve(2, tap_off + (size_t)tl * 4 + 16), who);
    A_CUDA(cudaMemcpyAsync(g_ctx.buf[2], &blob, sizeof blob, cudaMemcpyHostToDevice, g_ctx.stream), who);
    float* d_taps = nullptr;
    if (tl > 0) {
        d_taps = reinterpret_cast<float*>(static_cast<char*>(g_ctx.buf[2]) + tap_off);
        A_CUDA(cudaMemcpyAsync(d_taps, d->taps, (size_t)tl * 4, cudaMemcpyHostToDevice, g_ctx.stream), who);
    }
    const size_t sb = csdrb_fractional_decimator_bank_scratch_bytes(1, input_size, d->rate);
    A_CHECK(g_ctx.reserve(3, sb + 16), who);
    A_CHECK(csdrb_fractional_decimator_bank_ff((const float*)g_ctx.buf[0], 0, (float*)g_ctx.buf[1], 0, 1, input_size, d->rate, d->num_poly_points,
                                               d_taps, tl, (csdrb_fracdec_state_t*)g_ctx.buf[2], g_ctx.buf[3], g_ctx.cap[3], g_ctx.stream), who);
    A_CUDA(cudaMemcpyAsync(&blob, g_ctx.buf[2], sizeof blob, cudaMemcpyDeviceToHost, g_ctx.stream), who);
    A_SYNC(who);
    if (blob.st.output_size > 0) { A_DOWN(output, 1, (size_t)blob.st.output_size * 4, who); A_SYNC(who); }
    d->where = blob.st.where; d->input_processed = blob.st.input_processed; d->output_size = blob.st.output_size;
}

void fastagc_ff(fastagc_ff_t* a, float* output)
{
    const char* who = "fastagc_ff";
    const int n = a->input_size;
    if (n <= 0) return;
    A_BEGIN(who);
    A_UP(0, a->buffer_input, (size_t)n * 4, who);
    A_CHECK(g_ctx.reserve(1, (size_t)n * 4 + 16), who);
    A_CHECK(g_ctx.reserve(2, (size_t)n * 8 + 64 + 16), who);            // [state 64 B | hist1 | hist2]
    csdrb_fastagc_state_t st = {a->peak_1, a->peak_2, a->last_gain};
    char* b2 = static_cast<char*>(g_ctx.buf[2]);
    A_CUDA(cudaMemcpyAsync(b2, &st, sizeof st, cudaMemcpyHostToDevice, g_ctx.stream), who);
    A_CUDA(cudaMemcpyAsync(b2 + 64, a->buffer_1, (size_t)n * 4, cudaMemcpyHostToDevice, g_ctx.stream), who);
    A_CUDA(cudaMemcpyAsync(b2 + 64 + (size_t)n * 4, a->buffer_2, (size_t)n * 4, cudaMemcpyHostToDevice, g_ctx.stream), who);
    A_CHECK(g_ctx.reserve(3, 256), who);
    A_CHECK(csdrb_fastagc_bank_ff((const float*)g_ctx.buf[0], 0, (float*)g_ctx.buf[1], 0, 1, n, 1, a->reference, (csdrb_fastagc_state_t*)b2,
                                  (float*)(b2 + 64), g_ctx.buf[3], g_ctx.cap[3], g_ctx.stream), who);
    A_DOWN(output, 1, (size_t)n * 4, who);
    A_CUDA(cudaMemcpyAsync(&st, b2, sizeof st, cudaMemcpyDeviceToHost, g_ctx.stream), who);
    A_SYNC(who);
    // rotate the three caller-owned buffers exactly like libcsdr.c:981-989
    float* recycled = a->buffer_1;
    a->buffer_1 = a->buffer_2; a->buffer_2 = a->buffer_input; a->buffer_input = recycled;
    a->peak_1 = st.peak_1; a->peak_2 = st.peak_2; a->last_gain = st.last_gain;
}

void apply_precalculated_window_c(complexf* input, complexf* output, int size, float* windowt)
{
    const char* who = "apply_precalculated_window_c";
    if (size <= 0) return;
    A_BEGIN(who);
    A_UP(0, input, (size_t)size * 8, who);
    A_CHECK(g_ctx.reserve(1, (size_t)size * 8 + 16), who);
    A_UP(2, windowt, (size_t)size * 4, who);
    A_CHECK(csdrb_apply_window_rows_c((const complexf*)g_ctx.buf[0], (complexf*)g_ctx.buf[1], (const float*)g_ctx.buf[2], size, 1, g_ctx.stream), who);
    A_DOWN(output, 1, (size_t)size * 8, who);
    A_SYNC(who);
}

void apply_window_c(complexf* input, complexf* output, int size, window_t window)
{
    float* table = precalculate_window(size, window);                    // the table itself is one-off host work, like every filter design step
    apply_precalculated_window_c(input, output, size, table);
    free(table);
}

static void power_dropin(const char* who, const void* in, size_t in_bytes, float* out, int n, float add_db, int mode)
{
    if (n <= 0) return;
    A_BEGIN(who);
    A_UP(0, in, in_bytes, who);
    if (mode == 1) A_UP(1, out, (size_t)n * 4, who); else A_CHECK(g_ctx.reserve(1, (size_t)n * 4 + 16), who);
    int rc = launch_power(mode == 2 ? nullptr : (const float2*)g_ctx.buf[0], mode == 2 ? (const float*)g_ctx.buf[0] : nullptr, (float*)g_ctx.buf[1], n, add_db, mode, g_ctx.stream);
    A_CHECK(rc, who); counted(0, 1);
    A_DOWN(out, 1, (size_t)n * 4, who);
    A_SYNC(who);
}
void logpower_cf(complexf* input, float* output, int size, float add_db) { power_dropin("logpower_cf", input, (size_t)(size > 0 ? size : 0) * 8, output, size, add_db, 0); }
void accumulate_power_cf(complexf* input, float* output, int size) { power_dropin("accumulate_power_cf", input, (size_t)(size > 0 ? size : 0) * 8, output, size, 0.f, 1); }
void log_ff(float* input, float* output, int size, float add_db) { power_dropin("log_ff", input, (size_t)(size > 0 ? size : 0) * 4, output, size, add_db, 2); }

float shift_unroll_cc(complexf* input, complexf* output, int input_size, shift_unroll_data_t* d, float starting_phase)
{
    const char* who = "shift_unroll_cc";
    if (input_size <= 0) return starting_phase;
    if (!d || input_size > d->size) { set_error("input_size %d exceeds the table size %d", input_size, d ? d->size : 0); die(who); }
    A_BEGIN(who);
    A_UP(0, input, (size_t)input_size * 8, who);
    A_CHECK(g_ctx.reserve(1, (size_t)input_size * 8 + 16), who);
    // slot 2: [starting phase at +16 | dsin at +64 | dcos after it]
    const size_t tb = (size_t)d->size * 4;
    A_CHECK(g_ctx.reserve(2, 64 + 2 * tb + 16), who);
    char* b2 = static_cast<char*>(g_ctx.buf[2]);
    // one call = one chunk: the phase carried to the next call is one float multiply-add and a wrap -- done right here on the host
    float new_phase = starting_phase + input_size * d->phase_increment;
    while (new_phase > 3.14159265358979323846f) new_phase -= 2 * 3.14159265358979323846f;
    while (new_phase < -3.14159265358979323846f) new_phase += 2 * 3.14159265358979323846f;
    A_CUDA(cudaMemcpyAsync(b2 + 16, &starting_phase, 4, cudaMemcpyHostToDevice, g_ctx.stream), who);
    A_CUDA(cudaMemcpyAsync(b2 + 64, d->dsin, tb, cudaMemcpyHostToDevice, g_ctx.stream), who);
    A_CUDA(cudaMemcpyAsync(b2 + 64 + tb, d->dcos, tb, cudaMemcpyHostToDevice, g_ctx.stream), who);
    // single call = single chunk: chunk_phase[0] is the starting phase itself
    shift_unroll_bank_single(reinterpret_cast<const float2*>(g_ctx.buf[0]), reinterpret_cast<float2*>(g_ctx.buf[1]), input_size,
                             reinterpret_cast<const float*>(b2 + 64), reinterpret_cast<const float*>(b2 + 64 + tb), reinterpret_cast<const float*>(b2 + 16), g_ctx.stream);
    counted(0, 1);
    A_DOWN(output, 1, (size_t)input_size * 8, who);
    A_SYNC(who);
    return new_phase;
}

ima_adpcm_state_t encode_ima_adpcm_i16_u8(short* input, unsigned char* output, int input_length, ima_adpcm_state_t state)
{
    const char* who = "encode_ima_adpcm_i16_u8";
    if (input_length < 2) return state;
    A_BEGIN(who);
    A_UP(0, input, (size_t)input_length * 2, who);
    A_CHECK(g_ctx.reserve(1, (size_t)input_length / 2 + 16), who);
    A_UP(2, &state, sizeof state, who);
    A_CHECK(csdrb_encode_ima_adpcm_rows_i16_u8((const short*)g_ctx.buf[0], input_length, (unsigned char*)g_ctx.buf[1], input_length / 2, 1, input_length,
                                               (ima_adpcm_state_t*)g_ctx.buf[2], g_ctx.stream), who);
    A_DOWN(output, 1, (size_t)(input_length / 2), who);
    A_CUDA(cudaMemcpyAsync(&state, g_ctx.buf[2], sizeof state, cudaMemcpyDeviceToHost, g_ctx.stream), who);
    A_SYNC(who);
    return state;
}

void limit_ff(float* input, float* output, int input_size, float max_amplitude)
{
    const char* who = "limit_ff";
    if (input_size <= 0) return;
    A_BEGIN(who);
    A_UP(0, input, (size_t)input_size * 4, who);
    A_CHECK(g_ctx.reserve(1, (size_t)input_size * 4 + 16), who);
    A_CHECK(csdrb_limit_ff((const float*)g_ctx.buf[0], (float*)g_ctx.buf[1], input_size, max_amplitude, g_ctx.stream), who);
    A_DOWN(output, 1, (size_t)input_size * 4, who);
    A_SYNC(who);
}

float deemphasis_wfm_ff(float* input, float* output, int input_size, float tau, int sample_rate, float last_output)
{
    const char* who = "deemphasis_wfm_ff";
    if (input_size <= 0) return last_output;
    A_BEGIN(who);
    A_UP(0, input, (size_t)input_size * 4, who);
    A_CHECK(g_ctx.reserve(1, (size_t)input_size * 4 + 16), who);
    A_UP(2, &last_output, 4, who);
    A_CHECK(csdrb_deemphasis_wfm_bank_ff((const float*)g_ctx.buf[0], input_size, (float*)g_ctx.buf[1], input_size, 1, input_size, tau, sample_rate,
                                         (float*)g_ctx.buf[2], g_ctx.stream), who);
    A_DOWN(output, 1, (size_t)input_size * 4, who);
    A_SYNC(who);
    return output[input_size - 1];
}

int deemphasis_nfm_ff(float* input, float* output, int input_size, int sample_rate)
{
    const char* who = "deemphasis_nfm_ff";
    int taps_length = 0;
    if (!csdrb_deemphasis_nfm_taps(sample_rate, &taps_length)) return 0;          // libcsdr.c:1119: no table for this rate
    if (input_size - taps_length <= 0) return 0;
    A_BEGIN(who);
    A_UP(0, input, (size_t)input_size * 4, who);
    A_CHECK(g_ctx.reserve(1, (size_t)input_size * 4 + 16), who);
    int produced = csdrb_deemphasis_nfm_bank_ff((const float*)g_ctx.buf[0], input_size, (float*)g_ctx.buf[1], input_size, 1, input_size, sample_rate, 0.f, g_ctx.stream);
    A_CHECK(produced, who);
    A_DOWN(output, 1, (size_t)produced * 4, who);
    A_SYNC(who);
    return produced;
}

// ---- FFT abstraction ---------------------------------------------------------------------------------
struct csdrb_plan_impl { unsigned magic; int forward; };
static const unsigned kPlanMagic = 0xC5D2B200u;

FFT_PLAN_T* make_fft_c2c(int size, complexf* input, complexf* output, int forward, int benchmark)
{
    (void)benchmark;
    if (size < 2 || size > 16384 || (size & (size - 1))) {
        fprintf(stderr, "libcsdr_b200: make_fft_c2c: size %d unsupported (power of two, 2..16384)\n", size);
        return nullptr;
    }
    FFT_PLAN_T* p = (FFT_PLAN_T*)malloc(sizeof(FFT_PLAN_T));
    csdrb_plan_impl* impl = (csdrb_plan_impl*)malloc(sizeof(csdrb_plan_impl));
    impl->magic = kPlanMagic; impl->forward = forward ? 1 : 0;
    p->size = size; p->input = input; p->output = output; p->plan = impl;
    return p;
}

void fft_execute(FFT_PLAN_T* plan)
{
    const char* who = "fft_execute";
    if (!plan) return;
    if (!plan->plan || ((csdrb_plan_impl*)plan->plan)->magic != kPlanMagic) {
        set_error("plan was not created by libcsdr_b200's make_fft_c2c (r2c/c2r plans are outside the hot path)"); die(who);
    }
    A_BEGIN(who);
    const size_t bytes = (size_t)plan->size * 8;
    A_UP(0, plan->input, bytes, who);
    A_CHECK(g_ctx.reserve(1, bytes + 16), who);
    A_CHECK(csdrb_fft_c2c_batch((const complexf*)g_ctx.buf[0], plan->size, (complexf*)g_ctx.buf[1], plan->size, plan->size, 1,
                                ((csdrb_plan_impl*)plan->plan)->forward ? 0 : 1, g_ctx.stream), who);
    A_DOWN(plan->output, 1, bytes, who);
    A_SYNC(who);
}

void fft_destroy(FFT_PLAN_T* plan) { if (plan) { free(plan->plan); free(plan); } }
void* csdrb_fft_malloc(size_t bytes) { void* p = nullptr; return posix_memalign(&p, 64, bytes ? bytes : 64) ? nullptr : p; }
void csdrb_fft_free(void* p) { free(p); }

void apply_fir_fft_cc(FFT_PLAN_T* plan, FFT_PLAN_T* plan_inverse, complexf* taps_fft, complexf* last_overlap, int overlap_size)
{
    // libcsdr.c:814-849 in one fused kernel: the intermediate spectrum (plan->output) and product (plan_inverse->input)
    // never leave the GPU, so those two caller buffers are NOT written (no caller in the reference reads them).
    const char* who = "apply_fir_fft_cc";
    A_BEGIN(who);
    const int n = plan->size;
    const size_t bytes = (size_t)n * 8;
    A_UP(0, plan->input, bytes, who);
    A_CHECK(g_ctx.reserve(1, bytes + 16), who);
    A_UP(2, taps_fft, bytes, who);
    if (overlap_size > 0) A_UP(3, last_overlap, (size_t)overlap_size * 8, who); else A_CHECK(g_ctx.reserve(3, 64), who);
    int rc = launch_apply_fir_fft((const float2*)g_ctx.buf[0], (const float2*)g_ctx.buf[2], (const float2*)g_ctx.buf[3], overlap_size, (float2*)g_ctx.buf[1], n, g_ctx.stream);
    A_CHECK(rc, who); counted(0, 1);
    A_DOWN(plan_inverse->output, 1, bytes, who);
    A_SYNC(who);
}

decimating_shift_addition_status_t fastddc_inv_cc(complexf* input, complexf* output, fastddc_t* ddc, FFT_PLAN_T* plan_inverse, complexf* taps_fft,
                                                  decimating_shift_addition_status_t shift_stat)
{
    const char* who = "fastddc_inv_cc";
    (void)plan_inverse;
    {
        A_BEGIN(who);
        const size_t nb = (size_t)ddc->fft_size * 8;
        A_UP(0, input, nb, who);
        A_UP(2, taps_fft, nb, who);
        A_CHECK(g_ctx.reserve(1, (size_t)ddc->post_input_size * 8 + 64), who);
        struct { csdrb_fastddc_chan_t ch; int remain; float phase; int total; } blob =
            {{ddc->offsetbin, ddc->dsadata.sindelta, ddc->dsadata.cosdelta, ddc->dsadata.rate}, shift_stat.decimation_remain, shift_stat.starting_phase, 0};
        const size_t sb = csdrb_fastddc_inv_bank_scratch_bytes(1, 1);
        A_CHECK(g_ctx.reserve(3, 256 + sb), who);
        char* b3 = static_cast<char*>(g_ctx.buf[3]);
        A_CUDA(cudaMemcpyAsync(b3, &blob, sizeof blob, cudaMemcpyHostToDevice, g_ctx.stream), who);
        A_CHECK(csdrb_fastddc_inv_bank_cc((const complexf*)g_ctx.buf[0], 1, (const complexf*)g_ctx.buf[2], (const csdrb_fastddc_chan_t*)b3, 1, ddc,
                                          (int*)(b3 + offsetof(decltype(blob), remain)), (float*)(b3 + offsetof(decltype(blob), phase)),
                                          (complexf*)g_ctx.buf[1], ddc->post_input_size, (int*)(b3 + offsetof(decltype(blob), total)),
                                          b3 + 256, sb, g_ctx.stream), who);
        A_CUDA(cudaMemcpyAsync(&blob, b3, sizeof blob, cudaMemcpyDeviceToHost, g_ctx.stream), who);
        A_SYNC(who);
        if (blob.total > 0) { A_DOWN(output, 1, (size_t)blob.total * 8, who); A_SYNC(who); }
        shift_stat.decimation_remain = blob.remain; shift_stat.starting_phase = blob.phase; shift_stat.output_size = blob.total;
    }
    fft_swap_sides(input, ddc->fft_size);            // the reference leaves its input swapped in place (fastddc.c:123)
    return shift_stat;
}

}  // extern "C"
