// multi.cu -- one process, several GPUs: the shared-input DDC / NFM bank (BASELINE config 4) sliced over the devices of a node.
//
// What the reference does with nmux (one TCP fan-out of the wideband stream, nmux.cpp:246-353) plus one process chain per channel
// (ddcd_old.h:51-57), done in C behind the C ABI: the wideband block goes host -> GPU 0 once, travels to the other GPUs by ncclBroadcast
// over NVLink / NVSwitch, and every GPU runs its contiguous channel slice of the fused bank (csdrb_ddc_bank_*).  Channels are independent,
// so the broadcast is the only exchange step; outputs leave each GPU on its own PCIe link.
//
// Pipelining: submit(block k) only ENQUEUES -- H2D on the root's copy stream, the broadcast on one communication stream per device, the
// slice kernels and the D2H of the results on one compute stream per device -- and returns a ticket; collect(ticket) waits for that block's
// results.  Two wideband buffers per device alternate, so a caller that submits block k+1 before collecting block k has the broadcast of
// k+1 under the kernels of k.  The host buffers of a submitted block (input and output) belong to the library until its collect returns.
//
// NCCL is loaded with dlopen("libnccl.so.2") on first use: the library itself keeps no link-time dependency on it, and a single-device
// "multi" bank never touches it.
#include "common.cuh"
#include "kernels.h"
#include "csdr_b200.h"

#include <dlfcn.h>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace csdrb {

// the handful of NCCL entry points used here, with the types of nccl.h (2.x ABI: opaque comm pointer, int enums)
typedef struct ncclComm* ncclComm_t;
struct NcclApi {
    void* lib = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int /*ncclDataType_t*/, int /*root*/, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load()
    {
        if (lib) return true;
        const char* over = getenv("CSDRB_NCCL_LIB");                     // an explicit library path (tests substitute a memcpy stand-in on GPU-less hosts)
        lib = dlopen(over && *over ? over : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib && !(over && *over)) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) { set_error("multi bank: cannot load libnccl.so.2 (%s)", dlerror()); return false; }
#define CSDRB_NCCL_SYM(field, name) *reinterpret_cast<void**>(&field) = dlsym(lib, name); if (!field) { set_error("multi bank: %s missing in libnccl", name); return false; }
        CSDRB_NCCL_SYM(CommInitAll, "ncclCommInitAll") CSDRB_NCCL_SYM(CommDestroy, "ncclCommDestroy") CSDRB_NCCL_SYM(GroupStart, "ncclGroupStart")
        CSDRB_NCCL_SYM(GroupEnd, "ncclGroupEnd") CSDRB_NCCL_SYM(Broadcast, "ncclBroadcast") CSDRB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef CSDRB_NCCL_SYM
        return true;
    }
};
constexpr int kNcclFloat32 = 7;                                          // ncclFloat32, nccl.h

}  // namespace csdrb

using namespace csdrb;

struct csdrb_multi_bank_s {
    int ndev = 0, channels = 0, decimation = 0, taps_length = 0, demod = 0, max_block = 0;
    long dstride = 0;                                                    // device row stride of the outputs (elements)
    std::vector<int> dev, ch0, nch;
    std::vector<ncclComm_t> comm;
    std::vector<cudaStream_t> st, cst;                                   // compute / communication stream per device
    std::vector<csdrb_ddc_bank_t*> bank;
    std::vector<float2*> wide[2];                                        // two wideband buffers per device
    std::vector<void*> out[2];                                           // two result buffers per device
    std::vector<cudaEvent_t> ev_in[2], ev_read[2], ev_done[2];           // block arrived / kernel has read it / results are on the host
    int n_out[2] = {0, 0};
    long submitted = 0, collected = 0;
    NcclApi nccl;
    std::mutex mu;
};

#define M_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cuda_fail(e_, #call, __FILE__, __LINE__); return -1; } } while (0)

extern "C" {

void csdrb_multi_bank_destroy(csdrb_multi_bank_t* m)
{
    if (!m) return;
    for (int i = 0; i < m->ndev; i++) {
        if (cudaSetDevice(m->dev[(size_t)i]) != cudaSuccess) continue;
        if ((size_t)i < m->st.size() && m->st[(size_t)i]) cudaStreamSynchronize(m->st[(size_t)i]);
        if ((size_t)i < m->cst.size() && m->cst[(size_t)i]) cudaStreamSynchronize(m->cst[(size_t)i]);
        if ((size_t)i < m->bank.size() && m->bank[(size_t)i]) csdrb_ddc_bank_destroy(m->bank[(size_t)i]);
        for (int s = 0; s < 2; s++) {
            if ((size_t)i < m->wide[s].size()) cudaFree(m->wide[s][(size_t)i]);
            if ((size_t)i < m->out[s].size()) cudaFree(m->out[s][(size_t)i]);
            if ((size_t)i < m->ev_in[s].size() && m->ev_in[s][(size_t)i]) cudaEventDestroy(m->ev_in[s][(size_t)i]);
            if ((size_t)i < m->ev_read[s].size() && m->ev_read[s][(size_t)i]) cudaEventDestroy(m->ev_read[s][(size_t)i]);
            if ((size_t)i < m->ev_done[s].size() && m->ev_done[s][(size_t)i]) cudaEventDestroy(m->ev_done[s][(size_t)i]);
        }
        if ((size_t)i < m->comm.size() && m->comm[(size_t)i] && m->nccl.CommDestroy) m->nccl.CommDestroy(m->comm[(size_t)i]);
        if ((size_t)i < m->st.size() && m->st[(size_t)i]) cudaStreamDestroy(m->st[(size_t)i]);
        if ((size_t)i < m->cst.size() && m->cst[(size_t)i]) cudaStreamDestroy(m->cst[(size_t)i]);
    }
    delete m;
}

csdrb_multi_bank_t* csdrb_multi_bank_create(int ndev, const int* devices, int channels, const float* h_rates, int decimation, const float* h_taps,
                                            int taps_length, int demod, int chunk, int max_block)
{
    if (ndev <= 0 || channels < ndev || !h_rates || !h_taps || decimation <= 0 || taps_length <= 0 || max_block < taps_length) {
        set_error("multi bank create: bad argument (need 1 <= devices <= channels, max_block >= taps_length)"); return nullptr;
    }
    int prev = 0; cudaGetDevice(&prev);
    auto* m = new csdrb_multi_bank_s();
    m->ndev = ndev; m->channels = channels; m->decimation = decimation; m->taps_length = taps_length; m->demod = demod ? 1 : 0; m->max_block = (max_block + 1) & ~1;
    const int max_out = (max_block - taps_length) / decimation + 1;
    m->dstride = (max_out + 1) & ~1L;
    bool ok = true;
    for (int i = 0; i < ndev && ok; i++) {
        const int d = devices ? devices[i] : i;
        // contiguous slices, the first channels % ndev devices take one more (channel c -> GPU floor(c / (C/G)))
        const int base = channels / ndev, extra = channels % ndev;
        m->dev.push_back(d); m->nch.push_back(base + (i < extra ? 1 : 0)); m->ch0.push_back(i * base + (i < extra ? i : extra));
        ok = cudaSetDevice(d) == cudaSuccess;
        cudaStream_t s = nullptr, c = nullptr;
        ok = ok && cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess && cudaStreamCreateWithFlags(&c, cudaStreamNonBlocking) == cudaSuccess;
        m->st.push_back(s); m->cst.push_back(c);
        csdrb_ddc_bank_t* b = ok ? csdrb_ddc_bank_create(m->nch.back(), h_rates + m->ch0.back(), decimation, h_taps, taps_length, demod, chunk) : nullptr;
        ok = ok && b != nullptr;
        m->bank.push_back(b);
        for (int sidx = 0; sidx < 2; sidx++) {
            float2* w = nullptr; void* o = nullptr; cudaEvent_t e1 = nullptr, e2 = nullptr, e3 = nullptr;
            ok = ok && cudaMalloc(&w, sizeof(float2) * (size_t)m->max_block) == cudaSuccess;
            ok = ok && cudaMalloc(&o, (demod ? 4 : 8) * (size_t)m->dstride * (size_t)m->nch.back()) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&e1, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&e2, cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&e3, cudaEventDisableTiming) == cudaSuccess;
            m->wide[sidx].push_back(w); m->out[sidx].push_back(o); m->ev_in[sidx].push_back(e1); m->ev_read[sidx].push_back(e2); m->ev_done[sidx].push_back(e3);
        }
    }
    if (ok && ndev > 1) {
        ok = m->nccl.load();
        if (ok) {
            m->comm.assign((size_t)ndev, nullptr);
            const int rc = m->nccl.CommInitAll(m->comm.data(), ndev, m->dev.data());
            if (rc != 0) { set_error("multi bank create: ncclCommInitAll failed: %s", m->nccl.GetErrorString(rc)); ok = false; }
        }
    } else if (!ok && !csdrb_last_error()[0]) {
        set_error("multi bank create: CUDA object creation failed (%s)", cudaGetErrorString(cudaGetLastError()));
    }
    cudaSetDevice(prev);
    if (!ok) { csdrb_multi_bank_destroy(m); return nullptr; }
    return m;
}

int csdrb_multi_bank_devices(const csdrb_multi_bank_t* m) { return m ? m->ndev : -1; }
int csdrb_multi_bank_slice(const csdrb_multi_bank_t* m, int index, int* device, int* first_channel, int* channels)
{
    if (!m || index < 0 || index >= m->ndev) { set_error("multi bank slice: bad index"); return -1; }
    if (device) *device = m->dev[(size_t)index];
    if (first_channel) *first_channel = m->ch0[(size_t)index];
    if (channels) *channels = m->nch[(size_t)index];
    return 0;
}

int csdrb_multi_bank_set_rate(csdrb_multi_bank_t* m, int channel, float rate)
{
    if (!m || channel < 0 || channel >= m->channels) { set_error("multi bank set_rate: bad channel"); return -1; }
    // a retune re-chunks the NCO of its whole bank at the next block (csdrb_ddc_bank_set_rate); the other slices do the same, so that the result does
    // not depend on how the channels are sliced over devices
    int rc = -1;
    for (int i = 0; i < m->ndev; i++) {
        const size_t k = (size_t)i;
        if (channel >= m->ch0[k] && channel < m->ch0[k] + m->nch[k]) rc = csdrb_ddc_bank_set_rate(m->bank[k], channel - m->ch0[k], rate);
        else csdrb_ddc_bank_rechunk(m->bank[k]);
    }
    return rc;
}

// Enqueue one wideband block: returns a ticket (>= 0) for csdrb_multi_bank_collect, or a negative error.  At most two blocks may be in flight.
// h_out: [channels][out_stride] floats (demod) or complexf; rows of every slice are written by that slice's own device.
int csdrb_multi_bank_submit(csdrb_multi_bank_t* m, const complexf* h_wide, int input_size, void* h_out, long out_stride)
{
    if (!m || !h_wide || !h_out) { set_error("multi bank submit: null pointer"); return -1; }
    std::lock_guard<std::mutex> lk(m->mu);
    if (input_size > m->max_block || input_size < m->taps_length) { set_error("multi bank submit: block of %d samples (taps %d .. max_block %d)", input_size, m->taps_length, m->max_block); return -1; }
    if (m->submitted - m->collected >= 2) { set_error("multi bank submit: two blocks already in flight -- collect one first"); return -1; }
    const int slot = (int)(m->submitted & 1);
    const int n_out = (input_size - m->taps_length) / m->decimation + 1;
    if (out_stride < n_out) { set_error("multi bank submit: out_stride %ld < %d outputs", out_stride, n_out); return -1; }
    int prev = 0; cudaGetDevice(&prev);
    // 1. the block reaches device 0 over PCIe (after the kernel that last read this buffer), then every other device over NVLink
    M_CUDA(cudaSetDevice(m->dev[0]));
    M_CUDA(cudaStreamWaitEvent(m->cst[0], m->ev_read[slot][0], 0));
    M_CUDA(cudaMemcpyAsync(m->wide[slot][0], h_wide, sizeof(float2) * (size_t)input_size, cudaMemcpyHostToDevice, m->cst[0]));
    if (m->ndev > 1) {
        for (int i = 1; i < m->ndev; i++) { M_CUDA(cudaSetDevice(m->dev[(size_t)i])); M_CUDA(cudaStreamWaitEvent(m->cst[(size_t)i], m->ev_read[slot][(size_t)i], 0)); }
        int rc = m->nccl.GroupStart();
        for (int i = 0; i < m->ndev && rc == 0; i++)
            rc = m->nccl.Broadcast(m->wide[slot][0], m->wide[slot][(size_t)i], 2 * (size_t)input_size, kNcclFloat32, 0, m->comm[(size_t)i], m->cst[(size_t)i]);
        const int rc2 = m->nccl.GroupEnd();
        if (rc != 0 || rc2 != 0) { set_error("multi bank submit: ncclBroadcast failed: %s", m->nccl.GetErrorString(rc ? rc : rc2)); cudaSetDevice(prev); return -1; }
    }
    // 2. every device: wait for its copy, run its slice, send the results home
    const size_t esz = m->demod ? 4 : 8;
    for (int i = 0; i < m->ndev; i++) {
        const size_t k = (size_t)i;
        M_CUDA(cudaSetDevice(m->dev[k]));
        M_CUDA(cudaEventRecord(m->ev_in[slot][k], m->cst[k]));
        M_CUDA(cudaStreamWaitEvent(m->st[k], m->ev_in[slot][k], 0));
        const int rc = csdrb_ddc_bank_process(m->bank[k], reinterpret_cast<const complexf*>(m->wide[slot][k]), input_size, m->out[slot][k], m->dstride, m->st[k]);
        if (rc < 0) { cudaSetDevice(prev); return rc; }
        M_CUDA(cudaEventRecord(m->ev_read[slot][k], m->st[k]));
        M_CUDA(cudaMemcpy2DAsync(static_cast<char*>(h_out) + (size_t)m->ch0[k] * (size_t)out_stride * esz, (size_t)out_stride * esz, m->out[slot][k], (size_t)m->dstride * esz,
                                 (size_t)n_out * esz, (size_t)m->nch[k], cudaMemcpyDeviceToHost, m->st[k]));
        M_CUDA(cudaEventRecord(m->ev_done[slot][k], m->st[k]));
    }
    cudaSetDevice(prev);
    m->n_out[slot] = n_out;
    return (int)(m->submitted++ & 0x3fffffff);
}

// Wait for the block `ticket` names (tickets complete in order); returns its outputs per channel.
int csdrb_multi_bank_collect(csdrb_multi_bank_t* m, int ticket)
{
    if (!m) { set_error("multi bank collect: null pointer"); return -1; }
    std::lock_guard<std::mutex> lk(m->mu);
    if (m->collected >= m->submitted || ticket != (int)(m->collected & 0x3fffffff)) { set_error("multi bank collect: ticket %d is not the oldest block in flight", ticket); return -1; }
    const int slot = (int)(m->collected & 1);
    for (int i = 0; i < m->ndev; i++) M_CUDA(cudaEventSynchronize(m->ev_done[slot][(size_t)i]));
    m->collected++;
    return m->n_out[slot];
}

// One block, synchronously (submit + collect): the simple form for callers that do not pipeline.
int csdrb_multi_bank_process_host(csdrb_multi_bank_t* m, const complexf* h_wide, int input_size, void* h_out, long out_stride)
{
    const int t = csdrb_multi_bank_submit(m, h_wide, input_size, h_out, out_stride);
    return t < 0 ? t : csdrb_multi_bank_collect(m, t);
}

}  // extern "C"
