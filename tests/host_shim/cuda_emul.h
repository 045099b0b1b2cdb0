// cuda_emul.h -- TEST INFRASTRUCTURE: run the shipped CUDA kernels thread by thread on the CPU (CPU test tier only).
//
// Not a CPU fallback: nothing under csdr_b200/ can reach this file; it exists so that kernels which cannot be run in this container
// (no GPU) are still EXECUTED, with their real index arithmetic, barriers and rounding, before GPU minutes are spent on them.
// Model: one CTA at a time; every CUDA thread of the CTA is a ucontext fiber on one OS thread; __syncthreads() / __syncwarp() are
// real barriers between fibers (a fiber that reaches one yields until all live fibers of the CTA / warp have arrived);
// threadIdx/blockIdx/blockDim/gridDim are per-fiber values; dynamic shared memory is one zeroed buffer per CTA; IEEE single-precision
// intrinsics map onto the same operations (fmaf is a true fused multiply-add on the host as well).
// Not modelled: inline PTX (kernels that use it are not emulated), warp shuffles/votes, atomics, memory-model subtleties -- a data race
// a GPU could expose may go unnoticed here (fibers switch only at barriers).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <functional>
#include <vector>
#include <ucontext.h>

#include <cuda_runtime.h>                       // float2/float4/dim3/uint3, make_float2 ... (host-usable headers of the toolkit)

namespace cuda_emul {

struct Fiber {
    ucontext_t ctx;
    std::vector<unsigned char> stack;
    uint3 tid;
    bool done = false;
};

struct State {
    ucontext_t sched;
    std::vector<Fiber> fibers;
    Fiber* cur = nullptr;
    dim3 block, grid;
    uint3 bid;
    std::function<void()> body;
    int live = 0, arrived = 0;
    unsigned long generation = 0;
    std::vector<int> warp_live, warp_arrived;
    std::vector<unsigned long> warp_generation;
    std::vector<unsigned char> smem;
    long barriers = 0, switches = 0;
    int order_mode = 0;                       // 0 alternate, 1 reverse, 2 random, 3 forward
    unsigned long long rng = 88172645463325252ULL;
};
inline State& st()
{
    static State s;
    static bool init = false;
    if (!init) {
        init = true;
        const char* o = getenv("CUDA_EMUL_ORDER");
        if (o) s.order_mode = !strcmp(o, "reverse") ? 1 : !strcmp(o, "random") ? 2 : !strcmp(o, "forward") ? 3 : 0;
    }
    return s;
}

inline void yield() { State& s = st(); s.switches++; swapcontext(&s.cur->ctx, &s.sched); }

inline void sync_threads()
{
    State& s = st();
    const unsigned long my = s.generation;
    if (++s.arrived == s.live) { s.arrived = 0; s.generation++; s.barriers++; return; }
    while (s.generation == my) yield();
}

inline int linear_tid(const Fiber* f) { State& s = st(); return (int)(f->tid.x + s.block.x * (f->tid.y + s.block.y * f->tid.z)); }

inline void sync_warp()
{
    State& s = st();
    const int w = linear_tid(s.cur) / 32;
    const unsigned long my = s.warp_generation[w];
    if (++s.warp_arrived[w] == s.warp_live[w]) { s.warp_arrived[w] = 0; s.warp_generation[w]++; return; }
    while (s.warp_generation[w] == my) yield();
}

inline void trampoline()
{
    State& s = st();
    s.body();
    Fiber* f = s.cur;
    f->done = true;
    s.live--;
    s.warp_live[linear_tid(f) / 32]--;
    // a barrier the others are waiting in may have become complete by this thread's exit
    if (s.live > 0 && s.arrived == s.live) { s.arrived = 0; s.generation++; }
    const int w = linear_tid(f) / 32;
    if (s.warp_live[w] > 0 && s.warp_arrived[w] == s.warp_live[w]) { s.warp_arrived[w] = 0; s.warp_generation[w]++; }
    swapcontext(&f->ctx, &s.sched);
}

inline unsigned char* dyn_smem() { return st().smem.data(); }

// run `kernel(args...)` for every thread of every CTA of the grid
template <typename K, typename... A>
void launch(dim3 grid, dim3 block, size_t smem_bytes, K kernel, A... args)
{
    State& s = st();
    s.grid = grid; s.block = block;
    const int nthreads = (int)(block.x * block.y * block.z);
    const size_t stack_bytes = 256 * 1024;
    s.body = [=]() { kernel(args...); };
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                s.bid = make_uint3(bx, by, bz);
                s.smem.assign(smem_bytes + 64, 0);
                s.fibers.clear(); s.fibers.resize((size_t)nthreads);
                s.live = nthreads; s.arrived = 0; s.generation = 0;
                const int nwarps = (nthreads + 31) / 32;
                s.warp_live.assign((size_t)nwarps, 0); s.warp_arrived.assign((size_t)nwarps, 0); s.warp_generation.assign((size_t)nwarps, 0);
                for (int t = 0; t < nthreads; t++) {
                    Fiber& f = s.fibers[(size_t)t];
                    f.tid = make_uint3((unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y));
                    f.stack.resize(stack_bytes);
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = f.stack.data(); f.ctx.uc_stack.ss_size = stack_bytes; f.ctx.uc_link = nullptr;
                    makecontext(&f.ctx, (void (*)())trampoline, 0);
                    s.warp_live[(size_t)(t / 32)]++;
                }
                // Between two barriers the fibers run one after the other, so a missing barrier only shows when the victim happens to run
                // in the unlucky order: CUDA_EMUL_ORDER = forward | reverse | random (default: alternate forward / reverse every sweep,
                // which exposes read-after-write and write-after-read hazards in both directions).
                std::vector<int> order((size_t)nthreads);
                for (int t = 0; t < nthreads; t++) order[(size_t)t] = t;
                unsigned long sweep = 0;
                while (s.live > 0) {
                    if (s.order_mode == 1 || (s.order_mode == 0 && (sweep & 1))) std::reverse(order.begin(), order.end());
                    else if (s.order_mode == 2) for (int t = nthreads - 1; t > 0; t--) { s.rng = s.rng * 6364136223846793005ULL + 1442695040888963407ULL; std::swap(order[(size_t)t], order[(size_t)((s.rng >> 33) % (unsigned long)(t + 1))]); }
                    for (int k = 0; k < nthreads; k++) {
                        Fiber& f = s.fibers[(size_t)order[(size_t)k]];
                        if (f.done) continue;
                        s.cur = &f;
                        swapcontext(&s.sched, &f.ctx);
                    }
                    if (s.order_mode == 1 || (s.order_mode == 0 && (sweep & 1))) std::reverse(order.begin(), order.end());
                    sweep++;
                }
            }
}

}  // namespace cuda_emul

// ---- the CUDA surface the kernels use ----------------------------------------------------------------------------------------
#define threadIdx (::cuda_emul::st().cur->tid)
#define blockIdx (::cuda_emul::st().bid)
#define blockDim (::cuda_emul::st().block)
#define gridDim (::cuda_emul::st().grid)
#define __syncthreads() ::cuda_emul::sync_threads()
#define __syncwarp() ::cuda_emul::sync_warp()
#define CSDRB_DYN_SMEM(name) unsigned char* name = ::cuda_emul::dyn_smem()
#define __launch_bounds__(...)
#define __grid_constant__
#undef __global__
#define __global__
#undef __shared__
#define __shared__ static                       /* one CTA runs at a time, so a function-level static is CTA-shared */

template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline size_t __cvta_generic_to_shared(const void*) { return 0; }
