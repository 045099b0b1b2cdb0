// fft.cuh -- K7: block-level Stockham autosort FFT in shared memory (no cuFFT), power-of-two sizes.
//
// Stands behind the reference's FFT abstraction (fft_fftw.c:6-41: unnormalised DFT, exponent sign -1
// forward / +1 backward).  One CTA transforms one N-point signal that already sits in shared memory.
//
// Pass structure (validated against numpy in the design notes): radices are
// [2 or 4 (if log2 N is not a multiple of 3)] followed by radix-8 passes.  For a pass of radix R over
// sub-transforms of size Ns (Ns = product of the previous radices), butterfly j in [0, N/R):
//     k = j mod Ns;   v[r] = s[j + r*N/R] * W_N^(k*r*N/(Ns*R));   V = DFT_R(v);   s'[(j/Ns)*Ns*R + k + r*Ns] = V[r]
// All threads read their butterflies into registers, synchronise, then write: one buffer is enough
// (a thread holds at most 16 points per pass, so NT >= N/16 threads are required).
// Twiddles: for a radix-8 pass over sub-transforms of size NS the butterfly with k = j mod NS needs w^r, w = exp(-2*pi*i*k/(8*NS)).
// Three planes of a per-size table hold w^1, w^2, w^4 at index NS + k (the ranges [NS, 2NS) of the different passes do not
// overlap), computed in double on the host and rounded once to float; lanes with consecutive k read consecutive entries (the r01
// profile showed seven scattered LDGs per butterfly saturating L1TEX).  w^3, w^5, w^6, w^7 are one float product each, so every
// twiddle is within ~2 ulp and the whole transform stays ~1e-7*log2 N from the exact DFT.
#pragma once
#include "common.cuh"
#include <cmath>

namespace csdrb {

// Shared-memory layout: element i lives at i + 2*(i >> 4) (two pad slots per sixteen complex values, 12.5 % extra; even indices stay
// even so element pairs are 16-byte aligned).  A thread owns PAIRS of adjacent butterflies and moves them with 128-bit LDS/STS:
// in the first pass it writes 16 consecutive elements (lane stride 18 elements = 9 x 16 B, odd: conflict-free), in the later passes
// the pair lands at adjacent positions again (sub-transform sizes are even).
__host__ __device__ constexpr int fft_pad(int i) { return i + 2 * (i >> 4); }
__host__ __device__ constexpr int fft_smem_elems(int n) { return n + 2 * (n >> 4) + 2; }

template <bool INV>
__device__ __forceinline__ float2 cmul_w(float2 a, float2 w)
{
    // a * w (forward) or a * conj(w) (inverse)
    return INV ? make_float2(fmaf(a.x, w.x, a.y * w.y), fmaf(a.y, w.x, -a.x * w.y))
               : make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.y, w.x, a.x * w.y));
}
template <bool INV>
__device__ __forceinline__ float2 mul_mi(float2 a)       // a * (-i) forward, a * (+i) inverse
{
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

template <bool INV>
__device__ __forceinline__ void dft2(float2& a, float2& b) { float2 t = a; a = cadd(t, b); b = csub(t, b); }

template <bool INV>
__device__ __forceinline__ void dft4(float2& v0, float2& v1, float2& v2, float2& v3)
{
    const float2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = mul_mi<INV>(csub(v1, v3));
    v0 = cadd(a0, a2); v2 = csub(a0, a2); v1 = cadd(a1, a3); v3 = csub(a1, a3);
}

template <bool INV>
__device__ __forceinline__ void dft8(float2 (&v)[8])
{
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4<INV>(e0, e1, e2, e3);
    dft4<INV>(o0, o1, o2, o3);
    const float h = 0.70710678118654752440f;
    // W8^1 = h*(1 -+ i), W8^2 = -+i, W8^3 = h*(-1 -+ i)
    const float2 t1 = INV ? make_float2(h * (o1.x - o1.y), h * (o1.x + o1.y)) : make_float2(h * (o1.x + o1.y), h * (o1.y - o1.x));
    const float2 t2 = mul_mi<INV>(o2);
    const float2 t3 = INV ? make_float2(-h * (o3.x + o3.y), h * (o3.x - o3.y)) : make_float2(h * (o3.y - o3.x), -h * (o3.x + o3.y));
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, t1); v[5] = csub(e1, t1);
    v[2] = cadd(e2, t2); v[6] = csub(e2, t2);
    v[3] = cadd(e3, t3); v[7] = csub(e3, t3);
}

template <int R, bool INV>
__device__ __forceinline__ void dft_small(float2 (&v)[R])
{
    if constexpr (R == 2) dft2<INV>(v[0], v[1]);
    else if constexpr (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
    else dft8<INV>(v);
}

template <int N, int R, int NS, bool INV>
__device__ __forceinline__ void fft_butterfly(float2 (&v)[R], int j, const float2* __restrict__ tw)
{
    if constexpr (NS > 1) {
        static_assert(R == 8, "only the first pass may have a radix below 8");
        const int k = j % NS;
        const float2 w1 = __ldg(tw + NS + k), w2 = __ldg(tw + N + NS + k), w4 = __ldg(tw + 2 * N + NS + k);
        const float2 w3 = make_float2(fmaf(w1.x, w2.x, -w1.y * w2.y), fmaf(w1.x, w2.y, w1.y * w2.x));
        const float2 w5 = make_float2(fmaf(w1.x, w4.x, -w1.y * w4.y), fmaf(w1.x, w4.y, w1.y * w4.x));
        const float2 w6 = make_float2(fmaf(w2.x, w4.x, -w2.y * w4.y), fmaf(w2.x, w4.y, w2.y * w4.x));
        const float2 w7 = make_float2(fmaf(w3.x, w4.x, -w3.y * w4.y), fmaf(w3.x, w4.y, w3.y * w4.x));
        v[1] = cmul_w<INV>(v[1], w1); v[2] = cmul_w<INV>(v[2], w2); v[3] = cmul_w<INV>(v[3], w3);
        v[4] = cmul_w<INV>(v[4], w4); v[5] = cmul_w<INV>(v[5], w5); v[6] = cmul_w<INV>(v[6], w6);
        v[7] = cmul_w<INV>(v[7], w7);
    }
    dft_small<R, INV>(v);
}

template <int N, int NT, int R, int NS, bool INV>
__device__ __forceinline__ void fft_pass(float2* __restrict__ s, const float2* __restrict__ tw, int tid)
{
    constexpr int NB = N / R;
    constexpr int PER = (NB + NT - 1) / NT;
    static_assert(PER * R <= 16, "block_fft needs NT >= N/16 threads");
    constexpr bool PAIRED = (PER % 2 == 0) && (NB % (NT * PER) == 0);   // every thread owns whole pairs of adjacent butterflies
    float2 v[PER][R];
    if constexpr (PAIRED) {
#pragma unroll
        for (int b = 0; b < PER; b += 2) {
            const int j = tid * PER + b;                                // j even: (j, j+1) adjacent and 16-byte aligned in the padded array
#pragma unroll
            for (int r = 0; r < R; r++) {
                const float4 two = *reinterpret_cast<const float4*>(s + fft_pad(j + r * NB));
                v[b][r] = make_float2(two.x, two.y); v[b + 1][r] = make_float2(two.z, two.w);
            }
            fft_butterfly<N, R, NS, INV>(v[b], j, tw);
            fft_butterfly<N, R, NS, INV>(v[b + 1], j + 1, tw);
        }
        __syncthreads();
#pragma unroll
        for (int b = 0; b < PER; b += 2) {
            const int j = tid * PER + b;
            if constexpr (NS == 1) {
                // first pass: the two butterflies write 2R consecutive elements
#pragma unroll
                for (int r = 0; r < R; r += 2) {
                    *reinterpret_cast<float4*>(s + fft_pad(j * R + r)) = make_float4(v[b][r].x, v[b][r].y, v[b][r + 1].x, v[b][r + 1].y);
                    *reinterpret_cast<float4*>(s + fft_pad((j + 1) * R + r)) = make_float4(v[b + 1][r].x, v[b + 1][r].y, v[b + 1][r + 1].x, v[b + 1][r + 1].y);
                }
            } else {
                const int j0 = (j / NS) * NS * R + (j % NS);            // NS even: j and j+1 share the group, their outputs are adjacent
#pragma unroll
                for (int r = 0; r < R; r++)
                    *reinterpret_cast<float4*>(s + fft_pad(j0 + r * NS)) = make_float4(v[b][r].x, v[b][r].y, v[b + 1][r].x, v[b + 1][r].y);
            }
        }
        __syncthreads();
    } else {
#pragma unroll
        for (int b = 0; b < PER; b++) {
            const int j = tid + b * NT;
            if (NB % NT == 0 || j < NB) {
#pragma unroll
                for (int r = 0; r < R; r++) v[b][r] = s[fft_pad(j + r * NB)];
                fft_butterfly<N, R, NS, INV>(v[b], j, tw);
            }
        }
        __syncthreads();
#pragma unroll
        for (int b = 0; b < PER; b++) {
            const int j = tid + b * NT;
            if (NB % NT == 0 || j < NB) {
                const int j0 = (j / NS) * NS * R + (j % NS);
#pragma unroll
                for (int r = 0; r < R; r++) s[fft_pad(j0 + r * NS)] = v[b][r];
            }
        }
        __syncthreads();
    }
}

template <int N, int NT, int NS, bool INV>
__device__ __forceinline__ void fft_r8_passes(float2* __restrict__ s, const float2* __restrict__ tw, int tid)
{
    if constexpr (NS < N) {
        fft_pass<N, NT, 8, NS, INV>(s, tw, tid);
        fft_r8_passes<N, NT, NS * 8, INV>(s, tw, tid);
    }
}

constexpr int ilog2_c(int n) { return n <= 1 ? 0 : 1 + ilog2_c(n / 2); }

// In-place N-point transform of the PADDED array s (element i at s[fft_pad(i)], fft_smem_elems(N) slots) by a CTA of NT threads
// (all NT threads must call; a __syncthreads() must separate the last write to s from this call).
template <int N, int NT, bool INV>
__device__ __forceinline__ void block_fft(float2* __restrict__ s, const float2* __restrict__ tw, int tid)
{
    static_assert((N & (N - 1)) == 0 && N >= 2, "power of two sizes only");
    constexpr int LG = ilog2_c(N);
    if constexpr (LG % 3 == 1) { fft_pass<N, NT, 2, 1, INV>(s, tw, tid); fft_r8_passes<N, NT, 2, INV>(s, tw, tid); }
    else if constexpr (LG % 3 == 2) { fft_pass<N, NT, 4, 1, INV>(s, tw, tid); fft_r8_passes<N, NT, 4, INV>(s, tw, tid); }
    else fft_r8_passes<N, NT, 1, INV>(s, tw, tid);
}

// Prologue copy global -> padded shared array.  All of a thread's loads are issued before the first store: the r01 source-level
// profile had 58 % of the 4096-point kernel's stall samples on the first STS of a load-then-store loop (16 serial DRAM round trips).
template <int N, int NT, typename Fetch>
__device__ __forceinline__ void fft_stage_in(float2* __restrict__ s, int tid, Fetch fetch)
{
    constexpr int PER = (N + NT - 1) / NT;
    float2 t[PER];
#pragma unroll
    for (int k = 0; k < PER; k++) { const int i = tid + k * NT; t[k] = (N % NT == 0 || i < N) ? fetch(i) : make_float2(0.f, 0.f); }
#pragma unroll
    for (int k = 0; k < PER; k++) { const int i = tid + k * NT; if (N % NT == 0 || i < N) s[fft_pad(i)] = t[k]; }
}
// 128-bit variant for 16-byte aligned, fully dense rows (two samples per load; N >= 2*NT)
template <int N, int NT>
__device__ __forceinline__ void fft_stage_in_vec(float2* __restrict__ s, int tid, const float2* __restrict__ x)
{
    constexpr int PER = N / (2 * NT);
    float4 t[PER];
#pragma unroll
    for (int k = 0; k < PER; k++) t[k] = __ldg(reinterpret_cast<const float4*>(x) + tid + k * NT);
#pragma unroll
    for (int k = 0; k < PER; k++) *reinterpret_cast<float4*>(s + fft_pad(2 * (tid + k * NT))) = t[k];
}

// host: the three twiddle planes (w^1, w^2, w^4) of an n-point transform, 3*n entries; a radix-8 pass over sub-size NS reads index
// NS + k, k < NS.  Angles in double, rounded once to float.
inline void fft_fill_twiddles(int n, float2* h)
{
    for (long i = 0; i < 3L * n; i++) h[i] = make_float2(1.f, 0.f);
    int lg = 0; while ((1 << lg) < n) lg++;
    int ns = lg % 3 == 1 ? 2 : (lg % 3 == 2 ? 4 : 1);
    if (ns == 1) ns = 8;                                                // the first pass (NS = 1) needs no twiddles
    for (; ns < n; ns *= 8)
        for (int k = 0; k < ns; k++)
            for (int c = 0; c < 3; c++) {
                const double a = -2.0 * 3.14159265358979323846 * (double)((1 << c) * k) / (double)(ns * 8);
                h[(size_t)c * n + ns + k] = make_float2((float)cos(a), (float)sin(a));
            }
}

constexpr int fft_threads(int n) { return n / 16 < 32 ? 32 : n / 16; }
constexpr int FFT_MAX_N = 16384;

}  // namespace csdrb
