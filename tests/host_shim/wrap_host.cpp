// Host build of the SHIPPED device function csdrb::wrap_phase_pm_pi (csdr_b200/csrc/common.cuh) for the CPU test tier:
// the CUDA intrinsics it uses are mapped onto the same IEEE-754 single-precision operations (SSE, no contraction), everything
// else in the header is never instantiated on the host.  Built and driven by tests/test_phase_wrap_host.py.
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>

static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline size_t __cvta_generic_to_shared(const void*) { return 0; }

#include "../../csdr_b200/csrc/common.cuh"

extern "C" {
// the reference's loop, libcsdr_gpl.c:49-50 (PI is the float constant of libcsdr.h:65); guarded like the device code for |ph| >= 2^26
float wrap_loop(float ph)
{
    const float PI = (float)3.14159265358979323846;
    if (!(std::fabs(ph) < 67108864.f)) return ph;
    while (ph > PI) { volatile float t = ph - 2 * PI; ph = t; }
    while (ph < -PI) { volatile float t = ph + 2 * PI; ph = t; }
    return ph;
}
float wrap_fast(float ph) { return csdrb::wrap_phase_pm_pi(ph); }
// returns the number of mismatches over n inputs (first mismatching input in *bad)
long wrap_compare(const float* x, long n, float* bad)
{
    long miss = 0;
    for (long i = 0; i < n; i++) {
        const float a = wrap_loop(x[i]), b = wrap_fast(x[i]);
        if (__float_as_uint(a) != __float_as_uint(b)) { if (!miss && bad) *bad = x[i]; miss++; }
    }
    return miss;
}
}
