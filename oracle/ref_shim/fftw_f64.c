/*
 * float64 DFT provider behind the five fftwf_* entry points the reference calls
 * (fft_fftw.c:9,19,29,38,43).  TEST INFRASTRUCTURE ONLY (see fftw3.h in this directory).
 *
 * Semantics follow the FFTW3 manual: unnormalised transform, sign -1 = forward, +1 = backward,
 * out-of-place, plan bound to its in/out pointers; r2c emits n/2+1 bins, c2r consumes n/2+1 bins.
 * Power-of-two sizes use an iterative radix-2 decimation-in-time transform in double precision;
 * any other size falls back to the O(n^2) definition.  Results are rounded to float once, at the end.
 */
#include "fftw3.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { K_C2C, K_R2C, K_C2R };
struct oracle_fftwf_plan_s { int n, sign, kind; void *in, *out; double *w; };

static void dft_pow2(double *re, double *im, int n, int sign)
{
    for (int i = 1, j = 0; i < n; i++) {                 /* bit reversal */
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { double t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int s = 0; s < n; s += len)
            for (int k = 0; k < half; k++) {
                double ang = sign * 2.0 * M_PI * (double)(k * step) / (double)n;
                double wr = cos(ang), wi = sin(ang);
                double xr = re[s + k + half] * wr - im[s + k + half] * wi;
                double xi = re[s + k + half] * wi + im[s + k + half] * wr;
                re[s + k + half] = re[s + k] - xr; im[s + k + half] = im[s + k] - xi;
                re[s + k] += xr;                  im[s + k] += xi;
            }
    }
}

static void dft_any(double *re, double *im, int n, int sign)
{
    if (n > 0 && (n & (n - 1)) == 0) { dft_pow2(re, im, n, sign); return; }
    double *or_ = malloc(sizeof(double) * n), *oi = malloc(sizeof(double) * n);
    for (int k = 0; k < n; k++) {
        double sr = 0, si = 0;
        for (int t = 0; t < n; t++) {
            double ang = sign * 2.0 * M_PI * (double)(((long long)k * t) % n) / (double)n;
            sr += re[t] * cos(ang) - im[t] * sin(ang);
            si += re[t] * sin(ang) + im[t] * cos(ang);
        }
        or_[k] = sr; oi[k] = si;
    }
    memcpy(re, or_, sizeof(double) * n); memcpy(im, oi, sizeof(double) * n);
    free(or_); free(oi);
}

static fftwf_plan mk(int n, void *in, void *out, int sign, int kind)
{
    fftwf_plan p = malloc(sizeof(*p));
    p->n = n; p->in = in; p->out = out; p->sign = sign; p->kind = kind;
    p->w = malloc(sizeof(double) * 2 * (n > 0 ? n : 1));
    return p;
}
fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned f) { (void)f; return mk(n, in, out, sign, K_C2C); }
fftwf_plan fftwf_plan_dft_r2c_1d(int n, float *in, fftwf_complex *out, unsigned f) { (void)f; return mk(n, in, out, FFTW_FORWARD, K_R2C); }
fftwf_plan fftwf_plan_dft_c2r_1d(int n, fftwf_complex *in, float *out, unsigned f) { (void)f; return mk(n, in, out, FFTW_BACKWARD, K_C2R); }

void fftwf_execute(const fftwf_plan p)
{
    int n = p->n; double *re = p->w, *im = p->w + n;
    if (p->kind == K_C2C) {
        const float *x = p->in;
        for (int i = 0; i < n; i++) { re[i] = x[2 * i]; im[i] = x[2 * i + 1]; }
        dft_any(re, im, n, p->sign);
        float *y = p->out;
        for (int i = 0; i < n; i++) { y[2 * i] = (float)re[i]; y[2 * i + 1] = (float)im[i]; }
    } else if (p->kind == K_R2C) {
        const float *x = p->in;
        for (int i = 0; i < n; i++) { re[i] = x[i]; im[i] = 0; }
        dft_any(re, im, n, FFTW_FORWARD);
        float *y = p->out;
        for (int i = 0; i <= n / 2; i++) { y[2 * i] = (float)re[i]; y[2 * i + 1] = (float)im[i]; }
    } else {
        const float *x = p->in;
        for (int i = 0; i <= n / 2; i++) { re[i] = x[2 * i]; im[i] = x[2 * i + 1]; }
        for (int i = n / 2 + 1; i < n; i++) { re[i] = re[n - i]; im[i] = -im[n - i]; }
        im[0] = 0; if (n % 2 == 0) im[n / 2] = 0;
        dft_any(re, im, n, FFTW_BACKWARD);
        float *y = p->out;
        for (int i = 0; i < n; i++) y[i] = (float)re[i];
    }
}
void fftwf_destroy_plan(fftwf_plan p) { if (p) { free(p->w); free(p); } }
void *fftwf_malloc(size_t n) { void *q = NULL; return posix_memalign(&q, 64, n ? n : 64) ? NULL : q; }
void fftwf_free(void *p) { free(p); }
