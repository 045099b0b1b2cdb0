/*
 * Declarations-only stand-in for <fftw3.h>  --  TEST INFRASTRUCTURE, not product code.
 *
 * FFTW3 (single precision, -lfftw3f; reference Makefile:39) is a third-party dependency of the
 * reference that is not vendored under /root/reference and is not installed in this image.
 * The reference only touches it through fft_fftw.c:9,19,29,38,43 and the fft_malloc/fft_free
 * macros (fft_fftw.h:11-12).  This header declares exactly that surface; the definitions live in
 * fftw_f64.c (a float64 DFT rounded to float -- the mathematical definition of what FFTW computes).
 * libcsdr.h:385 uses FILE and relies on the real fftw3.h pulling <stdio.h> in, so we do too.
 */
#ifndef ORACLE_FFTW3_SHIM_H
#define ORACLE_FFTW3_SHIM_H
#include <stdio.h>
#include <stddef.h>

typedef float fftwf_complex[2];
typedef struct oracle_fftwf_plan_s *fftwf_plan;

#define FFTW_FORWARD  (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE  (0U)
#define FFTW_ESTIMATE (1U << 6)

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned flags);
fftwf_plan fftwf_plan_dft_r2c_1d(int n, float *in, fftwf_complex *out, unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_1d(int n, fftwf_complex *in, float *out, unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
void *fftwf_malloc(size_t n);
void fftwf_free(void *p);
#endif
