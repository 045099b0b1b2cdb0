/*
 * csdr_b200.h -- public C ABI of libcsdr_b200.so (B200 / sm_100a implementation of the csdr block-DSP hot path).
 *
 * Two layers, both plain C (System V x86-64, no C++ or torch types in any signature):
 *
 *   Part A  "libcsdr drop-in": the SAME names, argument meaning and return values as the reference
 *           library for the hot-path functions, taking HOST pointers and behaving synchronously, so
 *           the reference's own callers (csdr.c, test200.c) link against this library unchanged.
 *           Each prototype cites the reference declaration it replaces (file:line in ha7ilm/csdr @6ef2a742).
 *
 *   Part B  "bank API" (csdrb_*): the device-resident many-channel entry points the reference can only
 *           express as one process chain per channel (ddcd_old.h:51-61).  Pointers named d_* are DEVICE
 *           pointers, h_* are host pointers; `stream` is a cudaStream_t passed as void* (NULL = legacy
 *           default stream).  Calls are asynchronous on that stream unless stated otherwise.
 *
 * Error model: Part A keeps the reference's (no error codes); a CUDA failure prints to stderr and
 * aborts the process, which is the closest equivalent of the reference's behaviour on a fatal fault.
 * Part B returns >= 0 on success (usually an output count) and a negative value on failure;
 * csdrb_last_error() returns the message of the calling thread's last failure.
 * There is NO CPU fallback anywhere: without a usable CUDA device every compute entry point fails.
 */
#ifndef CSDR_B200_H
#define CSDR_B200_H
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif

/* =====================================================================================================
 * Part A -- libcsdr drop-in surface
 * =================================================================================================== */

typedef struct complexf_s { float i; float q; } complexf;                        /* libcsdr.h:46 */
typedef enum window_s { WINDOW_BOXCAR, WINDOW_BLACKMAN, WINDOW_HAMMING } window_t; /* libcsdr.h:70-75 */
#define WINDOW_DEFAULT WINDOW_HAMMING                                             /* libcsdr.h:77 */

/* filter design -- stays on the host, in C (libcsdr.h:85-92; libcsdr.c:57-174) */
void  firdes_lowpass_f(float *output, int length, float cutoff_rate, window_t window);
void  firdes_bandpass_c(complexf *output, int length, float lowcut, float highcut, window_t window);
float firdes_wkernel_blackman(float input);
float firdes_wkernel_hamming(float input);
float firdes_wkernel_boxcar(float input);
window_t firdes_get_window_from_string(char *input);
char *firdes_get_string_from_window(window_t window);
int   firdes_filter_len(float transition_bw);
int   log2n(int x);                                                               /* libcsdr.h:209 */
int   next_pow2(int x);                                                           /* libcsdr.h:210 */

/* sample-format conversion (libcsdr.h:220-227; libcsdr.c:2363-2401) */
void convert_u8_f(unsigned char *input, float *output, int input_size);
void convert_s16_f(short *input, float *output, int input_size);
void convert_i16_f(short *input, float *output, int input_size);
void convert_f_s16(float *input, short *output, int input_size);
void convert_f_i16(float *input, short *output, int input_size);

/* NCO shift by phasor recursion (libcsdr_gpl.h:26-46; libcsdr_gpl.c:27-52, 81-89, 126-160) */
typedef struct shift_addition_data_s { float sindelta; float cosdelta; float rate; } shift_addition_data_t;
shift_addition_data_t shift_addition_init(float rate);
float shift_addition_cc(complexf *input, complexf *output, int input_size, shift_addition_data_t d, float starting_phase);
typedef struct decimating_shift_addition_status_s { int decimation_remain; float starting_phase; int output_size; } decimating_shift_addition_status_t;
shift_addition_data_t decimating_shift_addition_init(float rate, int decimation);
decimating_shift_addition_status_t decimating_shift_addition_cc(complexf *input, complexf *output, int input_size,
        shift_addition_data_t d, int decimation, decimating_shift_addition_status_t s);

/* decimating FIR (libcsdr.h:104; libcsdr.c:528-549): returns the number of outputs written */
int fir_decimate_cc(complexf *input, complexf *output, int input_size, int decimation, float *taps, int taps_length);

/* FM demodulator (libcsdr.h:95; libcsdr.c:1040-1071): `temp` is accepted and ignored */
complexf fmdemod_quadri_cf(complexf *input, float *output, int input_size, float *temp, complexf last_sample);

/* fractional decimator (libcsdr.h:151-170; libcsdr.c:715-793) */
typedef struct fractional_decimator_ff_s {
    float where; int input_processed; int output_size; int num_poly_points;
    float *poly_precalc_denomiator; float *coeffs_buf; float *filtered_buf;
    int xifirst; int xilast; float rate; float *taps; int taps_length;
} fractional_decimator_ff_t;
fractional_decimator_ff_t fractional_decimator_ff_init(float rate, int num_poly_points, float *taps, int taps_length);
void fractional_decimator_ff(float *input, float *output, int input_size, fractional_decimator_ff_t *d);

/* block AGC (libcsdr.h:118-130; libcsdr.c:944-991) */
typedef struct fastagc_ff_s {
    float *buffer_1; float *buffer_2; float *buffer_input;
    float peak_1; float peak_2; int input_size; float reference; float last_gain;
} fastagc_ff_t;
void fastagc_ff(fastagc_ff_t *input, float *output);

/* audio tail of the WFM graph, SURVEY 8(f) rank 1 (libcsdr.h:100-105; libcsdr.c:1081-1097, 1130-1137) */
float deemphasis_wfm_ff(float *input, float *output, int input_size, float tau, int sample_rate, float last_output);
void  limit_ff(float *input, float *output, int input_size, float max_amplitude);
/* NFM audio tail (libcsdr.h:106; libcsdr.c:1099-1128, tables predefined.h:56-68): fixed FIR chosen by sample rate
 * (48000, 44100, 11025, 8000); returns the number of outputs = input_size - taps_length, 0 for any other rate */
int   deemphasis_nfm_ff(float *input, float *output, int input_size, int sample_rate);

/* waterfall / audio compression, SURVEY 8(f) rank 4 (ima_adpcm.h:35-41; ima_adpcm.c:95-150): IMA ADPCM, two samples per output byte */
typedef struct ImaState { int index; int previousValue; } ima_adpcm_state_t;
ima_adpcm_state_t encode_ima_adpcm_i16_u8(short *input, unsigned char *output, int input_length, ima_adpcm_state_t state);

/* spectrum side path and shift_unroll, SURVEY 8(f) ranks 3-4 (libcsdr.h:142-149, 199-207; libcsdr.c:1245-1276, 1296-1314, 283-320) */
float *precalculate_window(int size, window_t window);                                   /* host table, malloc'ed like the reference's */
void  apply_window_c(complexf *input, complexf *output, int size, window_t window);
void  apply_precalculated_window_c(complexf *input, complexf *output, int size, float *windowt);
void  logpower_cf(complexf *input, float *output, int size, float add_db);
void  accumulate_power_cf(complexf *input, float *output, int size);
void  log_ff(float *input, float *output, int size, float add_db);
typedef struct shift_unroll_data_s { float *dsin; float *dcos; float phase_increment; int size; } shift_unroll_data_t;
shift_unroll_data_t shift_unroll_init(float rate, int size);
float shift_unroll_cc(complexf *input, complexf *output, int input_size, shift_unroll_data_t *d, float starting_phase);
/* shift_math (libcsdr.h:171; libcsdr.c:186-209): cos/sin of a float phase advanced by one rounded addition per sample, wrapped to [0, 2*PI] */
float shift_math_cc(complexf *input, complexf *output, int input_size, float rate, float starting_phase);
/* shift_table (libcsdr.h:180-186; libcsdr.c:210-260): cos/sin from a quarter-wave table (host memory, made by shift_table_init).  Index
 * arithmetic as the reference's own build executes it; an index the source would read outside the table is clamped. */
typedef struct shift_table_data_s { float *table; int table_size; } shift_table_data_t;
shift_table_data_t shift_table_init(int table_size);
void  shift_table_deinit(shift_table_data_t table_data);
float shift_table_cc(complexf *input, complexf *output, int input_size, float rate, shift_table_data_t table_data, float starting_phase);
/* shift_addfast (libcsdr.h:189-197; libcsdr.c:307-317, 396-433): recursion advanced once per four samples; only input_size/4*4
 * samples of `output` are written, like the reference */
typedef struct shift_addfast_data_s { float dsin[4]; float dcos[4]; float phase_increment; } shift_addfast_data_t;
shift_addfast_data_t shift_addfast_init(float rate);
float shift_addfast_cc(complexf *input, complexf *output, int input_size, shift_addfast_data_t *d, float starting_phase);

/* FFT abstraction (fft_fftw.h:10-27; fft_fftw.c:6-46).  Callers read ->size/->input/->output directly
 * (libcsdr.c:822-835, fastddc.c:112-116), so the first three members keep the reference layout. */
struct fft_plan_s { int size; void *input; void *output; void *plan; };
#define FFT_PLAN_T struct fft_plan_s
FFT_PLAN_T *make_fft_c2c(int size, complexf *input, complexf *output, int forward, int benchmark);
void  fft_execute(FFT_PLAN_T *plan);
void  fft_destroy(FFT_PLAN_T *plan);
void *csdrb_fft_malloc(size_t bytes);            /* stands in for the fft_malloc macro (fft_fftw.h:11) */
void  csdrb_fft_free(void *p);
#define fft_malloc csdrb_fft_malloc
#define fft_free   csdrb_fft_free

/* overlap-add FFT filter step (libcsdr.h:211; libcsdr.c:814-849) */
void apply_fir_fft_cc(FFT_PLAN_T *plan, FFT_PLAN_T *plan_inverse, complexf *taps_fft, complexf *last_overlap, int overlap_size);

/* fastddc (fastddc.h:5-29; fastddc.c:38-166) */
typedef struct fastddc_s {
    int pre_decimation; int post_decimation; int taps_length; int taps_min_length; int overlap_length;
    int fft_size; int fft_inv_size; int input_size; int post_input_size;
    float pre_shift; int startbin; int v; int offsetbin; float post_shift; int output_scrape; int scrap;
    shift_addition_data_t dsadata;
} fastddc_t;
int  fastddc_init(fastddc_t *ddc, float transition_bw, int decimation, float shift_rate);
decimating_shift_addition_status_t fastddc_inv_cc(complexf *input, complexf *output, fastddc_t *ddc,
        FFT_PLAN_T *plan_inverse, complexf *taps_fft, decimating_shift_addition_status_t shift_stat);
void fastddc_print(fastddc_t *ddc, char *source);
void fft_swap_sides(complexf *io, int fft_size);

/* =====================================================================================================
 * Part B -- device-resident bank API
 * =================================================================================================== */

const char *csdrb_last_error(void);
const char *csdrb_version(void);
int  csdrb_device_count(void);                   /* < 0 when the CUDA runtime cannot be initialised */
int  csdrb_set_device(int device);
int  csdrb_stream_synchronize(void *stream);
long csdrb_kernel_launches(void);                /* kernels this library has launched so far (per process) */

/* Device memory and streams for hosts that do not want the CUDA headers (the csdr-bankd daemon is plain C on top of these).
 * Copies are asynchronous on `stream` (NULL = the default stream); host buffers should come from csdrb_host_alloc(). */
void *csdrb_device_alloc(size_t bytes);          /* zero-filled; NULL on failure (csdrb_last_error) */
void  csdrb_device_free(void *d_ptr);
void *csdrb_stream_create(void);
void  csdrb_stream_destroy(void *stream);
int csdrb_copy_h2d(void *d_dst, const void *h_src, size_t bytes, void *stream);
int csdrb_copy_d2h(void *h_dst, const void *d_src, size_t bytes, void *stream);
int csdrb_copy_d2d(void *d_dst, const void *d_src, size_t bytes, void *stream);
int csdrb_copy2d_d2d(void *d_dst, size_t dst_pitch_bytes, const void *d_src, size_t src_pitch_bytes, size_t width_bytes, size_t rows, void *stream);
int csdrb_copy2d_d2h(void *h_dst, size_t dst_pitch_bytes, const void *d_src, size_t src_pitch_bytes, size_t width_bytes, size_t rows, void *stream);
int csdrb_copy2d_h2d(void *d_dst, size_t dst_pitch_bytes, const void *h_src, size_t src_pitch_bytes, size_t width_bytes, size_t rows, void *stream);

/* K1 conversions on device buffers (16-byte aligned) */
int csdrb_convert_u8_f(const unsigned char *d_in, float *d_out, long n, void *stream);
int csdrb_convert_s16_f(const short *d_in, float *d_out, long n, void *stream);
int csdrb_convert_f_s16(const float *d_in, short *d_out, long n, void *stream);

/* K3 fir_decimate_cc bank: channel c reads d_in + c*in_stride (input_size samples) and writes
 * d_out + c*out_stride; all channels share h_taps (HOST pointer, copied into the launch).
 * Returns outputs per channel = input_size >= T ? (input_size - T)/D + 1 : 0  (reference libcsdr.c:537-547).
 * `variant` < 0 lets the library choose the tiling; >= 0 forces one (bench/tuning only). */
int csdrb_fir_decimate_bank_cc(const complexf *d_in, long in_stride, complexf *d_out, long out_stride, int channels,
                               int input_size, int decimation, const float *h_taps, int taps_length, int variant, void *stream);
int csdrb_fir_bank_variants(void);

/* Same operation on HOST buffers (the end-to-end path): streams the bank through the device in chunks of
 * `chunk_channels` channels (<= 0: automatic) on three streams so H2D, kernel and D2H overlap; synchronous.
 * Page-locked host memory (csdrb_host_alloc) is needed for full PCIe rate. */
int csdrb_fir_decimate_bank_cc_host(const complexf *h_in, long in_stride, complexf *h_out, long out_stride, int channels,
                                    int input_size, int decimation, const float *h_taps, int taps_length, int chunk_channels);
/* page-locked host memory placed on the NUMA node of the current device (CSDRB_NO_NUMA=1: wherever the calling thread happens to run) */
void *csdrb_host_alloc(size_t bytes);
void  csdrb_host_free(void *p);

/* convert_u8_f | fir_decimate_cc fused (libcsdr.c:2363-2366 + :528-549; the front of csdr-fm:41): the bank reads rtl_sdr-style interleaved unsigned 8-bit
 * I,Q -- 2 bytes per sample instead of 8 over HBM and PCIe -- and converts on the way into the FIR tile with the reference's own expression, so the
 * result equals convert_u8_f followed by fir_decimate_cc bit for bit in the conversion and like the cf32 bank in the filter.  in_stride counts
 * SAMPLES between rows.  Fused tilings: d=10 T<=200, d=50 T<=900 with in_stride % 8 == 0 and a 16-byte aligned d_in; other geometries convert into a
 * temporary and run the cf32 bank.  The _host form is the end-to-end call (chunked three-stream pipeline, see above). */
int csdrb_fir_decimate_bank_u8_cc(const unsigned char *d_in, long in_stride, complexf *d_out, long out_stride, int channels,
                                  int input_size, int decimation, const float *h_taps, int taps_length, void *stream);
int csdrb_fir_decimate_bank_u8_host(const unsigned char *h_in, long in_stride, complexf *h_out, long out_stride, int channels,
                                    int input_size, int decimation, const float *h_taps, int taps_length, int chunk_channels);

/* K4 fmdemod_quadri_cf bank: d_last_in[c] is the sample preceding channel c's block (NULL = zeros),
 * d_last_out[c] receives its last sample (may be NULL; must not alias d_last_in). */
int csdrb_fmdemod_quadri_bank_cf(const complexf *d_in, long in_stride, float *d_out, long out_stride, int channels,
                                 int input_size, const complexf *d_last_in, complexf *d_last_out, void *stream);

/* K2 shift_addition_cc bank.  in_stride == 0: every channel shifts the SAME wideband input (ddcd use case).
 * d_params[c] = shift_addition_init(rate_c) computed on the host (bit-exact sin/cos deltas); d_phase_io[c] is the
 * starting phase on entry and the phase to continue from on return.  `chunk` is how the reference caller cuts
 * the stream into shift_addition_cc() calls (the CLI uses 1024, csdr.c:911-918; <= 0 means one call).
 * Scratch: csdrb_shift_addition_bank_scratch_bytes() bytes of device memory. */
size_t csdrb_shift_addition_bank_scratch_bytes(int channels, int input_size, int chunk);
int csdrb_shift_addition_bank_cc(const complexf *d_in, long in_stride, complexf *d_out, long out_stride, int channels, int input_size,
                                 const shift_addition_data_t *d_params, float *d_phase_io, int chunk,
                                 void *d_scratch, size_t scratch_bytes, void *stream);
int csdrb_decimating_shift_addition_bank_cc(const complexf *d_in, long in_stride, complexf *d_out, long out_stride, int channels,
                                            int input_size, const shift_addition_data_t *d_params, int decimation,
                                            int *d_remain_io, float *d_phase_io, int *d_out_size, void *stream);

/* Fused shared-input DDC / NFM bank = shift_addition_cc(rate_c) | fir_decimate_cc(D, taps) [| fmdemod_quadri_cf] for `channels`
 * channels of ONE wideband block (ddcd_old.h:51-57 as one launch).  d_params / d_phase_io / chunk as in the shift bank, except
 * that chunks are counted on the absolute stream: the block starts `offset` samples into a chunk and d_phase_io[c] is the phase at
 * the START of that chunk; on return it is the phase at the start of the chunk containing sample n_out*decimation, where the caller
 * must start the next block (re-presenting the unconsumed tail exactly like fir_decimate_cc's callers do, csdr.c:1172-1174).
 * demod = 0: d_out is complexf [channels][out_stride] baseband; demod = 1: float [channels][out_stride] discriminator output,
 * d_last_in/d_last_out carry the previous baseband sample per channel (NULL = zeros / not wanted).
 * Returns outputs per channel, -2 when no fused kernel is compiled for (decimation, taps_length) -- use the unfused bank calls then. */
size_t csdrb_ddc_bank_scratch_bytes(int channels, int input_size, int chunk, int offset);
int csdrb_ddc_bank(const complexf *d_wide, int input_size, int channels, const shift_addition_data_t *d_params, float *d_phase_io,
                   int chunk, int offset, int decimation, const float *h_taps, int taps_length, int demod, void *d_out, long out_stride,
                   const complexf *d_last_in, complexf *d_last_out, void *d_scratch, size_t scratch_bytes, void *stream);

/* Streaming bank object over the fused kernel: owns rates, chunk phases, discriminator history and the position inside the current
 * NCO chunk, so a wideband stream is processed with one call per block; internally the serial phase-chain pre-pass of block k+1 runs on
 * a private stream while the caller's stream executes block k.  Contract of process(): it consumes n_out*decimation samples (the return
 * value is n_out); the next block must start there, i.e. the caller re-presents the unconsumed tail exactly like fir_decimate_cc's
 * callers do (csdr.c:1172-1174).  d_out is float [channels][out_stride] when the bank was created with demod = 1, else complexf. */
typedef struct csdrb_ddc_bank_s csdrb_ddc_bank_t;
csdrb_ddc_bank_t *csdrb_ddc_bank_create(int channels, const float *h_rates, int decimation, const float *h_taps, int taps_length, int demod, int chunk);
void csdrb_ddc_bank_destroy(csdrb_ddc_bank_t *bank);
/* retune one channel; effective from the first sample of the next block.  The phase is continuous across the retune: the bank closes the current NCO
 * chunk at that sample for every channel (a shorter shift_addition_cc call, libcsdr_gpl.c:48-50) and starts a fresh chunk there, like the reference CLI
 * re-initialises between two buffers (csdr.c:897-925) */
int  csdrb_ddc_bank_set_rate(csdrb_ddc_bank_t *bank, int channel, float rate);
int  csdrb_ddc_bank_rechunk(csdrb_ddc_bank_t *bank);                               /* close the NCO chunk at the next block's first sample, rates unchanged */
int  csdrb_ddc_bank_offset(const csdrb_ddc_bank_t *bank);                          /* samples of the current NCO chunk already consumed */
int  csdrb_ddc_bank_process(csdrb_ddc_bank_t *bank, const complexf *d_wide, int input_size, void *d_out, long out_stride, void *stream);

/* The same bank sliced over several GPUs of one node, driven from ONE process (what nmux + one process chain per channel do in the reference,
 * nmux.cpp:246-353, ddcd_old.h:51-57): contiguous channel slices per device, the wideband block goes host -> devices[0] once and on to the other
 * devices by ncclBroadcast (NCCL is loaded with dlopen on first use; a one-device bank never needs it).  submit() only enqueues (H2D, broadcast, slice
 * kernels, D2H of the results) and returns a ticket; collect(ticket) waits for that block.  Up to two blocks may be in flight, so a caller that submits
 * block k+1 before collecting block k has the broadcast of k+1 under the kernels of k.  h_wide / h_out of a submitted block belong to the library until
 * its collect returns (page-locked memory from csdrb_host_alloc for full PCIe rate).  h_out is [channels][out_stride] floats (demod = 1) or complexf.
 * Block contract as for csdrb_ddc_bank_process: a block consumes n_out*decimation samples, the caller re-presents the tail.  devices == NULL: 0..ndev-1. */
typedef struct csdrb_multi_bank_s csdrb_multi_bank_t;
csdrb_multi_bank_t *csdrb_multi_bank_create(int ndev, const int *devices, int channels, const float *h_rates, int decimation, const float *h_taps,
                                            int taps_length, int demod, int chunk, int max_block);
void csdrb_multi_bank_destroy(csdrb_multi_bank_t *bank);
int  csdrb_multi_bank_devices(const csdrb_multi_bank_t *bank);
int  csdrb_multi_bank_slice(const csdrb_multi_bank_t *bank, int index, int *device, int *first_channel, int *channels);
int  csdrb_multi_bank_set_rate(csdrb_multi_bank_t *bank, int channel, float rate);
int  csdrb_multi_bank_submit(csdrb_multi_bank_t *bank, const complexf *h_wide, int input_size, void *h_out, long out_stride);
int  csdrb_multi_bank_collect(csdrb_multi_bank_t *bank, int ticket);
int  csdrb_multi_bank_process_host(csdrb_multi_bank_t *bank, const complexf *h_wide, int input_size, void *h_out, long out_stride);   /* submit + collect */

/* audio tail banks: hard limiter (elementwise) and the 1-pole de-emphasis IIR (d_last_io[c] = previous output of channel c) */
int csdrb_limit_ff(const float *d_in, float *d_out, long n, float max_amplitude, void *stream);
int csdrb_deemphasis_wfm_bank_ff(const float *d_in, long in_stride, float *d_out, long out_stride, int channels, int input_size,
                                 float tau, int sample_rate, float *d_last_io, void *stream);

/* NFM de-emphasis bank: every row through the fixed FIR of `sample_rate`; returns outputs per row (input_size - taps_length),
 * 0 when the rate has no table.  limit_max > 0 fuses the preceding `limit_ff limit_max` of the NFM graph (README.md:87) into the load.
 * csdrb_deemphasis_nfm_taps() exposes the (host) table: NULL / *taps_length = 0 for an unknown rate. */
int csdrb_deemphasis_nfm_bank_ff(const float *d_in, long in_stride, float *d_out, long out_stride, int channels, int input_size,
                                 int sample_rate, float limit_max, void *stream);
const float *csdrb_deemphasis_nfm_taps(int sample_rate, int *taps_length);
/* the same kernel with caller-supplied (host) taps, taps_length <= 208: out[c][i] = sum_t taps[t] * in[c][i+t], i < input_size - taps_length
 * (e.g. a de-emphasis FIR designed for a sample rate the reference has no table for) */
int csdrb_fir_valid_bank_ff(const float *d_in, long in_stride, float *d_out, long out_stride, int channels, int input_size,
                            const float *taps, int taps_length, float limit_max, void *stream);

/* IMA ADPCM on device rows: d_state_io[r] carries row r's encoder state between calls.  csdrb_compress_fft_adpcm_rows_f_u8 is the waterfall line of
 * csdr.c:1745-1767 for many lines at once: each row of fft_size dB values -> (fft_size + 10) / 2 bytes, encoder state fresh per row. */
int csdrb_encode_ima_adpcm_rows_i16_u8(const short *d_in, long in_stride, unsigned char *d_out, long out_stride, int rows, int input_length,
                                       ima_adpcm_state_t *d_state_io, void *stream);
int csdrb_compress_fft_adpcm_rows_f_u8(const float *d_in, long in_stride, unsigned char *d_out, long out_stride, int rows, int fft_size, void *stream);

/* spectrum side path on device buffers: `rows` frames of `size` values share one window table; power modes as the reference's
 * logpower_cf / accumulate_power_cf (d_out is read-modify-write) / log_ff */
int csdrb_apply_window_rows_c(const complexf *d_in, complexf *d_out, const float *d_window, int size, long rows, void *stream);
int csdrb_logpower_cf(const complexf *d_in, float *d_out, long n, float add_db, void *stream);
int csdrb_accumulate_power_cf(const complexf *d_in, float *d_acc, long n, void *stream);
int csdrb_log_ff(const float *d_in, float *d_out, long n, float add_db, void *stream);
/* shift_math_cc bank: d_rates[c] is the plain rate argument; d_phase_io[c] the carried float phase.  The phase chain is sequential over the
 * whole block (one thread per channel walks it), the rotation itself runs fully parallel. */
size_t csdrb_shift_math_bank_scratch_bytes(int channels, int input_size);
int csdrb_shift_math_bank_cc(const complexf *d_in, long in_stride, complexf *d_out, long out_stride, int channels, int input_size,
                             const float *d_rates, float *d_phase_io, void *d_scratch, size_t scratch_bytes, void *stream);
/* shift_table_cc bank: like the shift_math bank plus the quarter-wave table in DEVICE memory (d_table, table_size floats) */
int csdrb_shift_table_bank_cc(const complexf *d_in, long in_stride, complexf *d_out, long out_stride, int channels, int input_size,
                              const float *d_rates, float *d_phase_io, const float *d_table, int table_size, void *d_scratch, size_t scratch_bytes,
                              void *stream);
/* shift_addfast_cc bank: d_params[c] = shift_addfast_init(rate_c); one reference call per `chunk` samples (csdr.c:781-791 uses 1024);
 * scratch as for the shift_addition bank (csdrb_shift_addition_bank_scratch_bytes) */
int csdrb_shift_addfast_bank_cc(const complexf *d_in, long in_stride, complexf *d_out, long out_stride, int channels, int input_size,
                                const shift_addfast_data_t *d_params, float *d_phase_io, int chunk, void *d_scratch, size_t scratch_bytes,
                                void *stream);

/* shift_unroll_cc bank: d_params as for the shift_addition bank (shift_addition_init(rate_c): its .rate is the same 2*rate), tables
 * d_dsin/d_dcos [channels][table_stride] from shift_unroll_init(rate_c, table_size); one reference call per table_size samples */
int csdrb_shift_unroll_bank_cc(const complexf *d_in, long in_stride, complexf *d_out, long out_stride, int channels, int input_size,
                               const shift_addition_data_t *d_params, const float *d_dsin, const float *d_dcos, long table_stride,
                               int table_size, float *d_phase_io, void *d_scratch, size_t scratch_bytes, void *stream);

/* K5 fractional_decimator_ff bank: d_state[c].where carries the reference's `where`; on return input_processed
 * and output_size are filled like fractional_decimator_ff() fills them (libcsdr.c:789-792). */
typedef struct csdrb_fracdec_state_s { float where; int input_processed; int output_size; } csdrb_fracdec_state_t;
size_t csdrb_fractional_decimator_bank_scratch_bytes(int channels, int input_size, float rate);
int csdrb_fractional_decimator_bank_ff(const float *d_in, long in_stride, float *d_out, long out_stride, int channels, int input_size,
                                       float rate, int num_poly_points, const float *d_taps, int taps_length,
                                       csdrb_fracdec_state_t *d_state, void *d_scratch, size_t scratch_bytes, void *stream);

/* K6 fastagc_ff bank: nblocks consecutive blocks of `block` samples per channel; d_hist is [channels][2][block]
 * (the reference's buffer_1, buffer_2; zero it at stream start), d_state[c] = {peak_1, peak_2, last_gain}. */
typedef struct csdrb_fastagc_state_s { float peak_1, peak_2, last_gain; } csdrb_fastagc_state_t;
size_t csdrb_fastagc_bank_scratch_bytes(int channels, int nblocks);
int csdrb_fastagc_bank_ff(const float *d_in, long in_stride, float *d_out, long out_stride, int channels, int block, int nblocks,
                          float reference, csdrb_fastagc_state_t *d_state, float *d_hist, void *d_scratch, size_t scratch_bytes, void *stream);

/* fastagc_ff | convert_f_s16 in one pass: d_out is short [channels][out_stride]; everything else as above */
int csdrb_fastagc_bank_f_s16(const float *d_in, long in_stride, short *d_out, long out_stride, int channels, int block, int nblocks,
                             float reference, csdrb_fastagc_state_t *d_state, float *d_hist, void *d_scratch, size_t scratch_bytes, void *stream);

/* K7 batched unnormalised c2c DFT (power-of-two size 2..16384), sign -1 forward / +1 inverse */
int csdrb_fft_c2c_batch(const complexf *d_in, long in_stride, complexf *d_out, long out_stride, int size, int batch, int inverse, void *stream);

/* K9 overlap-add FFT filter bank = bandpass_fir_fft_cc block loop (csdr.c:1872-1883) for many channels.
 * d_taps_fft: FFT of the zero-padded taps (taps_stride 0 = shared); d_tail_io [channels][fft_size] carries the
 * previous block's tail between calls (zero at stream start). nblocks blocks of input_size samples per channel. */
int csdrb_bandpass_fir_fft_bank_cc(const complexf *d_in, long in_stride, complexf *d_out, long out_stride, int channels, int fft_size,
                                   int input_size, int nblocks, const complexf *d_taps_fft, long taps_stride, complexf *d_tail_io, void *stream);

/* fastddc forward step (csdr.c:2288-2299): nblocks x input_size new samples -> nblocks x fft_size bins;
 * d_overlap_io [fft_size - input_size] carries the overlap between calls (zero at stream start). */
int csdrb_fastddc_fwd_cc(const complexf *d_in, complexf *d_spectra, complexf *d_overlap_io, int fft_size, int input_size, int nblocks, void *stream);

/* K8 fastddc_inv_cc bank: every channel c (its own d_taps_fft + c*fft_size, offsetbin and post-shift NCO) consumes the
 * same nblocks spectra.  d_remain_io/d_phase_io carry decimating_shift_addition_status_t between calls;
 * d_out_total[c] receives the samples written for channel c.  `geometry` is a HOST fastddc_t (fastddc_init). */
typedef struct csdrb_fastddc_chan_s { int offsetbin; float sindelta, cosdelta, rate; } csdrb_fastddc_chan_t;
size_t csdrb_fastddc_inv_bank_scratch_bytes(int channels, int nblocks);
int csdrb_fastddc_inv_bank_cc(const complexf *d_spectra, int nblocks, const complexf *d_taps_fft, const csdrb_fastddc_chan_t *d_chan,
                              int channels, const fastddc_t *geometry, int *d_remain_io, float *d_phase_io, complexf *d_out,
                              long out_stride, int *d_out_total, void *d_scratch, size_t scratch_bytes, void *stream);

/* The same bank as a plan object that OWNS the carried post-shift state (decimating_shift_addition_status_t per channel, fastddc.c:151-165 /
 * libcsdr_gpl.c:154-158) and a fixed nblocks: the data-independent half of run k+1 (state chain, post-shift phasors) is computed on a private
 * stream while run k's IFFT step and the caller's next forward FFT execute, so a run is fold + IFFT only.  Outputs are those of
 * csdrb_fastddc_inv_bank_cc bit for bit.  `chan` is a HOST array; geometries outside the fold path (fft_inv_size 64..1024, even
 * pre-decimation) are refused -- use the stateless call for them.  set_channel retunes one channel from the next run on (replace that
 * channel's row of d_taps_fft on your stream as well); get/set_state read and write the state the NEXT run starts from. */
typedef struct csdrb_fastddc_inv_plan csdrb_fastddc_inv_plan_t;
csdrb_fastddc_inv_plan_t *csdrb_fastddc_inv_plan_create(const csdrb_fastddc_chan_t *chan, int channels, const fastddc_t *geometry, int nblocks);
int csdrb_fastddc_inv_plan_run(csdrb_fastddc_inv_plan_t *plan, const complexf *d_spectra, const complexf *d_taps_fft, complexf *d_out, long out_stride,
                               int *d_out_total, void *stream);
int csdrb_fastddc_inv_plan_set_channel(csdrb_fastddc_inv_plan_t *plan, int channel, const csdrb_fastddc_chan_t *chan);
int csdrb_fastddc_inv_plan_get_state(csdrb_fastddc_inv_plan_t *plan, int *remain, float *phase);
int csdrb_fastddc_inv_plan_set_state(csdrb_fastddc_inv_plan_t *plan, const int *remain, const float *phase);
void csdrb_fastddc_inv_plan_destroy(csdrb_fastddc_inv_plan_t *plan);

#ifdef __cplusplus
}
#endif
#endif
