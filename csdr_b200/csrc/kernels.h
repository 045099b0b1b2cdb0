// kernels.h -- internal launcher prototypes (host side of each .cu).  Not part of the public C ABI.
#pragma once
#include <cuda_runtime.h>

namespace csdrb {

// K3 fir_decimate.cu
int launch_fir_decimate_bank(const float2* d_in, long in_stride, float2* d_out, long out_stride, int channels, int n_in,
                             int D, const float* h_taps, const float* d_taps, long taps_stride, int T, int variant,
                             cudaStream_t st);
int fir_bank_variant_count();
int launch_u8_rows_to_cf32(const unsigned char* d_in, long in_stride, float2* d_out, long out_stride, int channels, int n, cudaStream_t st);
int launch_fir_decimate_bank_u8(const unsigned char* d_in, long in_stride, float2* d_out, long out_stride, int channels, int n_in,
                                int D, const float* h_taps, int T, cudaStream_t st);

// K1/K4 elementwise.cu
int launch_convert_u8_f(const unsigned char* d_in, float* d_out, long n, cudaStream_t st);
int launch_convert_s16_f(const short* d_in, float* d_out, long n, cudaStream_t st);
int launch_convert_f_s16(const float* d_in, short* d_out, long n, cudaStream_t st);
int launch_fmdemod_quadri_bank(const float2* d_in, long in_stride, float* d_out, long out_stride, int channels, int n,
                               const float2* d_last_in, float2* d_last_out, cudaStream_t st);

int launch_adpcm_encode_rows(const short* d_in, long in_stride, unsigned char* d_out, long out_stride, int rows, int n, void* d_state_io, cudaStream_t st);
int launch_compress_fft_adpcm_rows(const float* d_in, long in_stride, unsigned char* d_out, long out_stride, int rows, int fft_size, cudaStream_t st);
int launch_limit_ff(const float* d_in, float* d_out, long n, float max_amplitude, cudaStream_t st);
int launch_deemphasis_wfm_bank(const float* d_in, long in_stride, float* d_out, long out_stride, int channels, int n, float tau, int sample_rate,
                               float* d_last_io, cudaStream_t st);

// deemphasis_nfm_ff: the tap tables live in host/firdes.c (public accessor, see include/csdr_b200.h)
constexpr int kNfmMaxTaps = 208;
extern "C" const float* csdrb_deemphasis_nfm_taps(int sample_rate, int* taps_length);
int launch_fir_valid_bank(const float* d_in, long in_stride, float* d_out, long out_stride, int channels, int n, const float* h_taps, int T,
                          float limit_max, cudaStream_t st);
int launch_deemphasis_nfm_bank(const float* d_in, long in_stride, float* d_out, long out_stride, int channels, int n, int sample_rate,
                               float limit_max, cudaStream_t st);

int launch_apply_window_rows(const float2* d_in, float2* d_out, const float* d_window, int size, long rows, cudaStream_t st);
int launch_power(const float2* d_in_c, const float* d_in_f, float* d_out, long n, float add_db, int mode, cudaStream_t st);
int launch_shift_unroll_bank(const float2* d_in, long in_stride, float2* d_out, long out_stride, int channels, int n,
                             const float* d_params, const float* d_dsin, const float* d_dcos, long table_stride, int table_size,
                             float* d_phase_io, void* d_scratch, size_t scratch_bytes, cudaStream_t st);

void shift_unroll_bank_single(const float2* d_in, float2* d_out, int n, const float* d_dsin, const float* d_dcos, const float* d_phase, cudaStream_t st);

// K2 shift.cu
size_t shift_bank_scratch_bytes(int channels, int n, int chunk);
int launch_shift_addition_bank(const float2* d_in, long in_stride, float2* d_out, long out_stride, int channels, int n,
                               const float* d_params, float* d_phase_io, int chunk, void* d_scratch, size_t scratch_bytes, cudaStream_t st);
int launch_shift_addfast_bank(const float2* d_in, long in_stride, float2* d_out, long out_stride, int channels, int n,
                              const float* d_params, float* d_phase_io, int chunk, void* d_scratch, size_t scratch_bytes, cudaStream_t st);
size_t shift_math_scratch_bytes(int channels, int n);
int launch_shift_math_bank(const float2* d_in, long in_stride, float2* d_out, long out_stride, int channels, int n, const float* d_rates,
                           float* d_phase_io, void* d_scratch, size_t scratch_bytes, cudaStream_t st);
int launch_shift_table_bank(const float2* d_in, long in_stride, float2* d_out, long out_stride, int channels, int n, const float* d_rates,
                            float* d_phase_io, const float* d_table, int table_size, void* d_scratch, size_t scratch_bytes, cudaStream_t st);
int launch_decimating_shift_bank(const float2* d_in, long in_stride, float2* d_out, long out_stride, int channels, int n,
                                 const float* d_params, int decimation, int* d_remain_io, float* d_phase_io, int* d_out_size, cudaStream_t st);

// K5/K6 audio.cu
size_t fracdec_scratch_bytes(int channels, int n, float rate);
int launch_fractional_decimator_bank(const float* d_in, long in_stride, float* d_out, long out_stride, int channels, int n,
                                     float rate, int num_poly_points, const float* d_taps, int taps_length, void* d_state,
                                     void* d_scratch, size_t scratch_bytes, cudaStream_t st);
size_t fastagc_scratch_bytes(int channels, int nblocks);
int launch_fastagc_bank_s16(const float* d_in, long in_stride, short* d_out, long out_stride, int channels, int block, int nblocks,
                            float reference, void* d_state, float* d_hist, void* d_scratch, size_t scratch_bytes, cudaStream_t st);
int launch_fastagc_bank(const float* d_in, long in_stride, float* d_out, long out_stride, int channels, int block, int nblocks,
                        float reference, void* d_state, float* d_hist, void* d_scratch, size_t scratch_bytes, cudaStream_t st);

// K7/K8/K9 fft.cu
int launch_fft_c2c_batch(const float2* d_in, long in_stride, float2* d_out, long out_stride, int n, int batch, int inverse, cudaStream_t st);
int launch_olafir_bank(const float2* d_in, long in_stride, float2* d_out, long out_stride, int channels, int fft_size, int input_size,
                       int nblocks, const float2* d_taps_fft, long taps_stride, float2* d_tail_io, int blocks_per_cta, cudaStream_t st);
int launch_apply_fir_fft(const float2* d_in, const float2* d_taps_fft, const float2* d_last_overlap, int overlap_size, float2* d_out,
                         int fft_size, cudaStream_t st);
int launch_fastddc_fwd(const float2* d_in, float2* d_spectra, float2* d_overlap_io, int fft_size, int input_size, int nblocks, cudaStream_t st);
size_t fastddc_inv_scratch_bytes(int channels, int nblocks);
int launch_fastddc_inv_bank(const float2* d_spectra, int nblocks, const float2* d_taps_fft, const void* d_chan, int channels,
                            int fft_size, int fft_inv_size, int pre_decimation, int scrap, int post_input_size, int post_decimation,
                            int* d_remain_io, float* d_phase_io, float2* d_out, long out_stride, int* d_out_total,
                            void* d_scratch, size_t scratch_bytes, cudaStream_t st);

// fastddc inverse bank with look-ahead (fft.cu): a plan owns the post-shift state and prepares the next run's chain + phasors during the current one
int fastddc_inv_plan_create(void** out_plan, const void* h_chan, int channels, int nblocks, int fft_size, int fft_inv_size, int pre_decimation, int scrap,
                            int post_input_size, int post_decimation);
int fastddc_inv_plan_run(void* plan, const float2* d_spectra, const float2* d_taps_fft, float2* d_out, long out_stride, int* d_out_total, cudaStream_t st);
int fastddc_inv_plan_set_channel(void* plan, int c, const void* h_chan_one);
int fastddc_inv_plan_get_state(void* plan, int* h_remain, float* h_phase);
int fastddc_inv_plan_set_state(void* plan, const int* h_remain, const float* h_phase);
void fastddc_inv_plan_destroy(void* plan);

// fused shared-input DDC bank, ddc_bank.cu
size_t ddc_bank_scratch_bytes(int channels, int input_size, int chunk, int offset);
size_t ddc_bank_tables_bytes(int channels);                            // persistent phase-wrap tables of a bank (phase_table.cuh), one per channel
int launch_ddc_rechunk(int channels, const float* d_params, float* d_phase_io, int n, cudaStream_t st);
int launch_ddc_tables(int channels, const float* d_params, int chunk, void* d_tables, cudaStream_t st);
int launch_ddc_prepass(int input_size, int channels, const float* d_params, float* d_phase_io, int chunk, int offset, int decimation,
                       int taps_length, void* d_scratch, size_t scratch_bytes, const void* d_tables, cudaStream_t st);   // d_tables NULL: built per call in d_scratch
int launch_ddc_main(const float2* d_wide, int input_size, int channels, const float* d_params, int chunk, int offset, int decimation,
                    const float* h_taps, int taps_length, int demod, void* d_out, long out_stride, const float2* d_last_in, float2* d_last_out,
                    const void* d_scratch, cudaStream_t st);
int launch_ddc_bank(const float2* d_wide, int input_size, int channels, const float* d_params, float* d_phase_io, int chunk, int offset,
                    int decimation, const float* h_taps, int taps_length, int demod, void* d_out, long out_stride,
                    const float2* d_last_in, float2* d_last_out, void* d_scratch, size_t scratch_bytes, int* launches, cudaStream_t st);

}  // namespace csdrb
