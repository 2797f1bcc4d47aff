// Internal declarations shared between the .cu translation units (not part of the public C ABI).
#pragma once
#include "common.cuh"
#include "../../include/nemo_b200.h"
#include <vector>

// pass twiddle tables of the shared-memory FFT (layout: fft.cuh pass_tw_offset), in long double (fft_kernels.cu)
void nb_fill_pass_twiddles(int N, int V, std::vector<long double>& re, std::vector<long double>& im);

// Bluestein passes for sides that are not powers of two (fft_bluestein.cu); same contracts as the power-of-two row / column
// passes of fft_kernels.cu
int nb_bs_length(int N);
int nb_bs_rows_fwd(int prec, const void* in, void* out, int B, int H, int W, int apodPix, double scale, cudaStream_t st);
int nb_bs_rows_inv(int prec, const void* in, void* out, int B, int H, int W, double scale, const double* norm, const unsigned char* mask,
                   int maskDiv, const void* zeroRef, cudaStream_t st);
int nb_bs_cols(int prec, const void* in, void* out, int B, int H, int Wc, int dir, cudaStream_t st);
int nb_bs_cols_apply(int prec, const void* fmaps, const void* filt, void* out, int nOut, int S, int perScale, int F, int H, int Wc,
                     cudaStream_t st);
int nb_bs_cols_window(int prec, void* data, int B, int H, int Wc, const void* wy, const void* wx, cudaStream_t st);

// compile-time mixed-radix passes for the 2-3-5-7-smooth sides listed in sfft_host.cuh (sfft_rows.cu, sfft_cols_f64/f32.cu);
// same contracts again
int nb_sfft_supported(int n);
int nb_sfft_rows_fwd(int prec, const void* in, void* out, int B, int H, int W, int apodPix, double scale, cudaStream_t st);
int nb_sfft_rows_inv(int prec, const void* in, void* out, int B, int H, int W, double scale, const double* norm, const unsigned char* mask,
                     int maskDiv, const void* zeroRef, int* zeroFlag, cudaStream_t st);
// TMA-fed column passes (sfft_cols_tma.cu, fp64): NB_OK when launched, 1 when the path does not apply (fall back), < 0 on error
int nb_sfft_cols_tma_f64(const void* in, void* out, int B, int H, int Wc, int dir, cudaStream_t st);
int nb_sfft_cols_window_tma_f64(void* data, int B, int H, int Wc, const void* wy, const void* wx, cudaStream_t st);
int nb_sfft_cols_apply_tma_f64(const void* fmaps, const void* filt, void* out, int nOut, int S, int perScale, int F, int H, int Wc,
                               cudaStream_t st);
int nb_sfft_cols_f64(const void* in, void* out, int B, int H, int Wc, int dir, cudaStream_t st);
int nb_sfft_cols_f32(const void* in, void* out, int B, int H, int Wc, int dir, cudaStream_t st);
int nb_sfft_cols_apply_f64(const void* fmaps, const void* filt, void* out, int nOut, int S, int perScale, int F, int H, int Wc, cudaStream_t st);
int nb_sfft_cols_apply_f32(const void* fmaps, const void* filt, void* out, int nOut, int S, int perScale, int F, int H, int Wc, cudaStream_t st);
int nb_sfft_cols_window_f64(void* data, int B, int H, int Wc, const void* wy, const void* wx, cudaStream_t st);
int nb_sfft_cols_window_f32(void* data, int B, int H, int Wc, const void* wy, const void* wx, cudaStream_t st);
// out of place, wx may be null (window in y only): the column half of the separable window undo
int nb_sfft_cols_window_io_f64(const void* in, void* out, int B, int H, int Wc, const void* wy, const void* wx, cudaStream_t st);
int nb_sfft_cols_window_io_f32(const void* in, void* out, int B, int H, int Wc, const void* wy, const void* wx, cudaStream_t st);
// real rows -> real rows: forward transform, times wx[min(k, W - k)] * scale, inverse, in one kernel (the row half of it)
int nb_sfft_rows_window(int prec, const void* in, void* out, int B, int H, int W, const void* wx, double scale, int restoreZeros, cudaStream_t st);
