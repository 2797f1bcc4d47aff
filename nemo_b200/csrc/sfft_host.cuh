// Host side shared by the sfft translation units: the list of lengths with a compile-time plan, and their twiddle tables.
#pragma once
#include "sfft.cuh"
#include <map>
#include <mutex>
#include <vector>

// X(N, R0, R1, R2, NT,  C0, C1, C2, CNT): length; pass radices and threads per transform of the ROW passes (about N / R
// threads, a thread does one butterfly per pass; two row pairs per CTA); pass radices and threads per column of the COLUMN
// passes (about N / 2R threads, two butterflies per pass and at most 16 values per thread: a CTA of 4 columns is then small
// enough in threads and registers for two CTAs per SM, one loading while the other computes)
// (-DSFFT_COL_VARIANT=1 on the nvcc command line selects the one-butterfly-per-thread column plans for experiments; nvcc
// splits option values at commas, so a plan cannot be passed directly)
#ifndef SFFT_COL_VARIANT
#define SFFT_COL_VARIANT 0
#endif
#if SFFT_COL_VARIANT == 1
#define SFFT_COL_1440 10, 12, 12, 144
#define SFFT_COL_840 7, 10, 12, 120
#else
#define SFFT_COL_1440 10, 12, 12, 144
#define SFFT_COL_840 7, 8, 15, 60
#endif
// row plan of the 1440-wide DR5 tile (-DSFFT_ROW_VARIANT=n for experiments): the scatter of a plan's FIRST pass writes with a
// stride of R values between lanes, a 2-way (R = 10) or 4-way (R = 12) bank conflict for doubles unless R is odd
#ifndef SFFT_ROW_VARIANT
#define SFFT_ROW_VARIANT 0
#endif
#if SFFT_ROW_VARIANT == 1
#define SFFT_ROW_1440 15, 8, 12, 120
#elif SFFT_ROW_VARIANT == 2
#define SFFT_ROW_1440 9, 16, 10, 160
#elif SFFT_ROW_VARIANT == 3
#define SFFT_ROW_1440 12, 10, 12, 144
#elif SFFT_ROW_VARIANT == 4
#define SFFT_ROW_1440 9, 10, 16, 160
#else
#define SFFT_ROW_1440 10, 12, 12, 144
#endif
// the inverse row pass runs its plan in one direction only, so its first radix can be odd (15: conflict-free scatter) at 120
// threads per row pair: 1.40 -> 1.09 ms per two launches of 128 images of 840 x 1440 (the same plan spills in the window kernel,
// whose reversed plan would start with radix 12 anyway: 1.54 -> 2.07 ms; 9 x 16 x 10 and 12 x 10 x 12 gain nothing)
#ifndef SFFT_ROW_INV_VARIANT
#define SFFT_ROW_INV_VARIANT 1
#endif
#if SFFT_ROW_INV_VARIANT == 1
#define SFFT_ROW_INV_1440 15, 8, 12, 120
#elif SFFT_ROW_INV_VARIANT == 2
#define SFFT_ROW_INV_1440 15, 12, 8, 120
#elif SFFT_ROW_INV_VARIANT == 3
#define SFFT_ROW_INV_1440 15, 8, 12, 180
#elif SFFT_ROW_INV_VARIANT == 4
#define SFFT_ROW_INV_1440 15, 6, 16, 120
#else
#define SFFT_ROW_INV_1440 SFFT_ROW_1440
#endif
// the window kernel (forward, multiply, reversed plan backwards) fed by the copy unit wants NO register cap: 15 x 8 x 12 at 120
// threads, two CTAs per SM, 254 registers, nothing spilled: 1.00 ms per 128 images of 840 x 1440; the same plan at three / four
// CTAs per SM (168 / 128 registers, 184 / 384 bytes spilled) 1.27 / 2.11 ms; 10 x 12 x 12 at 144 threads 1.39 (two CTAs) and 1.52
// (three); register-staged 1.55
#ifndef SFFT_ROW_WIN_VARIANT
#define SFFT_ROW_WIN_VARIANT 1
#endif
#if SFFT_ROW_WIN_VARIANT == 1
#define SFFT_ROW_WIN_1440 15, 8, 12, 120
#elif SFFT_ROW_WIN_VARIANT == 2
#define SFFT_ROW_WIN_1440 12, 8, 15, 120
#else
#define SFFT_ROW_WIN_1440 SFFT_ROW_1440
#endif
#define SFFT_APPLY(X, ...) X(__VA_ARGS__)
#define SFFT_APPLY2(X, N, A, B) SFFT_APPLY(X, N, A, B)
#define SFFT_FOR_SIZES_WIN(X) \
    SFFT_APPLY2(X, 1440, SFFT_ROW_WIN_1440,  SFFT_COL_1440) \
    SFFT_APPLY(X, 840, 7, 10, 12, 120,    SFFT_COL_840)  \
    X(420, 5, 7, 12, 84,      5, 7, 12, 42)    \
    X(320, 5, 8, 8, 64,       5, 8, 8, 32)     \
    X(300, 5, 6, 10, 60,      5, 6, 10, 30)
#define SFFT_FOR_SIZES_INV(X) \
    SFFT_APPLY2(X, 1440, SFFT_ROW_INV_1440,  SFFT_COL_1440) \
    SFFT_APPLY(X, 840, 7, 10, 12, 120,    SFFT_COL_840)  \
    X(420, 5, 7, 12, 84,      5, 7, 12, 42)    \
    X(320, 5, 8, 8, 64,       5, 8, 8, 32)     \
    X(300, 5, 6, 10, 60,      5, 6, 10, 30)
#define SFFT_FOR_SIZES(X) \
    SFFT_APPLY2(X, 1440, SFFT_ROW_1440,  SFFT_COL_1440) \
    SFFT_APPLY(X, 840, 7, 10, 12, 120,    SFFT_COL_840)  \
    X(420, 5, 7, 12, 84,      5, 7, 12, 42)    \
    X(320, 5, 8, 8, 64,       5, 8, 8, 32)     \
    X(300, 5, 6, 10, 60,      5, 6, 10, 30)

// knobs for experiments (tools/build_variant.sh)
#ifndef SFFT_ROW_G
#define SFFT_ROW_G 2          // row pairs per CTA
#endif
#ifndef SFFT_ROW_MINB
#define SFFT_ROW_MINB 2       // resident CTAs per SM asked of ptxas
#endif
#ifndef SFFT_COL_C
#define SFFT_COL_C 4          // adjacent columns per CTA
#endif
#ifndef SFFT_COL_MINB
#define SFFT_COL_MINB 2
#endif

// pass twiddles of a plan (layout: sfft::Plan::twOffset), cached per (precision, device, radices)
int nb_sfft_twiddles(int prec, int R0, int R1, int R2, const void** out);

template <class PL> static inline int sfft_tw(int prec, const void** out) {
    return nb_sfft_twiddles(prec, PL::radix(0), PL::radix(1), PL::NP > 2 ? PL::radix(2) : 1, out);
}

template <typename K> static inline int sfft_set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) NB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return NB_OK;
}
