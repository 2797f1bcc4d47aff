// Device core shared by the any-size FFT passes (fft_bluestein.cu: tables and row kernels; fft_bluestein_cols.cu: column
// kernels): the mixed-radix transform for 2-3-5-7-smooth lengths and Bluestein's chirp transform for every other length, both
// on the thread layout of fft.cuh (M / 16 threads per line, 16 values per thread).
#pragma once
#include "fft.cuh"
#include "nemo_b200_internal.h"
#include <math.h>

using namespace nbfft;

struct BsTables {
    const void* tw;
    const void* chirp;
    const void* bhat;
    int M;
};

// tables of a length-N transform (cached per precision, device and N); fft_bluestein.cu
int nb_get_bs_tables(int prec, int N, BsTables* out);

// ------------------------------------------------------------------------------------------------------
// Lengths whose only prime factors are 2, 3, 5 and 7 (1440 x 840 survey tiles: 2^5 3^2 5 and 2^3 3 5 7) do not need the chirp
// detour: a mixed-radix Stockham transform of length N itself, passes of radix 7 / 5 / 9 / 3 / 8 / 4 / 2 through the group's shared
// memory (two regions of N values inside the buffer Bluestein's M >= 2 N - 1 points need, so the kernels and their launch
// geometry are shared), about a sixth of the arithmetic.  Pass with sub-transform length Ns and radix R, butterfly j < N / R,
// k = j mod Ns: inputs x[j + r N / R] times exp(-2 pi i r k / (Ns R)), a natural-order DFT of size R, outputs to
// (j - k) R + k + r Ns.  Twiddles per pass as a block [R - 1][Ns]; the plan (number of passes, then radix and table offset
// of each) sits in device memory.
template <int R> __device__ __forceinline__ constexpr double mix_cos(int m) {
    return R == 3 ? (m == 0 ? 1.0 : -0.5)
         : R == 5 ? (m == 0 ? 1.0 : (m == 1 || m == 4) ? 0.30901699437494745 : -0.8090169943749475)
                  : (m == 0 ? 1.0 : (m == 1 || m == 6) ? 0.6234898018587335 : (m == 2 || m == 5) ? -0.2225209339563144 : -0.9009688679024191);
}
template <int R> __device__ __forceinline__ constexpr double mix_sin(int m) {
    return R == 3 ? (m == 0 ? 0.0 : m == 1 ? 0.8660254037844386 : -0.8660254037844386)
         : R == 5 ? (m == 0 ? 0.0 : m == 1 ? 0.9510565162951535 : m == 2 ? 0.5877852522924731 : m == 3 ? -0.5877852522924731 : -0.9510565162951535)
                  : (m == 0 ? 0.0 : m == 1 ? 0.7818314824680298 : m == 2 ? 0.9749279121818236 : m == 3 ? 0.4338837391175581
                     : m == 4 ? -0.4338837391175581 : m == 5 ? -0.9749279121818236 : -0.7818314824680298);
}

// natural-order DFT of R values in registers, DIR = -1 forward, +1 inverse (no scaling)
template <typename T, int R, int DIR>
__device__ __forceinline__ void dft_small(T (&xr)[R], T (&xi)[R]) {
    if constexpr (R == 2) {
        const T ar = xr[0], ai = xi[0];
        xr[0] = ar + xr[1]; xi[0] = ai + xi[1];
        xr[1] = ar - xr[1]; xi[1] = ai - xi[1];
    } else if constexpr (R == 4) {
        const T s02r = xr[0] + xr[2], s02i = xi[0] + xi[2], d02r = xr[0] - xr[2], d02i = xi[0] - xi[2];
        const T s13r = xr[1] + xr[3], s13i = xi[1] + xi[3], d13r = xr[1] - xr[3], d13i = xi[1] - xi[3];
        xr[0] = s02r + s13r; xi[0] = s02i + s13i;
        xr[2] = s02r - s13r; xi[2] = s02i - s13i;
        // forward: X1 = d02 - i d13, X3 = d02 + i d13; inverse the other way round
        if (DIR < 0) {
            xr[1] = d02r + d13i; xi[1] = d02i - d13r;
            xr[3] = d02r - d13i; xi[3] = d02i + d13r;
        } else {
            xr[1] = d02r - d13i; xi[1] = d02i + d13r;
            xr[3] = d02r + d13i; xi[3] = d02i - d13r;
        }
    } else {
        constexpr int Hh = (R - 1) / 2;
        T spr[Hh], spi[Hh], smr[Hh], smi[Hh];
#pragma unroll
        for (int r = 1; r <= Hh; ++r) {
            spr[r - 1] = xr[r] + xr[R - r]; spi[r - 1] = xi[r] + xi[R - r];
            smr[r - 1] = xr[r] - xr[R - r]; smi[r - 1] = xi[r] - xi[R - r];
        }
        const T x0r = xr[0], x0i = xi[0];
        T t0r = x0r, t0i = x0i;
#pragma unroll
        for (int r = 0; r < Hh; ++r) {
            t0r += spr[r];
            t0i += spi[r];
        }
        xr[0] = t0r; xi[0] = t0i;
#pragma unroll
        for (int k = 1; k <= Hh; ++k) {
            T ar = x0r, ai = x0i, br = 0, bi = 0;
#pragma unroll
            for (int r = 1; r <= Hh; ++r) {
                const T c = (T)mix_cos<R>((r * k) % R), sn = (T)mix_sin<R>((r * k) % R);
                ar += c * spr[r - 1]; ai += c * spi[r - 1];
                br += sn * smr[r - 1]; bi += sn * smi[r - 1];
            }
            // forward: X[k] = A - i B, X[R - k] = A + i B; inverse the other way round
            if (DIR < 0) {
                xr[k] = ar + bi; xi[k] = ai - br;
                xr[R - k] = ar - bi; xi[R - k] = ai + br;
            } else {
                xr[k] = ar - bi; xi[k] = ai + br;
                xr[R - k] = ar + bi; xi[R - k] = ai - br;
            }
        }
    }
}

// composite sizes, built in registers from the small ones: 8 and 16 by one decimation-in-time step over two half-size
// transforms (X[k] = E[k] + w^k O[k], X[k + R/2] = E[k] - w^k O[k]), 9 as 3 x 3 with the w9^(n2 k1) twiddles between
template <typename T, int R, int DIR>
__device__ __forceinline__ void dft_comp(T (&xr)[R], T (&xi)[R]) {
    if constexpr (R == 8 || R == 16) {
        constexpr int Hh = R / 2;
        T er[Hh], ei[Hh], orr[Hh], oi[Hh];
#pragma unroll
        for (int i = 0; i < Hh; ++i) {
            er[i] = xr[2 * i]; ei[i] = xi[2 * i];
            orr[i] = xr[2 * i + 1]; oi[i] = xi[2 * i + 1];
        }
        if constexpr (Hh == 4) {
            dft_small<T, 4, DIR>(er, ei);
            dft_small<T, 4, DIR>(orr, oi);
        } else {
            dft_comp<T, Hh, DIR>(er, ei);
            dft_comp<T, Hh, DIR>(orr, oi);
        }
#pragma unroll
        for (int k = 0; k < Hh; ++k) {
            const T c = nbfft::cos16<T>(k * (16 / R)), sn = nbfft::sin16<T>(k * (16 / R));
            const T wi = (DIR < 0) ? -sn : sn;
            const T pr = orr[k] * c - oi[k] * wi, pi = orr[k] * wi + oi[k] * c;
            xr[k] = er[k] + pr; xi[k] = ei[k] + pi;
            xr[k + Hh] = er[k] - pr; xi[k + Hh] = ei[k] - pi;
        }
    } else {   // R == 9
        T yr[3][3], yi[3][3];
#pragma unroll
        for (int n2 = 0; n2 < 3; ++n2) {
            T ar[3] = {xr[n2], xr[3 + n2], xr[6 + n2]}, ai[3] = {xi[n2], xi[3 + n2], xi[6 + n2]};
            dft_small<T, 3, DIR>(ar, ai);
#pragma unroll
            for (int k1 = 0; k1 < 3; ++k1) {
                const int e = n2 * k1;      // w9^e, e in {0, 1, 2, 4}
                if (e == 0) {
                    yr[n2][k1] = ar[k1]; yi[n2][k1] = ai[k1];
                } else {
                    const T c = (T)(e == 1 ? 0.766044443118978 : e == 2 ? 0.17364817766693036 : -0.9396926207859084);
                    const T sn = (T)(e == 1 ? 0.6427876096865394 : e == 2 ? 0.984807753012208 : 0.3420201433256687);
                    const T wi = (DIR < 0) ? -sn : sn;
                    yr[n2][k1] = ar[k1] * c - ai[k1] * wi;
                    yi[n2][k1] = ar[k1] * wi + ai[k1] * c;
                }
            }
        }
#pragma unroll
        for (int k1 = 0; k1 < 3; ++k1) {
            T ar[3] = {yr[0][k1], yr[1][k1], yr[2][k1]}, ai[3] = {yi[0][k1], yi[1][k1], yi[2][k1]};
            dft_small<T, 3, DIR>(ar, ai);
            xr[k1] = ar[0]; xi[k1] = ai[0];
            xr[k1 + 3] = ar[1]; xi[k1 + 3] = ai[1];
            xr[k1 + 6] = ar[2]; xi[k1 + 6] = ai[2];
        }
    }
}

template <typename T, int R, int DIR>
__device__ __forceinline__ void mix_pass(const T* __restrict__ aR, const T* __restrict__ aI, T* __restrict__ bR, T* __restrict__ bI,
                                         const typename V2<T>::type* __restrict__ tw, const int N, const int Ns, const int t, const int NT) {
    typedef typename V2<T>::type C2;
    const int nb = N / R;
    for (int j = t; j < nb; j += NT) {
        const int k = j % Ns;
        T xr[R], xi[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            xr[r] = aR[j + r * nb];
            xi[r] = aI[j + r * nb];
        }
        if (Ns > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) {
                const C2 w = __ldg(tw + (r - 1) * Ns + k);
                const T wr = w.x, wi = (DIR < 0) ? w.y : -w.y;
                const T ur = xr[r], ui = xi[r];
                xr[r] = ur * wr - ui * wi;
                xi[r] = ur * wi + ui * wr;
            }
        }
        if constexpr (R == 8 || R == 9 || R == 16) dft_comp<T, R, DIR>(xr, xi);
        else dft_small<T, R, DIR>(xr, xi);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            bR[j0 + r * Ns] = xr[r];
            bI[j0 + r * Ns] = xi[r];
        }
    }
}

// same contract as bluestein_run below
template <typename T, int M, int DIR, class Sync>
__device__ __forceinline__ void mixed_run(T (&re)[16], T (&im)[16], const int t, T* s_re, T* s_im,
                                          const typename V2<T>::type* __restrict__ mtw, const int* __restrict__ plan, const int N,
                                          Sync sync) {
    constexpr int NT = M / 16;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int n = t + q * NT;
        if (n < N) {
            s_re[n] = re[q];
            s_im[n] = im[q];
        }
    }
    sync();
    T *aR = s_re, *aI = s_im, *bR = s_re + N, *bI = s_im + N;
    const int nPass = plan[0];
    int Ns = 1;
    for (int p = 0; p < nPass; ++p) {
        const int R = plan[1 + 2 * p];
        const typename V2<T>::type* tw = mtw + plan[2 + 2 * p];
        switch (R) {
            case 2: mix_pass<T, 2, DIR>(aR, aI, bR, bI, tw, N, Ns, t, NT); break;
            case 3: mix_pass<T, 3, DIR>(aR, aI, bR, bI, tw, N, Ns, t, NT); break;
            case 4: mix_pass<T, 4, DIR>(aR, aI, bR, bI, tw, N, Ns, t, NT); break;
            case 5: mix_pass<T, 5, DIR>(aR, aI, bR, bI, tw, N, Ns, t, NT); break;
            case 7: mix_pass<T, 7, DIR>(aR, aI, bR, bI, tw, N, Ns, t, NT); break;
            case 8: mix_pass<T, 8, DIR>(aR, aI, bR, bI, tw, N, Ns, t, NT); break;
            default: mix_pass<T, 9, DIR>(aR, aI, bR, bI, tw, N, Ns, t, NT); break;
        }
        sync();
        T* x = aR; aR = bR; bR = x;
        x = aI; aI = bI; bI = x;
        Ns *= R;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int n = t + q * NT;
        if (n < N) {
            re[q] = aR[n];
            im[q] = aI[n];
        }
    }
    sync();    // the caller may reuse the buffer straight away
}

// ------------------------------------------------------------------------------------------------------
// The transform of one line held by a thread group: on entry re/im[q] = x[t + q * NT] for indices < N, 0 beyond; on exit
// re/im[q] = X[t + q * NT] for indices < N (undefined beyond).  DIR = -1: X[k] = sum x[n] exp(-2 pi i n k / N);
// DIR = +1: the un-normalised inverse sum.  s_re / s_im: npad(M) words each, private to the group.
template <typename T, int M, int DIR, class Sync>
__device__ __forceinline__ void bluestein_run(T (&re)[16], T (&im)[16], const int t, T* s_re, T* s_im,
                                              const typename V2<T>::type* __restrict__ tw,
                                              const typename V2<T>::type* __restrict__ chirp,
                                              const typename V2<T>::type* __restrict__ bhat, const int N, Sync sync) {
    typedef typename V2<T>::type C2;
    constexpr int NT = M / 16;
    if (bhat == nullptr) {   // a 2-3-5-7-smooth length: tw holds the mixed-radix pass tables, chirp points to the plan
        mixed_run<T, M, DIR>(re, im, t, s_re, s_im, tw, reinterpret_cast<const int*>(chirp), N, sync);
        return;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int n = t + q * NT;
        if (n < N) {
            const C2 c = __ldg(chirp + n);
            const T cr = c.x, ci = (DIR < 0) ? c.y : -c.y;
            const T xr = re[q], xi = im[q];
            re[q] = xr * cr - xi * ci;
            im[q] = xr * ci + xi * cr;
        } else {
            re[q] = 0;
            im[q] = 0;
        }
    }
    fft_run<T, M, -1>(re, im, t, s_re, s_im, tw, sync);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const C2 b = __ldg(bhat + t + q * NT);   // FFT of conj(b) is conj(bhat): b is even
        const T br = b.x, bi = (DIR < 0) ? b.y : -b.y;
        const T xr = re[q], xi = im[q];
        re[q] = xr * br - xi * bi;
        im[q] = xr * bi + xi * br;
    }
    sync();
    fft_run<T, M, +1>(re, im, t, s_re, s_im, tw, sync);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int k = t + q * NT;
        if (k < N) {
            const C2 c = __ldg(chirp + k);
            const T cr = c.x, ci = (DIR < 0) ? c.y : -c.y;
            const T xr = re[q], xi = im[q];
            re[q] = xr * cr - xi * ci;
            im[q] = xr * ci + xi * cr;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// host dispatch
template <typename K> static int bs_set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) NB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return NB_OK;
}

#define BS_DISPATCH_M(m, MACRO)                                                                  \
    switch (m) {                                                                                 \
        case 128: MACRO(128); break;                                                             \
        case 256: MACRO(256); break;                                                             \
        case 512: MACRO(512); break;                                                             \
        case 1024: MACRO(1024); break;                                                           \
        case 2048: MACRO(2048); break;                                                           \
        case 4096: MACRO(4096); break;                                                           \
        case 8192: MACRO(8192); break;                                                           \
        default: nb_set_error("Bluestein length %d not supported", m); return NB_ERR_UNSUPPORTED; \
    }
