// Mixed-radix Stockham transforms with COMPILE-TIME plans, for tile sides whose prime factors are 2, 3, 5, 7 (the 1440 x 840
// tiles of the ACT DR5 tiling, 300 x 420, ...): K1 / K4 / K8 on such tiles.  fft_anysize.cuh runs the same mathematics from a
// run-time plan on the thread layout of Bluestein's transform (M / 16 threads for M >= 2 N); here the length, the radices and
// the thread count are template parameters, so that
//   * a transform of N = R0 R1 R2 points takes three passes whose butterflies (radix <= 16, composite radices by the
//     prime-factor map, no internal twiddles) are fully unrolled in registers,
//   * NT threads carry one transform, thread t doing butterflies j = t, t + NT, ... < N / R of every pass: the thread count
//     follows N (144 threads for N = 1440), not the padded length,
//   * a thread's share of the data stays in registers from pass to pass; shared memory is only the exchange medium (ONE
//     buffer of N complex values per transform: scatter, barrier, gather, barrier),
//   * the first pass reads its inputs wherever the caller's loader finds them (global memory, with the taper / filter product
//     fused) and the last pass hands its outputs to the caller's functor in the order a coalesced store wants.
//
// Pass with sub-transform length Ns and radix R, butterfly j < N / R, k = j mod Ns: inputs x[j + r N / R] exp(-+2 pi i r k / (Ns R)),
// an R-point DFT, outputs to (j - k) R + k + r Ns.  Twiddles per pass as a block [R - 1][Ns], consecutive lanes reading
// consecutive entries.
#pragma once
#include "fft_anysize.cuh"

namespace sfft {

template <int N_, int R0, int R1, int R2 = 1> struct Plan {
    static constexpr int N = N_;
    static constexpr int NP = R2 > 1 ? 3 : 2;
    static_assert(R0 * R1 * R2 == N_, "radices must multiply to the length");
    static __host__ __device__ constexpr int radix(int p) { return p == 0 ? R0 : p == 1 ? R1 : R2; }
    static __host__ __device__ constexpr int ns(int p) { return p == 0 ? 1 : p == 1 ? R0 : R0 * R1; }
    static __host__ __device__ constexpr int twOffset(int p) {      // entries before pass p's block (pass 0 has none)
        return p <= 1 ? 0 : (R1 - 1) * R0;
    }
    static __host__ __device__ constexpr int twTotal() { return (R1 - 1) * R0 + (R2 > 1 ? (R2 - 1) * R0 * R1 : 0); }
};
// the same length with the passes in reverse order: its first pass reads x[j + r N / R_last], which is exactly what a thread
// holds after the last pass of Plan -- a forward transform followed by an inverse one needs no exchange in between
template <class PL> struct Reversed;
template <int N_, int R0, int R1, int R2> struct Reversed<Plan<N_, R0, R1, R2>> {
    typedef typename std::conditional<(R2 > 1), Plan<N_, R2, R1, R0>, Plan<N_, R1, R0, 1>>::type type;
};

__host__ __device__ constexpr int modinv(int a, int m) {
    for (int x = 1; x < m; ++x)
        if ((a * x) % m == 1) return x;
    return 1;
}

// natural-order DFT of R values in registers
template <typename T, int R, int DIR> __device__ __forceinline__ void dft(T (&xr)[R], T (&xi)[R]);

// R = A B with coprime A, B by the prime-factor (Good-Thomas) index maps: no twiddles between the two stages
template <typename T, int A, int B, int DIR> __device__ __forceinline__ void dft_pfa(T (&xr)[A * B], T (&xi)[A * B]) {
    constexpr int R = A * B;
    constexpr int IB = modinv(B % A, A), IA = modinv(A % B, B);
    T yr[A][B], yi[A][B];
#pragma unroll
    for (int n2 = 0; n2 < B; ++n2) {
        T ar[A], ai[A];
#pragma unroll
        for (int n1 = 0; n1 < A; ++n1) {
            ar[n1] = xr[(n1 * B + n2 * A) % R];
            ai[n1] = xi[(n1 * B + n2 * A) % R];
        }
        dft<T, A, DIR>(ar, ai);
#pragma unroll
        for (int k1 = 0; k1 < A; ++k1) {
            yr[k1][n2] = ar[k1];
            yi[k1][n2] = ai[k1];
        }
    }
#pragma unroll
    for (int k1 = 0; k1 < A; ++k1) {
        T br[B], bi[B];
#pragma unroll
        for (int n2 = 0; n2 < B; ++n2) {
            br[n2] = yr[k1][n2];
            bi[n2] = yi[k1][n2];
        }
        dft<T, B, DIR>(br, bi);
#pragma unroll
        for (int k2 = 0; k2 < B; ++k2) {
            xr[(k1 * B * IB + k2 * A * IA) % R] = br[k2];
            xi[(k1 * B * IB + k2 * A * IA) % R] = bi[k2];
        }
    }
}

template <typename T, int R, int DIR> __device__ __forceinline__ void dft(T (&xr)[R], T (&xi)[R]) {
    if constexpr (R == 2 || R == 3 || R == 4 || R == 5 || R == 7) dft_small<T, R, DIR>(xr, xi);
    else if constexpr (R == 8 || R == 9 || R == 16) dft_comp<T, R, DIR>(xr, xi);
    else if constexpr (R == 6) dft_pfa<T, 2, 3, DIR>(xr, xi);
    else if constexpr (R == 10) dft_pfa<T, 2, 5, DIR>(xr, xi);
    else if constexpr (R == 12) dft_pfa<T, 4, 3, DIR>(xr, xi);
    else if constexpr (R == 14) dft_pfa<T, 2, 7, DIR>(xr, xi);
    else if constexpr (R == 15) dft_pfa<T, 3, 5, DIR>(xr, xi);
    else static_assert(R == 2, "radix not implemented");
}

template <class PL, int NT, int P> struct Stage {
    static constexpr int R = PL::radix(P);
    static constexpr int NB = PL::N / R;            // butterflies of the pass
    static constexpr int IT = (NB + NT - 1) / NT;   // butterflies per thread
    static constexpr int NS = PL::ns(P);
    static constexpr int REGS = IT * R;
};

// element (i, r) of a stage's registers is value j + r NB of the pass input (before compute) with j = t + i NT
template <class PL, typename T, int NT, int DIR, int P>
__device__ __forceinline__ void stage_compute(T (&xr)[Stage<PL, NT, P>::REGS], T (&xi)[Stage<PL, NT, P>::REGS], const int t,
                                              const typename V2<T>::type* __restrict__ tw) {
    typedef Stage<PL, NT, P> S;
    typedef typename V2<T>::type C2;
#pragma unroll
    for (int i = 0; i < S::IT; ++i) {
        const int j = t + i * NT;
        if (S::IT * NT > S::NB && j >= S::NB) continue;
        T ar[S::R], ai[S::R];
#pragma unroll
        for (int r = 0; r < S::R; ++r) {
            ar[r] = xr[i * S::R + r];
            ai[r] = xi[i * S::R + r];
        }
        if constexpr (P > 0) {
            const int k = j % S::NS;
            const C2* twp = tw + PL::twOffset(P) + k;
#pragma unroll
            for (int r = 1; r < S::R; ++r) {
                const C2 w = __ldg(twp + (r - 1) * S::NS);
                const T wr = w.x, wi = (DIR < 0) ? w.y : -w.y;
                const T ur = ar[r], ui = ai[r];
                ar[r] = ur * wr - ui * wi;
                ai[r] = ur * wi + ui * wr;
            }
        }
        dft<T, S::R, DIR>(ar, ai);
#pragma unroll
        for (int r = 0; r < S::R; ++r) {
            xr[i * S::R + r] = ar[r];
            xi[i * S::R + r] = ai[r];
        }
    }
}

// outputs of pass P to their Stockham positions in the exchange buffer
template <class PL, typename T, int NT, int P>
__device__ __forceinline__ void stage_scatter(const T (&xr)[Stage<PL, NT, P>::REGS], const T (&xi)[Stage<PL, NT, P>::REGS], const int t,
                                              T* __restrict__ s_re, T* __restrict__ s_im) {
    typedef Stage<PL, NT, P> S;
#pragma unroll
    for (int i = 0; i < S::IT; ++i) {
        const int j = t + i * NT;
        if (S::IT * NT > S::NB && j >= S::NB) continue;
        const int k = j % S::NS;
        const int j0 = (j - k) * S::R + k;
#pragma unroll
        for (int r = 0; r < S::R; ++r) {
            s_re[j0 + r * S::NS] = xr[i * S::R + r];
            s_im[j0 + r * S::NS] = xi[i * S::R + r];
        }
    }
}

template <class PL, typename T, int NT, int P>
__device__ __forceinline__ void stage_gather(T (&xr)[Stage<PL, NT, P>::REGS], T (&xi)[Stage<PL, NT, P>::REGS], const int t,
                                             const T* __restrict__ s_re, const T* __restrict__ s_im) {
    typedef Stage<PL, NT, P> S;
#pragma unroll
    for (int i = 0; i < S::IT; ++i) {
        const int j = t + i * NT;
        if (S::IT * NT > S::NB && j >= S::NB) continue;
#pragma unroll
        for (int r = 0; r < S::R; ++r) {
            xr[i * S::R + r] = s_re[j + r * S::NB];
            xi[i * S::R + r] = s_im[j + r * S::NB];
        }
    }
}

// passes P .. NP-1 starting from pass P's inputs in registers; `out(xr, xi)` receives the last pass's registers: element
// (i, r) is X[j + r NB_last], j = t + i NT (valid while j < NB_last)
template <class PL, typename T, int NT, int DIR, int P, class Sync, class Out>
__device__ __forceinline__ void run_from(T (&xr)[Stage<PL, NT, P>::REGS], T (&xi)[Stage<PL, NT, P>::REGS], const int t, T* s_re, T* s_im,
                                         const typename V2<T>::type* __restrict__ tw, Sync sync, Out out) {
    stage_compute<PL, T, NT, DIR, P>(xr, xi, t, tw);
    if constexpr (P + 1 < PL::NP) {
        stage_scatter<PL, T, NT, P>(xr, xi, t, s_re, s_im);
        sync();
        T yr[Stage<PL, NT, P + 1>::REGS], yi[Stage<PL, NT, P + 1>::REGS];
        stage_gather<PL, T, NT, P + 1>(yr, yi, t, s_re, s_im);
        sync();
        run_from<PL, T, NT, DIR, P + 1>(yr, yi, t, s_re, s_im, tw, sync, out);
    } else {
        out(xr, xi);
    }
}

template <class PL, int NT> struct Last {
    typedef Stage<PL, NT, PL::NP - 1> S;
};

}  // namespace sfft
