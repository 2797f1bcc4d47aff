// Shared-memory Stockham FFT primitives, power-of-two lengths 16..8192, fp32 or fp64.
//
// One transform of length N is carried by NT = N/V threads; every thread keeps V complex values in
// registers (V = 16, or 8 where more resident warps matter more than fewer passes: V is deduced from the
// register arrays).  Before and after *every* pass thread t holds elements  t + q*NT  (q = 0..V-1), whatever
// the radix of the pass, so the first pass loads straight from global memory and the last one leaves results
// in natural order in registers (ready for a fused pointwise step, another transform, or a store).
// Passes are radix V (a DIF network with compile-time twiddles), with a smaller final radix when log2(N) is
// not a multiple of log2(V).  Between passes values are exchanged through SoA shared memory with one pad
// word per 16 elements.
#pragma once
#include "common.cuh"

namespace nbfft {

__host__ __device__ constexpr int bitrev(int i, int R) {
    int r = 0;
    for (int b = 1, rb = R >> 1; b < R; b <<= 1, rb >>= 1)
        if (i & b) r |= rb;
    return r;
}
__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }
__host__ __device__ constexpr int spad(int i) { return i + (i >> 4); }
__host__ __device__ constexpr int npad(int n) { return n + (n >> 4) + 1; }

template <int N, int Ns, int V = 16> struct PassRadix {
    static constexpr int REM = N / Ns;
    static constexpr int R = REM >= V ? V : REM;
};

template <typename T> __device__ __forceinline__ constexpr T cos16(int k) {  // cos(2 pi k / 16)
    return k == 0 ? (T)1.0
         : k == 1 ? (T)0.92387953251128673848
         : k == 2 ? (T)0.70710678118654752440
         : k == 3 ? (T)0.38268343236508977173
         : k == 4 ? (T)0.0
         : k == 5 ? (T)-0.38268343236508977173
         : k == 6 ? (T)-0.70710678118654752440
         : k == 7 ? (T)-0.92387953251128673848
                  : (T)-1.0;
}
template <typename T> __device__ __forceinline__ constexpr T sin16(int k) {  // sin(2 pi k / 16)
    return k == 0 ? (T)0.0
         : k == 1 ? (T)0.38268343236508977173
         : k == 2 ? (T)0.70710678118654752440
         : k == 3 ? (T)0.92387953251128673848
         : k == 4 ? (T)1.0
         : k == 5 ? (T)0.92387953251128673848
         : k == 6 ? (T)0.70710678118654752440
         : k == 7 ? (T)0.38268343236508977173
                  : (T)0.0;
}

// In-register DFT of the R values v[base + r*stride]; decimation in frequency, results bit-reversed:
// X[bitrev(i)] ends up in v[base + i*stride].  DIR = -1 forward, +1 inverse (no scaling).
template <typename T, int R, int DIR, int V>
__device__ __forceinline__ void dft_reg(T (&re)[V], T (&im)[V], const int base, const int stride) {
    constexpr int LOGR = ilog2(R);
#pragma unroll
    for (int s = 0; s < LOGR; ++s) {
        const int span = R >> (s + 1);
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if ((i & span) == 0) {
                const int e = (i % span) * (R / (2 * span)) * (16 / R);  // twiddle = W_16^e
                const int ia = base + i * stride, ib = base + (i + span) * stride;
                const T ar = re[ia], ai = im[ia], br = re[ib], bi = im[ib];
                re[ia] = ar + br;
                im[ia] = ai + bi;
                const T dr = ar - br, di = ai - bi;
                if (e == 0) {
                    re[ib] = dr;
                    im[ib] = di;
                } else if (e == 4) {
                    re[ib] = (DIR < 0) ? di : -di;
                    im[ib] = (DIR < 0) ? -dr : dr;
                } else {
                    const T c = cos16<T>(e);
                    const T sn = (DIR < 0) ? -sin16<T>(e) : sin16<T>(e);
                    re[ib] = dr * c - di * sn;
                    im[ib] = dr * sn + di * c;
                }
            }
        }
    }
}

// Pass twiddle table of an N-point transform carried V values per thread.  The pass that starts at sub-transform length
// Ns with radix R multiplies input r (r = 1..R-1) of the butterfly at position k = j mod Ns by exp(-2 pi i r k / (Ns R)).
// Stored per pass as the block [R-1][Ns] (k fastest), blocks in pass order; the first pass (Ns = 1) needs none.  The lanes
// of a warp hold consecutive k, so one load of a warp reads consecutive entries -- a table indexed by r * k * stride
// instead spreads every load over up to 32 sectors, and the row kernels were bound by exactly those L1 requests.
__host__ __device__ constexpr int pass_tw_offset(int N, int V, int NsTarget) {
    int off = 0;
    for (int Ns = 1; Ns < NsTarget;) {
        const int R = (N / Ns) >= V ? V : (N / Ns);
        if (Ns > 1) off += (R - 1) * Ns;
        Ns *= R;
    }
    return off;
}
__host__ __device__ constexpr int pass_tw_total(int N, int V) { return pass_tw_offset(N, V, N) > 0 ? pass_tw_offset(N, V, N) : 1; }

// twiddle + butterflies of the pass that starts at sub-transform length Ns.  tw: the pass table above.
template <typename T, int N, int DIR, int Ns, int V>
__device__ __forceinline__ void pass_compute(T (&re)[V], T (&im)[V], const int t,
                                             const typename V2<T>::type* __restrict__ tw) {
    constexpr int R = PassRadix<N, Ns, V>::R;
    constexpr int NT = N / V;
    constexpr int NB = V / R;
    constexpr int OFF = pass_tw_offset(N, V, Ns);
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        if (Ns > 1) {
            const int j = t + m * NT;
            const int k = j & (Ns - 1);
#pragma unroll
            for (int r = 1; r < R; ++r) {
                const typename V2<T>::type w = __ldg(tw + OFF + (r - 1) * Ns + k);
                const T wr = w.x;
                const T wi = (DIR < 0) ? w.y : -w.y;
                const int idx = m + r * NB;
                const T xr = re[idx], xi = im[idx];
                re[idx] = xr * wr - xi * wi;
                im[idx] = xr * wi + xi * wr;
            }
        }
        dft_reg<T, R, DIR, V>(re, im, m, NB);
    }
}

template <typename T, int N, int Ns, int V>
__device__ __forceinline__ void pass_scatter(const T (&re)[V], const T (&im)[V], const int t, T* s_re, T* s_im) {
    constexpr int R = PassRadix<N, Ns, V>::R;
    constexpr int NT = N / V;
    constexpr int NB = V / R;
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        const int j = t + m * NT;
        const int k = j & (Ns - 1);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int src = m + bitrev(r, R) * NB;
            const int pos = spad(j0 + r * Ns);
            s_re[pos] = re[src];
            s_im[pos] = im[src];
        }
    }
}

template <typename T, int N, int V>
__device__ __forceinline__ void gather_natural(T (&re)[V], T (&im)[V], const int t, const T* s_re, const T* s_im) {
    constexpr int NT = N / V;
#pragma unroll
    for (int q = 0; q < V; ++q) {
        const int pos = spad(t + q * NT);
        re[q] = s_re[pos];
        im[q] = s_im[pos];
    }
}

template <typename T, int N, int V>
__device__ __forceinline__ void scatter_natural(const T (&re)[V], const T (&im)[V], const int t, T* s_re, T* s_im) {
    constexpr int NT = N / V;
#pragma unroll
    for (int q = 0; q < V; ++q) {
        const int pos = spad(t + q * NT);
        s_re[pos] = re[q];
        s_im[pos] = im[q];
    }
}

template <typename T, int N, int DIR, int Ns, int V> struct Passes {
    template <class Sync>
    static __device__ __forceinline__ void run(T (&re)[V], T (&im)[V], const int t, T* s_re, T* s_im,
                                               const typename V2<T>::type* __restrict__ tw, Sync sync) {
        constexpr int R = PassRadix<N, Ns, V>::R;
        constexpr int NB = V / R;
        pass_compute<T, N, DIR, Ns, V>(re, im, t, tw);
        if constexpr (Ns * R < N) {
            pass_scatter<T, N, Ns, V>(re, im, t, s_re, s_im);
            sync();
            gather_natural<T, N, V>(re, im, t, s_re, s_im);
            sync();
            Passes<T, N, DIR, Ns * R, V>::run(re, im, t, s_re, s_im, tw, sync);
        } else {
            // last pass: output q = m + r*NB sits in register m + bitrev(r)*NB; permute to natural order
            T tr[V], ti[V];
#pragma unroll
            for (int m = 0; m < NB; ++m)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    tr[m + r * NB] = re[m + bitrev(r, R) * NB];
                    ti[m + r * NB] = im[m + bitrev(r, R) * NB];
                }
#pragma unroll
            for (int q = 0; q < V; ++q) {
                re[q] = tr[q];
                im[q] = ti[q];
            }
        }
    }
};

// Transform the N values spread over the NT = N/V threads of one group (thread t holds x[t + q*NT]).
// s_re/s_im: npad(N) words each, private to the group.  sync() must synchronise (at least) the group.
template <typename T, int N, int DIR, class Sync, int V>
__device__ __forceinline__ void fft_run(T (&re)[V], T (&im)[V], const int t, T* s_re, T* s_im,
                                        const typename V2<T>::type* __restrict__ tw, Sync sync) {
    static_assert(V == 8 || V == 16, "8 or 16 values per thread");
    static_assert(N >= V, "transform shorter than a thread's share");
    Passes<T, N, DIR, 1, V>::run(re, im, t, s_re, s_im, tw, sync);
}

struct BlockSync {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

}  // namespace nbfft
