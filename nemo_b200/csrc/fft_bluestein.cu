// 2-D real FFT passes for tile sides that are NOT powers of two (K1, K4, K8 on e.g. 1440 x 840 tiles).
//
// Padding a tile would change the answer (the apodisation and the filter are defined on the tile grid), so a length-N
// transform is done exactly with Bluestein's identity  n k = (n^2 + k^2 - (k - n)^2) / 2 :
//     X[k] = c[k] * sum_n (x[n] c[n]) * b[k - n],   c[n] = exp(-i pi n^2 / N),  b[n] = conj(c[n]),
// i.e. a cyclic convolution of length M >= 2N - 1, M a power of two, carried out with the register / shared-memory
// Stockham transforms of fft.cuh (two M-point FFTs per line and a pointwise product with the precomputed FFT of b).
// The kernels mirror fft_kernels.cu: rows pack two real rows into one complex transform; columns handle C adjacent
// columns per CTA; the same fusions (apodisation, filter multiply and frequency sum, pixel window, signalNorm / mask /
// zero-restore epilogue) are kept.
#include "fft.cuh"
#include "nemo_b200_internal.h"
#include <math.h>
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

using namespace nbfft;

// ------------------------------------------------------------------------------------------------------
// tables: tw[m] = exp(-2 pi i m / M); chirp[n] = exp(-i pi n^2 / N), n < N; bhat = FFT_M(b) / M with b[n] = conj(chirp[|n|])
// wrapped onto M (the 1/M of the inverse M-point transform is folded in here)
struct BsTables {
    const void* tw;
    const void* chirp;
    const void* bhat;
    int M;
};

static std::mutex g_bs_mutex;
static std::map<std::pair<int, int>, BsTables> g_bs_tables;   // (prec * 64 + device, N)

static void host_fft_pow2(std::vector<long double>& re, std::vector<long double>& im) {
    const int n = (int)re.size();
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            std::swap(re[i], re[j]);
            std::swap(im[i], im[j]);
        }
    }
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int len = 2; len <= n; len <<= 1) {
        for (int i = 0; i < n; i += len) {
            for (int k = 0; k < len / 2; ++k) {
                const long double a = -2.0L * pi * (long double)k / (long double)len;
                const long double wr = cosl(a), wi = sinl(a);
                const int p = i + k, q = i + k + len / 2;
                const long double xr = re[q] * wr - im[q] * wi, xi = re[q] * wi + im[q] * wr;
                re[q] = re[p] - xr;
                im[q] = im[p] - xi;
                re[p] += xr;
                im[p] += xi;
            }
        }
    }
}

int nb_bs_length(int N) {
    int M = 16;
    while (M < 2 * N - 1) M <<= 1;
    return M;
}

template <typename T> static int upload_c2(const std::vector<long double>& re, const std::vector<long double>& im, const void** out) {
    typedef typename V2<T>::type C2;
    std::vector<C2> h(re.size());
    for (size_t i = 0; i < re.size(); ++i) {
        h[i].x = (T)re[i];
        h[i].y = (T)im[i];
    }
    void* d = nullptr;
    NB_CHECK_CUDA(cudaMalloc(&d, sizeof(C2) * h.size()));
    NB_CHECK_CUDA(cudaMemcpy(d, h.data(), sizeof(C2) * h.size(), cudaMemcpyHostToDevice));
    *out = d;
    return NB_OK;
}

static int get_bs_tables(int prec, int N, BsTables* out) {
    std::lock_guard<std::mutex> lock(g_bs_mutex);
    int dev = 0;
    cudaGetDevice(&dev);
    auto key = std::make_pair(prec * 64 + dev, N);
    auto it = g_bs_tables.find(key);
    if (it != g_bs_tables.end()) {
        *out = it->second;
        return NB_OK;
    }
    const int M = nb_bs_length(N);
    const long double pi = 3.141592653589793238462643383279502884L;
    // 2-3-5-7-smooth lengths: mixed-radix plan and pass tables instead of the chirp tables (bhat stays null, see bluestein_run)
    {
        std::vector<int> radices;
        int rest = N;
        for (int f : {7, 5, 3}) while (rest % f == 0) { radices.push_back(f); rest /= f; }
        while (rest % 4 == 0) { radices.push_back(4); rest /= 4; }
        if (rest % 2 == 0) { radices.push_back(2); rest /= 2; }
        if (rest == 1 && !getenv("NEMO_B200_FORCE_BLUESTEIN")) {
            std::vector<int> plan(1 + 2 * radices.size());
            std::vector<long double> tr, ti;
            plan[0] = (int)radices.size();
            int Ns = 1;
            for (size_t p = 0; p < radices.size(); ++p) {
                const int R = radices[p];
                plan[1 + 2 * p] = R;
                plan[2 + 2 * p] = (int)tr.size();
                if (Ns > 1) {
                    for (int r = 1; r < R; ++r)
                        for (int k = 0; k < Ns; ++k) {
                            const long double a = 2.0L * pi * (long double)((long long)r * k) / (long double)((long long)Ns * R);
                            tr.push_back(cosl(a));
                            ti.push_back(-sinl(a));
                        }
                }
                Ns *= R;
            }
            if (tr.empty()) { tr.push_back(1.0L); ti.push_back(0.0L); }
            BsTables t;
            t.M = M;
            t.bhat = nullptr;
            int rc;
            if (prec == NB_PREC_F64) rc = upload_c2<double>(tr, ti, &t.tw);
            else rc = upload_c2<float>(tr, ti, &t.tw);
            if (rc != NB_OK) return rc;
            void* d = nullptr;
            NB_CHECK_CUDA(cudaMalloc(&d, sizeof(int) * plan.size()));
            NB_CHECK_CUDA(cudaMemcpy(d, plan.data(), sizeof(int) * plan.size(), cudaMemcpyHostToDevice));
            t.chirp = d;
            g_bs_tables[key] = t;
            *out = t;
            return NB_OK;
        }
    }
    std::vector<long double> cr(N), ci(N), br(M, 0.0L), bi(M, 0.0L), tr(M), ti(M);
    for (int n = 0; n < N; ++n) {
        const long long r = ((long long)n * n) % (2LL * N);   // n^2 mod 2N keeps the phase exact
        const long double a = pi * (long double)r / (long double)N;
        cr[n] = cosl(a);
        ci[n] = -sinl(a);
        br[n] = cr[n];
        bi[n] = -ci[n];
        if (n > 0) {
            br[M - n] = br[n];
            bi[M - n] = bi[n];
        }
    }
    host_fft_pow2(br, bi);
    for (int m = 0; m < M; ++m) {
        br[m] /= (long double)M;
        bi[m] /= (long double)M;
    }
    nb_fill_pass_twiddles(M, 16, tr, ti);     // pass tables of the M-point transforms (fft.cuh), 16 values per thread
    BsTables t;
    t.M = M;
    int rc;
    if (prec == NB_PREC_F64) {
        if ((rc = upload_c2<double>(tr, ti, &t.tw)) != NB_OK) return rc;
        if ((rc = upload_c2<double>(cr, ci, &t.chirp)) != NB_OK) return rc;
        if ((rc = upload_c2<double>(br, bi, &t.bhat)) != NB_OK) return rc;
    } else {
        if ((rc = upload_c2<float>(tr, ti, &t.tw)) != NB_OK) return rc;
        if ((rc = upload_c2<float>(cr, ci, &t.chirp)) != NB_OK) return rc;
        if ((rc = upload_c2<float>(br, bi, &t.bhat)) != NB_OK) return rc;
    }
    g_bs_tables[key] = t;
    *out = t;
    return NB_OK;
}

// ------------------------------------------------------------------------------------------------------
// Lengths whose only prime factors are 2, 3, 5 and 7 (1440 x 840 survey tiles: 2^5 3^2 5 and 2^3 3 5 7) do not need the chirp
// detour: a mixed-radix Stockham transform of length N itself, passes of radix 7 / 5 / 3 / 4 / 2 through the group's shared
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
        dft_small<T, R, DIR>(xr, xi);
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
            default: mix_pass<T, 7, DIR>(aR, aI, bR, bI, tw, N, Ns, t, NT); break;
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

template <typename T, int M> struct BsRowCfg {
    static constexpr int NT = M / 16;
    static constexpr int LD = npad(M) + 2;
    static constexpr int RAW = (96 * 1024) / (2 * LD * (int)sizeof(T));
    static constexpr int GPB0 = RAW < 1 ? 1 : RAW;                       // row pairs per CTA by shared memory
    static constexpr int GPB1 = (256 / NT) < 1 ? 1 : (256 / NT);         // ... and by threads
    static constexpr int GPB = GPB0 < GPB1 ? GPB0 : GPB1;
    static constexpr int THREADS = GPB * NT;
    static constexpr size_t SMEM = (size_t)2 * GPB * LD * sizeof(T);
};

// ------------------------------------------------------------------------------------------------------
// rows, real -> half spectrum; two rows per transform
template <typename T, int M>
__global__ void __launch_bounds__(BsRowCfg<T, M>::THREADS)
bs_rows_fwd_kernel(const T* __restrict__ in, typename V2<T>::type* __restrict__ out, int H, int W, long long nrows, NbTaper tp, T scale,
                   const typename V2<T>::type* __restrict__ tw, const typename V2<T>::type* __restrict__ chirp,
                   const typename V2<T>::type* __restrict__ bhat) {
    typedef BsRowCfg<T, M> Cfg;
    typedef typename V2<T>::type C2;
    constexpr int NT = Cfg::NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int g = threadIdx.x / NT, t = threadIdx.x % NT;
    const long long r0 = ((long long)blockIdx.x * Cfg::GPB + g) * 2, r1 = r0 + 1;
    T* s_re = smem + g * Cfg::LD;
    T* s_im = smem + (Cfg::GPB + g) * Cfg::LD;
    const int Wc = W / 2 + 1;
    T re[16], im[16];
    const bool v0 = r0 < nrows, v1 = r1 < nrows;
    const T ay0 = (v0 && tp.n > 0) ? taper_at<T>(tp, (int)(r0 % H), H) : (T)1;
    const T ay1 = (v1 && tp.n > 0) ? taper_at<T>(tp, (int)(r1 % H), H) : (T)1;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int n = t + q * NT;
        T a = 0, b = 0;
        if (n < W) {
            const T ax = (tp.n > 0) ? taper_at<T>(tp, n, W) : (T)1;
            if (v0) a = in[(size_t)r0 * W + n] * ay0 * ax;
            if (v1) b = in[(size_t)r1 * W + n] * ay1 * ax;
        }
        re[q] = a;
        im[q] = b;
    }
    bluestein_run<T, M, -1>(re, im, t, s_re, s_im, tw, chirp, bhat, W, BlockSync());
    __syncthreads();
    scatter_natural<T, M>(re, im, t, s_re, s_im);
    __syncthreads();
    const T h = (T)0.5;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int k = t + q * NT;
        if (k < Wc) {
            const int kc = (k == 0) ? 0 : W - k;
            const T zr = re[q], zi = im[q];
            const T cr = s_re[spad(kc)], ci = -s_im[spad(kc)];   // conj(Z[W - k])
            // A = (Z + conj(Zc)) / 2,  B = (Z - conj(Zc)) / (2 i)
            if (v0) out[(size_t)r0 * Wc + k] = mk2<T>(h * (zr + cr) * scale, h * (zi + ci) * scale);
            if (v1) out[(size_t)r1 * Wc + k] = mk2<T>(h * (zi - ci) * scale, -h * (zr - cr) * scale);
        }
    }
}

template <typename T> struct BsC2RParams {
    T scale;
    const double* norm;
    const unsigned char* mask;
    int maskDiv;
    const T* zeroRef;
};

// rows, half spectrum -> real (un-normalised inverse); two rows per transform; same epilogue as irfft_rows_kernel
template <typename T, int M>
__global__ void __launch_bounds__(BsRowCfg<T, M>::THREADS)
bs_rows_inv_kernel(const typename V2<T>::type* __restrict__ in, T* out, int H, int W, long long nrows, BsC2RParams<T> ep,
                   const typename V2<T>::type* __restrict__ tw, const typename V2<T>::type* __restrict__ chirp,
                   const typename V2<T>::type* __restrict__ bhat) {
    typedef BsRowCfg<T, M> Cfg;
    typedef typename V2<T>::type C2;
    constexpr int NT = Cfg::NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int g = threadIdx.x / NT, t = threadIdx.x % NT;
    const long long r0 = ((long long)blockIdx.x * Cfg::GPB + g) * 2, r1 = r0 + 1;
    T* s_re = smem + g * Cfg::LD;
    T* s_im = smem + (Cfg::GPB + g) * Cfg::LD;
    const int Wc = W / 2 + 1;
    const bool v0 = r0 < nrows, v1 = r1 < nrows;
    T re[16], im[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int k = t + q * NT;
        T zr = 0, zi = 0;
        if (k < W) {
            const bool low = k < Wc;
            const int kk = low ? k : W - k;
            C2 a = mk2<T>(0, 0), b = mk2<T>(0, 0);
            if (v0) a = in[(size_t)r0 * Wc + kk];
            if (v1) b = in[(size_t)r1 * Wc + kk];
            if (!low) {   // Hermitian extension
                a.y = -a.y;
                b.y = -b.y;
            }
            zr = a.x - b.y;   // Z = A + i B
            zi = a.y + b.x;
        }
        re[q] = zr;
        im[q] = zi;
    }
    bluestein_run<T, M, +1>(re, im, t, s_re, s_im, tw, chirp, bhat, W, BlockSync());
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        const long long row = rr == 0 ? r0 : r1;
        if (row >= nrows) continue;
        const long long img = row / H;
        const int y = (int)(row % H);
        T f = ep.scale;
        if (ep.norm) f *= (T)ep.norm[img];
        const unsigned char* mrow = ep.mask ? ep.mask + ((size_t)(img / ep.maskDiv) * H + y) * W : nullptr;
        const T* zrow = ep.zeroRef ? ep.zeroRef + (size_t)row * W : nullptr;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int n = t + q * NT;
            if (n < W) {
                T v = (rr == 0 ? re[q] : im[q]) * f;
                if (mrow) v *= (T)mrow[n];
                if (zrow && zrow[n] == (T)0) v = 0;
                out[(size_t)row * W + n] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// columns.  A CTA owns C adjacent columns for all H rows, one thread group per column.
template <typename T, int M> struct BsColCfg {
    static constexpr int NT = M / 16;
    static constexpr int BYTES_PER_COL = npad(M) * 2 * (int)sizeof(T);
    static constexpr int CSMEM = (144 * 1024) / BYTES_PER_COL;
    static constexpr int CTHR = 512 / NT > 0 ? 512 / NT : 1;
    static constexpr int CRAW = CSMEM < CTHR ? CSMEM : CTHR;
    static constexpr int C = CRAW >= 16 ? 16 : CRAW >= 8 ? 8 : CRAW >= 4 ? 4 : CRAW >= 2 ? 2 : 1;
    static constexpr int THREADS = C * NT;
    static constexpr int LD = npad(M) + 1;
    static constexpr size_t SMEM = (size_t)2 * C * LD * sizeof(T);
};

// MODE 0: plain transform in -> out (DIR); 1: out[o] = inverse of sum_f fmaps[m(o), f] * filt[o, f]; 2: in place forward,
// times wy[ky] wx[kx], inverse
template <typename T, int M, int MODE, int DIR, int F>
__global__ void __launch_bounds__(BsColCfg<T, M>::THREADS)
bs_cols_kernel(const typename V2<T>::type* __restrict__ in, const T* __restrict__ filt, typename V2<T>::type* __restrict__ out, int H,
               int Wc, int S, int perScale, const T* __restrict__ wy, const T* __restrict__ wx,
               const typename V2<T>::type* __restrict__ tw, const typename V2<T>::type* __restrict__ chirp,
               const typename V2<T>::type* __restrict__ bhat) {
    typedef BsColCfg<T, M> Cfg;
    typedef typename V2<T>::type C2;
    constexpr int NT = Cfg::NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int c = threadIdx.x % Cfg::C, t = threadIdx.x / Cfg::C;
    const int col = blockIdx.x * Cfg::C + c;
    const bool valid = col < Wc;
    T* s_re = smem + c * Cfg::LD;
    T* s_im = smem + (Cfg::C + c) * Cfg::LD;
    const int o = blockIdx.y;
    const size_t plane = (size_t)H * Wc;
    T re[16], im[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        re[q] = 0;
        im[q] = 0;
    }
    if (valid) {
        if (MODE == 1) {
            const int tile = perScale ? (o % S) : (o / S);
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const C2* fm = in + ((size_t)tile * F + f) * plane + col;
                const T* fl = filt + ((size_t)o * F + f) * plane + col;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int n = t + q * NT;
                    if (n < H) {
                        const C2 v = fm[(size_t)n * Wc];
                        const T gg = fl[(size_t)n * Wc];
                        re[q] += v.x * gg;
                        im[q] += v.y * gg;
                    }
                }
            }
        } else {
            const C2* src = in + (size_t)o * plane + col;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int n = t + q * NT;
                if (n < H) {
                    const C2 v = src[(size_t)n * Wc];
                    re[q] = v.x;
                    im[q] = v.y;
                }
            }
        }
    }
    if (MODE == 2) {
        bluestein_run<T, M, -1>(re, im, t, s_re, s_im, tw, chirp, bhat, H, BlockSync());
        const T wxc = valid ? wx[col] : (T)0;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int k = t + q * NT;
            const T w = (k < H) ? wy[k] * wxc : (T)0;
            re[q] *= w;
            im[q] *= w;
        }
        __syncthreads();
        bluestein_run<T, M, +1>(re, im, t, s_re, s_im, tw, chirp, bhat, H, BlockSync());
    } else {
        bluestein_run<T, M, (MODE == 1 ? +1 : DIR)>(re, im, t, s_re, s_im, tw, chirp, bhat, H, BlockSync());
    }
    if (valid) {
        C2* dst = out + (size_t)o * plane + col;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int k = t + q * NT;
            if (k < H) dst[(size_t)k * Wc] = mk2<T>(re[q], im[q]);
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

static void fill_taper(NbTaper& tp, int apodPix) {
    tp.n = apodPix;
    for (int i = 0; i < apodPix; ++i) tp.a[i] = (apodPix == 1) ? 0.0 : 0.5 * (1.0 - cos(M_PI * (double)i / (double)(apodPix - 1)));
    if (apodPix > 0) tp.a[apodPix - 1] = 1.0;
}

template <typename T>
static int bs_rows_fwd_t(const T* in, typename V2<T>::type* out, int B, int H, int W, int apodPix, double scale, cudaStream_t st) {
    typedef typename V2<T>::type C2;
    NB_REQUIRE(apodPix >= 0 && apodPix <= NB_MAX_TAPER && 2 * apodPix <= H && 2 * apodPix <= W, NB_ERR_ARG, "apodPix %d out of range", apodPix);
    NbTaper tp;
    fill_taper(tp, apodPix);
    BsTables tb;
    int rc;
    if ((rc = get_bs_tables(sizeof(T) == 8 ? NB_PREC_F64 : NB_PREC_F32, W, &tb)) != NB_OK) return rc;
    const long long nrows = (long long)B * H;
    const long long npairs = (nrows + 1) / 2;
#define BS_ROWS_FWD(MM)                                                                                                  \
    {                                                                                                                    \
        typedef BsRowCfg<T, MM> Cfg;                                                                                     \
        if ((rc = bs_set_smem(bs_rows_fwd_kernel<T, MM>, Cfg::SMEM)) != NB_OK) return rc;                                \
        const unsigned grid = (unsigned)((npairs + Cfg::GPB - 1) / Cfg::GPB);                                            \
        { NB_PROF_BEGIN(st); bs_rows_fwd_kernel<T, MM><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(in, out, H, W, nrows, tp, (T)scale, (const C2*)tb.tw, (const C2*)tb.chirp, (const C2*)tb.bhat); NB_PROF_END("bs_rows_fwd_kernel", st); } \
    }
    BS_DISPATCH_M(tb.M, BS_ROWS_FWD)
#undef BS_ROWS_FWD
    NB_CHECK_LAUNCH();
    return NB_OK;
}

template <typename T>
static int bs_rows_inv_t(const typename V2<T>::type* in, T* out, int B, int H, int W, BsC2RParams<T> ep, cudaStream_t st) {
    typedef typename V2<T>::type C2;
    BsTables tb;
    int rc;
    if ((rc = get_bs_tables(sizeof(T) == 8 ? NB_PREC_F64 : NB_PREC_F32, W, &tb)) != NB_OK) return rc;
    const long long nrows = (long long)B * H;
    const long long npairs = (nrows + 1) / 2;
#define BS_ROWS_INV(MM)                                                                                                  \
    {                                                                                                                    \
        typedef BsRowCfg<T, MM> Cfg;                                                                                     \
        if ((rc = bs_set_smem(bs_rows_inv_kernel<T, MM>, Cfg::SMEM)) != NB_OK) return rc;                                \
        const unsigned grid = (unsigned)((npairs + Cfg::GPB - 1) / Cfg::GPB);                                            \
        { NB_PROF_BEGIN(st); bs_rows_inv_kernel<T, MM><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(in, out, H, W, nrows, ep, (const C2*)tb.tw, (const C2*)tb.chirp, (const C2*)tb.bhat); NB_PROF_END("bs_rows_inv_kernel", st); } \
    }
    BS_DISPATCH_M(tb.M, BS_ROWS_INV)
#undef BS_ROWS_INV
    NB_CHECK_LAUNCH();
    return NB_OK;
}

template <typename T, int MODE, int DIR, int F>
static int bs_cols_t(const typename V2<T>::type* in, const T* filt, typename V2<T>::type* out, int nOut, int S, int perScale, int H, int Wc,
                     const T* wy, const T* wx, cudaStream_t st) {
    typedef typename V2<T>::type C2;
    BsTables tb;
    int rc;
    if ((rc = get_bs_tables(sizeof(T) == 8 ? NB_PREC_F64 : NB_PREC_F32, H, &tb)) != NB_OK) return rc;
#define BS_COLS(MM)                                                                                                      \
    {                                                                                                                    \
        typedef BsColCfg<T, MM> Cfg;                                                                                     \
        if ((rc = bs_set_smem(bs_cols_kernel<T, MM, MODE, DIR, F>, Cfg::SMEM)) != NB_OK) return rc;                      \
        dim3 grid((Wc + Cfg::C - 1) / Cfg::C, nOut);                                                                     \
        { NB_PROF_BEGIN(st); bs_cols_kernel<T, MM, MODE, DIR, F><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(in, filt, out, H, Wc, S, perScale, wy, wx, (const C2*)tb.tw, (const C2*)tb.chirp, (const C2*)tb.bhat); NB_PROF_END("bs_cols_kernel", st); } \
    }
    BS_DISPATCH_M(tb.M, BS_COLS)
#undef BS_COLS
    NB_CHECK_LAUNCH();
    return NB_OK;
}

// ------------------------------------------------------------------------------------------------------
// entry points used by fft_kernels.cu (declared in nemo_b200_internal.h)
int nb_bs_rows_fwd(int prec, const void* in, void* out, int B, int H, int W, int apodPix, double scale, cudaStream_t st) {
    if (prec == NB_PREC_F64) return bs_rows_fwd_t<double>((const double*)in, (double2*)out, B, H, W, apodPix, scale, st);
    return bs_rows_fwd_t<float>((const float*)in, (float2*)out, B, H, W, apodPix, scale, st);
}

int nb_bs_rows_inv(int prec, const void* in, void* out, int B, int H, int W, double scale, const double* norm, const unsigned char* mask,
                   int maskDiv, const void* zeroRef, cudaStream_t st) {
    if (prec == NB_PREC_F64) {
        BsC2RParams<double> ep = {scale, norm, mask, maskDiv < 1 ? 1 : maskDiv, (const double*)zeroRef};
        return bs_rows_inv_t<double>((const double2*)in, (double*)out, B, H, W, ep, st);
    }
    BsC2RParams<float> ep = {(float)scale, norm, mask, maskDiv < 1 ? 1 : maskDiv, (const float*)zeroRef};
    return bs_rows_inv_t<float>((const float2*)in, (float*)out, B, H, W, ep, st);
}

int nb_bs_cols(int prec, const void* in, void* out, int B, int H, int Wc, int dir, cudaStream_t st) {
    if (prec == NB_PREC_F64) {
        if (dir < 0) return bs_cols_t<double, 0, -1, 1>((const double2*)in, nullptr, (double2*)out, B, 1, 0, H, Wc, nullptr, nullptr, st);
        return bs_cols_t<double, 0, +1, 1>((const double2*)in, nullptr, (double2*)out, B, 1, 0, H, Wc, nullptr, nullptr, st);
    }
    if (dir < 0) return bs_cols_t<float, 0, -1, 1>((const float2*)in, nullptr, (float2*)out, B, 1, 0, H, Wc, nullptr, nullptr, st);
    return bs_cols_t<float, 0, +1, 1>((const float2*)in, nullptr, (float2*)out, B, 1, 0, H, Wc, nullptr, nullptr, st);
}

int nb_bs_cols_apply(int prec, const void* fmaps, const void* filt, void* out, int nOut, int S, int perScale, int F, int H, int Wc,
                     cudaStream_t st) {
#define BS_APPLY(T, C2T)                                                                                                             \
    switch (F) {                                                                                                                     \
        case 1: return bs_cols_t<T, 1, +1, 1>((const C2T*)fmaps, (const T*)filt, (C2T*)out, nOut, S, perScale, H, Wc, nullptr, nullptr, st); \
        case 2: return bs_cols_t<T, 1, +1, 2>((const C2T*)fmaps, (const T*)filt, (C2T*)out, nOut, S, perScale, H, Wc, nullptr, nullptr, st); \
        case 3: return bs_cols_t<T, 1, +1, 3>((const C2T*)fmaps, (const T*)filt, (C2T*)out, nOut, S, perScale, H, Wc, nullptr, nullptr, st); \
        default: return bs_cols_t<T, 1, +1, 4>((const C2T*)fmaps, (const T*)filt, (C2T*)out, nOut, S, perScale, H, Wc, nullptr, nullptr, st); \
    }
    if (prec == NB_PREC_F64) { BS_APPLY(double, double2) }
    BS_APPLY(float, float2)
#undef BS_APPLY
}

int nb_bs_cols_window(int prec, void* data, int B, int H, int Wc, const void* wy, const void* wx, cudaStream_t st) {
    if (prec == NB_PREC_F64)
        return bs_cols_t<double, 2, -1, 1>((const double2*)data, nullptr, (double2*)data, B, 1, 0, H, Wc, (const double*)wy, (const double*)wx, st);
    return bs_cols_t<float, 2, -1, 1>((const float2*)data, nullptr, (float2*)data, B, 1, 0, H, Wc, (const float*)wy, (const float*)wx, st);
}
