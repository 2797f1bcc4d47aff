// 2-D real FFT kernels (K1, K4, K8 of SURVEY.md section 2a) built on the primitives in fft.cuh.
//
// Layout: real maps [B, H, W] row-major; half spectra [B, H, Wc] complex, Wc = W/2 + 1 (kx = 0..W/2).
//   rows r2c : one group of W/32 threads per map row; even/odd samples packed as one complex value, an
//              (W/2)-point complex FFT, then the real-input untangle.  Apodisation (enmap.apod, reference
//              filters.py:528) and the 1/sqrt(HW) of enmap.fft are fused into the load / store.
//   columns  : a CTA owns C adjacent columns for all H rows (C*8 or C*16 contiguous bytes per row), one
//              group of H/16 threads per column.  Variants: plain; "apply" = sum over frequencies of
//              spectrum x filter fused into the load of the inverse transform (filters.py:851); "window" =
//              forward, multiply by the pixel window, inverse, without leaving registers (filters.py:103).
//   rows c2r : inverse of r2c with the output scale, per-image normalisation (signalNorm), mask multiply
//              and zero-restore fused into the store.
#include "fft.cuh"
#include "nemo_b200_internal.h"
#include <math.h>
#include <stdlib.h>
#include <map>
#include <mutex>
#include <vector>

using namespace nbfft;

// occupancy knobs (minimum resident CTAs per SM asked of ptxas, column-tile budget)
#ifndef FFT_ROW_V
#define FFT_ROW_V 8         // complex values per thread of the row passes: 8 halves the registers of 16 and doubles the resident warps
#endif
#ifndef FFT_ROW_MINB
#define FFT_ROW_MINB (FFT_ROW_V == 8 ? 3 : 2)   // resident CTAs per SM asked of ptxas (256 threads each)
#endif
#ifndef FFT_COL_MINB
#define FFT_COL_MINB 1
#endif
#ifndef FFT_COL_SMEM_KB
#define FFT_COL_SMEM_KB 144
#endif
#ifndef FFT_COL_THREADS
#define FFT_COL_THREADS 512
#endif

// ------------------------------------------------------------------------------------------------------
// twiddle tables: tw[m] = exp(-2 pi i m / N), computed in long double on the host, cached per (prec, N)
static std::mutex g_plan_mutex;
static std::map<std::pair<int, int>, void*> g_twiddles;

static int get_twiddles(int prec, int N, const void** out) {
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    int dev = 0;
    cudaGetDevice(&dev);
    auto key = std::make_pair(prec * 64 + dev, N);
    auto it = g_twiddles.find(key);
    if (it != g_twiddles.end()) {
        *out = it->second;
        return NB_OK;
    }
    void* d = nullptr;
    const long double twopi = 6.283185307179586476925286766559005768L;
    if (prec == NB_PREC_F64) {
        std::vector<double2> h(N);
        for (int m = 0; m < N; ++m) {
            long double a = twopi * (long double)m / (long double)N;
            h[m].x = (double)cosl(a);
            h[m].y = (double)(-sinl(a));
        }
        NB_CHECK_CUDA(cudaMalloc(&d, sizeof(double2) * N));
        NB_CHECK_CUDA(cudaMemcpy(d, h.data(), sizeof(double2) * N, cudaMemcpyHostToDevice));
    } else {
        std::vector<float2> h(N);
        for (int m = 0; m < N; ++m) {
            long double a = twopi * (long double)m / (long double)N;
            h[m].x = (float)cosl(a);
            h[m].y = (float)(-sinl(a));
        }
        NB_CHECK_CUDA(cudaMalloc(&d, sizeof(float2) * N));
        NB_CHECK_CUDA(cudaMemcpy(d, h.data(), sizeof(float2) * N, cudaMemcpyHostToDevice));
    }
    g_twiddles[key] = d;
    *out = d;
    return NB_OK;
}

// pass twiddle tables (fft.cuh: pass_tw_offset) of an N-point transform carried V values per thread, cached per (prec, N, V)
static std::map<std::pair<int, int>, void*> g_pass_twiddles;

void nb_fill_pass_twiddles(int N, int V, std::vector<long double>& re, std::vector<long double>& im) {
    const long double twopi = 6.283185307179586476925286766559005768L;
    re.assign(pass_tw_total(N, V), 1.0L);
    im.assign(pass_tw_total(N, V), 0.0L);
    for (int Ns = 1; Ns < N;) {
        const int R = (N / Ns) >= V ? V : (N / Ns);
        if (Ns > 1) {
            const int off = pass_tw_offset(N, V, Ns);
            for (int r = 1; r < R; ++r)
                for (int k = 0; k < Ns; ++k) {
                    const long double a = twopi * (long double)(r * k) / (long double)(Ns * R);
                    re[off + (r - 1) * Ns + k] = cosl(a);
                    im[off + (r - 1) * Ns + k] = -sinl(a);
                }
        }
        Ns *= R;
    }
}

static int get_pass_twiddles(int prec, int N, int V, const void** out) {
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    int dev = 0;
    cudaGetDevice(&dev);
    auto key = std::make_pair((prec * 64 + dev) * 32 + V, N);
    auto it = g_pass_twiddles.find(key);
    if (it != g_pass_twiddles.end()) {
        *out = it->second;
        return NB_OK;
    }
    std::vector<long double> re, im;
    nb_fill_pass_twiddles(N, V, re, im);
    const size_t n = re.size();
    void* d = nullptr;
    if (prec == NB_PREC_F64) {
        std::vector<double2> h(n);
        for (size_t m = 0; m < n; ++m) {
            h[m].x = (double)re[m];
            h[m].y = (double)im[m];
        }
        NB_CHECK_CUDA(cudaMalloc(&d, sizeof(double2) * n));
        NB_CHECK_CUDA(cudaMemcpy(d, h.data(), sizeof(double2) * n, cudaMemcpyHostToDevice));
    } else {
        std::vector<float2> h(n);
        for (size_t m = 0; m < n; ++m) {
            h[m].x = (float)re[m];
            h[m].y = (float)im[m];
        }
        NB_CHECK_CUDA(cudaMalloc(&d, sizeof(float2) * n));
        NB_CHECK_CUDA(cudaMemcpy(d, h.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
    }
    g_pass_twiddles[key] = d;
    *out = d;
    return NB_OK;
}

// columns per CTA: as many adjacent columns as fit a ~144 KB exchange buffer and 512 threads (one CTA per SM at 128
// registers per thread), up to 16: a row of the tile is then C * sizeof(complex) contiguous bytes -- 128 B, a whole L2 line,
// for fp64 at H = 1024 -- which is what the strided column walk needs to use DRAM bursts fully
template <typename T, int N> struct ColCfg {
    static constexpr int BYTES_PER_COL = npad(N) * 2 * (int)sizeof(T);
    static constexpr int CSMEM = (FFT_COL_SMEM_KB * 1024) / BYTES_PER_COL;
    static constexpr int CTHR = FFT_COL_THREADS / (N / 16) > 0 ? FFT_COL_THREADS / (N / 16) : 1;
    static constexpr int CRAW = CSMEM < CTHR ? CSMEM : CTHR;
    static constexpr int C = CRAW >= 16 ? 16 : CRAW >= 8 ? 8 : CRAW >= 4 ? 4 : CRAW >= 2 ? 2 : 1;
    static constexpr int NT = N / 16;
    static constexpr int THREADS = C * NT;
    // column stride in shared memory.  One access phase serves 128 bytes: BANKS = 32 floats or 16 doubles (a warp of doubles
    // takes two phases of 16 lanes).  Lanes run over the C columns first, then over BANKS / C consecutive positions, so a
    // stride == BANKS / C (mod BANKS) elements gives every lane of a phase its own bank
    static constexpr int BANKS = 128 / (int)sizeof(T);
    static constexpr int TV = (BANKS / C) > 0 ? (BANKS / C) : 1;
    static constexpr int LD = npad(N) + ((TV - npad(N)) % BANKS + BANKS) % BANKS;
    static constexpr size_t SMEM = (size_t)2 * C * LD * sizeof(T);
};

// ------------------------------------------------------------------------------------------------------
// column kernels
template <typename T, int N, int DIR>
__global__ void __launch_bounds__(ColCfg<T, N>::THREADS, FFT_COL_MINB)
fft_cols_kernel(const typename V2<T>::type* __restrict__ in, typename V2<T>::type* __restrict__ out, int Wc,
                const typename V2<T>::type* __restrict__ tw, T scale) {
    typedef ColCfg<T, N> Cfg;
    typedef typename V2<T>::type C2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int c = threadIdx.x % Cfg::C, t = threadIdx.x / Cfg::C;
    const int col = blockIdx.x * Cfg::C + c;
    const bool valid = col < Wc;
    T* s_re = smem + c * Cfg::LD;
    T* s_im = smem + (Cfg::C + c) * Cfg::LD;
    const size_t base = (size_t)blockIdx.y * N * Wc + col;
    T re[16], im[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        C2 v = mk2<T>(0, 0);
        if (valid) v = in[base + (size_t)(t + q * Cfg::NT) * Wc];
        re[q] = v.x;
        im[q] = v.y;
    }
    fft_run<T, N, DIR>(re, im, t, s_re, s_im, tw, BlockSync());
    if (valid) {
#pragma unroll
        for (int q = 0; q < 16; ++q) out[base + (size_t)(t + q * Cfg::NT) * Wc] = mk2<T>(re[q] * scale, im[q] * scale);
    }
}

// out[o] = inverse column FFT of sum_f fmaps[o / S, f] * filt[o, f]      (filt real, folded half plane)
template <typename T, int N, int F>
__global__ void __launch_bounds__(ColCfg<T, N>::THREADS, FFT_COL_MINB)
fft_cols_apply_kernel(const typename V2<T>::type* __restrict__ fmaps, const T* __restrict__ filt,
                      typename V2<T>::type* __restrict__ out, int Wc, int S, int perScale,
                      const typename V2<T>::type* __restrict__ tw) {
    typedef ColCfg<T, N> Cfg;
    typedef typename V2<T>::type C2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int c = threadIdx.x % Cfg::C, t = threadIdx.x / Cfg::C;
    const int col = blockIdx.x * Cfg::C + c;
    const bool valid = col < Wc;
    T* s_re = smem + c * Cfg::LD;
    T* s_im = smem + (Cfg::C + c) * Cfg::LD;
    const int o = blockIdx.y;
    const int tile = perScale ? (o % S) : (o / S);
    const size_t plane = (size_t)N * Wc;
    T re[16], im[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        re[q] = 0;
        im[q] = 0;
    }
    if (valid) {
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const C2* fm = fmaps + ((size_t)tile * F + f) * plane + col;
            const T* fl = filt + ((size_t)o * F + f) * plane + col;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const size_t off = (size_t)(t + q * Cfg::NT) * Wc;
                const C2 v = fm[off];
                const T g = fl[off];
                re[q] += v.x * g;
                im[q] += v.y * g;
            }
        }
    }
    fft_run<T, N, +1>(re, im, t, s_re, s_im, tw, BlockSync());
    if (valid) {
        const size_t base = (size_t)o * plane + col;
#pragma unroll
        for (int q = 0; q < 16; ++q) out[base + (size_t)(t + q * Cfg::NT) * Wc] = mk2<T>(re[q], im[q]);
    }
}

// in place: forward column FFT, multiply by wy[ky] * wx[kx], inverse column FFT (un-normalised: gains a factor N)
template <typename T, int N>
__global__ void __launch_bounds__(ColCfg<T, N>::THREADS, FFT_COL_MINB)
fft_cols_window_kernel(typename V2<T>::type* __restrict__ data, int Wc, const T* __restrict__ wy,
                       const T* __restrict__ wx, const typename V2<T>::type* __restrict__ tw) {
    typedef ColCfg<T, N> Cfg;
    typedef typename V2<T>::type C2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int c = threadIdx.x % Cfg::C, t = threadIdx.x / Cfg::C;
    const int col = blockIdx.x * Cfg::C + c;
    const bool valid = col < Wc;
    T* s_re = smem + c * Cfg::LD;
    T* s_im = smem + (Cfg::C + c) * Cfg::LD;
    const size_t base = (size_t)blockIdx.y * N * Wc + col;
    T re[16], im[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        C2 v = mk2<T>(0, 0);
        if (valid) v = data[base + (size_t)(t + q * Cfg::NT) * Wc];
        re[q] = v.x;
        im[q] = v.y;
    }
    fft_run<T, N, -1>(re, im, t, s_re, s_im, tw, BlockSync());
    const T wxc = valid ? wx[col] : (T)0;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const T w = wy[t + q * Cfg::NT] * wxc;
        re[q] *= w;
        im[q] *= w;
    }
    __syncthreads();
    fft_run<T, N, +1>(re, im, t, s_re, s_im, tw, BlockSync());
    if (valid) {
#pragma unroll
        for (int q = 0; q < 16; ++q) data[base + (size_t)(t + q * Cfg::NT) * Wc] = mk2<T>(re[q], im[q]);
    }
}

// ------------------------------------------------------------------------------------------------------
// row kernels.  W real samples per row, M = W/2 complex points, NT = M/V threads per row.
template <typename T, int W> struct RowCfg {
    static constexpr int V = FFT_ROW_V;
    static constexpr int M = W / 2;
    static constexpr int NT = M / V;
    static constexpr int LD = npad(M) + 2;
    // rows per CTA: aim for 256 threads
    static constexpr int RPB = (256 / NT) > 0 ? (256 / NT) : 1;
    static constexpr int THREADS = RPB * NT;
    static constexpr size_t SMEM = (size_t)2 * RPB * LD * sizeof(T);
};

template <typename T, int W>
__global__ void __launch_bounds__(RowCfg<T, W>::THREADS, FFT_ROW_MINB)
rfft_rows_kernel(const T* __restrict__ in, typename V2<T>::type* __restrict__ out, int H, long long nrows,
                 NbTaper tp, T scale, const typename V2<T>::type* __restrict__ twM,
                 const typename V2<T>::type* __restrict__ twW) {
    typedef RowCfg<T, W> Cfg;
    typedef typename V2<T>::type C2;
    constexpr int M = Cfg::M, NT = Cfg::NT, Wc = M + 1, V = Cfg::V;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int g = threadIdx.x / NT, t = threadIdx.x % NT;
    const long long row = (long long)blockIdx.x * Cfg::RPB + g;
    const bool valid = row < nrows;
    T* s_re = smem + g * Cfg::LD;
    T* s_im = smem + (Cfg::RPB + g) * Cfg::LD;
    T re[V], im[V];
    if (valid) {
        const C2* src = reinterpret_cast<const C2*>(in + (size_t)row * W);
        const int y = (int)(row % H);
        const T ay = (tp.n > 0) ? taper_at<T>(tp, y, H) : (T)1;
#pragma unroll
        for (int q = 0; q < V; ++q) {
            const int n = t + q * NT;
            C2 v = src[n];
            if (tp.n > 0) {
                v.x *= ay * taper_at<T>(tp, 2 * n, W);
                v.y *= ay * taper_at<T>(tp, 2 * n + 1, W);
            }
            re[q] = v.x;
            im[q] = v.y;
        }
    } else {
#pragma unroll
        for (int q = 0; q < V; ++q) {
            re[q] = 0;
            im[q] = 0;
        }
    }
    fft_run<T, M, -1>(re, im, t, s_re, s_im, twM, BlockSync());
    __syncthreads();
    scatter_natural<T, M>(re, im, t, s_re, s_im);
    __syncthreads();
    if (valid) {
        C2* dst = out + (size_t)row * Wc;
        const T h = (T)0.5;
#pragma unroll
        for (int q = 0; q < V; ++q) {
            const int k = t + q * NT;
            const int kc = (M - k) & (M - 1);
            const T zr = re[q], zi = im[q];
            const T cr = s_re[spad(kc)], ci = -s_im[spad(kc)];  // conj(Z[M-k])
            const C2 w = __ldg(twW + k);                         // exp(-2 pi i k / W)
            const T dr = zr - cr, di = zi - ci;
            const T a = w.x * dr - w.y * di, b = w.x * di + w.y * dr;  // w * d
            dst[k] = mk2<T>((h * (zr + cr) + h * b) * scale, (h * (zi + ci) - h * a) * scale);
            if (k == 0) dst[M] = mk2<T>((zr - zi) * scale, (T)0);
        }
    }
}

// epilogue parameters of the c2r kernel
template <typename T> struct C2RParams {
    T scale;                    // constant factor
    const double* norm;         // per-image factor (signalNorm), nullable
    const unsigned char* mask;  // multiply by mask[(img / maskDiv), pix], nullable
    int maskDiv;
    const T* zeroRef;           // output forced to 0 where zeroRef[img, pix] == 0, nullable (may alias out)
};

template <typename T, int W>
__global__ void __launch_bounds__(RowCfg<T, W>::THREADS, FFT_ROW_MINB)
irfft_rows_kernel(const typename V2<T>::type* __restrict__ in, T* out, int H, long long nrows, C2RParams<T> ep,
                  const typename V2<T>::type* __restrict__ twM, const typename V2<T>::type* __restrict__ twW) {
    typedef RowCfg<T, W> Cfg;
    typedef typename V2<T>::type C2;
    constexpr int M = Cfg::M, NT = Cfg::NT, Wc = M + 1, V = Cfg::V;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int g = threadIdx.x / NT, t = threadIdx.x % NT;
    const long long row = (long long)blockIdx.x * Cfg::RPB + g;
    const bool valid = row < nrows;
    T* s_re = smem + g * Cfg::LD;
    T* s_im = smem + (Cfg::RPB + g) * Cfg::LD;
    // every thread reads its 16 spectrum samples X[k] and their mirrors X[M - k] straight from global memory (the mirrors
    // are some other thread's samples of the same row, so they come from L1/L2): 2 V independent loads in flight per thread
    T re[V], im[V];
#pragma unroll
    for (int q = 0; q < V; ++q) {
        re[q] = 0;
        im[q] = 0;
    }
    if (valid) {
        const C2* src = in + (size_t)row * Wc;
        C2 xs[V], cs[V];
#pragma unroll
        for (int q = 0; q < V; ++q) {
            const int k = t + q * NT;
            xs[q] = src[k];
            cs[q] = src[M - k];
        }
#pragma unroll
        for (int q = 0; q < V; ++q) {
            const int k = t + q * NT;
            const T xr = xs[q].x, xi = xs[q].y;
            const T cr = cs[q].x, ci = -cs[q].y;                      // conj(X[M-k])
            const C2 w = __ldg(twW + k);                               // conj -> exp(+2 pi i k / W)
            const T dr = xr - cr, di = xi - ci;
            const T a = w.x * dr + w.y * di, b = w.x * di - w.y * dr;  // conj(w) * d
            re[q] = xr + cr - b;
            im[q] = xi + ci + a;
        }
    }
    __syncthreads();
    fft_run<T, M, +1>(re, im, t, s_re, s_im, twM, BlockSync());
    if (valid) {
        const long long img = row / H;
        T f = ep.scale;
        if (ep.norm) f *= (T)ep.norm[img];
        C2* dst = reinterpret_cast<C2*>(out + (size_t)row * W);
        const unsigned char* mrow = ep.mask ? ep.mask + ((size_t)(img / ep.maskDiv) * H + (row % H)) * W : nullptr;
        const C2* zrow = ep.zeroRef ? reinterpret_cast<const C2*>(ep.zeroRef + (size_t)row * W) : nullptr;
#pragma unroll
        for (int q = 0; q < V; ++q) {
            const int n = t + q * NT;
            T a = re[q] * f, b = im[q] * f;
            if (mrow) {
                a *= (T)mrow[2 * n];
                b *= (T)mrow[2 * n + 1];
            }
            if (zrow) {
                const C2 z = zrow[n];
                if (z.x == (T)0) a = 0;
                if (z.y == (T)0) b = 0;
            }
            dst[n] = mk2<T>(a, b);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// host dispatch
template <typename T> static int prec_of();
template <> int prec_of<float>() { return NB_PREC_F32; }
template <> int prec_of<double>() { return NB_PREC_F64; }

template <typename K> static int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) NB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return NB_OK;
}

#define NB_DISPATCH_N(n, MACRO)                                                             \
    switch (n) {                                                                            \
        case 32: MACRO(32); break;                                                          \
        case 64: MACRO(64); break;                                                          \
        case 128: MACRO(128); break;                                                        \
        case 256: MACRO(256); break;                                                        \
        case 512: MACRO(512); break;                                                        \
        case 1024: MACRO(1024); break;                                                      \
        case 2048: MACRO(2048); break;                                                      \
        case 4096: MACRO(4096); break;                                                      \
        default: nb_set_error("FFT length %d not supported (powers of two 32..4096)", n);   \
                 return NB_ERR_UNSUPPORTED;                                                 \
    }

static bool is_pow2_side(int n) {
    return n == 32 || n == 64 || n == 128 || n == 256 || n == 512 || n == 1024 || n == 2048 || n == 4096;
}

// 1: power of two (direct Stockham passes); 2: any other length in 64..4096 (Bluestein passes, fft_bluestein.cu); 0: no
int nb_fft_size_supported(int n) {
    if (is_pow2_side(n)) return 1;
    return (n >= 64 && n <= 4096) ? 2 : 0;
}

static bool tile_supported(int H, int W) {
    return nb_fft_size_supported(H) != 0 && nb_fft_size_supported(W) != 0;     // odd widths: Bluestein rows, Wc = (W + 1) / 2
}

template <typename T>
static int rfft_rows_t(const T* in, typename V2<T>::type* out, int B, int H, int W, int apodPix, double scale,
                       cudaStream_t st) {
    typedef typename V2<T>::type C2;
    NB_REQUIRE(apodPix >= 0 && apodPix <= NB_MAX_TAPER && 2 * apodPix <= H && 2 * apodPix <= W, NB_ERR_ARG,
               "apodPix %d out of range", apodPix);
    NbTaper tp;
    tp.n = apodPix;
    for (int i = 0; i < apodPix; ++i)
        tp.a[i] = (apodPix == 1) ? 0.0 : 0.5 * (1.0 - cos(M_PI * (double)i / (double)(apodPix - 1)));
    if (apodPix > 0) tp.a[apodPix - 1] = 1.0;  // np.linspace end point is exact; cos(pi) = -1 exactly
    const void *twM, *twW;
    int rc;
    if ((rc = get_pass_twiddles(prec_of<T>(), W / 2, FFT_ROW_V, &twM)) != NB_OK) return rc;
    if ((rc = get_twiddles(prec_of<T>(), W, &twW)) != NB_OK) return rc;
    const long long nrows = (long long)B * H;
#define ROWS_FWD(WW)                                                                                          \
    {                                                                                                         \
        typedef RowCfg<T, WW> Cfg;                                                                            \
        if ((rc = set_smem(rfft_rows_kernel<T, WW>, Cfg::SMEM)) != NB_OK) return rc;                          \
        const unsigned grid = (unsigned)((nrows + Cfg::RPB - 1) / Cfg::RPB);                                  \
        { NB_PROF_BEGIN(st); rfft_rows_kernel<T, WW><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(in, out, H, nrows, tp, (T)scale,      \
                                                                       (const C2*)twM, (const C2*)twW); NB_PROF_END("rfft_rows_kernel", st); }      \
    }
    switch (W) {
        case 64: ROWS_FWD(64); break;
        case 128: ROWS_FWD(128); break;
        case 256: ROWS_FWD(256); break;
        case 512: ROWS_FWD(512); break;
        case 1024: ROWS_FWD(1024); break;
        case 2048: ROWS_FWD(2048); break;
        case 4096: ROWS_FWD(4096); break;
        default: nb_set_error("map width %d not supported (powers of two 64..4096)", W); return NB_ERR_UNSUPPORTED;
    }
#undef ROWS_FWD
    NB_CHECK_LAUNCH();
    return NB_OK;
}

template <typename T>
static int irfft_rows_t(const typename V2<T>::type* in, T* out, int B, int H, int W, C2RParams<T> ep, cudaStream_t st) {
    typedef typename V2<T>::type C2;
    const void *twM, *twW;
    int rc;
    if ((rc = get_pass_twiddles(prec_of<T>(), W / 2, FFT_ROW_V, &twM)) != NB_OK) return rc;
    if ((rc = get_twiddles(prec_of<T>(), W, &twW)) != NB_OK) return rc;
    const long long nrows = (long long)B * H;
#define ROWS_INV(WW)                                                                                          \
    {                                                                                                         \
        typedef RowCfg<T, WW> Cfg;                                                                            \
        if ((rc = set_smem(irfft_rows_kernel<T, WW>, Cfg::SMEM)) != NB_OK) return rc;                         \
        const unsigned grid = (unsigned)((nrows + Cfg::RPB - 1) / Cfg::RPB);                                  \
        { NB_PROF_BEGIN(st); irfft_rows_kernel<T, WW><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(in, out, H, nrows, ep, (const C2*)twM, \
                                                                        (const C2*)twW); NB_PROF_END("irfft_rows_kernel", st); }                     \
    }
    switch (W) {
        case 64: ROWS_INV(64); break;
        case 128: ROWS_INV(128); break;
        case 256: ROWS_INV(256); break;
        case 512: ROWS_INV(512); break;
        case 1024: ROWS_INV(1024); break;
        case 2048: ROWS_INV(2048); break;
        case 4096: ROWS_INV(4096); break;
        default: nb_set_error("map width %d not supported (powers of two 64..4096)", W); return NB_ERR_UNSUPPORTED;
    }
#undef ROWS_INV
    NB_CHECK_LAUNCH();
    return NB_OK;
}

template <typename T>
static int fft_cols_t(const typename V2<T>::type* in, typename V2<T>::type* out, int B, int H, int Wc, int dir,
                      double scale, cudaStream_t st) {
    typedef typename V2<T>::type C2;
    const void* tw;
    int rc;
    if ((rc = get_pass_twiddles(prec_of<T>(), H, 16, &tw)) != NB_OK) return rc;
#define COLS(NN)                                                                                             \
    {                                                                                                        \
        typedef ColCfg<T, NN> Cfg;                                                                           \
        dim3 grid((Wc + Cfg::C - 1) / Cfg::C, B);                                                            \
        if (dir < 0) {                                                                                       \
            if ((rc = set_smem(fft_cols_kernel<T, NN, -1>, Cfg::SMEM)) != NB_OK) return rc;                  \
            { NB_PROF_BEGIN(st); fft_cols_kernel<T, NN, -1><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(in, out, Wc, (const C2*)tw, (T)scale); NB_PROF_END("fft_cols_kernel", st); } \
        } else {                                                                                             \
            if ((rc = set_smem(fft_cols_kernel<T, NN, +1>, Cfg::SMEM)) != NB_OK) return rc;                  \
            { NB_PROF_BEGIN(st); fft_cols_kernel<T, NN, +1><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(in, out, Wc, (const C2*)tw, (T)scale); NB_PROF_END("fft_cols_kernel", st); } \
        }                                                                                                    \
    }
    NB_DISPATCH_N(H, COLS)
#undef COLS
    NB_CHECK_LAUNCH();
    return NB_OK;
}

template <typename T, int F>
static int fft_cols_apply_tf(const typename V2<T>::type* fmaps, const T* filt, typename V2<T>::type* out, int nOut,
                             int S, int perScale, int H, int Wc, cudaStream_t st) {
    typedef typename V2<T>::type C2;
    const void* tw;
    int rc;
    if ((rc = get_pass_twiddles(prec_of<T>(), H, 16, &tw)) != NB_OK) return rc;
#define COLS_APPLY(NN)                                                                                       \
    {                                                                                                        \
        typedef ColCfg<T, NN> Cfg;                                                                           \
        dim3 grid((Wc + Cfg::C - 1) / Cfg::C, nOut);                                                         \
        if ((rc = set_smem(fft_cols_apply_kernel<T, NN, F>, Cfg::SMEM)) != NB_OK) return rc;                 \
        { NB_PROF_BEGIN(st); fft_cols_apply_kernel<T, NN, F><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(fmaps, filt, out, Wc, S, perScale, (const C2*)tw); NB_PROF_END("fft_cols_apply_kernel", st); } \
    }
    NB_DISPATCH_N(H, COLS_APPLY)
#undef COLS_APPLY
    NB_CHECK_LAUNCH();
    return NB_OK;
}

template <typename T>
static int fft_cols_window_t(typename V2<T>::type* data, int B, int H, int Wc, const T* wy, const T* wx, cudaStream_t st) {
    typedef typename V2<T>::type C2;
    const void* tw;
    int rc;
    if ((rc = get_pass_twiddles(prec_of<T>(), H, 16, &tw)) != NB_OK) return rc;
#define COLS_WIN(NN)                                                                                         \
    {                                                                                                        \
        typedef ColCfg<T, NN> Cfg;                                                                           \
        dim3 grid((Wc + Cfg::C - 1) / Cfg::C, B);                                                            \
        if ((rc = set_smem(fft_cols_window_kernel<T, NN>, Cfg::SMEM)) != NB_OK) return rc;                   \
        { NB_PROF_BEGIN(st); fft_cols_window_kernel<T, NN><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(data, Wc, wy, wx, (const C2*)tw); NB_PROF_END("fft_cols_window_kernel", st); } \
    }
    NB_DISPATCH_N(H, COLS_WIN)
#undef COLS_WIN
    NB_CHECK_LAUNCH();
    return NB_OK;
}

// ------------------------------------------------------------------------------------------------------
// C ABI
extern "C" {

// row / column passes by side: direct when the side is a power of two, Bluestein otherwise
static bool use_sfft(int n) {
    static const bool off = getenv("NEMO_B200_NO_SFFT") != nullptr;      // the run-time-plan passes of fft_bluestein*.cu instead
    return !off && nb_sfft_supported(n) != 0;
}

static int rows_fwd(int prec, const void* in, void* out, int B, int H, int W, int apodPix, double scale, cudaStream_t st) {
    if (use_sfft(W)) return nb_sfft_rows_fwd(prec, in, out, B, H, W, apodPix, scale, st);
    if (!is_pow2_side(W)) return nb_bs_rows_fwd(prec, in, out, B, H, W, apodPix, scale, st);
    if (prec == NB_PREC_F64) return rfft_rows_t<double>((const double*)in, (double2*)out, B, H, W, apodPix, scale, st);
    return rfft_rows_t<float>((const float*)in, (float2*)out, B, H, W, apodPix, scale, st);
}

// zeroFlags (may be null): [B] int32, 0 for an image in which out + (1 - mask) is nowhere exactly zero -- the pixels K7's edge
// filter looks for, which it then need not search (nb_snr_mask*).  Only the row kernels with a compile-time plan report it, as a
// by-product of their store; every other path answers 1 (unknown) for all images.
static int rows_inv(int prec, const void* in, void* out, int B, int H, int W, double scale, const double* norm, const unsigned char* mask,
                    int maskDiv, const void* zeroRef, cudaStream_t st, int* zeroFlags = nullptr) {
    if (maskDiv < 1) maskDiv = 1;
    const bool flagsHere = zeroFlags && mask && use_sfft(W);
    if (zeroFlags) NB_CHECK_CUDA(cudaMemsetAsync(zeroFlags, flagsHere ? 0 : 1, sizeof(int) * (size_t)B, st));
    if (use_sfft(W)) return nb_sfft_rows_inv(prec, in, out, B, H, W, scale, norm, mask, maskDiv, zeroRef, flagsHere ? zeroFlags : nullptr, st);
    if (!is_pow2_side(W)) return nb_bs_rows_inv(prec, in, out, B, H, W, scale, norm, mask, maskDiv, zeroRef, st);
    if (prec == NB_PREC_F64) {
        C2RParams<double> ep = {scale, norm, mask, maskDiv, (const double*)zeroRef};
        return irfft_rows_t<double>((const double2*)in, (double*)out, B, H, W, ep, st);
    }
    C2RParams<float> ep = {(float)scale, norm, mask, maskDiv, (const float*)zeroRef};
    return irfft_rows_t<float>((const float2*)in, (float*)out, B, H, W, ep, st);
}

static int cols_plain(int prec, const void* in, void* out, int B, int H, int Wc, int dir, cudaStream_t st) {
    if (use_sfft(H)) {
        if (prec == NB_PREC_F64) {
            const int rc = nb_sfft_cols_tma_f64(in, out, B, H, Wc, dir, st);
            if (rc != 1) return rc;
        }
        return prec == NB_PREC_F64 ? nb_sfft_cols_f64(in, out, B, H, Wc, dir, st) : nb_sfft_cols_f32(in, out, B, H, Wc, dir, st);
    }
    if (!is_pow2_side(H)) return nb_bs_cols(prec, in, out, B, H, Wc, dir, st);
    if (prec == NB_PREC_F64) return fft_cols_t<double>((const double2*)in, (double2*)out, B, H, Wc, dir, 1.0, st);
    return fft_cols_t<float>((const float2*)in, (float2*)out, B, H, Wc, dir, 1.0, st);
}

int nb_rfft2(int prec, const void* in, void* out, int B, int H, int W, int apodPix, double scale, cudaStream_t st) {
    NB_REQUIRE(in && out && B > 0, NB_ERR_ARG, "nb_rfft2: bad arguments");
    NB_REQUIRE(prec == NB_PREC_F64 || prec == NB_PREC_F32, NB_ERR_ARG, "nb_rfft2: bad precision %d", prec);
    NB_REQUIRE(tile_supported(H, W), NB_ERR_UNSUPPORTED, "nb_rfft2: tile %dx%d not supported (sides 64..4096)", H, W);
    int rc;
    if ((rc = rows_fwd(prec, in, out, B, H, W, apodPix, scale, st)) != NB_OK) return rc;
    return cols_plain(prec, out, out, B, H, W / 2 + 1, -1, st);
}

int nb_irfft2(int prec, void* spec, void* out, int B, int H, int W, double scale, const double* norm,
              const unsigned char* mask, int maskDiv, cudaStream_t st) {
    NB_REQUIRE(spec && out && B > 0, NB_ERR_ARG, "nb_irfft2: bad arguments");
    NB_REQUIRE(prec == NB_PREC_F64 || prec == NB_PREC_F32, NB_ERR_ARG, "nb_irfft2: bad precision %d", prec);
    NB_REQUIRE(tile_supported(H, W), NB_ERR_UNSUPPORTED, "nb_irfft2: tile %dx%d not supported (sides 64..4096)", H, W);
    int rc;
    if ((rc = cols_plain(prec, spec, spec, B, H, W / 2 + 1, +1, st)) != NB_OK) return rc;
    return rows_inv(prec, spec, out, B, H, W, scale, norm, mask, maskDiv, nullptr, st);
}

int nb_apply_filter(int prec, const void* fmaps, const void* filtH, void* scratch, void* out, int nOut, int S, int perScale,
                    int F, int H, int W, double scale, const double* norm, const unsigned char* mask, int* zeroFlags, cudaStream_t st) {
    NB_REQUIRE(fmaps && filtH && scratch && out && nOut > 0 && S > 0, NB_ERR_ARG, "nb_apply_filter: bad arguments");
    NB_REQUIRE(prec == NB_PREC_F64 || prec == NB_PREC_F32, NB_ERR_ARG, "nb_apply_filter: bad precision %d", prec);
    NB_REQUIRE(F >= 1 && F <= NB_MAX_FREQ, NB_ERR_UNSUPPORTED, "nb_apply_filter: %d frequencies not supported (1..%d)", F, NB_MAX_FREQ);
    NB_REQUIRE(tile_supported(H, W), NB_ERR_UNSUPPORTED, "nb_apply_filter: tile %dx%d not supported (sides 64..4096)", H, W);
    int rc;
    const int Wc = W / 2 + 1;
    if (use_sfft(H)) {
        rc = 1;
        if (prec == NB_PREC_F64) rc = nb_sfft_cols_apply_tma_f64(fmaps, filtH, scratch, nOut, S, perScale, F, H, Wc, st);
        if (rc == 1)
            rc = prec == NB_PREC_F64 ? nb_sfft_cols_apply_f64(fmaps, filtH, scratch, nOut, S, perScale, F, H, Wc, st)
                                     : nb_sfft_cols_apply_f32(fmaps, filtH, scratch, nOut, S, perScale, F, H, Wc, st);
    } else if (!is_pow2_side(H)) {
        rc = nb_bs_cols_apply(prec, fmaps, filtH, scratch, nOut, S, perScale, F, H, Wc, st);
    } else {
#define APPLY_F(T, C2T, FF) \
    rc = fft_cols_apply_tf<T, FF>((const C2T*)fmaps, (const T*)filtH, (C2T*)scratch, nOut, S, perScale, H, Wc, st)
#define APPLY_ALL(T, C2T)                   \
    switch (F) {                            \
        case 1: APPLY_F(T, C2T, 1); break;  \
        case 2: APPLY_F(T, C2T, 2); break;  \
        case 3: APPLY_F(T, C2T, 3); break;  \
        default: APPLY_F(T, C2T, 4); break; \
    }
        if (prec == NB_PREC_F64) {
            APPLY_ALL(double, double2)
        } else {
            APPLY_ALL(float, float2)
        }
#undef APPLY_ALL
#undef APPLY_F
    }
    if (rc != NB_OK) return rc;
    return rows_inv(prec, scratch, out, nOut, H, W, scale, norm, mask, S, nullptr, st, zeroFlags);
}

int nb_window_deconv(int prec, void* data, void* scratch, int B, int H, int W, const void* wy, const void* wx,
                     int restoreZeros, cudaStream_t st) {
    NB_REQUIRE(data && scratch && wy && wx && B > 0, NB_ERR_ARG, "nb_window_deconv: bad arguments");
    NB_REQUIRE(prec == NB_PREC_F64 || prec == NB_PREC_F32, NB_ERR_ARG, "nb_window_deconv: bad precision %d", prec);
    NB_REQUIRE(tile_supported(H, W), NB_ERR_UNSUPPORTED, "nb_window_deconv: tile %dx%d not supported (sides 64..4096)", H, W);
    int rc;
    const int Wc = W / 2 + 1;
    const double scale = 1.0 / ((double)H * (double)W);
    static const bool separable = !(getenv("NEMO_B200_WINDOW_SEPARABLE") && atoi(getenv("NEMO_B200_WINDOW_SEPARABLE")) == 0);
    if (separable && use_sfft(H) && use_sfft(W) && W % 2 == 0) {
        // the window is a product wy(ky) wx(kx): a filter along y on pairs of adjacent real columns (data -> scratch), then one
        // along x on pairs of real rows (scratch -> data, zeros of the old data restored); see sfft_rows_window_kernel
        rc = prec == NB_PREC_F64 ? nb_sfft_cols_window_io_f64(data, scratch, B, H, W / 2, wy, nullptr, st)
                                 : nb_sfft_cols_window_io_f32(data, scratch, B, H, W / 2, wy, nullptr, st);
        if (rc != NB_OK) return rc;
        return nb_sfft_rows_window(prec, scratch, data, B, H, W, wx, scale, restoreZeros, st);
    }
    if ((rc = rows_fwd(prec, data, scratch, B, H, W, 0, 1.0, st)) != NB_OK) return rc;
    if (use_sfft(H)) {
        rc = 1;
        if (prec == NB_PREC_F64) rc = nb_sfft_cols_window_tma_f64(scratch, B, H, Wc, wy, wx, st);
        if (rc == 1) rc = prec == NB_PREC_F64 ? nb_sfft_cols_window_f64(scratch, B, H, Wc, wy, wx, st) : nb_sfft_cols_window_f32(scratch, B, H, Wc, wy, wx, st);
    } else if (!is_pow2_side(H)) {
        rc = nb_bs_cols_window(prec, scratch, B, H, Wc, wy, wx, st);
    } else if (prec == NB_PREC_F64) {
        rc = fft_cols_window_t<double>((double2*)scratch, B, H, Wc, (const double*)wy, (const double*)wx, st);
    } else {
        rc = fft_cols_window_t<float>((float2*)scratch, B, H, Wc, (const float*)wy, (const float*)wx, st);
    }
    if (rc != NB_OK) return rc;
    return rows_inv(prec, scratch, data, B, H, W, scale, nullptr, nullptr, 1, restoreZeros ? data : nullptr, st);
}

}  // extern "C"
