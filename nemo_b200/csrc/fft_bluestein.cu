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
#include "fft_anysize.cuh"
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

// ------------------------------------------------------------------------------------------------------
// tables: tw[m] = exp(-2 pi i m / M); chirp[n] = exp(-i pi n^2 / N), n < N; bhat = FFT_M(b) / M with b[n] = conj(chirp[|n|])
// wrapped onto M (the 1/M of the inverse M-point transform is folded in here)

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

int nb_get_bs_tables(int prec, int N, BsTables* out) {
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
        // few passes: every pass costs a block barrier and a trip through shared memory
        for (int f : {7, 5, 9, 3}) while (rest % f == 0) { radices.push_back(f); rest /= f; }
        while (rest % 8 == 0) { radices.push_back(8); rest /= 8; }
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
__global__ void __launch_bounds__(BsRowCfg<T, M>::THREADS, 2)   // two CTAs per SM: left alone, ptxas takes 255 registers here
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
__global__ void __launch_bounds__(BsRowCfg<T, M>::THREADS, 2)   // two CTAs per SM: left alone, ptxas takes 255 registers here
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
    if ((rc = nb_get_bs_tables(sizeof(T) == 8 ? NB_PREC_F64 : NB_PREC_F32, W, &tb)) != NB_OK) return rc;
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
    if ((rc = nb_get_bs_tables(sizeof(T) == 8 ? NB_PREC_F64 : NB_PREC_F32, W, &tb)) != NB_OK) return rc;
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
