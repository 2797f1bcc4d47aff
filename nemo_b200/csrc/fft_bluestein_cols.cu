// Column passes of the any-size 2-D real FFT (see fft_bluestein.cu for the method and fft_anysize.cuh for the transforms): a
// translation unit of its own so that the two halves compile side by side.
#include "fft_anysize.cuh"

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
    // column stride == BANKS / C (mod BANKS) elements, BANKS = 128 bytes / sizeof(T): the C columns x BANKS / C consecutive
    // positions of one access phase then fall on distinct banks (as ColCfg in fft_kernels.cu)
    static constexpr int BANKS = 128 / (int)sizeof(T);
    static constexpr int TV = (BANKS / C) > 0 ? (BANKS / C) : 1;
    static constexpr int LD = npad(M) + 1 + ((TV - (npad(M) + 1)) % BANKS + BANKS) % BANKS;
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

template <typename T, int MODE, int DIR, int F>
static int bs_cols_t(const typename V2<T>::type* in, const T* filt, typename V2<T>::type* out, int nOut, int S, int perScale, int H, int Wc,
                     const T* wy, const T* wx, cudaStream_t st) {
    typedef typename V2<T>::type C2;
    BsTables tb;
    int rc;
    if ((rc = nb_get_bs_tables(sizeof(T) == 8 ? NB_PREC_F64 : NB_PREC_F32, H, &tb)) != NB_OK) return rc;
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

