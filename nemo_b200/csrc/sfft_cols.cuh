// Column passes of the 2-D real FFT for heights with a compile-time mixed-radix plan (sfft.cuh).  A CTA owns C adjacent
// columns of one image for all H rows (C x 16 contiguous bytes per row in fp64), NT threads per column.
//   MODE 0: plain transform in -> out (DIR)
//   MODE 1: out[o] = inverse transform of sum_f fmaps[m(o), f] * filt[o, f]   (K4: filters.py:851)
//   MODE 2: in place: forward, times wy[ky] wx[kx], inverse, the data never leaving registers between the two (K8: :103); the
//           inverse runs the reversed plan, whose first pass wants exactly what the forward's last pass left in each thread
#include "sfft_host.cuh"

using namespace sfft;

template <typename T, int N, int C> struct SColCfg {
    // column stride == BANKS / C (mod BANKS) elements, BANKS = 128 bytes / sizeof(T): the C columns x BANKS / C consecutive
    // positions of one access phase fall on distinct banks
    static constexpr int BANKS = 128 / (int)sizeof(T);
    static constexpr int TV = (BANKS / C) > 0 ? (BANKS / C) : 1;
    static constexpr int LD = N + ((TV - N) % BANKS + BANKS) % BANKS;
    static constexpr size_t SMEM = (size_t)2 * C * LD * sizeof(T);
};

template <typename T, class PL, int NT, int C, int MODE, int DIR, int F>
__global__ void __launch_bounds__(NT * C, SFFT_COL_MINB)
sfft_cols_kernel(const typename V2<T>::type* __restrict__ in, const T* __restrict__ filt, typename V2<T>::type* __restrict__ out,
                 int Wc, int S, int perScale, const T* __restrict__ wy, const T* __restrict__ wx,
                 const typename V2<T>::type* __restrict__ tw, const typename V2<T>::type* __restrict__ twRev) {
    typedef typename V2<T>::type C2;
    typedef SColCfg<T, PL::N, C> Cfg;
    typedef Stage<PL, NT, 0> S0;
    typedef typename Last<PL, NT>::S SL;
    typedef typename Reversed<PL>::type PR;
    constexpr int H = PL::N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int c = threadIdx.x % C, t = threadIdx.x / C;
    int cb = blockIdx.x, o = blockIdx.y;
    if (MODE == 1 && S > 1 && gridDim.y % S == 0) {
        // the images that multiply the same spectra (perScale: the same template for every tile; else the S scales of a tile)
        // take the same column block one after the other in launch order: the spectra columns come from DRAM once and from L2
        // for the others.  In image-major order the S scales of a tile re-read its spectra from DRAM (measured 3.7 GB of reads
        // per 128 images against 2.6 GB algorithmic).
        const unsigned L = blockIdx.y * gridDim.x + blockIdx.x;
        const unsigned nG = gridDim.y / S;
        if (perScale) {
            const unsigned tileIdx = L % nG, rest = L / nG;
            cb = rest % gridDim.x;
            o = tileIdx * S + rest / gridDim.x;
        } else {
            const unsigned s = L % S, rest = L / S;
            cb = rest % gridDim.x;
            o = (rest / gridDim.x) * S + s;
        }
    }
    const int col = cb * C + c;
    const bool valid = col < Wc;
    T* s_re = smem + c * Cfg::LD;
    T* s_im = smem + (C + c) * Cfg::LD;
    const size_t plane = (size_t)H * Wc;
    T xr[S0::REGS], xi[S0::REGS];
#pragma unroll
    for (int q = 0; q < S0::REGS; ++q) {
        xr[q] = 0;
        xi[q] = 0;
    }
    if (valid) {
        if (MODE == 1) {
            const int tile = perScale ? (o % S) : (o / S);
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const C2* fm = in + ((size_t)tile * F + f) * plane + col;
                const T* fl = filt + ((size_t)o * F + f) * plane + col;
#pragma unroll
                for (int i = 0; i < S0::IT; ++i) {
                    const int j = t + i * NT;
                    if (S0::IT * NT > S0::NB && j >= S0::NB) continue;
#pragma unroll
                    for (int r = 0; r < S0::R; ++r) {
                        const size_t n = (size_t)(j + r * S0::NB);
                        const C2 v = fm[n * Wc];
                        const T gg = fl[n * Wc];
                        xr[i * S0::R + r] += v.x * gg;
                        xi[i * S0::R + r] += v.y * gg;
                    }
                }
            }
        } else {
            const C2* src = in + (size_t)o * plane + col;
#pragma unroll
            for (int i = 0; i < S0::IT; ++i) {
                const int j = t + i * NT;
                if (S0::IT * NT > S0::NB && j >= S0::NB) continue;
#pragma unroll
                for (int r = 0; r < S0::R; ++r) {
                    const C2 v = src[(size_t)(j + r * S0::NB) * Wc];
                    xr[i * S0::R + r] = v.x;
                    xi[i * S0::R + r] = v.y;
                }
            }
        }
    }
    C2* dst = out + (size_t)o * plane + col;
    if constexpr (MODE == 2) {
        typedef typename Last<PR, NT>::S SRL;
        run_from<PL, T, NT, -1, 0>(xr, xi, t, s_re, s_im, tw, BlockSync(), [&](auto& yr, auto& yi) {
            const T wxc = valid ? (wx ? wx[col] : (T)1) : (T)0;      // wx == nullptr: a window in y only (separable undo)
#pragma unroll
            for (int i = 0; i < SL::IT; ++i) {
                const int j = t + i * NT;
                if (SL::IT * NT > SL::NB && j >= SL::NB) continue;
#pragma unroll
                for (int r = 0; r < SL::R; ++r) {
                    const T w = wy[j + r * SL::NB] * wxc;
                    yr[i * SL::R + r] *= w;
                    yi[i * SL::R + r] *= w;
                }
            }
            run_from<PR, T, NT, +1, 0>(yr, yi, t, s_re, s_im, twRev, BlockSync(), [&](auto& zr, auto& zi) {
                if (!valid) return;
#pragma unroll
                for (int i = 0; i < SRL::IT; ++i) {
                    const int j = t + i * NT;
                    if (SRL::IT * NT > SRL::NB && j >= SRL::NB) continue;
#pragma unroll
                    for (int r = 0; r < SRL::R; ++r) dst[(size_t)(j + r * SRL::NB) * Wc] = mk2<T>(zr[i * SRL::R + r], zi[i * SRL::R + r]);
                }
            });
        });
    } else {
        run_from<PL, T, NT, (MODE == 1 ? +1 : DIR), 0>(xr, xi, t, s_re, s_im, tw, BlockSync(), [&](auto& yr, auto& yi) {
            if (!valid) return;
#pragma unroll
            for (int i = 0; i < SL::IT; ++i) {
                const int j = t + i * NT;
                if (SL::IT * NT > SL::NB && j >= SL::NB) continue;
#pragma unroll
                for (int r = 0; r < SL::R; ++r) dst[(size_t)(j + r * SL::NB) * Wc] = mk2<T>(yr[i * SL::R + r], yi[i * SL::R + r]);
            }
        });
    }
}

template <typename T, int MODE, int DIR, int F>
static int sfft_cols_t(const typename V2<T>::type* in, const T* filt, typename V2<T>::type* out, int nOut, int S, int perScale, int H, int Wc,
                       const T* wy, const T* wx, cudaStream_t st) {
    typedef typename V2<T>::type C2;
    const int prec = sizeof(T) == 8 ? NB_PREC_F64 : NB_PREC_F32;
    constexpr int C = SFFT_COL_C;
    int rc;
    const void *tw, *twRev = nullptr;
#define SFFT_CASE(NN, RR0, RR1, RR2, RNT, R0, R1, R2, NT)                                                                              \
    if (H == NN) {                                                                                                        \
        typedef Plan<NN, R0, R1, R2> PL;                                                                                  \
        typedef SColCfg<T, NN, C> Cfg;                                                                                    \
        if ((rc = sfft_tw<PL>(prec, &tw)) != NB_OK) return rc;                                                            \
        if (MODE == 2 && (rc = sfft_tw<typename Reversed<PL>::type>(prec, &twRev)) != NB_OK) return rc;                   \
        if ((rc = sfft_set_smem(sfft_cols_kernel<T, PL, NT, C, MODE, DIR, F>, Cfg::SMEM)) != NB_OK) return rc;            \
        dim3 grid((Wc + C - 1) / C, nOut);                                                                                \
        { NB_PROF_BEGIN(st); sfft_cols_kernel<T, PL, NT, C, MODE, DIR, F><<<grid, NT * C, Cfg::SMEM, st>>>(in, filt, out, Wc, S, perScale, wy, wx, (const C2*)tw, (const C2*)twRev); NB_PROF_END("sfft_cols_kernel", st); } \
        NB_CHECK_LAUNCH();                                                                                                \
        return NB_OK;                                                                                                     \
    }
    SFFT_FOR_SIZES(SFFT_CASE)
#undef SFFT_CASE
    nb_set_error("sfft: height %d has no compile-time plan", H);
    return NB_ERR_UNSUPPORTED;
}

#define SFFT_COLS_ENTRY(T, C2T, SUFFIX)                                                                                                   \
    int nb_sfft_cols_##SUFFIX(const void* in, void* out, int B, int H, int Wc, int dir, cudaStream_t st) {                                \
        if (dir < 0) return sfft_cols_t<T, 0, -1, 1>((const C2T*)in, nullptr, (C2T*)out, B, 1, 0, H, Wc, nullptr, nullptr, st);            \
        return sfft_cols_t<T, 0, +1, 1>((const C2T*)in, nullptr, (C2T*)out, B, 1, 0, H, Wc, nullptr, nullptr, st);                         \
    }                                                                                                                                     \
    int nb_sfft_cols_apply_##SUFFIX(const void* fmaps, const void* filt, void* out, int nOut, int S, int perScale, int F, int H, int Wc,  \
                                    cudaStream_t st) {                                                                                    \
        switch (F) {                                                                                                                      \
            case 1: return sfft_cols_t<T, 1, +1, 1>((const C2T*)fmaps, (const T*)filt, (C2T*)out, nOut, S, perScale, H, Wc, nullptr, nullptr, st); \
            case 2: return sfft_cols_t<T, 1, +1, 2>((const C2T*)fmaps, (const T*)filt, (C2T*)out, nOut, S, perScale, H, Wc, nullptr, nullptr, st); \
            case 3: return sfft_cols_t<T, 1, +1, 3>((const C2T*)fmaps, (const T*)filt, (C2T*)out, nOut, S, perScale, H, Wc, nullptr, nullptr, st); \
            default: return sfft_cols_t<T, 1, +1, 4>((const C2T*)fmaps, (const T*)filt, (C2T*)out, nOut, S, perScale, H, Wc, nullptr, nullptr, st); \
        }                                                                                                                                 \
    }                                                                                                                                     \
    int nb_sfft_cols_window_##SUFFIX(void* data, int B, int H, int Wc, const void* wy, const void* wx, cudaStream_t st) {                 \
        return sfft_cols_t<T, 2, -1, 1>((const C2T*)data, nullptr, (C2T*)data, B, 1, 0, H, Wc, (const T*)wy, (const T*)wx, st);            \
    }                                                                                                                                     \
    int nb_sfft_cols_window_io_##SUFFIX(const void* in, void* out, int B, int H, int Wc, const void* wy, const void* wx, cudaStream_t st) { \
        return sfft_cols_t<T, 2, -1, 1>((const C2T*)in, nullptr, (C2T*)out, B, 1, 0, H, Wc, (const T*)wy, (const T*)wx, st);               \
    }
