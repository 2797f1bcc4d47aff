// Column passes of the 2-D real FFT fed by the TMA unit (fp64; heights with a compile-time mixed-radix plan, sfft.cuh).
//
// The register-staged kernels of sfft_cols.cuh spend most of a tile's time in its two global-memory phases: measured per image
// of 840 x 721 columns, a load + transform + store pass takes 5.8 us, the transform alone 2.4 us (second transform of the
// window pass, which never leaves registers), and the filter application, which loads three times as much, 8.4 us.  Here a CTA
// is persistent and its column tiles (C adjacent columns x H rows = a 2-D box of the [rows, 2 Wc] tensor of doubles) arrive by
// cp.async.bulk.tensor (SASS UTMALDG) in a staging buffer, completion tracked by an mbarrier; the buffer goes back to the copy
// unit as soon as the first pass has gathered from it, so the next tile streams in under the passes and stores of this one.
//
//   sfft_cols_tma_kernel<MODE 0 | 2>   plain transform / window pass (forward, x wy wx, inverse), two CTAs of C x NT threads per SM
//   sfft_cols_apply_tma_kernel<F>      out[o] = inverse transform of sum_f spectra[m(o), f] * filt[o, f]  (K4, filters.py:851).
//       A CTA owns a column block of ONE set of spectra (a tile's maps; with perScale a scale's templates) and walks over the
//       images that use it (the tile's S scales; the nT tiles): the F spectra tiles are fetched once and stay in shared
//       memory, only the filter tile (F x H x C reals, Ampere-style cp.async since its rows are only 8-byte aligned) streams per
//       image.  107 + 54 + 54 KB of shared memory: one CTA of 480 threads per SM, every global load asynchronous.
#include "sfft_host.cuh"
#include "tma.cuh"

using namespace sfft;

constexpr int sfft_box_rows(int H) {      // largest divisor of H that a TMA box can have (<= 256 rows)
    int best = 1;
    for (int d = 1; d <= 256; ++d)
        if (H % d == 0) best = d;
    return best;
}

template <typename T, int N, int C> struct SColTmaCfg {
    static constexpr int BANKS = 128 / (int)sizeof(T);
    static constexpr int TV = (BANKS / C) > 0 ? (BANKS / C) : 1;
    static constexpr int LD = N + ((TV - N) % BANKS + BANKS) % BANKS;        // as SColCfg: conflict-free column stride
    static constexpr size_t EXCH = (size_t)2 * C * LD * sizeof(T);
    static constexpr size_t TILE = (size_t)N * C * 2 * sizeof(T);            // one spectra tile [N][C] complex
    static constexpr int BOXR = sfft_box_rows(N);
    static constexpr int NBOX = N / BOXR;
};

__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(tma::smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------------
// (the twiddle pointers are laundered through an empty asm statement in every iteration: otherwise the compiler hoists the
// loop-invariant twiddle LOADS out of the persistent loop and keeps dozens of them in registers -- 1100 bytes of spills)
template <typename P> __device__ __forceinline__ P opaque_ptr(P p) {
    asm volatile("" : "+l"(p));
    return p;
}

// Plain transform (MODE 0) / window pass (MODE 2).  Like the filter application below it runs the plan with one butterfly per
// thread and pass (NT about H / 7: 480 threads for four columns of 840 rows, at most 12 values per thread) as ONE CTA per SM:
// with the loads asynchronous nothing is gained from a second CTA that loads while the first computes, and the 16-values-per-
// thread plan of the register-staged kernel, which needs exactly the 128 registers two CTAs per SM allow, spilled 500 to 1200
// bytes per thread as soon as the persistent loop was wrapped around it.  NSTG staging buffers, each with its own mbarrier:
// while tile k is transformed, tile k + 1 has landed or is landing and tile k + 2 has just been requested.
template <typename T, class PL, int NT, int C, int MODE, int DIR, int NSTG>
__global__ void __launch_bounds__(NT * C, 1)
sfft_cols_tma_kernel(const __grid_constant__ CUtensorMap tmIn, typename V2<T>::type* __restrict__ out, int Wc, int nImg,
                     const T* __restrict__ wy0, const T* __restrict__ wx, const typename V2<T>::type* __restrict__ tw0,
                     const typename V2<T>::type* __restrict__ twRev0) {
    typedef typename V2<T>::type C2;
    typedef SColTmaCfg<T, PL::N, C> Cfg;
    typedef Stage<PL, NT, 0> S0;
    typedef typename Last<PL, NT>::S SL;
    typedef typename Reversed<PL>::type PR;
    constexpr int H = PL::N;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* exch = reinterpret_cast<T*>(smem_raw + NSTG * Cfg::TILE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + NSTG * Cfg::TILE + Cfg::EXCH);
    const int c = threadIdx.x % C, t = threadIdx.x / C;
    T* s_re = exch + c * Cfg::LD;
    T* s_im = exch + (C + c) * Cfg::LD;
    const int nCB = (Wc + C - 1) / C;
    const int total = nImg * nCB;
    const size_t plane = (size_t)H * Wc;

    auto issue = [&](int L, int sidx) {
        const int o = L / nCB, cb = L - o * nCB;
        C2* stg = reinterpret_cast<C2*>(smem_raw + (size_t)sidx * Cfg::TILE);
        tma::mbar_arrive_expect_tx(bars + sidx, (uint32_t)Cfg::TILE);
#pragma unroll
        for (int b = 0; b < Cfg::NBOX; ++b)
            tma::tma_load_2d(stg + (size_t)b * Cfg::BOXR * C, &tmIn, cb * C * 2, o * H + b * Cfg::BOXR, bars + sidx);
    };
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NSTG; ++q) tma::mbar_init(bars + q, 1);
        tma::fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NSTG; ++q) {
            const int Lq = (int)blockIdx.x + q * (int)gridDim.x;
            if (Lq < total) issue(Lq, q);
        }
    }
    int sidx = 0;
    uint32_t phase = 0;        // parity of the current use of buffer sidx (all buffers are on the same round)
    for (int L = (int)blockIdx.x; L < total; L += (int)gridDim.x) {
        const int o = L / nCB, cb = L - o * nCB;
        const int col = cb * C + c;
        const bool valid = col < Wc;
        // pointers to loop-invariant tables are laundered in every iteration (see opaque_ptr)
        const C2* tw = opaque_ptr(tw0);
        const C2* twRev = opaque_ptr(twRev0);
        const T* wy = opaque_ptr(wy0);
        (void)twRev;
        (void)wy;
        const C2* stage = reinterpret_cast<const C2*>(smem_raw + (size_t)sidx * Cfg::TILE);
        tma::mbar_wait(bars + sidx, phase);
        T xr[S0::REGS], xi[S0::REGS];
#pragma unroll
        for (int i = 0; i < S0::IT; ++i) {
            const int j = t + i * NT;
#pragma unroll
            for (int r = 0; r < S0::R; ++r) {
                C2 v = mk2<T>(0, 0);
                if (S0::IT * NT == S0::NB || j < S0::NB) v = stage[(size_t)(j + r * S0::NB) * C + c];
                xr[i * S0::R + r] = v.x;
                xi[i * S0::R + r] = v.y;
            }
        }
        __syncthreads();      // the staging buffer has been read by everyone
        if (threadIdx.x == 0) {
            const int next = L + NSTG * (int)gridDim.x;
            if (next < total) {
                tma::fence_proxy_async();
                issue(next, sidx);
            }
        }
        if (++sidx == NSTG) {
            sidx = 0;
            phase ^= 1u;
        }
        C2* dst = out + (size_t)o * plane + col;
        if constexpr (MODE == 2) {
            typedef typename Last<PR, NT>::S SRL;
            run_from<PL, T, NT, -1, 0>(xr, xi, t, s_re, s_im, tw, BlockSync(), [&](auto& yr, auto& yi) {
                const T wxc = valid ? wx[col] : (T)0;
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
            run_from<PL, T, NT, DIR, 0>(xr, xi, t, s_re, s_im, tw, BlockSync(), [&](auto& yr, auto& yi) {
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
}

// ------------------------------------------------------------------------------------------------------
template <typename T, class PL, int NT, int C, int F>
__global__ void __launch_bounds__(NT * C, 1)
sfft_cols_apply_tma_kernel(const __grid_constant__ CUtensorMap tmSpec, const T* __restrict__ filt, typename V2<T>::type* __restrict__ out,
                           int Wc, int S, int nG, int perScale, const typename V2<T>::type* __restrict__ tw0) {
    typedef typename V2<T>::type C2;
    typedef SColTmaCfg<T, PL::N, C> Cfg;
    typedef Stage<PL, NT, 0> S0;
    typedef typename Last<PL, NT>::S SL;
    constexpr int H = PL::N;
    constexpr size_t FTILE = (size_t)F * H * C * sizeof(T);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    C2* spec = reinterpret_cast<C2*>(smem_raw);                               // [F][H][C] complex
    T* fstage = reinterpret_cast<T*>(smem_raw + F * Cfg::TILE);               // [F][H][C] real
    T* exch = reinterpret_cast<T*>(smem_raw + F * Cfg::TILE + FTILE);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + F * Cfg::TILE + FTILE + Cfg::EXCH);
    const int c = threadIdx.x % C, t = threadIdx.x / C;
    T* s_re = exch + c * Cfg::LD;
    T* s_im = exch + (C + c) * Cfg::LD;
    const int nCB = (Wc + C - 1) / C;
    // a unit = (set of spectra g, column block); its images: o = g S + i, i < S (a tile's scales), or with perScale
    // o = i S + g, i < nT (the tiles that use template g)
    const int nSets = perScale ? S : nG;
    const int nInner = perScale ? nG : S;
    const int units = nSets * nCB;
    const size_t plane = (size_t)H * Wc;

    auto issue_spec = [&](int u) {
        const int g = u / nCB, cb = u - g * nCB;
        tma::mbar_arrive_expect_tx(bar, (uint32_t)(F * Cfg::TILE));
#pragma unroll
        for (int f = 0; f < F; ++f)
#pragma unroll
            for (int b = 0; b < Cfg::NBOX; ++b)
                tma::tma_load_2d(spec + ((size_t)f * H + (size_t)b * Cfg::BOXR) * C, &tmSpec, cb * C * 2, (g * F + f) * H + b * Cfg::BOXR, bar);
    };
    // every thread copies its share of the filter tile of image o: element e = (f, row, column) in the tile's own order
    auto issue_filt = [&](int u, int i) {
        const int g = u / nCB, cb = u - g * nCB;
        const int o = perScale ? i * S + g : g * S + i;
        const T* base = filt + (size_t)o * F * plane + (size_t)cb * C;
        const int ncol = (Wc - cb * C) < C ? (Wc - cb * C) : C;
        for (int e = (int)threadIdx.x; e < F * H * C; e += NT * C) {
            const int cc = e % C, row = e / C;          // row runs over f * H + y: the F planes of an image are adjacent
            if (cc < ncol) cp_async8(fstage + e, base + (size_t)row * Wc + cc);
        }
        cp_async_commit();
    };

    if (threadIdx.x == 0) {
        tma::mbar_init(bar, 1);
        tma::fence_mbar_init();
    }
    __syncthreads();
    int u = (int)blockIdx.x;
    if (u < units) {
        if (threadIdx.x == 0) issue_spec(u);
        issue_filt(u, 0);
    }
    uint32_t phase = 0;
    for (; u < units; u += (int)gridDim.x) {
        const int g = u / nCB, cb = u - g * nCB;
        const int col = cb * C + c;
        const bool valid = col < Wc;
        tma::mbar_wait(bar, phase);
        phase ^= 1u;
        for (int i = 0; i < nInner; ++i) {
            const int o = perScale ? i * S + g : g * S + i;
            const C2* tw = opaque_ptr(tw0);
            cp_async_wait_all();
            __syncthreads();      // everybody's share of the filter tile has landed
            T xr[S0::REGS], xi[S0::REGS];
#pragma unroll
            for (int q = 0; q < S0::REGS; ++q) {
                xr[q] = 0;
                xi[q] = 0;
            }
            if (valid) {
#pragma unroll
                for (int f = 0; f < F; ++f) {
#pragma unroll
                    for (int ii = 0; ii < S0::IT; ++ii) {
                        const int j = t + ii * NT;
                        if (S0::IT * NT > S0::NB && j >= S0::NB) continue;
#pragma unroll
                        for (int r = 0; r < S0::R; ++r) {
                            const size_t n = (size_t)f * H + (size_t)(j + r * S0::NB);
                            const C2 v = spec[n * C + c];
                            const T gg = fstage[n * C + c];
                            xr[ii * S0::R + r] += v.x * gg;
                            xi[ii * S0::R + r] += v.y * gg;
                        }
                    }
                }
            }
            __syncthreads();      // the filter tile (and, after the unit's last image, the spectra tiles) can be overwritten
            {
                const bool lastOfUnit = (i + 1 == nInner);
                const int un = lastOfUnit ? u + (int)gridDim.x : u;
                if (un < units) {
                    if (lastOfUnit && threadIdx.x == 0) {
                        tma::fence_proxy_async();
                        issue_spec(un);
                    }
                    issue_filt(un, lastOfUnit ? 0 : i + 1);
                }
            }
            C2* dst = out + (size_t)o * plane + col;
            run_from<PL, T, NT, +1, 0>(xr, xi, t, s_re, s_im, tw, BlockSync(), [&](auto& yr, auto& yi) {
                if (!valid) return;
#pragma unroll
                for (int ii = 0; ii < SL::IT; ++ii) {
                    const int j = t + ii * NT;
                    if (SL::IT * NT > SL::NB && j >= SL::NB) continue;
#pragma unroll
                    for (int r = 0; r < SL::R; ++r) dst[(size_t)(j + r * SL::NB) * Wc] = mk2<T>(yr[ii * SL::R + r], yi[ii * SL::R + r]);
                }
            });
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// host side.  Return value: NB_OK when the pass was launched, 1 when this path does not apply (the caller then uses the
// register-staged kernels), a negative code on error.
#define SFFT_TMA_NA 1
bool sfft_use_tma();

// Off by default: measured on 8 tiles x 16 scales of 840 x 1440 (fp64) the three column passes take 3.34 ms with these kernels
// against 2.15 ms register-staged -- one CTA of 480 threads per SM runs every pass in lock step (in the passes of radix 10 and
// 12 only 84 and 70 of a column's 120 threads have a butterfly), which costs more than the hidden load latency gains.
// NEMO_B200_SFFT_COLS_TMA = all | plain | apply switches them on (A/B timing, tools/fft_variants.sh).
static bool tma_cols_applicable(const void* p, int Wc, bool apply) {
    // tensor-map constraints for rows of Wc complex doubles: 16-byte pitch (always), 16-byte aligned base
    const char* e = getenv("NEMO_B200_SFFT_COLS_TMA");
    if (!e) return false;
    const bool on = !strcmp(e, "all") || !strcmp(e, "1") || (apply ? !strcmp(e, "apply") : !strcmp(e, "plain"));
    return on && sfft_use_tma() && ((uintptr_t)p & 15) == 0 && Wc >= 1;
}

template <int MODE, int DIR>
static int sfft_cols_tma_t(const double2* in, double2* out, int B, int H, int Wc, const double* wy, const double* wx, cudaStream_t st) {
    typedef double T;
    typedef double2 C2;
    constexpr int C = SFFT_COL_C;
    constexpr int NSTG = 2;
    if (!tma_cols_applicable(in, Wc, false) || (long long)B * H >= (1LL << 31)) return SFFT_TMA_NA;
    int rc;
    const void *tw, *twRev = nullptr;
    // (the row plan of the length: one butterfly per thread and pass)
#define SFFT_CASE(NN, R0, R1, R2, NT, CR0, CR1, CR2, CNT)                                                                    \
    if (H == NN) {                                                                                                        \
        typedef Plan<NN, R0, R1, R2> PL;                                                                                  \
        typedef SColTmaCfg<T, NN, C> Cfg;                                                                                 \
        const size_t smem = NSTG * Cfg::TILE + Cfg::EXCH + 8 * NSTG + 8;                                                  \
        if (smem > 227 * 1024 || NT * C > 1024) return SFFT_TMA_NA;                                                       \
        if ((rc = sfft_tw<PL>(NB_PREC_F64, &tw)) != NB_OK) return rc;                                                     \
        if (MODE == 2 && (rc = sfft_tw<typename Reversed<PL>::type>(NB_PREC_F64, &twRev)) != NB_OK) return rc;            \
        CUtensorMap tm;                                                                                                   \
        if ((rc = nb_make_tmap_2d(&tm, in, 8, (unsigned long long)B * H, 2ULL * Wc, 16ULL * Wc, Cfg::BOXR, 2 * C)) != NB_OK) return rc; \
        if ((rc = sfft_set_smem(sfft_cols_tma_kernel<T, PL, NT, C, MODE, DIR, NSTG>, smem)) != NB_OK) return rc;          \
        const long long total = (long long)B * ((Wc + C - 1) / C);                                                        \
        if (total >= (1LL << 30)) return SFFT_TMA_NA;                                                                     \
        const long long cap = nb_sm_count();                                                                              \
        const unsigned grid = (unsigned)(total < cap ? total : cap);                                                      \
        { NB_PROF_BEGIN(st); sfft_cols_tma_kernel<T, PL, NT, C, MODE, DIR, NSTG><<<grid, NT * C, smem, st>>>(tm, out, Wc, B, wy, wx, (const C2*)tw, (const C2*)twRev); NB_PROF_END("sfft_cols_kernel", st); } \
        NB_CHECK_LAUNCH();                                                                                                \
        return NB_OK;                                                                                                     \
    }
    SFFT_FOR_SIZES(SFFT_CASE)
#undef SFFT_CASE
    return SFFT_TMA_NA;
}

template <int F>
static int sfft_cols_apply_tma_t(const double2* fmaps, const double* filt, double2* out, int nOut, int S, int perScale, int H, int Wc,
                                 cudaStream_t st) {
    typedef double T;
    typedef double2 C2;
    constexpr int C = SFFT_COL_C;
    if (!tma_cols_applicable(fmaps, Wc, true) || S < 1 || nOut % S != 0) return SFFT_TMA_NA;
    const int nG = nOut / S;
    const long long nSets = perScale ? S : nG;
    if (nSets * F * H >= (1LL << 31)) return SFFT_TMA_NA;
    int rc;
    const void* tw;
    // the row plan of a length (one butterfly per thread and pass, about twice the threads of the column plan): the single CTA
    // of an SM then has 15 warps for rows of 840
#define SFFT_CASE(NN, R0, R1, R2, NT, CR0, CR1, CR2, CNT)                                                                    \
    if (H == NN) {                                                                                                        \
        typedef Plan<NN, R0, R1, R2> PL;                                                                                  \
        typedef SColTmaCfg<T, NN, C> Cfg;                                                                                 \
        const size_t smem = F * Cfg::TILE + (size_t)F * NN * C * sizeof(T) + Cfg::EXCH + 16;                              \
        if (smem > 227 * 1024 || NT * C > 1024) return SFFT_TMA_NA;                                                       \
        if ((rc = sfft_tw<PL>(NB_PREC_F64, &tw)) != NB_OK) return rc;                                                     \
        CUtensorMap tm;                                                                                                   \
        if ((rc = nb_make_tmap_2d(&tm, fmaps, 8, (unsigned long long)nSets * F * H, 2ULL * Wc, 16ULL * Wc, Cfg::BOXR, 2 * C)) != NB_OK) return rc; \
        if ((rc = sfft_set_smem(sfft_cols_apply_tma_kernel<T, PL, NT, C, F>, smem)) != NB_OK) return rc;                  \
        const long long units = nSets * ((Wc + C - 1) / C);                                                               \
        const long long cap = nb_sm_count();                                                                              \
        const unsigned grid = (unsigned)(units < cap ? units : cap);                                                      \
        { NB_PROF_BEGIN(st); sfft_cols_apply_tma_kernel<T, PL, NT, C, F><<<grid, NT * C, smem, st>>>(tm, filt, out, Wc, S, nG, perScale, (const C2*)tw); NB_PROF_END("sfft_cols_kernel", st); } \
        NB_CHECK_LAUNCH();                                                                                                \
        return NB_OK;                                                                                                     \
    }
    SFFT_FOR_SIZES(SFFT_CASE)
#undef SFFT_CASE
    return SFFT_TMA_NA;
}

int nb_sfft_cols_tma_f64(const void* in, void* out, int B, int H, int Wc, int dir, cudaStream_t st) {
    if (dir < 0) return sfft_cols_tma_t<0, -1>((const double2*)in, (double2*)out, B, H, Wc, nullptr, nullptr, st);
    return sfft_cols_tma_t<0, +1>((const double2*)in, (double2*)out, B, H, Wc, nullptr, nullptr, st);
}

int nb_sfft_cols_window_tma_f64(void* data, int B, int H, int Wc, const void* wy, const void* wx, cudaStream_t st) {
    return sfft_cols_tma_t<2, -1>((const double2*)data, (double2*)data, B, H, Wc, (const double*)wy, (const double*)wx, st);
}

int nb_sfft_cols_apply_tma_f64(const void* fmaps, const void* filt, void* out, int nOut, int S, int perScale, int F, int H, int Wc,
                               cudaStream_t st) {
    switch (F) {
        case 1: return sfft_cols_apply_tma_t<1>((const double2*)fmaps, (const double*)filt, (double2*)out, nOut, S, perScale, H, Wc, st);
        case 2: return sfft_cols_apply_tma_t<2>((const double2*)fmaps, (const double*)filt, (double2*)out, nOut, S, perScale, H, Wc, st);
        default: return SFFT_TMA_NA;      // three or four spectra tiles do not fit beside the exchange buffer
    }
}
