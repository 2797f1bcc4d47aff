// Row passes of the 2-D real FFT for widths with a compile-time mixed-radix plan (sfft.cuh): real rows -> half spectra with the
// cosine taper and 1/sqrt(HW) fused into load / store (K1), half spectra -> real rows with signalNorm, scale, mask and
// zero-restore fused into the store (K4, K8).  Two real rows ride one complex transform of length W.
#include "sfft_host.cuh"
#include "tma.cuh"
#include <math.h>
#include <stdlib.h>

using namespace sfft;

// ------------------------------------------------------------------------------------------------------
static std::mutex g_sfft_mutex;
static std::map<std::vector<int>, void*> g_sfft_tw;

int nb_sfft_twiddles(int prec, int R0, int R1, int R2, const void** out) {
    std::lock_guard<std::mutex> lock(g_sfft_mutex);
    int dev = 0;
    cudaGetDevice(&dev);
    std::vector<int> key = {prec, dev, R0, R1, R2};
    auto it = g_sfft_tw.find(key);
    if (it != g_sfft_tw.end()) {
        *out = it->second;
        return NB_OK;
    }
    const long double twopi = 6.283185307179586476925286766559005768L;
    std::vector<long double> re, im;
    int Ns = R0;
    for (int R : {R1, R2}) {
        if (R <= 1) break;
        for (int r = 1; r < R; ++r)
            for (int k = 0; k < Ns; ++k) {
                const long double a = twopi * (long double)((long long)r * k) / (long double)((long long)Ns * R);
                re.push_back(cosl(a));
                im.push_back(-sinl(a));
            }
        Ns *= R;
    }
    void* d = nullptr;
    if (prec == NB_PREC_F64) {
        std::vector<double2> h(re.size());
        for (size_t i = 0; i < re.size(); ++i) { h[i].x = (double)re[i]; h[i].y = (double)im[i]; }
        NB_CHECK_CUDA(cudaMalloc(&d, sizeof(double2) * h.size()));
        NB_CHECK_CUDA(cudaMemcpy(d, h.data(), sizeof(double2) * h.size(), cudaMemcpyHostToDevice));
    } else {
        std::vector<float2> h(re.size());
        for (size_t i = 0; i < re.size(); ++i) { h[i].x = (float)re[i]; h[i].y = (float)im[i]; }
        NB_CHECK_CUDA(cudaMalloc(&d, sizeof(float2) * h.size()));
        NB_CHECK_CUDA(cudaMemcpy(d, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice));
    }
    g_sfft_tw[key] = d;
    *out = d;
    return NB_OK;
}

int nb_sfft_supported(int n) {
#define SFFT_CASE(NN, R0, R1, R2, NT, CR0, CR1, CR2, CNT) if (n == NN) return 1;
    SFFT_FOR_SIZES(SFFT_CASE)
#undef SFFT_CASE
    return 0;
}

// ------------------------------------------------------------------------------------------------------
template <typename T, int N, int G> struct SRowCfg {
    static constexpr int LD = N + 2;
    static constexpr size_t SMEM = (size_t)2 * G * LD * sizeof(T);
};

template <typename T, class PL, int NT, int G>
__global__ void __launch_bounds__(NT * G, SFFT_ROW_MINB)
sfft_rows_fwd_kernel(const T* __restrict__ in, typename V2<T>::type* __restrict__ out, int H, long long nrows, NbTaper tp, T scale,
                     const typename V2<T>::type* __restrict__ tw) {
    typedef typename V2<T>::type C2;
    typedef SRowCfg<T, PL::N, G> Cfg;
    typedef Stage<PL, NT, 0> S0;
    typedef typename Last<PL, NT>::S SL;
    constexpr int W = PL::N, Wc = W / 2 + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int g = threadIdx.x / NT, t = threadIdx.x % NT;
    const long long r0 = ((long long)blockIdx.x * G + g) * 2, r1 = r0 + 1;
    T* s_re = smem + g * Cfg::LD;
    T* s_im = smem + (G + g) * Cfg::LD;
    const bool v0 = r0 < nrows, v1 = r1 < nrows;
    const T ay0 = (v0 && tp.n > 0) ? taper_at<T>(tp, (int)(r0 % H), H) : (T)1;
    const T ay1 = (v1 && tp.n > 0) ? taper_at<T>(tp, (int)(r1 % H), H) : (T)1;
    const T* row0 = in + (size_t)(v0 ? r0 : 0) * W;
    const T* row1 = in + (size_t)(v1 ? r1 : 0) * W;
    T xr[S0::REGS], xi[S0::REGS];
#pragma unroll
    for (int i = 0; i < S0::IT; ++i) {
        const int j = t + i * NT;
#pragma unroll
        for (int r = 0; r < S0::R; ++r) {
            T a = 0, b = 0;
            if (S0::IT * NT == S0::NB || j < S0::NB) {
                const int n = j + r * S0::NB;
                a = row0[n];
                b = row1[n];
                if (tp.n > 0) {
                    const T ax = taper_at<T>(tp, n, W);
                    a *= ay0 * ax;
                    b *= ay1 * ax;
                }
            }
            xr[i * S0::R + r] = v0 ? a : (T)0;
            xi[i * S0::R + r] = v1 ? b : (T)0;
        }
    }
    run_from<PL, T, NT, -1, 0>(xr, xi, t, s_re, s_im, tw, BlockSync(), [&](auto& yr, auto& yi) {
#pragma unroll
        for (int i = 0; i < SL::IT; ++i) {
            const int j = t + i * NT;
            if (SL::IT * NT > SL::NB && j >= SL::NB) continue;
#pragma unroll
            for (int r = 0; r < SL::R; ++r) {
                s_re[j + r * SL::NB] = yr[i * SL::R + r];
                s_im[j + r * SL::NB] = yi[i * SL::R + r];
            }
        }
    });
    __syncthreads();
    // A = (Z + conj(Zc)) / 2,  B = (Z - conj(Zc)) / (2 i),  Zc = Z[W - k]
    const T h = (T)0.5 * scale;
    C2* o0 = out + (size_t)r0 * Wc;
    C2* o1 = out + (size_t)r1 * Wc;
    for (int k = t; k < Wc; k += NT) {
        const int kc = (k == 0) ? 0 : W - k;
        const T zr = s_re[k], zi = s_im[k];
        const T cr = s_re[kc], ci = -s_im[kc];
        if (v0) o0[k] = mk2<T>(h * (zr + cr), h * (zi + ci));
        if (v1) o1[k] = mk2<T>(h * (zi - ci), -h * (zr - cr));
    }
}

template <typename T> struct SC2RParams {
    T scale;
    const double* norm;
    const unsigned char* mask;
    int maskDiv;
    const T* zeroRef;
};

template <typename T, class PL, int NT, int G>
__global__ void __launch_bounds__(NT * G, SFFT_ROW_MINB)
sfft_rows_inv_kernel(const typename V2<T>::type* __restrict__ in, T* out, int H, long long nrows, SC2RParams<T> ep,
                     const typename V2<T>::type* __restrict__ tw) {
    typedef typename V2<T>::type C2;
    typedef SRowCfg<T, PL::N, G> Cfg;
    typedef Stage<PL, NT, 0> S0;
    typedef typename Last<PL, NT>::S SL;
    constexpr int W = PL::N, Wc = W / 2 + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int g = threadIdx.x / NT, t = threadIdx.x % NT;
    const long long r0 = ((long long)blockIdx.x * G + g) * 2, r1 = r0 + 1;
    T* s_re = smem + g * Cfg::LD;
    T* s_im = smem + (G + g) * Cfg::LD;
    const bool v0 = r0 < nrows, v1 = r1 < nrows;
    // what the store epilogue needs from global memory -- the mask bytes and which pixels of the zero reference (possibly the
    // output itself: in-place window undo) are zero -- is requested now, together with the spectra: read after the transform it
    // would cost every CTA a second exposed memory latency.  Kept as packed bytes / one bit per register slot.
    static_assert(SL::REGS <= 32, "one bit per register slot");
    constexpr int MW = (SL::REGS + 3) / 4;
    unsigned mpack[2][MW];
    unsigned zbits[2] = {0u, 0u};
    const unsigned H32 = (unsigned)H;
    unsigned imgOf[2], yOf[2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        const unsigned row = (unsigned)(rr == 0 ? r0 : r1);       // fewer than 2^32 rows in a batch (checked by the host)
        imgOf[rr] = row / H32;
        yOf[rr] = row - imgOf[rr] * H32;
#pragma unroll
        for (int w = 0; w < MW; ++w) mpack[rr][w] = 0x01010101u;
        if (!(rr == 0 ? v0 : v1)) continue;
        const unsigned char* mrow = ep.mask ? ep.mask + ((size_t)(imgOf[rr] / (unsigned)ep.maskDiv) * H + yOf[rr]) * W : nullptr;
        const T* zrow = ep.zeroRef ? ep.zeroRef + (size_t)row * W : nullptr;
        if (mrow) {
#pragma unroll
            for (int i = 0; i < SL::IT; ++i) {
                const int j = t + i * NT;
#pragma unroll
                for (int r = 0; r < SL::R; ++r) {
                    const int q = i * SL::R + r;
                    const unsigned m = (SL::IT * NT == SL::NB || j < SL::NB) ? (unsigned)mrow[j + r * SL::NB] : 1u;
                    mpack[rr][q >> 2] = (mpack[rr][q >> 2] & ~(0xffu << (8 * (q & 3)))) | (m << (8 * (q & 3)));
                }
            }
        }
        if (zrow) {
#pragma unroll
            for (int i = 0; i < SL::IT; ++i) {
                const int j = t + i * NT;
#pragma unroll
                for (int r = 0; r < SL::R; ++r) {
                    if ((SL::IT * NT == SL::NB || j < SL::NB) && zrow[j + r * SL::NB] == (T)0) zbits[rr] |= 1u << (i * SL::R + r);
                }
            }
        }
    }
    // the two half spectra, read once and coalesced, become the full Hermitian input Z = A + i B in the exchange buffer
    {
        const C2* a0 = in + (size_t)(v0 ? r0 : 0) * Wc;
        const C2* b0 = in + (size_t)(v1 ? r1 : 0) * Wc;
        constexpr int KT = (Wc + NT - 1) / NT;
        C2 av[KT], bv[KT];
#pragma unroll
        for (int i = 0; i < KT; ++i) {      // every load in flight before the first store
            const int k = t + i * NT;
            av[i] = (v0 && k < Wc) ? a0[k] : mk2<T>(0, 0);
            bv[i] = (v1 && k < Wc) ? b0[k] : mk2<T>(0, 0);
        }
#pragma unroll
        for (int i = 0; i < KT; ++i) {
            const int k = t + i * NT;
            if (k >= Wc) continue;
            const C2 a = av[i], b = bv[i];
            s_re[k] = a.x - b.y;
            s_im[k] = a.y + b.x;
            if (k > 0 && 2 * k < W) {   // Z[W - k] = conj(A[k]) + i conj(B[k])
                s_re[W - k] = a.x + b.y;
                s_im[W - k] = b.x - a.y;
            }
        }
    }
    __syncthreads();
    T xr[S0::REGS], xi[S0::REGS];
    stage_gather<PL, T, NT, 0>(xr, xi, t, s_re, s_im);
    __syncthreads();
    run_from<PL, T, NT, +1, 0>(xr, xi, t, s_re, s_im, tw, BlockSync(), [&](auto& yr, auto& yi) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const long long row = rr == 0 ? r0 : r1;
            if (row >= nrows) continue;
            T f = ep.scale;
            if (ep.norm) f *= (T)ep.norm[imgOf[rr]];
            T* orow = out + (size_t)row * W;
#pragma unroll
            for (int i = 0; i < SL::IT; ++i) {
                const int j = t + i * NT;
                if (SL::IT * NT > SL::NB && j >= SL::NB) continue;
#pragma unroll
                for (int r = 0; r < SL::R; ++r) {
                    const int q = i * SL::R + r;
                    T v = (rr == 0 ? yr[q] : yi[q]) * f;
                    if (ep.mask) v *= (T)((mpack[rr][q >> 2] >> (8 * (q & 3))) & 0xffu);
                    if ((zbits[rr] >> q) & 1u) v = 0;
                    orow[j + r * SL::NB] = v;
                }
            }
        }
    });
}


// ------------------------------------------------------------------------------------------------------
// Persistent row kernels fed by the TMA unit.  A CTA walks over tiles of G row pairs (tile = blockIdx.x, + gridDim.x, ...).  The
// two rows of a pair are adjacent in memory, so a pair is ONE bulk copy (cp.async.bulk, SASS UBLKCP) into a staging buffer, issued
// by one thread and tracked by an mbarrier -- no thread waits on its own global loads.  The staging buffer is handed back to the
// copy unit as soon as the first pass has gathered its inputs from it: the next tile's rows stream in underneath the three
// passes and the stores of the current one.  In the register-staged kernels above 70 % of the stall samples sat on the two
// exposed global-load phases (mask / zero pattern, then spectra), at 25 % active warps.
//
// The per-pixel bytes the inverse kernel's store needs (the point-source mask, or the "was exactly zero" map that the forward
// kernel of the window undo writes) travel the same way, into one of two small byte buffers (the epilogue of tile k reads its
// bytes while tile k + 1's are arriving).  That needs the two rows of a pair to be adjacent in the byte map too and 16-byte
// aligned: even H, W a multiple of 8 -- otherwise the host falls back to the kernels above.
//
// G = 1 by default: 144 threads for rows of 1440, three CTAs per SM (the persistent loop needs ~110 registers; with 288-thread
// CTAs two of them put five warps on one SM sub-partition, which caps a thread at 96 registers and spilled 550 bytes).
#ifndef SFFT_ROW_TMA_G
#define SFFT_ROW_TMA_G 1
#endif
template <typename T, int N, int G> struct SRowTmaCfg {
    static constexpr int LD = N + 2;
    static constexpr size_t EXCH = (size_t)2 * G * LD * sizeof(T);                     // s_re / s_im of G transforms
    static constexpr size_t STAGE_INV = (size_t)G * 2 * (N / 2 + 1) * 2 * sizeof(T);   // G pairs of half spectra
    static constexpr size_t STAGE_FWD = (size_t)G * 2 * N * sizeof(T);                 // G pairs of real rows
    static constexpr size_t BYTES = ((size_t)G * 2 * N + 15) / 16 * 16;                // one byte buffer
    static constexpr size_t SMEM_INV = STAGE_INV + EXCH + 2 * BYTES + 16;
    static constexpr size_t SMEM_FWD = STAGE_FWD + EXCH + 16;
    static constexpr int THREADS(int NT) { return NT * G; }
};
// resident CTAs per SM asked of ptxas: at most 16 warps, so that no sub-partition holds more than four (128 registers each)
constexpr int sfft_row_tma_minb(int threads) { return 16 / ((threads + 31) / 32) > 4 ? 4 : (16 / ((threads + 31) / 32) < 1 ? 1 : 16 / ((threads + 31) / 32)); }

template <typename T, class PL, int NT, int G>
__global__ void __launch_bounds__(NT * G, sfft_row_tma_minb(NT * G))
sfft_rows_inv_tma_kernel(const typename V2<T>::type* __restrict__ in, T* out, int H, int nrows, int ntiles, SC2RParams<T> ep,
                         const unsigned char* __restrict__ zeroBytes, const typename V2<T>::type* __restrict__ tw) {
    typedef typename V2<T>::type C2;
    typedef SRowTmaCfg<T, PL::N, G> Cfg;
    typedef Stage<PL, NT, 0> S0;
    typedef typename Last<PL, NT>::S SL;
    constexpr int W = PL::N, Wc = W / 2 + 1;
    constexpr uint32_t PAIR_BYTES = 2u * Wc * (uint32_t)sizeof(C2);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    C2* stage = reinterpret_cast<C2*>(smem_raw);
    T* exch = reinterpret_cast<T*>(smem_raw + Cfg::STAGE_INV);
    unsigned char* bytes = smem_raw + Cfg::STAGE_INV + Cfg::EXCH;       // [2][G][2][W]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + Cfg::STAGE_INV + Cfg::EXCH + 2 * Cfg::BYTES);
    const int g = threadIdx.x / NT, t = threadIdx.x % NT;
    T* s_re = exch + g * Cfg::LD;
    T* s_im = exch + (G + g) * Cfg::LD;
    const C2* pa = stage + (size_t)g * 2 * Wc;      // this group's pair: rows A (first) and B (second)
    const C2* pb = pa + Wc;
    // at most one of the two byte maps is present (mask: filter application; zero map: window undo)
    const unsigned char* bsrc = ep.mask ? ep.mask : zeroBytes;
    const unsigned H32 = (unsigned)H;
    const unsigned bdiv = ep.mask ? (unsigned)ep.maskDiv : 1u;

    // full pairs of a tile arrive by bulk copy; a pair with only its first row inside the batch (odd row count: the very last
    // pair) is read by its group's threads
    auto issue = [&](int tile, int par) {
        uint32_t total = 0;
#pragma unroll
        for (int gg = 0; gg < G; ++gg)
            if ((tile * G + gg) * 2 + 1 < nrows) total += PAIR_BYTES + (bsrc ? 2u * W : 0u);
        if (total == 0) return;
        tma::mbar_arrive_expect_tx(bar, total);
#pragma unroll
        for (int gg = 0; gg < G; ++gg) {
            const int q0 = (tile * G + gg) * 2;
            if (q0 + 1 >= nrows) continue;
            tma::bulk_g2s(stage + (size_t)gg * 2 * Wc, in + (size_t)q0 * Wc, PAIR_BYTES, bar);
            if (bsrc) {
                const unsigned img = (unsigned)q0 / H32, y = (unsigned)q0 - img * H32;
                tma::bulk_g2s(bytes + (size_t)par * Cfg::BYTES + (size_t)gg * 2 * W, bsrc + ((size_t)(img / bdiv) * H + y) * W, 2u * W, bar);
            }
        }
    };

    if (threadIdx.x == 0) {
        tma::mbar_init(bar, 1);
        tma::fence_mbar_init();
    }
    __syncthreads();
    int tile = (int)blockIdx.x;
    if (threadIdx.x == 0 && tile < ntiles) issue(tile, 0);
    uint32_t phase = 0;
    int par = 0;
    for (; tile < ntiles; tile += (int)gridDim.x, par ^= 1) {
        const int r0 = (tile * G + g) * 2, r1 = r0 + 1;
        const bool v0 = r0 < nrows, v1 = r1 < nrows;
        if (tile * G * 2 + 1 < nrows) {
            tma::mbar_wait(bar, phase);
            phase ^= 1u;
        }
        // first-pass inputs straight from the staged half spectra: Z[n] = A[n] + i B[n] for n <= W / 2, and
        // Z[n] = conj(A[W - n]) + i conj(B[W - n]) above (Hermitian halves of two real rows)
        T xr[S0::REGS], xi[S0::REGS];
        const C2* ga = in + (size_t)(v0 ? r0 : 0) * Wc;
#pragma unroll
        for (int i = 0; i < S0::IT; ++i) {
            const int j = t + i * NT;
#pragma unroll
            for (int r = 0; r < S0::R; ++r) {
                T zr = 0, zi = 0;
                if (S0::IT * NT == S0::NB || j < S0::NB) {
                    const int n = j + r * S0::NB;
                    const bool low = 2 * n <= W;
                    const int m = low ? n : W - n;
                    const T sg = low ? (T)1 : (T)-1;
                    if (v1) {
                        const C2 a = pa[m], b = pb[m];
                        zr = a.x - sg * b.y;
                        zi = sg * a.y + b.x;
                    } else if (v0) {
                        const C2 a = ga[m];
                        zr = a.x;
                        zi = sg * a.y;
                    }
                }
                xr[i * S0::R + r] = zr;
                xi[i * S0::R + r] = zi;
            }
        }
        __syncthreads();      // every thread has read the staging buffer: it can take the next tile
        if (threadIdx.x == 0) {
            const int next = tile + (int)gridDim.x;
            if (next < ntiles) {
                tma::fence_proxy_async();
                issue(next, par ^ 1);
            }
        }
        const unsigned char* sb = bytes + (size_t)par * Cfg::BYTES + (size_t)g * 2 * W;
        run_from<PL, T, NT, +1, 0>(xr, xi, t, s_re, s_im, tw, BlockSync(), [&](auto& yr, auto& yi) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int row = rr == 0 ? r0 : r1;
                if (row >= nrows) continue;
                T f = ep.scale;
                if (ep.norm) f *= (T)ep.norm[(unsigned)row / H32];
                T* orow = out + (size_t)row * W;
                // a lone first row (odd row count) has no staged bytes: its mask / zero bytes come from global memory
                const unsigned char* brow = nullptr;
                if (bsrc) {
                    if (v1) brow = sb + rr * W;
                    else {
                        const unsigned img = (unsigned)row / H32, y = (unsigned)row - img * H32;
                        brow = bsrc + ((size_t)(img / bdiv) * H + y) * W;
                    }
                }
#pragma unroll
                for (int i = 0; i < SL::IT; ++i) {
                    const int j = t + i * NT;
                    if (SL::IT * NT > SL::NB && j >= SL::NB) continue;
#pragma unroll
                    for (int r = 0; r < SL::R; ++r) {
                        const int q = i * SL::R + r;
                        const int n = j + r * SL::NB;
                        T v = (rr == 0 ? yr[q] : yi[q]) * f;
                        if (ep.mask) v *= (T)brow[n];
                        else if (zeroBytes && brow[n]) v = 0;
                        orow[n] = v;
                    }
                }
            }
        });
        __syncthreads();      // the exchange buffer is free for the next tile's first scatter
    }
}

template <typename T, class PL, int NT, int G>
__global__ void __launch_bounds__(NT * G, sfft_row_tma_minb(NT * G))
sfft_rows_fwd_tma_kernel(const T* __restrict__ in, typename V2<T>::type* __restrict__ out, int H, int nrows, int ntiles,
                         NbTaper tp, T scale, unsigned char* __restrict__ zeroOut, const typename V2<T>::type* __restrict__ tw) {
    typedef typename V2<T>::type C2;
    typedef SRowTmaCfg<T, PL::N, G> Cfg;
    typedef Stage<PL, NT, 0> S0;
    typedef typename Last<PL, NT>::S SL;
    constexpr int W = PL::N, Wc = W / 2 + 1;
    constexpr uint32_t PAIR_BYTES = 2u * W * (uint32_t)sizeof(T);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* stage = reinterpret_cast<T*>(smem_raw);
    T* exch = reinterpret_cast<T*>(smem_raw + Cfg::STAGE_FWD);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + Cfg::STAGE_FWD + Cfg::EXCH);
    const int g = threadIdx.x / NT, t = threadIdx.x % NT;
    T* s_re = exch + g * Cfg::LD;
    T* s_im = exch + (G + g) * Cfg::LD;
    const T* p0 = stage + (size_t)g * 2 * W;
    const T* p1 = p0 + W;

    auto issue = [&](int tile) {
        uint32_t total = 0;
#pragma unroll
        for (int gg = 0; gg < G; ++gg)
            if ((tile * G + gg) * 2 + 1 < nrows) total += PAIR_BYTES;
        if (total == 0) return;
        tma::mbar_arrive_expect_tx(bar, total);
#pragma unroll
        for (int gg = 0; gg < G; ++gg) {
            const int q0 = (tile * G + gg) * 2;
            if (q0 + 1 < nrows) tma::bulk_g2s(stage + (size_t)gg * 2 * W, in + (size_t)q0 * W, PAIR_BYTES, bar);
        }
    };
    if (threadIdx.x == 0) {
        tma::mbar_init(bar, 1);
        tma::fence_mbar_init();
    }
    __syncthreads();
    int tile = (int)blockIdx.x;
    if (threadIdx.x == 0 && tile < ntiles) issue(tile);
    uint32_t phase = 0;
    for (; tile < ntiles; tile += (int)gridDim.x) {
        const int r0 = (tile * G + g) * 2, r1 = r0 + 1;
        const bool v0 = r0 < nrows, v1 = r1 < nrows;
        const T ay0 = (v0 && tp.n > 0) ? taper_at<T>(tp, r0 % H, H) : (T)1;
        const T ay1 = (v1 && tp.n > 0) ? taper_at<T>(tp, r1 % H, H) : (T)1;
        if (tile * G * 2 + 1 < nrows) {
            tma::mbar_wait(bar, phase);
            phase ^= 1u;
        }
        const T* grow = in + (size_t)(v0 ? r0 : 0) * W;
        unsigned char* z0 = zeroOut ? zeroOut + (size_t)r0 * W : nullptr;
        T xr[S0::REGS], xi[S0::REGS];
#pragma unroll
        for (int i = 0; i < S0::IT; ++i) {
            const int j = t + i * NT;
#pragma unroll
            for (int r = 0; r < S0::R; ++r) {
                T a = 0, b = 0;
                if (S0::IT * NT == S0::NB || j < S0::NB) {
                    const int n = j + r * S0::NB;
                    if (v1) {
                        a = p0[n];
                        b = p1[n];
                    } else if (v0) {
                        a = grow[n];
                    }
                    if (z0) {      // which pixels are exactly zero: the inverse kernel of the window undo restores them
                        if (v0) z0[n] = (a == (T)0) ? 1 : 0;
                        if (v1) z0[W + n] = (b == (T)0) ? 1 : 0;
                    }
                    if (tp.n > 0) {
                        const T ax = taper_at<T>(tp, n, W);
                        a *= ay0 * ax;
                        b *= ay1 * ax;
                    }
                }
                xr[i * S0::R + r] = v0 ? a : (T)0;
                xi[i * S0::R + r] = v1 ? b : (T)0;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int next = tile + (int)gridDim.x;
            if (next < ntiles) {
                tma::fence_proxy_async();
                issue(next);
            }
        }
        run_from<PL, T, NT, -1, 0>(xr, xi, t, s_re, s_im, tw, BlockSync(), [&](auto& yr, auto& yi) {
#pragma unroll
            for (int i = 0; i < SL::IT; ++i) {
                const int j = t + i * NT;
                if (SL::IT * NT > SL::NB && j >= SL::NB) continue;
#pragma unroll
                for (int r = 0; r < SL::R; ++r) {
                    s_re[j + r * SL::NB] = yr[i * SL::R + r];
                    s_im[j + r * SL::NB] = yi[i * SL::R + r];
                }
            }
        });
        __syncthreads();
        const T h = (T)0.5 * scale;
        C2* o0 = out + (size_t)r0 * Wc;
        C2* o1 = out + (size_t)r1 * Wc;
        for (int k = t; k < Wc; k += NT) {
            const int kc = (k == 0) ? 0 : W - k;
            const T zr = s_re[k], zi = s_im[k];
            const T cr = s_re[kc], ci = -s_im[kc];
            if (v0) o0[k] = mk2<T>(h * (zr + cr), h * (zi + ci));
            if (v1) o1[k] = mk2<T>(h * (zi - ci), -h * (zr - cr));
        }
        __syncthreads();
    }
}

// the TMA-fed persistent kernels are the default; NEMO_B200_SFFT_NO_TMA=1 selects the register-staged ones (A/B timing)
bool sfft_use_tma() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("NEMO_B200_SFFT_NO_TMA");
        on = (e && e[0] == '1') ? 0 : 1;
    }
    return on == 1;
}

static void sfft_fill_taper(NbTaper& tp, int apodPix) {
    tp.n = apodPix;
    for (int i = 0; i < apodPix; ++i) tp.a[i] = (apodPix == 1) ? 0.0 : 0.5 * (1.0 - cos(M_PI * (double)i / (double)(apodPix - 1)));
    if (apodPix > 0) tp.a[apodPix - 1] = 1.0;
}

template <typename T>
static int sfft_rows_fwd_t(const T* in, typename V2<T>::type* out, int B, int H, int W, int apodPix, double scale, unsigned char* zeroOut,
                           cudaStream_t st) {
    typedef typename V2<T>::type C2;
    NB_REQUIRE(apodPix >= 0 && apodPix <= NB_MAX_TAPER && 2 * apodPix <= H && 2 * apodPix <= W, NB_ERR_ARG, "apodPix %d out of range", apodPix);
    NbTaper tp;
    sfft_fill_taper(tp, apodPix);
    const int prec = sizeof(T) == 8 ? NB_PREC_F64 : NB_PREC_F32;
    const long long nrows = (long long)B * H;
    const long long npairs = (nrows + 1) / 2;
    constexpr int G = SFFT_ROW_G;
    constexpr int TG = SFFT_ROW_TMA_G;
    const bool tmaOk = sfft_use_tma() && nrows < (1LL << 30);
    NB_REQUIRE(!zeroOut || tmaOk, NB_ERR_UNSUPPORTED, "sfft: the zero map is written by the TMA row kernels only");
    int rc;
    const void* tw;
#define SFFT_CASE(NN, R0, R1, R2, NT, CR0, CR1, CR2, CNT)                                                                              \
    if (W == NN) {                                                                                                        \
        typedef Plan<NN, R0, R1, R2> PL;                                                                                  \
        typedef SRowCfg<T, NN, G> Cfg;                                                                                    \
        if ((rc = sfft_tw<PL>(prec, &tw)) != NB_OK) return rc;                                                            \
        if (tmaOk) {                                                                                                      \
            typedef SRowTmaCfg<T, NN, TG> TCfg;                                                                           \
            if ((rc = sfft_set_smem(sfft_rows_fwd_tma_kernel<T, PL, NT, TG>, TCfg::SMEM_FWD)) != NB_OK) return rc;        \
            const long long ntiles = (npairs + TG - 1) / TG;                                                              \
            const long long cap = (long long)nb_sm_count() * sfft_row_tma_minb(NT * TG);                                  \
            const unsigned pgrid = (unsigned)(ntiles < cap ? ntiles : cap);                                               \
            { NB_PROF_BEGIN(st); sfft_rows_fwd_tma_kernel<T, PL, NT, TG><<<pgrid, NT * TG, TCfg::SMEM_FWD, st>>>(in, out, H, (int)nrows, (int)ntiles, tp, (T)scale, zeroOut, (const C2*)tw); NB_PROF_END("sfft_rows_fwd_kernel", st); } \
            NB_CHECK_LAUNCH();                                                                                            \
            return NB_OK;                                                                                                 \
        }                                                                                                                 \
        if ((rc = sfft_set_smem(sfft_rows_fwd_kernel<T, PL, NT, G>, Cfg::SMEM)) != NB_OK) return rc;                      \
        const unsigned grid = (unsigned)((npairs + G - 1) / G);                                                           \
        { NB_PROF_BEGIN(st); sfft_rows_fwd_kernel<T, PL, NT, G><<<grid, NT * G, Cfg::SMEM, st>>>(in, out, H, nrows, tp, (T)scale, (const C2*)tw); NB_PROF_END("sfft_rows_fwd_kernel", st); } \
        NB_CHECK_LAUNCH();                                                                                                \
        return NB_OK;                                                                                                     \
    }
    SFFT_FOR_SIZES(SFFT_CASE)
#undef SFFT_CASE
    nb_set_error("sfft: width %d has no compile-time plan", W);
    return NB_ERR_UNSUPPORTED;
}

template <typename T>
static int sfft_rows_inv_t(const typename V2<T>::type* in, T* out, int B, int H, int W, SC2RParams<T> ep, const unsigned char* zeroBytes,
                           cudaStream_t st) {
    typedef typename V2<T>::type C2;
    const int prec = sizeof(T) == 8 ? NB_PREC_F64 : NB_PREC_F32;
    const long long nrows = (long long)B * H;
    NB_REQUIRE(nrows < (1LL << 31), NB_ERR_ARG, "sfft: %lld rows in one call (at most 2^31 - 1)", nrows);
    const long long npairs = (nrows + 1) / 2;
    constexpr int G = SFFT_ROW_G;
    constexpr int TG = SFFT_ROW_TMA_G;
    // the TMA kernel stages the per-pixel bytes as pairs of rows: see nb_sfft_rows_bytes_ok; it has no path for a zero
    // reference given as values (the register-staged kernel reads those)
    const bool bytesWanted = ep.mask != nullptr || zeroBytes != nullptr;
    const bool tmaOk = sfft_use_tma() && nrows < (1LL << 30) && ep.zeroRef == nullptr && !(ep.mask && zeroBytes) &&
                       (!bytesWanted || nb_sfft_rows_bytes_ok(H, W));
    NB_REQUIRE(!zeroBytes || tmaOk, NB_ERR_UNSUPPORTED, "sfft: the zero map is read by the TMA row kernels only");
    int rc;
    const void* tw;
#define SFFT_CASE(NN, R0, R1, R2, NT, CR0, CR1, CR2, CNT)                                                                              \
    if (W == NN) {                                                                                                        \
        typedef Plan<NN, R0, R1, R2> PL;                                                                                  \
        typedef SRowCfg<T, NN, G> Cfg;                                                                                    \
        if ((rc = sfft_tw<PL>(prec, &tw)) != NB_OK) return rc;                                                            \
        if (tmaOk) {                                                                                                      \
            typedef SRowTmaCfg<T, NN, TG> TCfg;                                                                           \
            if ((rc = sfft_set_smem(sfft_rows_inv_tma_kernel<T, PL, NT, TG>, TCfg::SMEM_INV)) != NB_OK) return rc;        \
            const long long ntiles = (npairs + TG - 1) / TG;                                                              \
            const long long cap = (long long)nb_sm_count() * sfft_row_tma_minb(NT * TG);                                  \
            const unsigned pgrid = (unsigned)(ntiles < cap ? ntiles : cap);                                               \
            { NB_PROF_BEGIN(st); sfft_rows_inv_tma_kernel<T, PL, NT, TG><<<pgrid, NT * TG, TCfg::SMEM_INV, st>>>(in, out, H, (int)nrows, (int)ntiles, ep, zeroBytes, (const C2*)tw); NB_PROF_END("sfft_rows_inv_kernel", st); } \
            NB_CHECK_LAUNCH();                                                                                            \
            return NB_OK;                                                                                                 \
        }                                                                                                                 \
        if ((rc = sfft_set_smem(sfft_rows_inv_kernel<T, PL, NT, G>, Cfg::SMEM)) != NB_OK) return rc;                      \
        const unsigned grid = (unsigned)((npairs + G - 1) / G);                                                           \
        { NB_PROF_BEGIN(st); sfft_rows_inv_kernel<T, PL, NT, G><<<grid, NT * G, Cfg::SMEM, st>>>(in, out, H, nrows, ep, (const C2*)tw); NB_PROF_END("sfft_rows_inv_kernel", st); } \
        NB_CHECK_LAUNCH();                                                                                                \
        return NB_OK;                                                                                                     \
    }
    SFFT_FOR_SIZES(SFFT_CASE)
#undef SFFT_CASE
    nb_set_error("sfft: width %d has no compile-time plan", W);
    return NB_ERR_UNSUPPORTED;
}

int nb_sfft_rows_fwd(int prec, const void* in, void* out, int B, int H, int W, int apodPix, double scale, unsigned char* zeroOut,
                     cudaStream_t st) {
    if (prec == NB_PREC_F64) return sfft_rows_fwd_t<double>((const double*)in, (double2*)out, B, H, W, apodPix, scale, zeroOut, st);
    return sfft_rows_fwd_t<float>((const float*)in, (float2*)out, B, H, W, apodPix, scale, zeroOut, st);
}

// can the TMA row kernels stage a [.., H, W] byte map (mask / zero map) as pairs of rows: the rows of a pair adjacent (even H:
// pairs never straddle two images) and every pair 16-byte aligned (2 W a multiple of 16); and are those kernels in use at all
int nb_sfft_rows_bytes_ok(int H, int W) { return sfft_use_tma() && (H % 2) == 0 && (W % 8) == 0 && nb_sfft_supported(W); }

int nb_sfft_rows_inv(int prec, const void* in, void* out, int B, int H, int W, double scale, const double* norm, const unsigned char* mask,
                     int maskDiv, const void* zeroRef, const unsigned char* zeroBytes, cudaStream_t st) {
    if (prec == NB_PREC_F64) {
        SC2RParams<double> ep = {scale, norm, mask, maskDiv < 1 ? 1 : maskDiv, (const double*)zeroRef};
        return sfft_rows_inv_t<double>((const double2*)in, (double*)out, B, H, W, ep, zeroBytes, st);
    }
    SC2RParams<float> ep = {(float)scale, norm, mask, maskDiv < 1 ? 1 : maskDiv, (const float*)zeroRef};
    return sfft_rows_inv_t<float>((const float2*)in, (float*)out, B, H, W, ep, zeroBytes, st);
}
