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
    int* zeroFlag;      // [images] or null: set to 1 for an image with a pixel where out + (1 - mask) == 0 (what K7's edge filter looks for)
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
            bool hasZero = false;
#pragma unroll
            for (int i = 0; i < SL::IT; ++i) {
                const int j = t + i * NT;
                if (SL::IT * NT > SL::NB && j >= SL::NB) continue;
#pragma unroll
                for (int r = 0; r < SL::R; ++r) {
                    const int q = i * SL::R + r;
                    T v = (rr == 0 ? yr[q] : yi[q]) * f;
                    if (ep.mask) {
                        const unsigned mb = (mpack[rr][q >> 2] >> (8 * (q & 3))) & 0xffu;
                        v *= (T)mb;
                        hasZero |= (v + (T)(1 - (int)mb)) == (T)0;
                    }
                    if ((zbits[rr] >> q) & 1u) v = 0;
                    orow[j + r * SL::NB] = v;
                }
            }
            if (hasZero && ep.mask && ep.zeroFlag) ep.zeroFlag[imgOf[rr]] = 1;
        }
    });
}


// ------------------------------------------------------------------------------------------------------
// Persistent inverse row kernel fed by the TMA unit.  A CTA walks over tiles of G row pairs (tile = blockIdx.x, + gridDim.x,
// ...).  The two rows of a pair are adjacent in memory, so a pair of half spectra is ONE bulk copy (cp.async.bulk, SASS UBLKCP)
// into a staging buffer, issued by one thread and tracked by an mbarrier -- no thread waits on its own global loads.  The
// staging buffer is handed back to the copy unit as soon as the first pass has gathered its inputs from it: the next tile's
// rows stream in underneath the three passes and the stores of the current one.  In the register-staged kernel above 70 % of
// the stall samples sat on its two exposed global-load phases (mask / zero pattern, then spectra), at 25 % active warps.
//
// What the store needs per pixel travels the same way: the rows of the zero reference (the data before the window undo; turned
// into one bit per register slot while the first pass gathers, so that their buffer is free again at once), or the rows of the
// byte mask (two small buffers: the epilogue of tile k reads its bytes while tile k + 1's arrive).  That needs the two rows of a
// pair to be adjacent there too and 16-byte aligned: for the mask, even H and W a multiple of 8 -- otherwise the host uses the
// kernel above.  Measured on 840 x 1440 tiles (144 images per launch pair): 1.92 -> 1.46 ms.  The forward row pass was tried
// the same way and is FASTER register-staged (0.71 against 0.84 ms: it was the one FFT kernel above half of the HBM roofline
// already, and the staged copy adds a trip through shared memory), so it keeps its kernel.
//
// G = 1: 144 threads for rows of 1440, three CTAs per SM (the persistent loop needs ~110 registers; with 288-thread CTAs two of
// them put five warps on one SM sub-partition, which caps a thread at 96 registers and spilled 550 bytes; G = 2: 1.59 ms).
#ifndef SFFT_ROW_TMA_G
#define SFFT_ROW_TMA_G 1
#endif
template <typename T, int N, int G> struct SRowTmaCfg {
    static constexpr int LD = N + 2;
    static constexpr size_t EXCH = (size_t)2 * G * LD * sizeof(T);                     // s_re / s_im of G transforms
    static constexpr size_t STAGE = (size_t)G * 2 * (N / 2 + 1) * 2 * sizeof(T);       // G pairs of half spectra
    static constexpr size_t ZREF = (size_t)G * 2 * N * sizeof(T);                      // G pairs of zero-reference rows
    static constexpr size_t BYTES = ((size_t)G * 2 * N + 15) / 16 * 16;                // one byte-mask buffer
    static constexpr size_t smem(bool zref, bool mask) { return STAGE + EXCH + (zref ? ZREF : 0) + (mask ? 2 * BYTES : 0) + 16; }
};
// resident CTAs per SM asked of ptxas: at most 16 warps, so that no sub-partition holds more than four (128 registers each)
#ifdef SFFT_ROW_TMA_MINB_OVERRIDE
constexpr int sfft_row_tma_minb(int) { return SFFT_ROW_TMA_MINB_OVERRIDE; }
constexpr int sfft_row_tma_minb_default(int threads) {
#else
constexpr int sfft_row_tma_minb(int threads) {
#endif
    return 16 / ((threads + 31) / 32) > 4 ? 4 : (16 / ((threads + 31) / 32) < 1 ? 1 : 16 / ((threads + 31) / 32)); }

// EP: 0 nothing per pixel, 1 byte mask, 2 zero reference
template <typename T, class PL, int NT, int G, int EP>
__global__ void __launch_bounds__(NT * G, sfft_row_tma_minb(NT * G))
sfft_rows_inv_tma_kernel(const typename V2<T>::type* __restrict__ in, T* out, int H, int nrows, int ntiles, SC2RParams<T> ep,
                         const typename V2<T>::type* __restrict__ tw0) {
    typedef typename V2<T>::type C2;
    typedef SRowTmaCfg<T, PL::N, G> Cfg;
    typedef Stage<PL, NT, 0> S0;
    typedef typename Last<PL, NT>::S SL;
    constexpr int W = PL::N, Wc = W / 2 + 1;
    constexpr uint32_t PAIR_BYTES = 2u * Wc * (uint32_t)sizeof(C2);
    constexpr uint32_t EXTRA_BYTES = EP == 1 ? 2u * W : EP == 2 ? 2u * W * (uint32_t)sizeof(T) : 0u;
    static_assert(SL::REGS <= 32, "one bit per register slot");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    C2* stage = reinterpret_cast<C2*>(smem_raw);
    T* exch = reinterpret_cast<T*>(smem_raw + Cfg::STAGE);
    unsigned char* extra = smem_raw + Cfg::STAGE + Cfg::EXCH;         // EP 1: bytes [2][G][2][W]; EP 2: zero reference [G][2][W]
    uint64_t* bar = reinterpret_cast<uint64_t*>(extra + (EP == 1 ? 2 * Cfg::BYTES : EP == 2 ? Cfg::ZREF : 0));
    const int g = threadIdx.x / NT, t = threadIdx.x % NT;
    T* s_re = exch + g * Cfg::LD;
    T* s_im = exch + (G + g) * Cfg::LD;
    const C2* pa = stage + (size_t)g * 2 * Wc;      // this group's pair: rows A (first) and B (second)
    const C2* pb = pa + Wc;
    const unsigned H32 = (unsigned)H;

    // full pairs of a tile arrive by bulk copy; a pair with only its first row inside the batch (odd row count: the very last
    // pair) is read by its group's threads
    auto issue = [&](int tile, int par) {
        uint32_t total = 0;
#pragma unroll
        for (int gg = 0; gg < G; ++gg)
            if ((tile * G + gg) * 2 + 1 < nrows) total += PAIR_BYTES + EXTRA_BYTES;
        if (total == 0) return;
        tma::mbar_arrive_expect_tx(bar, total);
#pragma unroll
        for (int gg = 0; gg < G; ++gg) {
            const int q0 = (tile * G + gg) * 2;
            if (q0 + 1 >= nrows) continue;
            tma::bulk_g2s(stage + (size_t)gg * 2 * Wc, in + (size_t)q0 * Wc, PAIR_BYTES, bar);
            if (EP == 1) {
                const unsigned img = (unsigned)q0 / H32, y = (unsigned)q0 - img * H32;
                tma::bulk_g2s(extra + (size_t)par * Cfg::BYTES + (size_t)gg * 2 * W, ep.mask + ((size_t)(img / (unsigned)ep.maskDiv) * H + y) * W,
                              EXTRA_BYTES, bar);
            } else if (EP == 2) {
                tma::bulk_g2s(extra + (size_t)gg * 2 * W * sizeof(T), ep.zeroRef + (size_t)q0 * W, EXTRA_BYTES, bar);
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
        // (laundered per iteration: the compiler would otherwise hoist the loop-invariant twiddle loads out of the loop and
        // hold them in registers)
        const C2* tw = tw0;
        asm volatile("" : "+l"(tw));
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
        // which of this thread's output pixels must stay exactly zero: one bit per register slot of the last pass
        unsigned zbits[2] = {0u, 0u};
        if (EP == 2) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                if (!(rr == 0 ? v0 : v1)) continue;
                // (a lone first row has no staged copy: it is read from global memory)
                const T* zrow = v1 ? reinterpret_cast<const T*>(extra) + ((size_t)g * 2 + rr) * W : ep.zeroRef + (size_t)r0 * W;
#pragma unroll
                for (int i = 0; i < SL::IT; ++i) {
                    const int j = t + i * NT;
#pragma unroll
                    for (int r = 0; r < SL::R; ++r)
                        if ((SL::IT * NT == SL::NB || j < SL::NB) && zrow[j + r * SL::NB] == (T)0) zbits[rr] |= 1u << (i * SL::R + r);
                }
            }
        }
        __syncthreads();      // every thread has read the staging buffers: they can take the next tile
        if (threadIdx.x == 0) {
            const int next = tile + (int)gridDim.x;
            if (next < ntiles) {
                tma::fence_proxy_async();
                issue(next, par ^ 1);
            }
        }
        const unsigned char* sb = extra + (size_t)par * Cfg::BYTES + (size_t)g * 2 * W;
        run_from<PL, T, NT, +1, 0>(xr, xi, t, s_re, s_im, tw, BlockSync(), [&](auto& yr, auto& yi) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int row = rr == 0 ? r0 : r1;
                if (row >= nrows) continue;
                T f = ep.scale;
                if (ep.norm) f *= (T)ep.norm[(unsigned)row / H32];
                T* orow = out + (size_t)row * W;
                bool hasZero = false;
                const unsigned char* brow = nullptr;
                if (EP == 1) {
                    if (v1) brow = sb + rr * W;
                    else {
                        const unsigned img = (unsigned)row / H32, y = (unsigned)row - img * H32;
                        brow = ep.mask + ((size_t)(img / (unsigned)ep.maskDiv) * H + y) * W;
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
                        if (EP == 1) {
                            const unsigned mb = brow[n];
                            v *= (T)mb;
                            hasZero |= (v + (T)(1 - (int)mb)) == (T)0;
                        }
                        if (EP == 2 && ((zbits[rr] >> q) & 1u)) v = 0;
                        orow[n] = v;
                    }
                }
                if (EP == 1 && hasZero && ep.zeroFlag) ep.zeroFlag[(unsigned)row / H32] = 1;
            }
        });
        __syncthreads();      // the exchange buffer is free for the next tile's first scatter
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
static int sfft_rows_fwd_t(const T* in, typename V2<T>::type* out, int B, int H, int W, int apodPix, double scale, cudaStream_t st) {
    typedef typename V2<T>::type C2;
    NB_REQUIRE(apodPix >= 0 && apodPix <= NB_MAX_TAPER && 2 * apodPix <= H && 2 * apodPix <= W, NB_ERR_ARG, "apodPix %d out of range", apodPix);
    NbTaper tp;
    sfft_fill_taper(tp, apodPix);
    const int prec = sizeof(T) == 8 ? NB_PREC_F64 : NB_PREC_F32;
    const long long nrows = (long long)B * H;
    const long long npairs = (nrows + 1) / 2;
    constexpr int G = SFFT_ROW_G;
    int rc;
    const void* tw;
#define SFFT_CASE(NN, R0, R1, R2, NT, CR0, CR1, CR2, CNT)                                                                              \
    if (W == NN) {                                                                                                        \
        typedef Plan<NN, R0, R1, R2> PL;                                                                                  \
        typedef SRowCfg<T, NN, G> Cfg;                                                                                    \
        if ((rc = sfft_tw<PL>(prec, &tw)) != NB_OK) return rc;                                                            \
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

template <typename T, class PL, int NT, int EP>
static int sfft_rows_inv_tma_launch(const typename V2<T>::type* in, T* out, int H, long long nrows, SC2RParams<T> ep, const void* tw,
                                    cudaStream_t st) {
    typedef typename V2<T>::type C2;
    constexpr int TG = SFFT_ROW_TMA_G;
    typedef SRowTmaCfg<T, PL::N, TG> TCfg;
    const size_t smem = TCfg::smem(EP == 2, EP == 1);
    int rc;
    if ((rc = sfft_set_smem(sfft_rows_inv_tma_kernel<T, PL, NT, TG, EP>, smem)) != NB_OK) return rc;
    const long long npairs = (nrows + 1) / 2;
    const long long ntiles = (npairs + TG - 1) / TG;
    const long long cap = (long long)nb_sm_count() * sfft_row_tma_minb(NT * TG);
    const unsigned pgrid = (unsigned)(ntiles < cap ? ntiles : cap);
    { NB_PROF_BEGIN(st); sfft_rows_inv_tma_kernel<T, PL, NT, TG, EP><<<pgrid, NT * TG, smem, st>>>(in, out, H, (int)nrows, (int)ntiles, ep, (const C2*)tw); NB_PROF_END("sfft_rows_inv_kernel", st); }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

template <typename T>
static int sfft_rows_inv_t(const typename V2<T>::type* in, T* out, int B, int H, int W, SC2RParams<T> ep, cudaStream_t st) {
    typedef typename V2<T>::type C2;
    const int prec = sizeof(T) == 8 ? NB_PREC_F64 : NB_PREC_F32;
    const long long nrows = (long long)B * H;
    NB_REQUIRE(nrows < (1LL << 31), NB_ERR_ARG, "sfft: %lld rows in one call (at most 2^31 - 1)", nrows);
    const long long npairs = (nrows + 1) / 2;
    constexpr int G = SFFT_ROW_G;
    // the TMA-fed kernel stages what the store needs as pairs of rows: half spectra and zero-reference rows are always adjacent
    // and 16-byte aligned (W even); the byte mask needs even H (pairs inside one image) and W a multiple of 8.  A call with
    // both a mask and a zero reference (none on the path) goes to the register-staged kernel.
    const bool maskOk = (H % 2) == 0 && (W % 8) == 0;
    const bool tmaOk = sfft_use_tma() && nrows < (1LL << 30) && !(ep.mask && ep.zeroRef) && (!ep.mask || maskOk);
    int rc;
    const void* tw;
#define SFFT_CASE(NN, R0, R1, R2, NT, CR0, CR1, CR2, CNT)                                                                              \
    if (W == NN) {                                                                                                        \
        typedef Plan<NN, R0, R1, R2> PL;                                                                                  \
        typedef SRowCfg<T, NN, G> Cfg;                                                                                    \
        if ((rc = sfft_tw<PL>(prec, &tw)) != NB_OK) return rc;                                                            \
        if (tmaOk) {                                                                                                      \
            if (ep.mask) return sfft_rows_inv_tma_launch<T, PL, NT, 1>(in, out, H, nrows, ep, tw, st);                    \
            if (ep.zeroRef) return sfft_rows_inv_tma_launch<T, PL, NT, 2>(in, out, H, nrows, ep, tw, st);                 \
            return sfft_rows_inv_tma_launch<T, PL, NT, 0>(in, out, H, nrows, ep, tw, st);                                 \
        }                                                                                                                 \
        if ((rc = sfft_set_smem(sfft_rows_inv_kernel<T, PL, NT, G>, Cfg::SMEM)) != NB_OK) return rc;                      \
        const unsigned grid = (unsigned)((npairs + G - 1) / G);                                                           \
        { NB_PROF_BEGIN(st); sfft_rows_inv_kernel<T, PL, NT, G><<<grid, NT * G, Cfg::SMEM, st>>>(in, out, H, nrows, ep, (const C2*)tw); NB_PROF_END("sfft_rows_inv_kernel", st); } \
        NB_CHECK_LAUNCH();                                                                                                \
        return NB_OK;                                                                                                     \
    }
    SFFT_FOR_SIZES_INV(SFFT_CASE)
#undef SFFT_CASE
    nb_set_error("sfft: width %d has no compile-time plan", W);
    return NB_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------------
// Row half of the separable window undo (K8: enmap.apply_window(data, pow = -1), filters.py:101-104).  The pixel window is a
// product wy(ky) wx(kx), so dividing the 2-D spectrum by it is a filter along y followed by a filter along x, each of them a
// transform, a REAL and EVEN multiplier and the inverse transform.  Such a multiplier commutes with the two-for-one packing:
// with z = a + i b for two real rows a, b the product w Z is the transform of (w * a) + i (w * b), so the kernel transforms the
// packed pair, multiplies by wx[min(k, W - k)], runs the reversed plan backwards (whose first pass wants exactly what the
// forward's last pass left in each thread: no exchange in between) and stores the two real rows -- the half spectra never exist.
// Together with the column kernel in MODE 2 on pairs of adjacent real columns (sfft_cols.cuh) the undo moves 40 bytes per pixel
// instead of the 56 of forward rows / column window / inverse rows.  out may hold the data before the undo: where it is exactly
// zero the result is zero (restoreZeros; the pattern is read up front, one bit per register slot, as in the inverse kernel).
#ifndef SFFT_ROWW_G
#define SFFT_ROWW_G 2
#endif
#ifndef SFFT_ROWW_MINB
#define SFFT_ROWW_MINB 3      // two row pairs per CTA, three CTAs per SM: 1.54 ms per 128 images of 840 x 1440 (2.00 with two CTAs; one pair per CTA 1.69)
#endif
template <typename T, class PL, int NT, int G>
__global__ void __launch_bounds__(NT * G, SFFT_ROWW_MINB)
sfft_rows_window_kernel(const T* __restrict__ in, T* out, long long nrows, const T* __restrict__ wx, T scale, int restoreZeros,
                        const typename V2<T>::type* __restrict__ tw, const typename V2<T>::type* __restrict__ twRev) {
    typedef SRowCfg<T, PL::N, G> Cfg;
    typedef typename Reversed<PL>::type PR;
    typedef Stage<PL, NT, 0> S0;
    typedef typename Last<PL, NT>::S SL;
    typedef typename Last<PR, NT>::S SRL;
    constexpr int W = PL::N;
    static_assert(SRL::REGS <= 32, "one bit per register slot");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int g = threadIdx.x / NT, t = threadIdx.x % NT;
    const long long r0 = ((long long)blockIdx.x * G + g) * 2, r1 = r0 + 1;
    T* s_re = smem + g * Cfg::LD;
    T* s_im = smem + (G + g) * Cfg::LD;
    const bool v0 = r0 < nrows, v1 = r1 < nrows;
    T* o0 = out + (size_t)(v0 ? r0 : 0) * W;
    T* o1 = out + (size_t)(v1 ? r1 : 0) * W;
    unsigned zb0 = 0u, zb1 = 0u;
    if (restoreZeros) {
#pragma unroll
        for (int i = 0; i < SRL::IT; ++i) {
            const int j = t + i * NT;
#pragma unroll
            for (int r = 0; r < SRL::R; ++r) {
                if (SRL::IT * NT == SRL::NB || j < SRL::NB) {
                    const int n = j + r * SRL::NB;
                    if (v0 && o0[n] == (T)0) zb0 |= 1u << (i * SRL::R + r);
                    if (v1 && o1[n] == (T)0) zb1 |= 1u << (i * SRL::R + r);
                }
            }
        }
    }
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
                const int k = j + r * SL::NB;
                const T w = wx[2 * k <= W ? k : W - k] * scale;
                yr[i * SL::R + r] *= w;
                yi[i * SL::R + r] *= w;
            }
        }
        run_from<PR, T, NT, +1, 0>(yr, yi, t, s_re, s_im, twRev, BlockSync(), [&](auto& zr, auto& zi) {
#pragma unroll
            for (int i = 0; i < SRL::IT; ++i) {
                const int j = t + i * NT;
                if (SRL::IT * NT > SRL::NB && j >= SRL::NB) continue;
#pragma unroll
                for (int r = 0; r < SRL::R; ++r) {
                    const int q = i * SRL::R + r, n = j + r * SRL::NB;
                    if (v0) o0[n] = ((zb0 >> q) & 1u) ? (T)0 : zr[q];
                    if (v1) o1[n] = ((zb1 >> q) & 1u) ? (T)0 : zi[q];
                }
            }
        });
    });
}

// The same, fed by the copy unit: a persistent CTA per resident slot walks over row pairs; the pair of input rows (adjacent in
// memory: one cp.async.bulk) and the pair of rows of the old data (for the zero pattern) land in two staging buffers tracked by an
// mbarrier, and the copy of the next pair is issued as soon as every thread has gathered from them -- under the six passes and the
// store of the current pair.  In the register-staged kernel the long-scoreboard stall (its exposed loads) was 17 warp-cycles per
// issued instruction.  Needs an even number of rows (every pair complete).
#ifndef SFFT_ROWW_TMA_MINB
#define SFFT_ROWW_TMA_MINB 2
#endif
template <typename T, class PL, int NT, bool restoreZeros>
__global__ void __launch_bounds__(NT, SFFT_ROWW_TMA_MINB)
sfft_rows_window_tma_kernel(const T* __restrict__ in, T* out, int nrows, int ntiles, const T* __restrict__ wx0, T scale,
                            const typename V2<T>::type* __restrict__ tw0, const typename V2<T>::type* __restrict__ twRev0) {
    typedef typename V2<T>::type C2;
    typedef typename Reversed<PL>::type PR;
    typedef Stage<PL, NT, 0> S0;
    typedef typename Last<PL, NT>::S SL;
    typedef typename Last<PR, NT>::S SRL;
    constexpr int W = PL::N, LD = W + 2;
    constexpr uint32_t PAIR_BYTES = 2u * W * (uint32_t)sizeof(T);
    static_assert(SRL::REGS <= 32, "one bit per register slot");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* stage = reinterpret_cast<T*>(smem_raw);                       // [2][W] input rows
    T* zref = stage + 2 * W;                                         // [2][W] rows of the old data
    T* s_re = zref + 2 * W;
    T* s_im = s_re + LD;
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_im + LD);
    const int t = threadIdx.x;
    auto issue = [&](int tile) {
        const size_t off = (size_t)tile * 2 * W;
        tma::mbar_arrive_expect_tx(bar, restoreZeros ? 2u * PAIR_BYTES : PAIR_BYTES);
        tma::bulk_g2s(stage, in + off, PAIR_BYTES, bar);
        if (restoreZeros) tma::bulk_g2s(zref, out + off, PAIR_BYTES, bar);
    };
    if (t == 0) {
        tma::mbar_init(bar, 1);
        tma::fence_mbar_init();
    }
    __syncthreads();
    int tile = (int)blockIdx.x;
    if (t == 0 && tile < ntiles) issue(tile);
    uint32_t phase = 0;
#pragma unroll 1
    for (; tile < ntiles; tile += (int)gridDim.x) {
        // (laundered per iteration: the compiler would otherwise hoist the loop-invariant twiddle loads out of the loop)
        const C2* tw = tw0;
        const C2* twRev = twRev0;
        const T* wx = wx0;
        asm volatile("" : "+l"(tw));
        asm volatile("" : "+l"(twRev));
        asm volatile("" : "+l"(wx));
        tma::mbar_wait(bar, phase);
        phase ^= 1u;
        T xr[S0::REGS], xi[S0::REGS];
#pragma unroll
        for (int i = 0; i < S0::IT; ++i) {
            const int j = t + i * NT;
#pragma unroll
            for (int r = 0; r < S0::R; ++r) {
                T a = 0, b = 0;
                if (S0::IT * NT == S0::NB || j < S0::NB) {
                    const int n = j + r * S0::NB;
                    a = stage[n];
                    b = stage[W + n];
                }
                xr[i * S0::R + r] = a;
                xi[i * S0::R + r] = b;
            }
        }
        unsigned zb0 = 0u, zb1 = 0u;
        if (restoreZeros) {
#pragma unroll
            for (int i = 0; i < SRL::IT; ++i) {
                const int j = t + i * NT;
#pragma unroll
                for (int r = 0; r < SRL::R; ++r) {
                    if (SRL::IT * NT == SRL::NB || j < SRL::NB) {
                        const int n = j + r * SRL::NB;
                        if (zref[n] == (T)0) zb0 |= 1u << (i * SRL::R + r);
                        if (zref[W + n] == (T)0) zb1 |= 1u << (i * SRL::R + r);
                    }
                }
            }
        }
        __syncthreads();      // every thread has read the staging buffers: they can take the next pair
        if (t == 0) {
            const int next = tile + (int)gridDim.x;
            if (next < ntiles) {
                tma::fence_proxy_async();
                issue(next);
            }
        }
        T* o0 = out + (size_t)tile * 2 * W;
        T* o1 = o0 + W;
        run_from<PL, T, NT, -1, 0>(xr, xi, t, s_re, s_im, tw, BlockSync(), [&](auto& yr, auto& yi) {
#pragma unroll
            for (int i = 0; i < SL::IT; ++i) {
                const int j = t + i * NT;
                if (SL::IT * NT > SL::NB && j >= SL::NB) continue;
#pragma unroll
                for (int r = 0; r < SL::R; ++r) {
                    const int k = j + r * SL::NB;
                    const T w = wx[2 * k <= W ? k : W - k] * scale;
                    yr[i * SL::R + r] *= w;
                    yi[i * SL::R + r] *= w;
                }
            }
            run_from<PR, T, NT, +1, 0>(yr, yi, t, s_re, s_im, twRev, BlockSync(), [&](auto& zr, auto& zi) {
#pragma unroll
                for (int i = 0; i < SRL::IT; ++i) {
                    const int j = t + i * NT;
                    if (SRL::IT * NT > SRL::NB && j >= SRL::NB) continue;
#pragma unroll
                    for (int r = 0; r < SRL::R; ++r) {
                        const int q = i * SRL::R + r, n = j + r * SRL::NB;
                        o0[n] = ((zb0 >> q) & 1u) ? (T)0 : zr[q];
                        o1[n] = ((zb1 >> q) & 1u) ? (T)0 : zi[q];
                    }
                }
            });
        });
        __syncthreads();      // the exchange buffer is free for the next pair's first scatter
    }
}

template <typename T>
static int sfft_rows_window_t(const T* in, T* out, int B, int H, int W, const T* wx, double scale, int restoreZeros, cudaStream_t st) {
    typedef typename V2<T>::type C2;
    const int prec = sizeof(T) == 8 ? NB_PREC_F64 : NB_PREC_F32;
    const long long nrows = (long long)B * H;
    const long long npairs = (nrows + 1) / 2;
    constexpr int G = SFFT_ROWW_G;
    static const bool noWinTma = getenv("NEMO_B200_ROWW_NO_TMA") != nullptr;
    const bool tmaOk = sfft_use_tma() && !noWinTma && (nrows % 2) == 0 && nrows < (1LL << 30) && (W % 2) == 0;
    int rc;
    const void *tw, *twRev;
#define SFFT_CASE(NN, R0, R1, R2, NT, CR0, CR1, CR2, CNT)                                                                              \
    if (W == NN) {                                                                                                        \
        typedef Plan<NN, R0, R1, R2> PL;                                                                                  \
        typedef SRowCfg<T, NN, G> Cfg;                                                                                    \
        if ((rc = sfft_tw<PL>(prec, &tw)) != NB_OK) return rc;                                                            \
        if ((rc = sfft_tw<typename Reversed<PL>::type>(prec, &twRev)) != NB_OK) return rc;                                \
        if (tmaOk) {                                                                                                      \
            const size_t smem = ((size_t)4 * NN + 2 * (NN + 2)) * sizeof(T) + 16;                                         \
            if ((rc = sfft_set_smem(sfft_rows_window_tma_kernel<T, PL, NT, true>, smem)) != NB_OK) return rc;             \
            if ((rc = sfft_set_smem(sfft_rows_window_tma_kernel<T, PL, NT, false>, smem)) != NB_OK) return rc;            \
            const long long cap = (long long)nb_sm_count() * SFFT_ROWW_TMA_MINB;                                          \
            const unsigned pgrid = (unsigned)(npairs < cap ? npairs : cap);                                               \
            { NB_PROF_BEGIN(st);                                                                                          \
              if (restoreZeros) sfft_rows_window_tma_kernel<T, PL, NT, true><<<pgrid, NT, smem, st>>>(in, out, (int)nrows, (int)npairs, wx, (T)scale, (const C2*)tw, (const C2*)twRev); \
              else sfft_rows_window_tma_kernel<T, PL, NT, false><<<pgrid, NT, smem, st>>>(in, out, (int)nrows, (int)npairs, wx, (T)scale, (const C2*)tw, (const C2*)twRev); \
              NB_PROF_END("sfft_rows_window_kernel", st); }                                                               \
            NB_CHECK_LAUNCH();                                                                                            \
            return NB_OK;                                                                                                 \
        }                                                                                                                 \
        if ((rc = sfft_set_smem(sfft_rows_window_kernel<T, PL, NT, G>, Cfg::SMEM)) != NB_OK) return rc;                   \
        const unsigned grid = (unsigned)((npairs + G - 1) / G);                                                           \
        { NB_PROF_BEGIN(st); sfft_rows_window_kernel<T, PL, NT, G><<<grid, NT * G, Cfg::SMEM, st>>>(in, out, nrows, wx, (T)scale, restoreZeros, (const C2*)tw, (const C2*)twRev); NB_PROF_END("sfft_rows_window_kernel", st); } \
        NB_CHECK_LAUNCH();                                                                                                \
        return NB_OK;                                                                                                     \
    }
    SFFT_FOR_SIZES_WIN(SFFT_CASE)
#undef SFFT_CASE
    nb_set_error("sfft: width %d has no compile-time plan", W);
    return NB_ERR_UNSUPPORTED;
}

int nb_sfft_rows_window(int prec, const void* in, void* out, int B, int H, int W, const void* wx, double scale, int restoreZeros, cudaStream_t st) {
    if (prec == NB_PREC_F64) return sfft_rows_window_t<double>((const double*)in, (double*)out, B, H, W, (const double*)wx, scale, restoreZeros, st);
    return sfft_rows_window_t<float>((const float*)in, (float*)out, B, H, W, (const float*)wx, scale, restoreZeros, st);
}

int nb_sfft_rows_fwd(int prec, const void* in, void* out, int B, int H, int W, int apodPix, double scale, cudaStream_t st) {
    if (prec == NB_PREC_F64) return sfft_rows_fwd_t<double>((const double*)in, (double2*)out, B, H, W, apodPix, scale, st);
    return sfft_rows_fwd_t<float>((const float*)in, (float2*)out, B, H, W, apodPix, scale, st);
}

int nb_sfft_rows_inv(int prec, const void* in, void* out, int B, int H, int W, double scale, const double* norm, const unsigned char* mask,
                     int maskDiv, const void* zeroRef, int* zeroFlag, cudaStream_t st) {
    if (prec == NB_PREC_F64) {
        SC2RParams<double> ep = {scale, norm, mask, maskDiv < 1 ? 1 : maskDiv, (const double*)zeroRef, zeroFlag};
        return sfft_rows_inv_t<double>((const double2*)in, (double*)out, B, H, W, ep, st);
    }
    SC2RParams<float> ep = {(float)scale, norm, mask, maskDiv < 1 ? 1 : maskDiv, (const float*)zeroRef, zeroFlag};
    return sfft_rows_inv_t<float>((const float2*)in, (float*)out, B, H, W, ep, st);
}
