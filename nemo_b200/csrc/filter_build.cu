// Matched-filter construction kernels (K2, K3, K5 of SURVEY.md section 2a).
//   nb_xpower_smooth : noiseCov[i][j] = gaussian_filter(Re(F_i conj F_j), (3,3))      filters.py:568-587
//   nb_build_filter  : filt[:,k] = inv(noiseCov[:,:,k]) @ (w * |fSignal[:,k]|)         filters.py:624-630
//   nb_fold_filter   : (filt(k) + filt(-k)) / 2 on the half plane
//   nb_template_norm : argmax / max / spline value at the centre -> signalNorm         filters.py:653-662, 686-687
//   nb_plane_values_at : per-frequency filtered-template values at one pixel            filters.py:664-670
// All arithmetic here is fp64 whatever the storage type of the spectra.
#include "nemo_b200_internal.h"
#include "spline.cuh"
#include "gauss.cuh"
#include <math.h>
#include <stdlib.h>

// ------------------------------------------------------------------------------------------------------
// K2 pass 1: one CTA per (image, ky).  The cross-power row over the full kx range is rebuilt from the half
// spectrum (kx > W/2 comes from row -ky at W - kx), kept in shared memory as fp64, and smoothed along kx.
template <typename T, int F>
__global__ void __launch_bounds__(256) xpower_smooth_x_kernel(const typename V2<T>::type* __restrict__ fmaps,
                                                              double* __restrict__ tmp, int H, int W, GaussW g) {
    typedef typename V2<T>::type C2;
    constexpr int P = F * (F + 1) / 2;
    extern __shared__ double s_np[];  // [P][W]
    const int ky = blockIdx.x, b = blockIdx.y;
    const int Wc = W / 2 + 1;
    const int kym = (H - ky) % H;
    const C2* base = fmaps + (size_t)b * F * H * Wc;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        const int row = (x < Wc) ? ky : kym;
        const int col = (x < Wc) ? x : (W - x);
        double re[F], im[F];
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const C2 v = base[((size_t)f * H + row) * Wc + col];
            re[f] = (double)v.x;
            im[f] = (double)v.y;
        }
        int p = 0;
#pragma unroll
        for (int i = 0; i < F; ++i)
#pragma unroll
            for (int j = i; j < F; ++j) {
                s_np[p * W + x] = re[i] * re[j] + im[i] * im[j];
                ++p;
            }
    }
    __syncthreads();
    double* out = tmp + ((size_t)b * P * H + ky) * W;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const double* s = s_np + p * W;
            double acc = 0;
            if (x >= g.r && x + g.r < W) {
                for (int d = -g.r; d <= g.r; ++d) acc += g.w[d + g.r] * s[x + d];
            } else {
                for (int d = -g.r; d <= g.r; ++d) acc += g.w[d + g.r] * s[reflect_idx(x + d, W)];
            }
            out[(size_t)p * H * W + x] = acc;
        }
    }
}

// K2 pass 2 (also the column pass of K12): smooth along y.  CTA tile 32 columns x TY rows, halo in shared memory.
#define GY_TX 32
#define GY_TY 64
__global__ void __launch_bounds__(256) smooth_y_f64_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                           int H, int W, GaussW g) {
    extern __shared__ double s_t[];  // [(GY_TY + 2r)][GY_TX]
    const int x0 = blockIdx.x * GY_TX, y0 = blockIdx.y * GY_TY;
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int tx = threadIdx.x % GY_TX, ty = threadIdx.x / GY_TX;  // 32 x 8
    const int rows = GY_TY + 2 * g.r;
    const int x = x0 + tx;
    for (int i = ty; i < rows; i += 8) {
        const int y = reflect_idx(y0 - g.r + i, H);
        s_t[i * GY_TX + tx] = (x < W) ? in[plane + (size_t)y * W + x] : 0.0;
    }
    __syncthreads();
    if (x >= W) return;
    for (int i = ty; i < GY_TY; i += 8) {
        const int y = y0 + i;
        if (y >= H) break;
        double acc = 0;
        for (int d = 0; d <= 2 * g.r; ++d) acc += g.w[d] * s_t[(i + d) * GY_TX + tx];
        out[plane + (size_t)y * W + x] = acc;
    }
}

// Register-blocked versions of the two passes for a radius known at compile time (sigma = 3 -> radius 12, the only
// value K2 uses): a thread computes RB = 8 consecutive outputs along the smoothing axis from RB + 2 R inputs it reads once,
// with the weights in registers -- 4 shared-memory reads per output instead of 25 (the plain kernels above were bound by
// the shared-memory pipe at ~80 % of its peak).  Same products in the same order, so the results are bit-identical.
#define GS_RB 8
__host__ __device__ constexpr int gs_pad(int i) { return i + (i >> 3); }   // one pad word per 8: lanes 8 apart hit different banks

template <typename T, int F, int R>
__global__ void __launch_bounds__(512) xpower_smooth_x_rb_kernel(const typename V2<T>::type* __restrict__ fmaps,
                                                                 double* __restrict__ tmp, int H, int W, GaussW g) {
    typedef typename V2<T>::type C2;
    constexpr int P = F * (F + 1) / 2;
    extern __shared__ double s_np[];  // [P][LD], position x + R holds kx = x, reflected halos on both sides (W rounded up to 8)
    const int LD = gs_pad((W + GS_RB - 1) / GS_RB * GS_RB + 2 * R) + 1;   // the last thread of a row reads a full block of 8 + 2 R inputs
    const int ky = blockIdx.x, b = blockIdx.y;
    const int Wc = W / 2 + 1;
    const int kym = (H - ky) % H;
    const C2* base = fmaps + (size_t)b * F * H * Wc;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        const int row = (x < Wc) ? ky : kym;
        const int col = (x < Wc) ? x : (W - x);
        double re[F], im[F];
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const C2 v = base[((size_t)f * H + row) * Wc + col];
            re[f] = (double)v.x;
            im[f] = (double)v.y;
        }
        int p = 0;
#pragma unroll
        for (int i = 0; i < F; ++i)
#pragma unroll
            for (int j = i; j < F; ++j) {
                s_np[p * LD + gs_pad(x + R)] = re[i] * re[j] + im[i] * im[j];
                ++p;
            }
    }
    __syncthreads();
    for (int h = threadIdx.x; h < 2 * R * P; h += blockDim.x) {     // scipy 'reflect': position -1 - h mirrors h, W + h mirrors W - 1 - h
        const int p = h / (2 * R), k = h % (2 * R);
        double* s = s_np + p * LD;
        if (k < R) s[gs_pad(R - 1 - k)] = s[gs_pad(R + k)];
        else s[gs_pad(W + R + (k - R))] = s[gs_pad(W + R - 1 - (k - R))];
    }
    __syncthreads();
    double w[2 * R + 1];
#pragma unroll
    for (int d = 0; d <= 2 * R; ++d) w[d] = g.w[d];
    double* out = tmp + ((size_t)b * P * H + ky) * W;
    for (int x0 = threadIdx.x * GS_RB; x0 < W; x0 += blockDim.x * GS_RB) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const double* s = s_np + p * LD;
            double acc[GS_RB];
#pragma unroll
            for (int o = 0; o < GS_RB; ++o) acc[o] = 0;
#pragma unroll
            for (int d = 0; d < GS_RB + 2 * R; ++d) {
                const double v = s[gs_pad(x0 + d)];
#pragma unroll
                for (int o = 0; o < GS_RB; ++o)
                    if (d - o >= 0 && d - o <= 2 * R) acc[o] += w[d - o] * v;
            }
            double* dst = out + (size_t)p * H * W + x0;
            if (x0 + GS_RB <= W) {
#pragma unroll
                for (int o = 0; o < GS_RB; o += 2) *reinterpret_cast<double2*>(dst + o) = make_double2(acc[o], acc[o + 1]);
            } else {
#pragma unroll
                for (int o = 0; o < GS_RB; ++o)
                    if (x0 + o < W) dst[o] = acc[o];
            }
        }
    }
}

template <int R>
__global__ void __launch_bounds__(512) smooth_y_rb_kernel(const double* __restrict__ in, double* __restrict__ out, int H, int W, GaussW g) {
    constexpr int TY = 16 * GS_RB;     // rows per CTA: 16 thread rows x 8 outputs
    extern __shared__ double s_t[];    // [TY + 2 R][32]
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * TY;
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x = x0 + tx;
    for (int i = ty; i < TY + 2 * R; i += 16) {
        const int y = reflect_idx(y0 - R + i, H);
        s_t[i * 32 + tx] = (x < W) ? in[plane + (size_t)y * W + x] : 0.0;
    }
    __syncthreads();
    if (x >= W) return;
    double w[2 * R + 1];
#pragma unroll
    for (int d = 0; d <= 2 * R; ++d) w[d] = g.w[d];
    double acc[GS_RB];
#pragma unroll
    for (int o = 0; o < GS_RB; ++o) acc[o] = 0;
    const int i0 = ty * GS_RB;
#pragma unroll
    for (int d = 0; d < GS_RB + 2 * R; ++d) {
        const double v = s_t[(i0 + d) * 32 + tx];
#pragma unroll
        for (int o = 0; o < GS_RB; ++o)
            if (d - o >= 0 && d - o <= 2 * R) acc[o] += w[d - o] * v;
    }
#pragma unroll
    for (int o = 0; o < GS_RB; ++o) {
        const int y = y0 + i0 + o;
        if (y < H) out[plane + (size_t)y * W + x] = acc[o];
    }
}

// ------------------------------------------------------------------------------------------------------
// K3.  Solve the F x F system with partial pivoting (LAPACK-style); an exactly zero pivot marks the matrix
// singular and yields zeros, which is what the reference's "except: continue" leaves in self.filt.
template <int F> __device__ __forceinline__ void solve_sym(const double* c /*P packed upper*/, const double* rhs, double* x) {
    double A[F][F], bb[F];
    int p = 0;
#pragma unroll
    for (int i = 0; i < F; ++i)
#pragma unroll
        for (int j = i; j < F; ++j) {
            A[i][j] = c[p];
            A[j][i] = c[p];
            ++p;
        }
#pragma unroll
    for (int i = 0; i < F; ++i) bb[i] = rhs[i];
    bool singular = false;
#pragma unroll
    for (int k = 0; k < F; ++k) {
        // pivot search
        int piv = k;
        double best = fabs(A[k][k]);
#pragma unroll
        for (int i = k + 1; i < F; ++i) {
            const double v = fabs(A[i][k]);
            if (v > best) {
                best = v;
                piv = i;
            }
        }
#pragma unroll
        for (int i = k + 1; i < F; ++i) {
            if (i == piv) {
#pragma unroll
                for (int j = 0; j < F; ++j) {
                    const double t = A[k][j];
                    A[k][j] = A[i][j];
                    A[i][j] = t;
                }
                const double t = bb[k];
                bb[k] = bb[i];
                bb[i] = t;
            }
        }
        if (A[k][k] == 0.0) singular = true;
        // explicit round-to-nearest multiplies / subtractions: no FMA contraction, so every instance of this solve
        // (a mode and its mirror are solved by different threads) rounds identically
#pragma unroll
        for (int i = k + 1; i < F; ++i) {
            const double m = __ddiv_rn(A[i][k], A[k][k]);
#pragma unroll
            for (int j = k + 1; j < F; ++j) A[i][j] = __dsub_rn(A[i][j], __dmul_rn(m, A[k][j]));
            bb[i] = __dsub_rn(bb[i], __dmul_rn(m, bb[k]));
        }
    }
#pragma unroll
    for (int i = F - 1; i >= 0; --i) {
        double s = bb[i];
#pragma unroll
        for (int j = i + 1; j < F; ++j) s = __dsub_rn(s, __dmul_rn(A[i][j], x[j]));
        x[i] = __ddiv_rn(s, A[i][i]);
    }
    if (singular) {
#pragma unroll
        for (int i = 0; i < F; ++i) x[i] = 0.0;
    }
}

// |v| of a complex template value.  np.abs is hypot; CUDA's fp64 hypot is ~90 instructions (scaling against overflow), which
// made it the largest item of the filter build.  sqrt(x^2 + y^2) with one fused multiply-add is within 1 ulp of it wherever the
// squares neither underflow nor overflow; outside that range (and for an exactly zero value) hypot itself is used.
__device__ __forceinline__ double cabs_fast(double x, double y) {
    const double s = fma(x, x, y * y);
    if (!(s > 1e-280 && s < 1e280)) return hypot(x, y);
    return sqrt(s);
}

struct FreqW {
    double w[NB_MAX_FREQ];
};

template <typename T, int F>
__global__ void __launch_bounds__(256) build_filter_kernel(const double* __restrict__ cov,
                                                           const typename V2<T>::type* __restrict__ fsig, FreqW fw,
                                                           double* __restrict__ filtFull, T* __restrict__ filtH, int S,
                                                           int sigPerTile, int H, int W) {
    typedef typename V2<T>::type C2;
    constexpr int P = F * (F + 1) / 2;
    const int Wc = W / 2 + 1;
    const int kx = blockIdx.x * blockDim.x + threadIdx.x;
    const int ky = blockIdx.y;
    const int o = blockIdx.z;
    if (kx >= Wc) return;
    const int tile = o / S;
    const int sidx = sigPerTile ? o : (o % S);
    const int kym = (H - ky) % H, kxm = (W - kx) % W;
    const double* cv = cov + (size_t)tile * P * H * W;
    double rhs[F], rhs2[F], c1[P], c2[P], x1[F], x2[F];
    // the mirror mode -k has the conjugate template value, hence the same modulus, except on the self-conjugate
    // columns kx = 0 and (even W) W/2 where it is a stored element of its own (equal only up to FFT rounding)
    const bool ownMirror = (kx == 0 || 2 * kx == W);
#pragma unroll
    for (int f = 0; f < F; ++f) {
        const C2 v = fsig[(((size_t)sidx * F + f) * H + ky) * Wc + kx];
        rhs[f] = __dmul_rn(fw.w[f], cabs_fast((double)v.x, (double)v.y));  // np.abs of a complex number is hypot
        rhs2[f] = rhs[f];
        if (ownMirror) {
            const C2 vm = fsig[(((size_t)sidx * F + f) * H + kym) * Wc + kx];
            rhs2[f] = __dmul_rn(fw.w[f], cabs_fast((double)vm.x, (double)vm.y));
        }
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
        c1[p] = cv[((size_t)p * H + ky) * W + kx];
        c2[p] = cv[((size_t)p * H + kym) * W + kxm];
    }
    solve_sym<F>(c1, rhs, x1);
    solve_sym<F>(c2, rhs2, x2);
#pragma unroll
    for (int f = 0; f < F; ++f) {
        filtH[(((size_t)o * F + f) * H + ky) * Wc + kx] = (T)(0.5 * __dadd_rn(x1[f], x2[f]));
        if (filtFull) {
            double* ff = filtFull + ((size_t)o * F + f) * H * W;
            ff[(size_t)ky * W + kx] = x1[f];
            if (kx > 0 && 2 * kx != W) ff[(size_t)kym * W + kxm] = x2[f];
        }
    }
}

// The same solve for ALL filter scales of a tile in one thread: the covariance of a mode (and of its mirror) is read and
// factorised once, every scale then costs its template modulus, the recorded row exchanges and multipliers applied to its
// right-hand side, and the back substitution.  The operations on a right-hand side are exactly those of solve_sym (same
// multipliers, same order, explicit round-to-nearest), so the filters are bit-identical to build_filter_kernel's; what goes
// away is 2 P covariance loads, the pivot search, the elimination divisions and the index arithmetic per scale -- the
// one-scale kernel was issue-bound (about 600 instructions per mode and scale, a quarter of them fp64).
template <int F> struct LUFactor {
    double U[F][F];      // upper triangle after elimination (rows exchanged)
    double m[F][F];      // multipliers m[i][k], i > k
    int piv[F];          // row exchanged with row k at step k
    bool singular;
};

template <int F> __device__ __forceinline__ void lu_factor(const double* c /*P packed upper*/, LUFactor<F>& lu) {
    int p = 0;
#pragma unroll
    for (int i = 0; i < F; ++i)
#pragma unroll
        for (int j = i; j < F; ++j) {
            lu.U[i][j] = c[p];
            lu.U[j][i] = c[p];
            ++p;
        }
    lu.singular = false;
#pragma unroll
    for (int k = 0; k < F; ++k) {
        int piv = k;
        double best = fabs(lu.U[k][k]);
#pragma unroll
        for (int i = k + 1; i < F; ++i) {
            const double v = fabs(lu.U[i][k]);
            if (v > best) {
                best = v;
                piv = i;
            }
        }
        lu.piv[k] = piv;
#pragma unroll
        for (int i = k + 1; i < F; ++i) {
            if (i == piv) {
#pragma unroll
                for (int j = 0; j < F; ++j) {
                    const double t = lu.U[k][j];
                    lu.U[k][j] = lu.U[i][j];
                    lu.U[i][j] = t;
                }
            }
        }
        if (lu.U[k][k] == 0.0) lu.singular = true;
#pragma unroll
        for (int i = k + 1; i < F; ++i) {
            const double mm = __ddiv_rn(lu.U[i][k], lu.U[k][k]);
            lu.m[i][k] = mm;
#pragma unroll
            for (int j = k + 1; j < F; ++j) lu.U[i][j] = __dsub_rn(lu.U[i][j], __dmul_rn(mm, lu.U[k][j]));
        }
    }
}

template <int F> __device__ __forceinline__ void lu_solve(const LUFactor<F>& lu, const double* rhs, double* x) {
    double bb[F];
#pragma unroll
    for (int i = 0; i < F; ++i) bb[i] = rhs[i];
#pragma unroll
    for (int k = 0; k < F; ++k) {
#pragma unroll
        for (int i = k + 1; i < F; ++i) {
            if (i == lu.piv[k]) {
                const double t = bb[k];
                bb[k] = bb[i];
                bb[i] = t;
            }
        }
#pragma unroll
        for (int i = k + 1; i < F; ++i) bb[i] = __dsub_rn(bb[i], __dmul_rn(lu.m[i][k], bb[k]));
    }
#pragma unroll
    for (int i = F - 1; i >= 0; --i) {
        double s = bb[i];
#pragma unroll
        for (int j = i + 1; j < F; ++j) s = __dsub_rn(s, __dmul_rn(lu.U[i][j], x[j]));
        x[i] = __ddiv_rn(s, lu.U[i][i]);
    }
    if (lu.singular) {
#pragma unroll
        for (int i = 0; i < F; ++i) x[i] = 0.0;
    }
}

template <typename T, int F>
__global__ void __launch_bounds__(128) build_filter_multi_kernel(const double* __restrict__ cov,
                                                                 const typename V2<T>::type* __restrict__ fsig, FreqW fw,
                                                                 double* __restrict__ filtFull, T* __restrict__ filtH, int S,
                                                                 int sigPerTile, int H, int W) {
    typedef typename V2<T>::type C2;
    constexpr int P = F * (F + 1) / 2;
    const int Wc = W / 2 + 1;
    const int kx = blockIdx.x * blockDim.x + threadIdx.x;
    const int ky = blockIdx.y;
    const int tile = blockIdx.z;
    if (kx >= Wc) return;
    const int kym = (H - ky) % H, kxm = (W - kx) % W;
    const double* cv = cov + (size_t)tile * P * H * W;
    const bool ownMirror = (kx == 0 || 2 * kx == W);       // see build_filter_kernel
    const bool hasMirrorOut = (kx > 0 && 2 * kx != W);
    LUFactor<F> lu1, lu2;
    {
        double c1[P], c2[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            c1[p] = cv[((size_t)p * H + ky) * W + kx];
            c2[p] = cv[((size_t)p * H + kym) * W + kxm];
        }
        lu_factor<F>(c1, lu1);
        lu_factor<F>(c2, lu2);
    }
    const size_t planeH = (size_t)H * Wc, planeF = (size_t)H * W;
    const size_t offH = (size_t)ky * Wc + kx, offHm = (size_t)kym * Wc + kx;
    const C2* sg = fsig + (sigPerTile ? (size_t)tile * S * F * planeH : (size_t)0);
    T* dst = filtH + (size_t)tile * S * F * planeH + offH;
    double* ff = filtFull ? filtFull + (size_t)tile * S * F * planeF : nullptr;
#pragma unroll 2
    for (int s = 0; s < S; ++s) {
        double rhs[F], rhs2[F], x1[F], x2[F];
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const C2 v = sg[((size_t)s * F + f) * planeH + offH];
            rhs[f] = __dmul_rn(fw.w[f], cabs_fast((double)v.x, (double)v.y));
            rhs2[f] = rhs[f];
            if (ownMirror) {
                const C2 vm = sg[((size_t)s * F + f) * planeH + offHm];
                rhs2[f] = __dmul_rn(fw.w[f], cabs_fast((double)vm.x, (double)vm.y));
            }
        }
        lu_solve<F>(lu1, rhs, x1);
        lu_solve<F>(lu2, rhs2, x2);
#pragma unroll
        for (int f = 0; f < F; ++f) {
            dst[((size_t)s * F + f) * planeH] = (T)(0.5 * __dadd_rn(x1[f], x2[f]));
            if (ff) {
                double* g = ff + ((size_t)s * F + f) * planeF;
                g[(size_t)ky * W + kx] = x1[f];
                if (hasMirrorOut) g[(size_t)kym * W + kxm] = x2[f];
            }
        }
    }
}

// Templates shared by the tiles of a batch (the usual case): a thread keeps the right-hand sides of up to SC scales of its mode
// in registers -- every template modulus is read and computed ONCE per batch -- and walks over the tiles: covariance of the mode
// and of its mirror (the next tile's already requested), two factorisations, SC x 2 substitutions, SC x F stores.  DRAM traffic
// per batch drops from nT x (templates + covariance + filters) to templates + nT x (covariance + filters): the per-tile kernel
// above re-read the 310 MB of template spectra (16 scales x 2 frequencies of 840 x 1440) for every tile, 2.7 of its 4 GB.  The
// two self-conjugate columns (kx = 0 and, for even W, W / 2), whose mirror mode has a template value of its own, are redone by
// build_filter_selfmirror_kernel afterwards.
template <typename T, int F, int SC>
__global__ void __launch_bounds__(128) build_filter_batch_kernel(const double* __restrict__ cov,
                                                                 const typename V2<T>::type* __restrict__ fsig, FreqW fw,
                                                                 double* __restrict__ filtFull, T* __restrict__ filtH, int S, int nT,
                                                                 int H, int W) {
    typedef typename V2<T>::type C2;
    constexpr int P = F * (F + 1) / 2;
    // right-hand sides of the thread's mode: [scale][frequency][thread] (a thread only ever touches its own column of the
    // array, so no barrier is needed; kept out of registers so that the scale loop need not be unrolled -- fully unrolled it took
    // 246 registers, two CTAs per SM)
    __shared__ double s_rhs[SC * F][128];
    const int Wc = W / 2 + 1;
    const int kx = blockIdx.x * blockDim.x + threadIdx.x;
    const int ky = blockIdx.y;
    const int s0 = blockIdx.z * SC;
    if (kx >= Wc) return;
    const int ns = (S - s0) < SC ? (S - s0) : SC;
    const int kym = (H - ky) % H, kxm = (W - kx) % W;
    const bool hasMirrorOut = (kx > 0 && 2 * kx != W);
    const size_t planeH = (size_t)H * Wc, planeF = (size_t)H * W;
    const size_t offH = (size_t)ky * Wc + kx;
    const size_t off1 = (size_t)ky * W + kx, off2 = (size_t)kym * W + kxm;
    {
        const C2* sg = fsig + (size_t)s0 * F * planeH + offH;
#pragma unroll 8
        for (int q = 0; q < ns * F; ++q) {
            const C2 v = sg[(size_t)q * planeH];
            s_rhs[q][threadIdx.x] = __dmul_rn(fw.w[q % F], cabs_fast((double)v.x, (double)v.y));
        }
    }
    double c1[P], c2[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        c1[p] = cov[(size_t)p * planeF + off1];
        c2[p] = cov[(size_t)p * planeF + off2];
    }
    for (int tile = 0; tile < nT; ++tile) {
        LUFactor<F> lu1, lu2;
        lu_factor<F>(c1, lu1);
        lu_factor<F>(c2, lu2);
        if (tile + 1 < nT) {
            const double* cv = cov + (size_t)(tile + 1) * P * planeF;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                c1[p] = cv[(size_t)p * planeF + off1];
                c2[p] = cv[(size_t)p * planeF + off2];
            }
        }
        T* dst = filtH + ((size_t)tile * S + s0) * F * planeH + offH;
        double* ff = filtFull ? filtFull + ((size_t)tile * S + s0) * F * planeF : nullptr;
#pragma unroll 2
        for (int s = 0; s < ns; ++s) {
            double rhs[F], x1[F], x2[F];
#pragma unroll
            for (int f = 0; f < F; ++f) rhs[f] = s_rhs[s * F + f][threadIdx.x];
            lu_solve<F>(lu1, rhs, x1);
            lu_solve<F>(lu2, rhs, x2);
#pragma unroll
            for (int f = 0; f < F; ++f) {
                dst[(size_t)f * planeH] = (T)(0.5 * __dadd_rn(x1[f], x2[f]));
                if (ff) {
                    ff[(size_t)f * planeF + off1] = x1[f];
                    if (hasMirrorOut) ff[(size_t)f * planeF + off2] = x2[f];
                }
            }
            dst += (size_t)F * planeH;
            if (ff) ff += (size_t)F * planeF;
        }
    }
}

// the self-conjugate columns of build_filter_kernel's grid only: blockIdx.y selects kx = 0 or W / 2, threads run over ky
template <typename T, int F>
__global__ void __launch_bounds__(128) build_filter_selfmirror_kernel(const double* __restrict__ cov,
                                                                      const typename V2<T>::type* __restrict__ fsig, FreqW fw,
                                                                      double* __restrict__ filtFull, T* __restrict__ filtH, int S,
                                                                      int sigPerTile, int H, int W) {
    typedef typename V2<T>::type C2;
    constexpr int P = F * (F + 1) / 2;
    const int Wc = W / 2 + 1;
    const int ky = blockIdx.x * blockDim.x + threadIdx.x;
    const int kx = blockIdx.y == 0 ? 0 : W / 2;
    const int o = blockIdx.z;
    if (ky >= H) return;
    const int tile = o / S;
    const int sidx = sigPerTile ? o : (o % S);
    const int kym = (H - ky) % H;
    const double* cv = cov + (size_t)tile * P * H * W;
    double rhs[F], rhs2[F], c1[P], c2[P], x1[F], x2[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
        const C2 v = fsig[(((size_t)sidx * F + f) * H + ky) * Wc + kx];
        const C2 vm = fsig[(((size_t)sidx * F + f) * H + kym) * Wc + kx];
        rhs[f] = __dmul_rn(fw.w[f], cabs_fast((double)v.x, (double)v.y));
        rhs2[f] = __dmul_rn(fw.w[f], cabs_fast((double)vm.x, (double)vm.y));
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
        c1[p] = cv[((size_t)p * H + ky) * W + kx];
        c2[p] = cv[((size_t)p * H + kym) * W + kx];       // (W - kx) % W == kx on these columns
    }
    solve_sym<F>(c1, rhs, x1);
    solve_sym<F>(c2, rhs2, x2);
#pragma unroll
    for (int f = 0; f < F; ++f) {
        filtH[(((size_t)o * F + f) * H + ky) * Wc + kx] = (T)(0.5 * __dadd_rn(x1[f], x2[f]));
        if (filtFull) filtFull[((size_t)o * F + f) * H * W + (size_t)ky * W + kx] = x1[f];
    }
}

// ------------------------------------------------------------------------------------------------------
// K2 + K3 fused (nb_build_filter_fused): the smoothed covariance never goes to HBM.  A CTA owns a tile of up to FB_T x FB_T
// modes of the half plane and the tile of their mirror modes -k (whose covariance the folded filter also needs; with scipy's
// 'reflect' boundary the smoothed covariance is NOT point-symmetric near the array edges, so both are computed).  Row ky = 0 and
// column kx = 0 are tiles of their own, which makes every tile and its mirror a contiguous rectangle of the full plane.
// Per plane (i, j): cross power on the rectangle plus a halo of the smoothing radius, rebuilt from the half spectrum, smoothed
// along kx then ky in shared memory; then the F x F solve for every filter scale of the tile.
#define FB_T 32

template <typename T, int F>
__device__ __forceinline__ double cross_power_at(const typename V2<T>::type* __restrict__ base, int i, int j, int y, int x, int H,
                                                 int W, int Wc) {
    typedef typename V2<T>::type C2;
    // full-plane mode (y, x): stored at (y, x) for x < Wc, else the conjugate of ((H - y) % H, W - x) -- same cross power
    const int row = (x < Wc) ? y : (H - y) % H;
    const int col = (x < Wc) ? x : (W - x);
    const C2 a = base[((size_t)i * H + row) * Wc + col];
    if (i == j) return (double)a.x * (double)a.x + (double)a.y * (double)a.y;
    const C2 b = base[((size_t)j * H + row) * Wc + col];
    return (double)a.x * (double)b.x + (double)a.y * (double)b.y;
}

// smoothed plane (i, j) over rows ra .. ra + nr - 1, cols ca .. ca + nc - 1 of the full plane -> s_out[nr][FB_T]
template <typename T, int F>
__device__ __forceinline__ void smooth_region(const typename V2<T>::type* __restrict__ base, int i, int j, int ra, int nr, int ca, int nc,
                                              int H, int W, const GaussW& g, double* s_np, double* s_tx, double* s_out) {
    const int Wc = W / 2 + 1;
    const int r = g.r;
    const int hr = nr + 2 * r, hc = nc + 2 * r;
    for (int e = threadIdx.x; e < hr * hc; e += blockDim.x) {
        const int rr = e / hc, cc = e - rr * hc;
        const int y = reflect_idx(ra - r + rr, H), x = reflect_idx(ca - r + cc, W);
        s_np[e] = cross_power_at<T, F>(base, i, j, y, x, H, W, Wc);
    }
    __syncthreads();
    if (r == 12) {
        // sigma = 3 (the reference's fixed kernelSize): register-blocked passes, every input is read from shared memory once per
        // 8 (along kx) or 4 (along ky) outputs instead of once per output
        constexpr int R = 12, BX = 8, BY = 4;
        double wreg[2 * R + 1];
#pragma unroll
        for (int d = 0; d <= 2 * R; ++d) wreg[d] = g.w[d];
        const int segs = (nc + BX - 1) / BX;
        for (int task = threadIdx.x; task < hr * segs; task += blockDim.x) {
            const int rr = task / segs, c0 = (task - rr * segs) * BX;
            const double* src = s_np + rr * hc + c0;
            double in[BX + 2 * R];
#pragma unroll
            for (int k = 0; k < BX + 2 * R; ++k) in[k] = (c0 + k < hc) ? src[k] : 0.0;
#pragma unroll
            for (int o = 0; o < BX; ++o) {
                double acc = 0;
#pragma unroll
                for (int d = 0; d <= 2 * R; ++d) acc += wreg[d] * in[o + d];
                if (c0 + o < nc) s_tx[rr * FB_T + c0 + o] = acc;
            }
        }
        __syncthreads();
        const int segsY = (nr + BY - 1) / BY;
        for (int task = threadIdx.x; task < segsY * nc; task += blockDim.x) {
            const int sy = task / nc, cc = task - sy * nc;
            const int r0 = sy * BY;
            double in[BY + 2 * R];
#pragma unroll
            for (int k = 0; k < BY + 2 * R; ++k) in[k] = (r0 + k < hr) ? s_tx[(r0 + k) * FB_T + cc] : 0.0;
#pragma unroll
            for (int o = 0; o < BY; ++o) {
                double acc = 0;
#pragma unroll
                for (int d = 0; d <= 2 * R; ++d) acc += wreg[d] * in[o + d];
                if (r0 + o < nr) s_out[(r0 + o) * FB_T + cc] = acc;
            }
        }
        __syncthreads();
        return;
    }
    for (int e = threadIdx.x; e < hr * nc; e += blockDim.x) {
        const int rr = e / nc, cc = e - rr * nc;
        const double* src = s_np + rr * hc + cc;
        double acc = 0;
        for (int d = 0; d <= 2 * r; ++d) acc += g.w[d] * src[d];
        s_tx[rr * FB_T + cc] = acc;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nr * nc; e += blockDim.x) {
        const int rr = e / nc, cc = e - rr * nc;
        double acc = 0;
        for (int d = 0; d <= 2 * r; ++d) acc += g.w[d] * s_tx[(rr + d) * FB_T + cc];
        s_out[rr * FB_T + cc] = acc;
    }
    __syncthreads();
}

template <typename T, int F>
__global__ void __launch_bounds__(256) build_filter_fused_kernel(const typename V2<T>::type* __restrict__ fmaps,
                                                                 const typename V2<T>::type* __restrict__ fsig, FreqW fw,
                                                                 double* __restrict__ filtFull, T* __restrict__ filtH, int S,
                                                                 int sigPerTile, int H, int W, GaussW g) {
    typedef typename V2<T>::type C2;
    constexpr int P = F * (F + 1) / 2;
    extern __shared__ double s_fb[];
    const int Wc = W / 2 + 1;
    const int r = g.r;
    double* s_np = s_fb;                                         // [(FB_T + 2r)][(FB_T + 2r)]
    double* s_tx = s_np + (FB_T + 2 * r) * (FB_T + 2 * r);       // [(FB_T + 2r)][FB_T]
    double* s_c1 = s_tx + (FB_T + 2 * r) * FB_T;                 // [P][FB_T][FB_T]  covariance at the tile's modes
    double* s_c2 = s_c1 + P * FB_T * FB_T;                       // [P][FB_T][FB_T]  ... at their mirrors
    const int tile = blockIdx.z;
    // tile 0 along an axis is the single index 0; tile k >= 1 covers 1 + (k - 1) FB_T .. (clipped)
    const int x0 = blockIdx.x == 0 ? 0 : 1 + (blockIdx.x - 1) * FB_T;
    const int nx = blockIdx.x == 0 ? 1 : min(FB_T, Wc - x0);
    const int y0 = blockIdx.y == 0 ? 0 : 1 + (blockIdx.y - 1) * FB_T;
    const int ny = blockIdx.y == 0 ? 1 : min(FB_T, H - y0);
    // mirror rectangle: rows (H - ky) % H, cols (W - kx) % W; ascending starts
    const int my0 = blockIdx.y == 0 ? 0 : H - (y0 + ny - 1);
    const int mx0 = blockIdx.x == 0 ? 0 : W - (x0 + nx - 1);
    const C2* base = fmaps + (size_t)tile * F * H * Wc;
    int p = 0;
#pragma unroll
    for (int i = 0; i < F; ++i)
#pragma unroll
        for (int j = i; j < F; ++j) {
            smooth_region<T, F>(base, i, j, y0, ny, x0, nx, H, W, g, s_np, s_tx, s_c1 + p * FB_T * FB_T);
            smooth_region<T, F>(base, i, j, my0, ny, mx0, nx, H, W, g, s_np, s_tx, s_c2 + p * FB_T * FB_T);
            ++p;
        }
    for (int e = threadIdx.x; e < ny * nx; e += blockDim.x) {
        const int ii = e / nx, jj = e - ii * nx;
        const int ky = y0 + ii, kx = x0 + jj;
        const int kym = (H - ky) % H, kxm = (W - kx) % W;
        // position of the mirror mode inside the mirror rectangle (it runs backwards)
        const int mi = (blockIdx.y == 0) ? 0 : (ny - 1 - ii), mj = (blockIdx.x == 0) ? 0 : (nx - 1 - jj);
        double c1[P], c2[P];
#pragma unroll
        for (int q = 0; q < P; ++q) {
            c1[q] = s_c1[q * FB_T * FB_T + ii * FB_T + jj];
            c2[q] = s_c2[q * FB_T * FB_T + mi * FB_T + mj];
        }
        const bool ownMirror = (kx == 0 || 2 * kx == W);
        for (int sIdx = 0; sIdx < S; ++sIdx) {
            const int o = tile * S + sIdx;
            const int sg = sigPerTile ? o : sIdx;
            double rhs[F], rhs2[F], x1[F], x2[F];
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const C2 v = fsig[(((size_t)sg * F + f) * H + ky) * Wc + kx];
                rhs[f] = __dmul_rn(fw.w[f], cabs_fast((double)v.x, (double)v.y));
                rhs2[f] = rhs[f];
                if (ownMirror) {
                    const C2 vm = fsig[(((size_t)sg * F + f) * H + kym) * Wc + kx];
                    rhs2[f] = __dmul_rn(fw.w[f], cabs_fast((double)vm.x, (double)vm.y));
                }
            }
            solve_sym<F>(c1, rhs, x1);
            solve_sym<F>(c2, rhs2, x2);
#pragma unroll
            for (int f = 0; f < F; ++f) {
                filtH[(((size_t)o * F + f) * H + ky) * Wc + kx] = (T)(0.5 * __dadd_rn(x1[f], x2[f]));
                if (filtFull) {
                    double* ff = filtFull + ((size_t)o * F + f) * H * W;
                    ff[(size_t)ky * W + kx] = x1[f];
                    if (kx > 0 && 2 * kx < W) ff[(size_t)kym * W + kxm] = x2[f];
                }
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) fold_filter_kernel(const double* __restrict__ filtFull, T* __restrict__ filtH,
                                                          int H, int W) {
    const int Wc = W / 2 + 1;
    const int kx = blockIdx.x * blockDim.x + threadIdx.x;
    const int ky = blockIdx.y;
    if (kx >= Wc) return;
    const double* ff = filtFull + (size_t)blockIdx.z * H * W;
    const int kym = (H - ky) % H, kxm = (W - kx) % W;
    filtH[((size_t)blockIdx.z * H + ky) * Wc + kx] = (T)(0.5 * (ff[(size_t)ky * W + kx] + ff[(size_t)kym * W + kxm]));
}

// ------------------------------------------------------------------------------------------------------
// K5.  First maximum in raster order and optional spline value at (cy, cx).  Stage 1: TN_PARTS CTAs per image each reduce a
// contiguous slice to (value, first index); stage 2: one CTA per image reduces the partials and evaluates the spline.
#define TN_PARTS 64

__device__ __forceinline__ void argmax_combine(double& best, long long& bi, double ov, long long oi) {
    if (ov > best || (ov == best && oi < bi)) {
        best = ov;
        bi = oi;
    }
}

__device__ __forceinline__ void argmax_block_reduce(double& best, long long& bi, long long n, double* s_v, long long* s_i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        argmax_combine(best, bi, ov, oi);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        s_v[wid] = best;
        s_i[wid] = bi;
    }
    __syncthreads();
    if (wid == 0) {
        const int nw = blockDim.x >> 5;
        best = (lane < nw) ? s_v[lane] : -INFINITY;
        bi = (lane < nw) ? s_i[lane] : n;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            argmax_combine(best, bi, ov, oi);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) template_argmax_partial_kernel(const T* __restrict__ map, long long n, double* __restrict__ pVal,
                                                                      long long* __restrict__ pIdx) {
    const int b = blockIdx.y;
    const T* img = map + (size_t)b * n;
    const long long per = (n + TN_PARTS - 1) / TN_PARTS;
    const long long i0 = (long long)blockIdx.x * per, i1 = min(i0 + per, n);
    double best = -INFINITY;
    long long bi = n;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const double v = (double)img[i];
        if (v > best) {  // strided scan visits increasing i, so the first maximum per thread is kept
            best = v;
            bi = i;
        }
    }
    __shared__ double s_v[32];
    __shared__ long long s_i[32];
    argmax_block_reduce(best, bi, n, s_v, s_i);
    if (threadIdx.x == 0) {
        pVal[(size_t)b * TN_PARTS + blockIdx.x] = best;
        pIdx[(size_t)b * TN_PARTS + blockIdx.x] = bi;
    }
}

template <typename T>
__global__ void __launch_bounds__(64) template_norm_kernel(const T* __restrict__ map, int H, int W, int mode,
                                                           double cy, double cx, double numer, const double* __restrict__ pVal,
                                                           const long long* __restrict__ pIdx,
                                                           long long* __restrict__ argmaxIdx, double* __restrict__ maxVal,
                                                           double* __restrict__ centreVal, double* __restrict__ signalNorm) {
    const int b = blockIdx.x;
    const T* img = map + (size_t)b * H * W;
    const long long n = (long long)H * W;
    double best = -INFINITY;
    long long bi = n;
    for (int i = threadIdx.x; i < TN_PARTS; i += blockDim.x) argmax_combine(best, bi, pVal[(size_t)b * TN_PARTS + i], pIdx[(size_t)b * TN_PARTS + i]);
    __shared__ double s_v[32];
    __shared__ long long s_i[32];
    __shared__ double s_centre;
    argmax_block_reduce(best, bi, n, s_v, s_i);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (wid == 0 && lane == 0) {
        argmaxIdx[b] = bi;
        maxVal[b] = best;
        s_v[0] = best;
    }
    if (wid == 1 && mode == 1) {
        const double v = spline_eval_warp<T>(img, H, W, cy, cx, 1.0);
        if (lane == 0) s_centre = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double denom = s_v[0];
        if (mode == 1) {
            denom = s_centre;
            if (centreVal) centreVal[b] = s_centre;
        }
        signalNorm[b] = numer / denom;
    }
}

// value at one pixel (py, px) of Re(sum over the full plane of G(k) e^{+2 pi i k.p / N}) for a Hermitian G given
// on the half plane: columns 0 and W/2 count once, the others twice.  The phase separates, exp(i a ky) exp(i b kx): the
// column factors (with the 1 / 2 weights) are tabulated once per CTA in shared memory, the row factor once per row.
// PV_CHUNKS CTAs per (frequency, image) each take a band of rows; a second kernel adds the partial sums in a fixed order.
#define PV_CHUNKS 16

template <typename T>
__global__ void __launch_bounds__(256) plane_values_partial_kernel(const typename V2<T>::type* __restrict__ fmaps,
                                                                   const T* __restrict__ filtH, const long long* __restrict__ idx,
                                                                   double* __restrict__ partial, int F, int H, int W, int S) {
    typedef typename V2<T>::type C2;
    extern __shared__ double s_col[];   // [Wc][2]: weight * (cos, sin) of 2 pi kx px / W
    int chunk = blockIdx.x, f = blockIdx.y, b = blockIdx.z;
    if (S > 1 && gridDim.z % S == 0) {
        // images that share a template spectrum (the same scale of every tile) next to each other in launch order, so that the
        // template band comes from DRAM once and from L2 for the other tiles (the set of all templates, 310 MB for 16 scales
        // x 2 frequencies of 840 x 1440, does not survive in L2 from one tile to the next)
        const unsigned L = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        const unsigned nT = gridDim.z / S;
        const unsigned tileIdx = L % nT;
        unsigned rest = L / nT;
        chunk = rest % gridDim.x;
        rest /= gridDim.x;
        f = rest % gridDim.y;
        b = tileIdx * S + rest / gridDim.y;
    }
    const int Wc = W / 2 + 1;
    const long long p = idx[b];
    const int py = (int)(p / W), px = (int)(p % W);
    const size_t base = ((size_t)b * F + f) * H * Wc;                        // filter of image b
    const size_t mbase = ((size_t)(S > 0 ? b % S : b) * F + f) * H * Wc;    // its template spectrum
    for (int kx = threadIdx.x; kx < Wc; kx += blockDim.x) {
        // phase reduced exactly in integers before the division
        double sn, cs;
        sincospi(2.0 * (double)(((long long)kx * px) % W) / (double)W, &sn, &cs);
        const double wgt = (kx == 0 || 2 * kx == W) ? 1.0 : 2.0;
        s_col[2 * kx] = wgt * cs;
        s_col[2 * kx + 1] = wgt * sn;
    }
    // row factors of this CTA's band, once per CTA (every thread used to evaluate sincospi and a 64-bit modulo per row: three
    // quarters of the kernel's instructions)
    const int rows = (H + PV_CHUNKS - 1) / PV_CHUNKS;
    const int ky0 = chunk * rows, ky1 = min(ky0 + rows, H);
    double* s_row = s_col + 2 * Wc;      // [rows][2]: (cos, sin) of 2 pi ky py / H
    for (int r = threadIdx.x; r < ky1 - ky0; r += blockDim.x) {
        double rs, rc;
        sincospi(2.0 * (double)(((long long)(ky0 + r) * py) % H) / (double)H, &rs, &rc);
        s_row[2 * r] = rc;
        s_row[2 * r + 1] = rs;
    }
    __syncthreads();
    double acc = 0;
    for (int ky = ky0; ky < ky1; ++ky) {
        const double rc = s_row[2 * (ky - ky0)], rs = s_row[2 * (ky - ky0) + 1];
        const C2* fm = fmaps + mbase + (size_t)ky * Wc;
        const T* fl = filtH + base + (size_t)ky * Wc;
        double rowRe = 0, rowIm = 0;     // sum over kx of G * colFactor
        for (int kx = threadIdx.x; kx < Wc; kx += blockDim.x) {
            const C2 v = fm[kx];
            const double g = (double)fl[kx];
            const double gr = (double)v.x * g, gi = (double)v.y * g;
            const double cr = s_col[2 * kx], ci = s_col[2 * kx + 1];
            rowRe += gr * cr - gi * ci;
            rowIm += gr * ci + gi * cr;
        }
        acc += rowRe * rc - rowIm * rs;   // real part of (row sum) * exp(i phase_y)
    }
    __shared__ double s[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s[i];
        partial[((size_t)b * F + f) * PV_CHUNKS + chunk] = t;
    }
}

// total != nullptr: only the first F - 1 planes of every image were summed; the last one is what is left of the image's total
// (the planes add up to the filtered template, whose value at the pixel the caller already has: by linearity of the transform)
__global__ void plane_values_sum_kernel(const double* __restrict__ partial, double* __restrict__ out, int B, int F,
                                        const double* __restrict__ total) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double others = 0;
    for (int f = 0; f < F; ++f) {
        if (total && f == F - 1) {
            out[(size_t)b * F + f] = total[b] - others;
            break;
        }
        double t = 0;
        for (int c = 0; c < PV_CHUNKS; ++c) t += partial[((size_t)b * F + f) * PV_CHUNKS + c];
        out[(size_t)b * F + f] = t;
        others += t;
    }
}

// ------------------------------------------------------------------------------------------------------
template <typename T, int F>
static int xpower_smooth_tf(const void* fmaps, double* cov, double* tmp, int B, int H, int W, const GaussW& g, cudaStream_t st) {
    typedef typename V2<T>::type C2;
    constexpr int P = F * (F + 1) / 2;
    constexpr int R = 12;     // gauss_radius(3.0): the register-blocked kernels are compiled for it
    if (g.r == R && W >= 2 * R && (W % 2) == 0) {
        const size_t smem = (size_t)P * (gs_pad((W + GS_RB - 1) / GS_RB * GS_RB + 2 * R) + 1) * sizeof(double);
        NB_REQUIRE(smem <= 200 * 1024, NB_ERR_UNSUPPORTED, "nb_xpower_smooth: F=%d, W=%d needs too much shared memory", F, W);
        if (smem > 48 * 1024)
            NB_CHECK_CUDA(cudaFuncSetAttribute(xpower_smooth_x_rb_kernel<T, F, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int threads = ((W + GS_RB - 1) / GS_RB + 31) / 32 * 32;
        if (threads > 512) threads = 512;
        { NB_PROF_BEGIN(st); xpower_smooth_x_rb_kernel<T, F, R><<<dim3(H, B), threads, smem, st>>>((const C2*)fmaps, tmp, H, W, g); NB_PROF_END("xpower_smooth_x_kernel", st); }
        NB_CHECK_LAUNCH();
        const size_t smem2 = (size_t)(16 * GS_RB + 2 * R) * 32 * sizeof(double);
        { NB_PROF_BEGIN(st); smooth_y_rb_kernel<R><<<dim3((W + 31) / 32, (H + 16 * GS_RB - 1) / (16 * GS_RB), B * P), 512, smem2, st>>>(tmp, cov, H, W, g); NB_PROF_END("smooth_y_f64_kernel", st); }
        NB_CHECK_LAUNCH();
        return NB_OK;
    }
    const size_t smem = (size_t)P * W * sizeof(double);
    NB_REQUIRE(smem <= 200 * 1024, NB_ERR_UNSUPPORTED, "nb_xpower_smooth: F=%d, W=%d needs too much shared memory", F, W);
    if (smem > 48 * 1024)
        NB_CHECK_CUDA(cudaFuncSetAttribute(xpower_smooth_x_kernel<T, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { NB_PROF_BEGIN(st); xpower_smooth_x_kernel<T, F><<<dim3(H, B), 256, smem, st>>>((const C2*)fmaps, tmp, H, W, g); NB_PROF_END("xpower_smooth_x_kernel", st); }
    NB_CHECK_LAUNCH();
    const size_t smem2 = (size_t)(GY_TY + 2 * g.r) * GY_TX * sizeof(double);
    if (smem2 > 48 * 1024)
        NB_CHECK_CUDA(cudaFuncSetAttribute(smooth_y_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    { NB_PROF_BEGIN(st); smooth_y_f64_kernel<<<dim3((W + GY_TX - 1) / GY_TX, (H + GY_TY - 1) / GY_TY, B * P), 256, smem2, st>>>(tmp, cov, H, W, g); NB_PROF_END("smooth_y_f64_kernel", st); }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

template <typename T, int F>
static int build_filter_tf(const double* cov, const void* fsig, const double* w, double* filtFull, void* filtH, int nOut,
                           int S, int sigPerTile, int H, int W, cudaStream_t st) {
    typedef typename V2<T>::type C2;
    FreqW fw;
    for (int f = 0; f < NB_MAX_FREQ; ++f) fw.w[f] = f < F ? w[f] : 0.0;
    const int Wc = W / 2 + 1;
    if (!sigPerTile && !getenv("NEMO_B200_BUILD_FILTER_ONE_SCALE") && !getenv("NEMO_B200_BUILD_FILTER_PER_TILE")) {
        // templates shared by the tiles: right-hand sides in registers, tiles walked by the thread, self-mirror columns redone
        constexpr int SC = 16 / (F <= 2 ? 1 : 2);
        const int nT = nOut / S;
        { NB_PROF_BEGIN(st); build_filter_batch_kernel<T, F, SC><<<dim3((Wc + 127) / 128, H, (S + SC - 1) / SC), 128, 0, st>>>(cov, (const C2*)fsig, fw, filtFull, (T*)filtH, S,
                                                                             nT, H, W); NB_PROF_END("build_filter_kernel", st); }
        NB_CHECK_LAUNCH();
        { NB_PROF_BEGIN(st); build_filter_selfmirror_kernel<T, F><<<dim3((H + 127) / 128, (W % 2 == 0) ? 2 : 1, nOut), 128, 0, st>>>(cov, (const C2*)fsig, fw, filtFull, (T*)filtH, S,
                                                                             0, H, W); NB_PROF_END("build_filter_selfmirror_kernel", st); }
        NB_CHECK_LAUNCH();
        return NB_OK;
    }
    if (S > 1 && !getenv("NEMO_B200_BUILD_FILTER_ONE_SCALE")) {
        // all scales of a tile per thread (the profiler keeps the kernel's family name)
        { NB_PROF_BEGIN(st); build_filter_multi_kernel<T, F><<<dim3((Wc + 127) / 128, H, nOut / S), 128, 0, st>>>(cov, (const C2*)fsig, fw, filtFull, (T*)filtH, S,
                                                                             sigPerTile, H, W); NB_PROF_END("build_filter_kernel", st); }
        NB_CHECK_LAUNCH();
        return NB_OK;
    }
    { NB_PROF_BEGIN(st); build_filter_kernel<T, F><<<dim3((Wc + 255) / 256, H, nOut), 256, 0, st>>>(cov, (const C2*)fsig, fw, filtFull, (T*)filtH, S,
                                                                             sigPerTile, H, W); NB_PROF_END("build_filter_kernel", st); }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

template <typename T, int F>
static int build_filter_fused_tf(const void* fmaps, const void* fsig, const double* w, double* filtFull, void* filtH, int nOut, int S,
                                 int sigPerTile, int H, int W, const GaussW& g, cudaStream_t st) {
    typedef typename V2<T>::type C2;
    constexpr int P = F * (F + 1) / 2;
    FreqW fw;
    for (int f = 0; f < NB_MAX_FREQ; ++f) fw.w[f] = f < F ? w[f] : 0.0;
    const int Wc = W / 2 + 1;
    const int side = FB_T + 2 * g.r;
    const size_t smem = sizeof(double) * ((size_t)side * side + (size_t)side * FB_T + 2 * (size_t)P * FB_T * FB_T);
    NB_REQUIRE(smem <= 220 * 1024, NB_ERR_UNSUPPORTED, "nb_build_filter_fused: smoothing radius %d too large", g.r);
    if (smem > 48 * 1024)
        NB_CHECK_CUDA(cudaFuncSetAttribute(build_filter_fused_kernel<T, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(1 + (Wc - 1 + FB_T - 1) / FB_T, 1 + (H - 1 + FB_T - 1) / FB_T, nOut / S);
    { NB_PROF_BEGIN(st); build_filter_fused_kernel<T, F><<<grid, 256, smem, st>>>((const C2*)fmaps, (const C2*)fsig, fw, filtFull, (T*)filtH, S, sigPerTile, H, W, g); NB_PROF_END("build_filter_fused_kernel", st); }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

extern "C" {

int nb_build_filter_fused(int prec, const void* fmaps, const void* fsig, const double* w, double* filtFull, void* filtH, int nOut,
                          int S, int sigPerTile, int F, int H, int W, double sigma, cudaStream_t st) {
    NB_REQUIRE(fmaps && fsig && w && filtH && nOut > 0 && S > 0 && nOut % S == 0 && (W % 2) == 0, NB_ERR_ARG,
               "nb_build_filter_fused: bad arguments");
    NB_REQUIRE(F >= 1 && F <= NB_MAX_FREQ, NB_ERR_UNSUPPORTED, "nb_build_filter_fused: %d frequencies not supported", F);
    GaussW g;
    int rc = make_gauss(sigma, &g);
    if (rc != NB_OK) return rc;
    NB_REQUIRE(g.r < H && g.r < W, NB_ERR_UNSUPPORTED, "nb_build_filter_fused: map smaller than the smoothing kernel");
#define BFF(T)                                                                                                              \
    switch (F) {                                                                                                            \
        case 1: return build_filter_fused_tf<T, 1>(fmaps, fsig, w, filtFull, filtH, nOut, S, sigPerTile, H, W, g, st);      \
        case 2: return build_filter_fused_tf<T, 2>(fmaps, fsig, w, filtFull, filtH, nOut, S, sigPerTile, H, W, g, st);      \
        case 3: return build_filter_fused_tf<T, 3>(fmaps, fsig, w, filtFull, filtH, nOut, S, sigPerTile, H, W, g, st);      \
        default: return build_filter_fused_tf<T, 4>(fmaps, fsig, w, filtFull, filtH, nOut, S, sigPerTile, H, W, g, st);     \
    }
    if (prec == NB_PREC_F64) { BFF(double) }
    if (prec == NB_PREC_F32) { BFF(float) }
#undef BFF
    nb_set_error("nb_build_filter_fused: bad precision %d", prec);
    return NB_ERR_ARG;
}

int nb_xpower_smooth(int prec, const void* fmaps, double* cov, double* tmp, int B, int F, int H, int W, double sigma,
                     cudaStream_t st) {
    NB_REQUIRE(fmaps && cov && tmp && B > 0 && H > 0 && W > 0, NB_ERR_ARG, "nb_xpower_smooth: bad arguments");
    NB_REQUIRE(F >= 1 && F <= NB_MAX_FREQ, NB_ERR_UNSUPPORTED, "nb_xpower_smooth: %d frequencies not supported", F);
    GaussW g;
    int rc = make_gauss(sigma, &g);
    if (rc != NB_OK) return rc;
    NB_REQUIRE(g.r < H && g.r < W, NB_ERR_UNSUPPORTED, "nb_xpower_smooth: map smaller than the smoothing kernel");
#define XP(T)                                                                         \
    switch (F) {                                                                      \
        case 1: return xpower_smooth_tf<T, 1>(fmaps, cov, tmp, B, H, W, g, st);       \
        case 2: return xpower_smooth_tf<T, 2>(fmaps, cov, tmp, B, H, W, g, st);       \
        case 3: return xpower_smooth_tf<T, 3>(fmaps, cov, tmp, B, H, W, g, st);       \
        default: return xpower_smooth_tf<T, 4>(fmaps, cov, tmp, B, H, W, g, st);      \
    }
    if (prec == NB_PREC_F64) { XP(double) }
    if (prec == NB_PREC_F32) { XP(float) }
#undef XP
    nb_set_error("nb_xpower_smooth: bad precision %d", prec);
    return NB_ERR_ARG;
}

int nb_build_filter(int prec, const double* cov, const void* fsig, const double* w, double* filtFull, void* filtH,
                    int nOut, int S, int sigPerTile, int F, int H, int W, cudaStream_t st) {
    NB_REQUIRE(cov && fsig && w && filtH && nOut > 0 && S > 0 && nOut % S == 0, NB_ERR_ARG, "nb_build_filter: bad arguments");
    NB_REQUIRE(F >= 1 && F <= NB_MAX_FREQ, NB_ERR_UNSUPPORTED, "nb_build_filter: %d frequencies not supported", F);
#define BF(T)                                                                                             \
    switch (F) {                                                                                          \
        case 1: return build_filter_tf<T, 1>(cov, fsig, w, filtFull, filtH, nOut, S, sigPerTile, H, W, st); \
        case 2: return build_filter_tf<T, 2>(cov, fsig, w, filtFull, filtH, nOut, S, sigPerTile, H, W, st); \
        case 3: return build_filter_tf<T, 3>(cov, fsig, w, filtFull, filtH, nOut, S, sigPerTile, H, W, st); \
        default: return build_filter_tf<T, 4>(cov, fsig, w, filtFull, filtH, nOut, S, sigPerTile, H, W, st); \
    }
    if (prec == NB_PREC_F64) { BF(double) }
    if (prec == NB_PREC_F32) { BF(float) }
#undef BF
    nb_set_error("nb_build_filter: bad precision %d", prec);
    return NB_ERR_ARG;
}

int nb_fold_filter(int prec, const double* filtFull, void* filtH, int n, int H, int W, cudaStream_t st) {
    NB_REQUIRE(filtFull && filtH && n > 0, NB_ERR_ARG, "nb_fold_filter: bad arguments");
    const int Wc = W / 2 + 1;
    dim3 grid((Wc + 255) / 256, H, n);
    if (prec == NB_PREC_F64)
        { NB_PROF_BEGIN(st); fold_filter_kernel<double><<<grid, 256, 0, st>>>(filtFull, (double*)filtH, H, W); NB_PROF_END("fold_filter_kernel", st); }
    else if (prec == NB_PREC_F32)
        { NB_PROF_BEGIN(st); fold_filter_kernel<float><<<grid, 256, 0, st>>>(filtFull, (float*)filtH, H, W); NB_PROF_END("fold_filter_kernel", st); }
    else {
        nb_set_error("nb_fold_filter: bad precision %d", prec);
        return NB_ERR_ARG;
    }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

int nb_template_norm(int prec, const void* map, int B, int H, int W, int mode, double cy, double cx, double numer,
                     long long* argmaxIdx, double* maxVal, double* centreVal, double* signalNorm, void* scratch,
                     cudaStream_t st) {
    NB_REQUIRE(map && argmaxIdx && maxVal && signalNorm && scratch && B > 0, NB_ERR_ARG, "nb_template_norm: bad arguments");
    double* pVal = (double*)scratch;                                       // [B, TN_PARTS]
    long long* pIdx = (long long*)(pVal + (size_t)B * TN_PARTS);           // [B, TN_PARTS]
    const long long n = (long long)H * W;
    if (prec == NB_PREC_F64) {
        { NB_PROF_BEGIN(st); template_argmax_partial_kernel<double><<<dim3(TN_PARTS, B), 256, 0, st>>>((const double*)map, n, pVal, pIdx); NB_PROF_END("template_argmax_partial_kernel", st); }
        NB_CHECK_LAUNCH();
        { NB_PROF_BEGIN(st); template_norm_kernel<double><<<B, 64, 0, st>>>((const double*)map, H, W, mode, cy, cx, numer, pVal, pIdx, argmaxIdx, maxVal, centreVal, signalNorm); NB_PROF_END("template_norm_kernel", st); }
    } else if (prec == NB_PREC_F32) {
        { NB_PROF_BEGIN(st); template_argmax_partial_kernel<float><<<dim3(TN_PARTS, B), 256, 0, st>>>((const float*)map, n, pVal, pIdx); NB_PROF_END("template_argmax_partial_kernel", st); }
        NB_CHECK_LAUNCH();
        { NB_PROF_BEGIN(st); template_norm_kernel<float><<<B, 64, 0, st>>>((const float*)map, H, W, mode, cy, cx, numer, pVal, pIdx, argmaxIdx, maxVal, centreVal, signalNorm); NB_PROF_END("template_norm_kernel", st); }
    } else {
        nb_set_error("nb_template_norm: bad precision %d", prec);
        return NB_ERR_ARG;
    }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

int nb_plane_values_at(int prec, const void* fmaps, const void* filtH, const long long* idx, double* out, int B, int S,
                       int F, int H, int W, const double* total, double* scratch, cudaStream_t st) {
    NB_REQUIRE(fmaps && filtH && idx && out && scratch && B > 0 && F > 0, NB_ERR_ARG, "nb_plane_values_at: bad arguments");
    const size_t smem = sizeof(double) * 2 * ((size_t)(W / 2 + 1) + (size_t)((H + PV_CHUNKS - 1) / PV_CHUNKS));
    NB_REQUIRE(smem <= 48 * 1024, NB_ERR_UNSUPPORTED, "nb_plane_values_at: tile too wide");
    const int nSum = total ? F - 1 : F;      // planes summed over the Fourier plane; with `total` the last one is the remainder
    dim3 grid(PV_CHUNKS, nSum > 0 ? nSum : 1, B);
    if (nSum == 0) {
    } else if (prec == NB_PREC_F64) {
        { NB_PROF_BEGIN(st); plane_values_partial_kernel<double><<<grid, 256, smem, st>>>((const double2*)fmaps, (const double*)filtH, idx, scratch, F, H, W, S); NB_PROF_END("plane_values_partial_kernel", st); }
    } else if (prec == NB_PREC_F32) {
        { NB_PROF_BEGIN(st); plane_values_partial_kernel<float><<<grid, 256, smem, st>>>((const float2*)fmaps, (const float*)filtH, idx, scratch, F, H, W, S); NB_PROF_END("plane_values_partial_kernel", st); }
    } else {
        nb_set_error("nb_plane_values_at: bad precision %d", prec);
        return NB_ERR_ARG;
    }
    NB_CHECK_LAUNCH();
    { NB_PROF_BEGIN(st); plane_values_sum_kernel<<<(B + 255) / 256, 256, 0, st>>>(scratch, out, B, F, total); NB_PROF_END("plane_values_sum_kernel", st); }
    NB_CHECK_LAUNCH();
    return NB_OK;
}

}  // extern "C"
