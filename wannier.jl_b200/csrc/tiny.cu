// tiny.cu — one trial point of the optimiser's line search in ONE cooperative launch, for the problem sizes the
// reference's own benchmark suite runs (benchmark/bench_maxloc.jl: Si2_valence, 4 WFs, 64 k-points;
// bench_disentangle.jl: Si2, 12 bands -> 8 WFs) and their neighbourhood: at most 24 k complex multiply-adds per
// k-point (4 x 4 ... 12 x 12, 18 -> 9), one CTA per k-point while the grid is co-resident (592 k-points on a B200) and
// up to four k-points per CTA, in slots of shared memory, beyond (8 WFs on a 12 x 12 x 10 grid) — where it measures
// 1.1-1.8 x faster than the chain of general kernels (tools/debug_tiny_crossover.py, profiles/r4y_tiny_crossover.txt).
//
// At those sizes one evaluation is ~0.3-2 MFLOP: the arithmetic of a whole trial point
//     xn = retract(x + a s);  f, G = spread and gradient at xn;  gn = project(G, xn);  dphi = <gn, s>
// (max_localize.jl:13-23 / disentangle.jl:586-614 behind Optim's retract! / project_tangent!, call sites
// max_localize.jl:62-63, disentangle.jl:687-690) is a few microseconds, but as the chain of general kernels it is
// 13 dependent launches of 3-28 us each (profiles/r4_summary.md: 72 us per trial point even as a CUDA graph).
// Here a CTA owns its k-points for the whole trial point and keeps everything of them in shared memory — the factors
// (X_k, Y_k) or U_k, the nn rotated overlaps MU_kb, diag N_kb, the gradient — and the grid meets three times
// (cooperative groups grid sync): after the retraction (neighbours read U_{k+b}), after the per-k partial sums (every
// CTA then forms the centres r itself, in the same fixed order, so all agree bitwise), and before CTA 0 adds up the
// inner product and publishes f, dphi, min |N_nn| and the retraction flag to page-locked host memory.
//
// The arithmetic restates, per k-point, what the general kernels do (same formulas, same reference lines):
//   retraction        retract_small_kernel (stiefel_small.cu): Newton-Schulz polar factor
//   U = Y X           xy_to_u_small_kernel (disentangle.jl:351-356)
//   MU, diag N        compute_MU_UtMU! (spread.jl:164-195); only diag(N) is needed (spread.jl:307-312, 459-470)
//   partial sums      pair_reduce_kernel (omega! / center!, spread.jl:284-322, 565-594)
//   spread scalars    finalize_kernel (spread.jl:324-356, 246-252)
//   gradient          grad_kernel (omega_grad!, spread.jl:425-478)
//   G_X, G_Y, masks   gu_to_gxy_small_kernel (disentangle.jl:481-501)
//   projection        project_small_kernel (Optim project_tangent!: G - X (X^H G + G^H X)/2)
// tests/test_gpu_parity.py::test_tiny_trial_kernel_matches_general_path compares the two paths iterate by iterate;
// the reference's trajectory goldens (max_iter = 4) and its converged optimum run through this kernel by default.
#include "common.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define WB_TINY_MAXB 32   /* bands */
#define WB_TINY_MAXW 16   /* Wannier functions: the polar iteration keeps ncol^2 / 128 elements per thread in registers */
#define WB_TINY_THREADS 128

struct TinyArgs {
    int mode, nb, nw, nk, nn;
    int kper;                      // k-points per CTA (CTA c owns k = c, c + gridDim.x, ...)
    int two_level;                 // the global sums are formed once (CTA j adds column j) instead of in every CTA
    double* gsums;                 // [nsums] (ctx->d_sums)
    int mu_global;                 // the rotated overlaps MU_kb wait for phase 3 in global memory (L2), not in the slots
    cplx* MUg;                     // [pairs][nb*nw] (ctx->d_MU)
    const int32_t* kpb;
    const double* bvec;
    const double* wb;
    const cplx* M;                 // plain overlaps [pairs][nb*nb] column-major
    const uint8_t* frozen;         // [nk][nb] or nullptr
    const int32_t* nfrozen;        // [nk]
    int has_penalty;
    double lambda;
    const double* r0;
    double wsum;
    int split_valid;
    const cplx* x;                 // base point
    const cplx* s;                 // direction (nullptr: evaluate at retract(x))
    double alpha;
    cplx* xn;                      // trial point (may alias x when s == nullptr)
    cplx* gn;                      // projected gradient at xn
    cplx* U;                       // [nk][nb*nw] gauges of the trial point (mode 0: == xn)
    double* part;                  // [nsums][nk]
    double* pmin;                  // [nk]
    double* dpart;                 // [nk] per-k inner products <gn_k, s_k>
    double* res;                   // [nres]
    double* dmin;
    int* flag;                     // device word: epoch of the last failed retraction
    int epoch;
    double* hpin;                  // page-locked host scratch (h_pinned)
};

namespace {

__device__ __forceinline__ double tiny_block_sum(double v, double* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < WB_TINY_THREADS / 32; w++) t += sh[w];
    return t;
}

// (The helpers are __noinline__: the kernel runs once per launch on cold instruction caches, and with three inlined
// copies of each iteration it measured 8 % slower end to end.)
// A (nrow x ncol, column-major, ld = nrow, shared memory) <- its polar factor; P: ncol x ncol scratch, T: nrow x ncol
// scratch.  The iteration of retract_small_kernel: stop when ||A^H A - I||_F < tol was measured BEFORE the last update
// (which takes err to ~err^2); first pass scaled by ||A^H A||_F^(-1/2) when sigma_max^2 may exceed 3.
__device__ __noinline__ bool tiny_polar(cplx* A, cplx* P, cplx* T, int nrow, int ncol, double* red) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    bool converged = false;
    cplx* cur = A;
    cplx* nxt = T;
    for (int it = 0; it < 100 && !converged; it++) {
        // P = cur^H cur (each thread keeps its elements in registers until they are scaled), ||P - I||_F, ||P||_F
        cplx pv[(WB_TINY_MAXW * WB_TINY_MAXW + WB_TINY_THREADS - 1) / WB_TINY_THREADS];
        double e2 = 0.0, f2 = 0.0;
#pragma unroll
        for (int q = 0; q < (WB_TINY_MAXW * WB_TINY_MAXW + WB_TINY_THREADS - 1) / WB_TINY_THREADS; q++) {
            const int e = tid + q * WB_TINY_THREADS;
            if (e < ncol * ncol) {
                const int i = e % ncol, j = e / ncol;
                cplx acc = make_double2(0.0, 0.0);
                for (int m = 0; m < nrow; m++) cfma_conj(acc, cur[i * nrow + m], cur[j * nrow + m]);
                pv[q] = acc;
                const double dx = acc.x - (i == j ? 1.0 : 0.0);
                e2 += dx * dx + acc.y * acc.y;
                f2 += acc.x * acc.x + acc.y * acc.y;
            }
        }
        e2 = warp_sum(e2);
        f2 = warp_sum(f2);
        __syncthreads();   // (red may still be read from the previous use)
        if (lane == 0) { red[warp] = e2; red[4 + warp] = f2; }
        __syncthreads();
        const double err = sqrt((red[0] + red[1]) + (red[2] + red[3]));
        const double fro = sqrt((red[4] + red[5]) + (red[6] + red[7]));
        if (!(err == err)) break;   // NaN input
        converged = err < 2e-8;
        const double sc = (it == 0 && err >= 1.9) ? rsqrt(fro) : 1.0;
        const double s3 = 0.5 * sc * sc * sc;
#pragma unroll
        for (int q = 0; q < (WB_TINY_MAXW * WB_TINY_MAXW + WB_TINY_THREADS - 1) / WB_TINY_THREADS; q++) {
            const int e = tid + q * WB_TINY_THREADS;
            if (e < ncol * ncol) {
                const int i = e % ncol, j = e / ncol;
                P[e] = make_double2((i == j ? 1.5 * sc : 0.0) - s3 * pv[q].x, -s3 * pv[q].y);
            }
        }
        __syncthreads();
        for (int e = tid; e < nrow * ncol; e += WB_TINY_THREADS) {
            const int m = e % nrow, n = e / nrow;
            cplx acc = make_double2(0.0, 0.0);
            for (int l = 0; l < ncol; l++) cfma(acc, cur[l * nrow + m], P[n * ncol + l]);
            nxt[e] = acc;
        }
        __syncthreads();
        cplx* t = cur; cur = nxt; nxt = t;
    }
    if (cur != A) {
        for (int e = tid; e < nrow * ncol; e += WB_TINY_THREADS) A[e] = cur[e];
        __syncthreads();
    }
    return converged;
}

// The same iteration run by ONE warp (barriers become __syncwarp, reductions shuffles) for matrices of at most 128
// elements with ncol <= 8: at those sizes a block-wide step is all barrier latency, and the two factors of the (X, Y)
// parametrisation can iterate side by side on two warps.  P and T are private to the calling warp.
__device__ __forceinline__ bool tiny_polar_fits_warp(int nrow, int ncol) { return nrow * ncol <= 128 && ncol <= 8; }
__device__ __noinline__ bool tiny_polar_warp(cplx* A, cplx* P, cplx* T, int nrow, int ncol) {
    const int lane = threadIdx.x & 31;
    bool converged = false;
    cplx* cur = A;
    cplx* nxt = T;
    for (int it = 0; it < 100 && !converged; it++) {
        cplx pv[2];
        double e2 = 0.0, f2 = 0.0;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int e = lane + 32 * q;
            if (e < ncol * ncol) {
                const int i = e % ncol, j = e / ncol;
                cplx acc = make_double2(0.0, 0.0);
                for (int m = 0; m < nrow; m++) cfma_conj(acc, cur[i * nrow + m], cur[j * nrow + m]);
                pv[q] = acc;
                const double dx = acc.x - (i == j ? 1.0 : 0.0);
                e2 += dx * dx + acc.y * acc.y;
                f2 += acc.x * acc.x + acc.y * acc.y;
            }
        }
        const double err = sqrt(warp_sum(e2));
        const double fro = sqrt(warp_sum(f2));
        if (!(err == err)) break;   // NaN input
        converged = err < 2e-8;
        const double sc = (it == 0 && err >= 1.9) ? rsqrt(fro) : 1.0;
        const double s3 = 0.5 * sc * sc * sc;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int e = lane + 32 * q;
            if (e < ncol * ncol) {
                const int i = e % ncol, j = e / ncol;
                P[e] = make_double2((i == j ? 1.5 * sc : 0.0) - s3 * pv[q].x, -s3 * pv[q].y);
            }
        }
        __syncwarp();
        for (int e = lane; e < nrow * ncol; e += 32) {
            const int m = e % nrow, n = e / nrow;
            cplx acc = make_double2(0.0, 0.0);
            for (int l = 0; l < ncol; l++) cfma(acc, cur[l * nrow + m], P[n * ncol + l]);
            nxt[e] = acc;
        }
        __syncwarp();
        cplx* t = cur; cur = nxt; nxt = t;
    }
    if (cur != A) {
        for (int e = lane; e < nrow * ncol; e += 32) A[e] = cur[e];
        __syncwarp();
    }
    return converged;
}

// G (nrow x ncol, shared) <- G - X (X^H G + G^H X)/2; S: ncol x ncol scratch, T: nrow x ncol scratch
__device__ __noinline__ void tiny_project(const cplx* X, cplx* G, cplx* S, cplx* T, int nrow, int ncol) {
    const int tid = threadIdx.x;
    for (int e = tid; e < ncol * ncol; e += WB_TINY_THREADS) {
        const int i = e % ncol, j = e / ncol;
        cplx acc = make_double2(0.0, 0.0);
        for (int m = 0; m < nrow; m++) cfma_conj(acc, X[i * nrow + m], G[j * nrow + m]);
        S[e] = acc;
    }
    __syncthreads();
    for (int e = tid; e < ncol * ncol; e += WB_TINY_THREADS) {
        const int i = e % ncol, j = e / ncol;
        if (i > j) continue;
        const cplx a = S[i + j * ncol], b = S[j + i * ncol];
        const cplx h = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
        S[i + j * ncol] = h;
        S[j + i * ncol] = cconj(h);
    }
    __syncthreads();
    for (int e = tid; e < nrow * ncol; e += WB_TINY_THREADS) {
        const int m = e % nrow, n = e / nrow;
        cplx acc = make_double2(0.0, 0.0);
        for (int l = 0; l < ncol; l++) cfma(acc, X[l * nrow + m], S[n * ncol + l]);
        T[e] = make_double2(G[e].x - acc.x, G[e].y - acc.y);
    }
    __syncthreads();
    for (int e = tid; e < nrow * ncol; e += WB_TINY_THREADS) G[e] = T[e];
    __syncthreads();
}

__global__ void __launch_bounds__(WB_TINY_THREADS) tiny_trial_kernel(TinyArgs a) {
    cg::grid_group grid = cg::this_grid();
    const int nb = a.nb, nw = a.nw, nn = a.nn, nk = a.nk, kper = a.kper;
    const int tid = threadIdx.x;
    const int mat = nb * nw, nxx = nw * nw;
    const int vlen = a.mode == 1 ? nxx + mat : mat;      // complex elements of the optimisation variable per k-point
    const int nsums = 4 * nw + 2;

    // shared memory: `kper` slots of per-k state that lives through the whole trial point (the CTA owns the k-points
    // blockIdx.x, blockIdx.x + gridDim.x, ...), then scratch that every slot re-uses
    extern __shared__ __align__(16) unsigned char tiny_smem[];
    // slot: U_k, X_k, Y_k, diag N_kb [nn][nw] (+ MU_kb [nn][mat] unless they wait in global memory)
    const int state_c = 2 * mat + nxx + nn * nw + (a.mu_global ? 0 : nn * mat);
    cplx* st0 = reinterpret_cast<cplx*>(tiny_smem);
    cplx* sP = st0 + (size_t)kper * state_c;         // [nxx]   scratch
    cplx* sT = sP + nxx;                             // [mat]   scratch
    cplx* sG = sT + mat;                             // [mat]   G_k
    cplx* sGX = sG + mat;                            // [nxx]
    cplx* sMk = sGX + nxx;                           // [nn][nb*nb] M_kb staging
    cplx* sUp = sMk + nn * nb * nb;                  // [nn][mat]   U_{k+b} staging (later: G_Y)
    cplx* sMUscr = sUp + nn * mat;                   // [nn][mat]   MU_kb of the current k-point when they live in global memory
    double* dst0 = reinterpret_cast<double*>(sMUscr + (a.mu_global ? nn * mat : 0));
    double* ssum = dst0 + (size_t)kper * 2 * nn * nw;   // [nsums]
    double* sfro = ssum + nsums;                     // [nn]
    double* red = sfro + nn;                         // [8]
    __shared__ int s_ok, s_fail;
    if (tid == 0) s_fail = 0;
    __syncthreads();
#define SLOT_POINTERS(slot)                                                                   \
    cplx* sU = st0 + (size_t)(slot) * state_c;            /* [mat]  U_k               */      \
    cplx* sX = sU + mat;                                  /* [nxx]  X_k (mode 1)      */      \
    cplx* sY = sX + nxx;                                  /* [mat]  Y_k (mode 1)      */      \
    cplx* sZ = sY + mat;                                  /* [nn][nw] diag N_kb, later the gradient coefficients */ \
    cplx* sMU = a.mu_global ? sMUscr : sZ + nn * nw;      /* [nn][mat]                */      \
    double* sphi = dst0 + (size_t)(slot) * 2 * nn * nw;   /* [nn][nw]                 */      \
    double* sa2 = sphi + nn * nw;                         /* [nn][nw]                 */
#ifdef WB_TINY_TIMING
    unsigned long long tt[8];
#define TT(i) do { if (blockIdx.x == 0 && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt[i])); } while (0)
#else
#define TT(i)
#endif
    TT(0);

    // ---- phase 1: trial point and its retraction -------------------------------------------------------------
    const double alpha = a.s ? a.alpha : 0.0;
    bool ok = true;
    for (int slot = 0, k = blockIdx.x; slot < kper && k < nk; slot++, k += gridDim.x) {
        cplx* sU = st0 + (size_t)slot * state_c;
        cplx* sX = sU + mat;
        cplx* sY = sX + nxx;
        const cplx* xk = a.x + (long long)k * vlen;
        const cplx* sk = a.s ? a.s + (long long)k * vlen : nullptr;
        cplx* xnk = a.xn + (long long)k * vlen;
        if (a.mode == 1) {
            for (int e = tid; e < vlen; e += WB_TINY_THREADS) {
                cplx v = xk[e];
                if (sk) { const cplx d = sk[e]; v.x = fma(alpha, d.x, v.x); v.y = fma(alpha, d.y, v.y); }
                if (e < nxx) sX[e] = v; else sY[e - nxx] = v;
            }
            __syncthreads();
            if (tiny_polar_fits_warp(nb, nw)) {   // both factors side by side: warp 0 iterates on X_k, warp 1 on Y_k
                if (tid < 32) { if (!tiny_polar_warp(sX, sP, sGX, nw, nw)) s_fail = 1; }
                else if (tid < 64) { if (!tiny_polar_warp(sY, sG, sT, nb, nw)) s_fail = 1; }
                __syncthreads();
            } else {
                ok = tiny_polar(sX, sP, sT, nw, nw, red) && ok;
                ok = tiny_polar(sY, sP, sT, nb, nw, red) && ok;
            }
            for (int e = tid; e < vlen; e += WB_TINY_THREADS) xnk[e] = e < nxx ? sX[e] : sY[e - nxx];
            for (int e = tid; e < mat; e += WB_TINY_THREADS) {   // U = Y X
                const int m = e % nb, n = e / nb;
                cplx acc = make_double2(0.0, 0.0);
                for (int l = 0; l < nw; l++) cfma(acc, sY[l * nb + m], sX[n * nw + l]);
                sU[e] = acc;
                a.U[(long long)k * mat + e] = acc;
            }
        } else {
            for (int e = tid; e < mat; e += WB_TINY_THREADS) {
                cplx v = xk[e];
                if (sk) { const cplx d = sk[e]; v.x = fma(alpha, d.x, v.x); v.y = fma(alpha, d.y, v.y); }
                sU[e] = v;
            }
            __syncthreads();
            if (tiny_polar_fits_warp(nb, nw)) {
                if (tid < 32) { if (!tiny_polar_warp(sU, sP, sT, nb, nw)) s_fail = 1; }
                __syncthreads();
            } else {
                ok = tiny_polar(sU, sP, sT, nb, nw, red) && ok;
            }
            for (int e = tid; e < mat; e += WB_TINY_THREADS) xnk[e] = sU[e];   // mode 0: a.U == a.xn
        }
        __syncthreads();
    }
    if ((!ok || s_fail) && tid == 0) atomicMax(a.flag, a.epoch);
    TT(1);
    __threadfence();
    grid.sync();
    TT(2);

    // ---- phase 2: MU_kb = M_kb U_{k+b}, diag N_kb, per-k partial sums -----------------------------------------
    for (int slot = 0, k = blockIdx.x; slot < kper && k < nk; slot++, k += gridDim.x) {
        cplx* sU = st0 + (size_t)slot * state_c;
        cplx* sZ = sU + 2 * mat + nxx;
        cplx* sMU = a.mu_global ? sMUscr : sZ + nn * nw;
        double* sphi = dst0 + (size_t)slot * 2 * nn * nw;
        double* sa2 = sphi + nn * nw;
        // all nn neighbours at once: one exposure of the load latency, one atan2 per thread
        for (int e = tid; e < nn * nb * nb; e += WB_TINY_THREADS) sMk[e] = a.M[(long long)k * nn * nb * nb + e];
        for (int e = tid; e < nn * mat; e += WB_TINY_THREADS) {
            const int b = e / mat;
            sUp[e] = a.U[(long long)a.kpb[(long long)k * nn + b] * mat + (e - b * mat)];
        }
        __syncthreads();
        for (int e = tid; e < nn * mat; e += WB_TINY_THREADS) {
            const int b = e / mat, r = e - b * mat, m = r % nb, n = r / nb;
            const cplx* Mb = sMk + b * nb * nb;
            const cplx* Ub = sUp + b * mat;
            cplx acc = make_double2(0.0, 0.0);
            for (int l = 0; l < nb; l++) cfma(acc, Mb[l * nb + m], Ub[n * nb + l]);
            sMU[e] = acc;
            if (a.mu_global) a.MUg[(long long)k * nn * mat + e] = acc;
        }
        __syncthreads();
        for (int b = tid >> 5; b < nn; b += WB_TINY_THREADS / 32) {   // ||MU_kb||_F^2: a warp per pair, fixed order
            double f2 = 0.0;
            for (int e = tid & 31; e < mat; e += 32) { const cplx v = sMU[b * mat + e]; f2 = fma(v.x, v.x, fma(v.y, v.y, f2)); }
            f2 = warp_sum(f2);
            if ((tid & 31) == 0) sfro[b] = f2;
        }
        for (int e = tid; e < nn * nw; e += WB_TINY_THREADS) {
            const int b = e / nw, n = e - b * nw;
            const cplx* mu = sMU + b * mat + n * nb;
            cplx d = make_double2(0.0, 0.0);
            for (int m = 0; m < nb; m++) cfma_conj(d, sU[n * nb + m], mu[m]);
            sZ[e] = d;
            sphi[e] = atan2(d.y, d.x);             // imaglog, src/utils/linalg.jl:15
            sa2[e] = d.x * d.x + d.y * d.y;
        }
        __syncthreads();
        double sd = 0.0, mn = 1e300;
        for (int n = tid; n < nw; n += WB_TINY_THREADS) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            for (int b = 0; b < nn; b++) {
                const double phi = sphi[b * nw + n], a2 = sa2[b * nw + n], w = a.wb[b];
                const double* bv = a.bvec + ((long long)k * nn + b) * 3;
                s0 += w * bv[0] * phi;
                s1 += w * bv[1] * phi;
                s2 += w * bv[2] * phi;
                s3 += w * (1.0 - a2 + phi * phi);
                sd += w * a2;
                mn = fmin(mn, a2);
            }
            a.part[(long long)(3 * n + 0) * nk + k] = s0;
            a.part[(long long)(3 * n + 1) * nk + k] = s1;
            a.part[(long long)(3 * n + 2) * nk + k] = s2;
            a.part[(long long)(3 * nw + n) * nk + k] = s3;
        }
        sd = tiny_block_sum(sd, red);
        mn = warp_min(mn);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = mn;
        __syncthreads();
        if (tid == 0) {
            double m = 1e300;
            for (int w = 0; w < WB_TINY_THREADS / 32; w++) m = fmin(m, red[w]);
            double sf = 0.0;
            for (int b = 0; b < nn; b++) sf += a.wb[b] * sfro[b];
            a.part[(long long)(4 * nw) * nk + k] = sf;
            a.part[(long long)(4 * nw + 1) * nk + k] = sd;
            a.pmin[k] = sqrt(m);
        }
        __syncthreads();
    }
    TT(3);
    __threadfence();
    grid.sync();
    TT(4);

    // ---- phase 3: global sums (every CTA, same order), spread scalars, gradient of k, projection, <gn_k, s_k> --
    if (a.two_level) {
        // many k-points: re-adding all partial sums in every CTA is O(nk^2) loads.  CTA j < nsums adds column j (all
        // its threads, fixed order), one more grid sync, then everybody reads the nsums results.
        for (int j = blockIdx.x; j < nsums; j += gridDim.x) {
            const double* p = a.part + (long long)j * nk;
            double t = 0.0;
            for (int q = tid; q < nk; q += WB_TINY_THREADS) t += p[q];
            t = tiny_block_sum(t, red);
            if (tid == 0) a.gsums[j] = t;
        }
        __threadfence();
        grid.sync();
        for (int j = tid; j < nsums; j += WB_TINY_THREADS) ssum[j] = a.gsums[j];
    } else {
        for (int j = tid; j < nsums; j += WB_TINY_THREADS) {   // four interleaved partial sums: the loads pipeline
            double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
            const double* p = a.part + (long long)j * nk;
            int q = 0;
            for (; q + 3 < nk; q += 4) { t0 += p[q]; t1 += p[q + 1]; t2 += p[q + 2]; t3 += p[q + 3]; }
            for (; q < nk; q++) t0 += p[q];
            ssum[j] = (t0 + t1) + (t2 + t3);
        }
    }
    __syncthreads();
    const double inv = 1.0 / (double)nk;
    if (blockIdx.x == 0) {   // finalize_kernel
        for (int n = tid; n < nw; n += WB_TINY_THREADS) {
            const double rx = -ssum[3 * n + 0] * inv, ry = -ssum[3 * n + 1] * inv, rz = -ssum[3 * n + 2] * inv;
            a.res[8 + n] = ssum[3 * nw + n] * inv - (rx * rx + ry * ry + rz * rz);
            a.res[8 + nw + 3 * n + 0] = rx;
            a.res[8 + nw + 3 * n + 1] = ry;
            a.res[8 + nw + 3 * n + 2] = rz;
        }
        __syncthreads();
        if (tid == 0) {
            double Om = 0.0, Oc = 0.0;
            for (int n = 0; n < nw; n++) {
                Om += a.res[8 + n];
                if (a.has_penalty) {
                    const double dx = a.res[8 + nw + 3 * n + 0] - a.r0[3 * n + 0];
                    const double dy = a.res[8 + nw + 3 * n + 1] - a.r0[3 * n + 1];
                    const double dz = a.res[8 + nw + 3 * n + 2] - a.r0[3 * n + 2];
                    Oc += a.lambda * (dx * dx + dy * dy + dz * dz);
                }
            }
            const double sumF = ssum[4 * nw], sumD = ssum[4 * nw + 1];
            const double OI = ((double)nw * a.wsum * (double)nk - sumF) * inv;
            const double OOD = (sumF - sumD) * inv;
            const double Ot = Om - OI;
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            a.res[0] = Om;
            a.res[1] = a.split_valid ? OI : nan;
            a.res[2] = a.split_valid ? OOD : nan;
            a.res[3] = a.split_valid ? Ot - OOD : nan;
            a.res[4] = a.split_valid ? Ot : nan;
            a.res[5] = Om + Oc;
            a.res[6] = Oc;
            a.res[7] = 0.0;
        }
    }
    for (int slot = 0, k = blockIdx.x; slot < kper && k < nk; slot++, k += gridDim.x) {
        SLOT_POINTERS(slot)
        const cplx* sk = a.s ? a.s + (long long)k * vlen : nullptr;
        if (a.mu_global)   // (written by this CTA in phase 2: L2, or still L1)
            for (int e = tid; e < nn * mat; e += WB_TINY_THREADS) sMU[e] = a.MUg[(long long)k * nn * mat + e];
        // gradient coefficients (grad_kernel): c = 4 w_b / nk * (t - conj z), t = -i q / z, q = imaglog z + b . r~
        for (int e = tid; e < nn * nw; e += WB_TINY_THREADS) {
            const int b = e / nw, n = e % nw;
            const cplx z = sZ[e];
            double rx = -ssum[3 * n + 0] * inv, ry = -ssum[3 * n + 1] * inv, rz = -ssum[3 * n + 2] * inv;
            if (a.has_penalty) {   // spread.jl:227
                rx -= a.lambda * (rx - a.r0[3 * n + 0]);
                ry -= a.lambda * (ry - a.r0[3 * n + 1]);
                rz -= a.lambda * (rz - a.r0[3 * n + 2]);
            }
            const double* bv = a.bvec + ((long long)k * nn + b) * 3;
            const double q = sphi[e] + (rx * bv[0] + ry * bv[1] + rz * bv[2]);
            const double sq = q / sa2[e];
            const double f = 4.0 * a.wb[b] * inv;
            sZ[e] = make_double2(f * (-sq * z.y - z.x), f * (-sq * z.x + z.y));
        }
        __syncthreads();
        for (int e = tid; e < mat; e += WB_TINY_THREADS) {
            const int n = e / nb;
            cplx acc = make_double2(0.0, 0.0);
            for (int b = 0; b < nn; b++) cfma(acc, sMU[b * mat + e], sZ[b * nw + n]);
            sG[e] = acc;
        }
        __syncthreads();
        cplx* gk = a.gn + (long long)k * vlen;
        double dot = 0.0;
        if (a.mode == 1) {
            // G_X = Y^H G_U, G_Y = G_U X^H with the frozen masks (disentangle.jl:481-501)
            for (int e = tid; e < nxx; e += WB_TINY_THREADS) {
                const int i = e % nw, n = e / nw;
                cplx acc = make_double2(0.0, 0.0);
                for (int m = 0; m < nb; m++) cfma_conj(acc, sY[i * nb + m], sG[n * nb + m]);
                sGX[e] = acc;
            }
            const int nf = a.frozen ? a.nfrozen[k] : 0;
            for (int e = tid; e < mat; e += WB_TINY_THREADS) {
                const int m = e % nb, l = e / nb;
                cplx acc = make_double2(0.0, 0.0);
                if (!(a.frozen && (a.frozen[(long long)k * nb + m] || l < nf)))
                    for (int n = 0; n < nw; n++) cfma(acc, sG[n * nb + m], cconj(sX[n * nw + l]));
                sUp[e] = acc;   // G_Y (the U_{k+b} staging buffer is free now)
            }
            __syncthreads();
            tiny_project(sX, sGX, sP, sT, nw, nw);
            tiny_project(sY, sUp, sP, sT, nb, nw);
            for (int e = tid; e < vlen; e += WB_TINY_THREADS) {
                const cplx g = e < nxx ? sGX[e] : sUp[e - nxx];
                gk[e] = g;
                if (sk) { const cplx d = sk[e]; dot = fma(g.x, d.x, fma(g.y, d.y, dot)); }
            }
        } else {
            tiny_project(sU, sG, sP, sT, nb, nw);
            for (int e = tid; e < mat; e += WB_TINY_THREADS) {
                const cplx g = sG[e];
                gk[e] = g;
                if (sk) { const cplx d = sk[e]; dot = fma(g.x, d.x, fma(g.y, d.y, dot)); }
            }
        }
        dot = tiny_block_sum(dot, red);
        if (tid == 0) a.dpart[k] = dot;
        __syncthreads();
    }
    TT(5);
    __threadfence();
    grid.sync();
    TT(6);

    // ---- phase 4: CTA 0 publishes ----------------------------------------------------------------------------
    if (blockIdx.x == 0) {
        double d0 = 0.0, d1 = 0.0, mn = 1e300;
        for (int q = tid; q < nk; q += 2 * WB_TINY_THREADS) {
            d0 += a.dpart[q]; mn = fmin(mn, a.pmin[q]);
            if (q + WB_TINY_THREADS < nk) { d1 += a.dpart[q + WB_TINY_THREADS]; mn = fmin(mn, a.pmin[q + WB_TINY_THREADS]); }
        }
        const double d = tiny_block_sum(d0 + d1, red);
        mn = warp_min(mn);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = mn;
        __syncthreads();
        if (tid == 0) {
            double m = 1e300;
            for (int w = 0; w < WB_TINY_THREADS / 32; w++) m = fmin(m, red[w]);
            *a.dmin = m;
            s_ok = *reinterpret_cast<volatile int*>(a.flag) == a.epoch ? 0 : 1;
            a.hpin[16] = a.res[5];
            a.hpin[17] = m;
            a.hpin[WB_PIN_CROSS] = d;
            *reinterpret_cast<int*>(a.hpin + 8) = s_ok ? 0 : 1;
#ifdef WB_TINY_TIMING
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt[7]));
            for (int i = 0; i < 7; i++) a.hpin[48 + i] = (double)(tt[i + 1] - tt[i]);
#endif
        }
    }
#undef SLOT_POINTERS
}

size_t tiny_smem_bytes(int nb, int nw, int nn, int kper, bool mu_global) {
    const size_t mat = (size_t)nb * nw, nxx = (size_t)nw * nw;
    const size_t state_c = 2 * mat + nxx + (size_t)nn * nw + (mu_global ? 0 : (size_t)nn * mat);
    const size_t scratch_c = 2 * nxx + 2 * mat + (size_t)nn * nb * nb + (size_t)nn * mat + (mu_global ? (size_t)nn * mat : 0);
    const size_t dbl = (size_t)kper * 2 * nn * nw + (4 * (size_t)nw + 2) + nn + 8;
    return ((size_t)kper * state_c + scratch_c) * sizeof(cplx) + dbl * sizeof(double);
}

}  // namespace

// Is this context's shape served by the fused kernel?  (decided once, at creation: the plain overlaps must be kept)
bool wb_tiny_shape(int nb, int nw, int nk_total, int nk, int nn) {
    static const bool off = [] { const char* e = getenv("WB200_TINY"); return e && e[0] == '0'; }();
    // CUDA-core FP64 on 128 threads per k-point: measured against the general kernels (profiles/r4s_tiny_crossover.txt)
    // it wins by 1.15-1.34 x up to nn nb^2 nw = 14 k complex multiply-adds per k-point (4 x 4 ... 12 x 12), by 1.18 x at
    // 23 k (18 -> 9), and loses 2 % at 33 k (16 x 16): the line is drawn at 24 k
    const double cmacs_per_k = (double)nn * nb * nb * nw;
    return !off && nb <= WB_TINY_MAXB && nw <= WB_TINY_MAXW && nw <= nb && nn <= 32 && nk == nk_total && cmacs_per_k <= 24e3;
}

// mode 0: U on (Stiefel nb x nw)^nk with nb == nw handled as one factor; mode 1: (X, Y).  s may be nullptr.
bool wb_tiny_usable(wb200_ctx* ctx, int mode) {
    if (!ctx->tiny_ok || ctx->allreduce || ctx->peer_self || ctx->profiling || !ctx->d_M) return false;
    if (mode != 0 && mode != 1) return false;
    if (ctx->tiny_blocks == 0) {   // grid and k-points per CTA such that the whole grid is co-resident: once per ctx
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        if (!coop || cudaFuncSetAttribute(tiny_trial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
            cudaGetLastError();
            ctx->tiny_ok = 0;
            return false;
        }
        // One k-point per CTA while the grid is co-resident (register-limited to 4 CTAs of 128 threads per SM: 592
        // k-points on a B200), then two, three, ... k-points per CTA in slots of shared memory (WB200_TINY_KPER forces a
        // minimum, WB200_TINY_MAX_KPER caps it; measured crossover: tools/debug_tiny_crossover.py, profiles/r4_summary.md).
        int kmin = 1, kmax = 4;
        if (const char* e = getenv("WB200_TINY_KPER")) { kmin = atoi(e) > 1 ? atoi(e) : 1; kmax = 64; }
        if (const char* e = getenv("WB200_TINY_MAX_KPER")) kmax = atoi(e) > 0 ? atoi(e) : kmax;
        ctx->tiny_ok = 0;
        // first choice: the rotated overlaps of a k-point stay in its slot; second: they wait in global memory (L2)
        for (int mug = 0; mug < 2 && !ctx->tiny_ok; mug++)
            for (int kper = kmin; kper <= kmax; kper++) {
                const size_t smem = tiny_smem_bytes(ctx->nb, ctx->nw, ctx->nn, kper, mug != 0);
                if (smem > 200 * 1024) break;
                int per_sm = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tiny_trial_kernel, WB_TINY_THREADS, smem) != cudaSuccess) {
                    cudaGetLastError();
                    break;
                }
                const int grid = (ctx->nk + kper - 1) / kper;
                // work per CTA: measured wins up to 23 k complex multiply-adds (18 -> 9, one k-point per CTA) and at
                // 18 k (12 -> 8, two per CTA), a loss at 117 k (18 -> 9, five per CTA): the line is drawn at 40 k
                if (kper > 1 && (double)kper * ctx->nn * ctx->nb * ctx->nb * ctx->nw > 40e3) break;
                if (per_sm * ctx->num_sms >= grid) {
                    ctx->tiny_ok = 1; ctx->tiny_blocks = grid; ctx->tiny_kper = kper; ctx->tiny_mu_global = mug;
                    break;
                }
            }
        if (!ctx->tiny_ok) return false;
    }
    return true;
}

// Enqueues one trial point; the results are in h_pinned[16] (f), [17] (min |N_nn|), [WB_PIN_CROSS] (dphi) and
// h_pinned[8] (int: retraction failed) after the next wait on ctx->stream.
int wb_tiny_trial(wb200_ctx* ctx, int mode, const cplx* x, const cplx* s, double alpha, cplx* xn, cplx* gn) {
    const int nb = ctx->nb, nw = ctx->nw, nk = ctx->nk, nn = ctx->nn;
    if (!ctx->d_tiny) {   // [part | dpart | flag]; the other buffers are the context's own
        const size_t bytes = sizeof(double) * ((size_t)nk + 8);
        WB_CUDA(ctx, wb_dev_alloc(&ctx->d_tiny, bytes));
        WB_CUDA(ctx, cudaMemsetAsync(ctx->d_tiny, 0, bytes, ctx->stream));
        ctx->tiny_epoch = 0;
    }
    if (mode == 1 && !ctx->d_U) { ctx->err = "wb_tiny_trial: no gauge buffer"; return WB200_ERR_ARG; }
    TinyArgs a{};
    a.mode = mode; a.nb = nb; a.nw = nw; a.nk = nk; a.nn = nn; a.kper = ctx->tiny_kper;
    a.mu_global = ctx->tiny_mu_global; a.MUg = ctx->d_MU;
    a.two_level = ctx->nk > ctx->num_sms ? 1 : 0; a.gsums = ctx->d_sums;   // (by nk, not by the grid: the result does not depend on k-points per CTA)
    a.kpb = ctx->d_kpb; a.bvec = ctx->d_bvec; a.wb = ctx->d_wb; a.M = ctx->d_M;
    a.frozen = mode == 1 ? ctx->d_frozen : nullptr; a.nfrozen = ctx->d_nfrozen;
    a.has_penalty = ctx->has_penalty; a.lambda = ctx->lambda; a.r0 = ctx->d_r0; a.wsum = ctx->wsum;
    ctx->split_valid = nb == nw;
    a.split_valid = ctx->split_valid;
    a.x = x; a.s = s; a.alpha = alpha; a.xn = xn; a.gn = gn;
    a.U = mode == 1 ? ctx->d_U : xn;
    a.part = ctx->d_part; a.pmin = ctx->d_pmin; a.dpart = ctx->d_tiny;
    a.res = ctx->d_res; a.dmin = ctx->d_min;
    a.flag = reinterpret_cast<int*>(ctx->d_tiny + nk);
    a.epoch = ++ctx->tiny_epoch;
    a.hpin = ctx->h_pinned;
    void* args[] = {&a};
    const size_t smem = tiny_smem_bytes(nb, nw, nn, ctx->tiny_kper, ctx->tiny_mu_global != 0);
    WB_CUDA(ctx, cudaLaunchCooperativeKernel((void*)tiny_trial_kernel, dim3(ctx->tiny_blocks), dim3(WB_TINY_THREADS), args, smem,
                                             ctx->stream));
    WB_LAUNCH_CHECK(ctx);
    ctx->retract_pending = 1;   // h_pinned[8] carries the flag of this launch
    return WB200_OK;
}
