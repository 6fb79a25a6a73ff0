// stiefel_small.cu — fused per-matrix Stiefel kernels for small shapes (nrow, ncol <= 64).
//
// One CTA owns one matrix and keeps it in shared memory for the whole operation, so a retraction
// (all Newton-Schulz iterations, per-matrix convergence test included) or a tangent projection is
// ONE launch with no host round trip, instead of ~15 batched-GEMM launches and one synchronisation
// per iteration.  This is what the optimiser loop of max_localize / disentangle spends its time
// in for the reference-sized problems (n_wann ~ 4..32), where everything is launch-bound.
//
// Replaces Optim.Stiefel_SVD retract! / project_tangent! (third-party Optim.jl; call sites
// src/localization/max_localize.jl:62-63, src/localization/disentangle.jl:687-690).
// Shared-memory leading dimensions are padded by one complex element: a warp reading 32
// consecutive columns then touches 8 distinct 16-byte bank groups, i.e. the minimum four wavefronts.
#include "common.cuh"

#define WB_SMALL_MAX 64
#define WB_SMALL_THREADS 256

namespace {

__device__ __forceinline__ double block_sum(double v, double* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < WB_SMALL_THREADS / 32; w++) t += sh[w];
    return t;
}

// X <- polar(X) by Newton-Schulz, converged per matrix to ||X^H X - I||_F < tol before the last update
__global__ void __launch_bounds__(WB_SMALL_THREADS) retract_small_kernel(int nrow, int ncol, int ld,
                                                                        long long stride,
                                                                        cplx* __restrict__ X, int maxit,
                                                                        double tol, int* __restrict__ flag) {
    extern __shared__ cplx sm[];
    const int lds = nrow + 1;
    cplx* sX = sm;                    // [ncol][lds]
    cplx* sP = sm + ncol * lds;       // [ncol][ncol]
    __shared__ double red[WB_SMALL_THREADS / 32];
    cplx* g = X + (long long)blockIdx.x * stride;
    const int tid = threadIdx.x;
    for (int e = tid; e < nrow * ncol; e += WB_SMALL_THREADS) sX[(e / nrow) * lds + (e % nrow)] = g[(e / nrow) * (long long)ld + (e % nrow)];
    __syncthreads();
    bool converged = false;
    for (int it = 0; it < maxit && !converged; it++) {
        double e2 = 0.0, f2 = 0.0;
        for (int e = tid; e < ncol * ncol; e += WB_SMALL_THREADS) {
            const int i = e % ncol, j = e / ncol;
            cplx acc = make_double2(0.0, 0.0);
            const cplx* xi = sX + i * lds;
            const cplx* xj = sX + j * lds;
            for (int m = 0; m < nrow; m++) cfma_conj(acc, xi[m], xj[m]);
            sP[e] = acc;
            const double dx = acc.x - (i == j ? 1.0 : 0.0);
            e2 += dx * dx + acc.y * acc.y;
            f2 += acc.x * acc.x + acc.y * acc.y;
        }
        const double err = sqrt(block_sum(e2, red));
        const double fro = sqrt(block_sum(f2, red));
        if (!(err == err)) break;    // NaN input: leave and flag
        converged = err < tol;       // the update below takes err -> ~err^2
        // sigma_max^2 <= 1 + err: Newton-Schulz needs sigma^2 < 3; otherwise (first pass only)
        // scale by ||P||_F^(-1/2) so that all singular values are <= 1
        const double s = (it == 0 && err >= 1.9) ? rsqrt(fro) : 1.0;
        const double s3 = 0.5 * s * s * s;
        for (int e = tid; e < ncol * ncol; e += WB_SMALL_THREADS) {
            const int i = e % ncol, j = e / ncol;
            cplx v = sP[e];
            v.x = (i == j ? 1.5 * s : 0.0) - s3 * v.x;
            v.y = -s3 * v.y;
            sP[e] = v;
        }
        __syncthreads();
        // X <- X T, staged through registers (at most 64*64/256 = 16 outputs per thread)
        cplx out[WB_SMALL_MAX * WB_SMALL_MAX / WB_SMALL_THREADS];
#pragma unroll
        for (int q = 0; q < WB_SMALL_MAX * WB_SMALL_MAX / WB_SMALL_THREADS; q++) {
            const int e = tid + q * WB_SMALL_THREADS;
            if (e < nrow * ncol) {
                const int m = e % nrow, n = e / nrow;
                cplx acc = make_double2(0.0, 0.0);
                const cplx* tn = sP + n * ncol;
                for (int l = 0; l < ncol; l++) cfma(acc, sX[l * lds + m], tn[l]);
                out[q] = acc;
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < WB_SMALL_MAX * WB_SMALL_MAX / WB_SMALL_THREADS; q++) {
            const int e = tid + q * WB_SMALL_THREADS;
            if (e < nrow * ncol) sX[(e / nrow) * lds + (e % nrow)] = out[q];
        }
        __syncthreads();
    }
    for (int e = tid; e < nrow * ncol; e += WB_SMALL_THREADS) g[(e / nrow) * (long long)ld + (e % nrow)] = sX[(e / nrow) * lds + (e % nrow)];
    if (!converged && tid == 0) atomicMax(flag, 1);
}

// G <- G - X (X^H G + G^H X)/2
__global__ void __launch_bounds__(WB_SMALL_THREADS) project_small_kernel(int nrow, int ncol, int ld,
                                                                        long long stride,
                                                                        const cplx* __restrict__ X,
                                                                        cplx* __restrict__ G) {
    extern __shared__ cplx sm[];
    const int lds = nrow + 1;
    cplx* sX = sm;
    cplx* sG = sm + ncol * lds;
    cplx* sS = sm + 2 * ncol * lds;   // [ncol][ncol]
    const cplx* gx = X + (long long)blockIdx.x * stride;
    cplx* gg = G + (long long)blockIdx.x * stride;
    const int tid = threadIdx.x;
    for (int e = tid; e < nrow * ncol; e += WB_SMALL_THREADS) {
        const int m = e % nrow, n = e / nrow;
        sX[n * lds + m] = gx[n * (long long)ld + m];
        sG[n * lds + m] = gg[n * (long long)ld + m];
    }
    __syncthreads();
    for (int e = tid; e < ncol * ncol; e += WB_SMALL_THREADS) {
        const int i = e % ncol, j = e / ncol;
        cplx acc = make_double2(0.0, 0.0);
        const cplx* xi = sX + i * lds;
        const cplx* gj = sG + j * lds;
        for (int m = 0; m < nrow; m++) cfma_conj(acc, xi[m], gj[m]);
        sS[e] = acc;
    }
    __syncthreads();
    for (int e = tid; e < ncol * ncol; e += WB_SMALL_THREADS) {
        const int i = e % ncol, j = e / ncol;
        if (i > j) continue;
        const cplx a = sS[i + j * ncol], b = sS[j + i * ncol];
        const cplx h = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
        sS[i + j * ncol] = h;
        sS[j + i * ncol] = cconj(h);
    }
    __syncthreads();
    for (int e = tid; e < nrow * ncol; e += WB_SMALL_THREADS) {
        const int m = e % nrow, n = e / nrow;
        cplx acc = make_double2(0.0, 0.0);
        const cplx* sn = sS + n * ncol;
        for (int l = 0; l < ncol; l++) cfma(acc, sX[l * lds + m], sn[l]);
        const cplx o = sG[n * lds + m];
        gg[n * (long long)ld + m] = make_double2(o.x - acc.x, o.y - acc.y);
    }
}


// ---------------------------------------------------------------------------------------------
// n <= 32: the same two operations on the FP64 tensor pipe.
// The kernels above read two 16-byte shared-memory operands per complex multiply-add and run at
// shared-memory bandwidth.  DMMA.8x8x4 needs one operand element per lane per 512 flops, so for the
// sizes that dominate max_localize at the reference-like shapes (n_wann <= 32) a CTA of 8 warps
// owns FOUR matrices (a warp pair each, warp h of the pair computes columns 16 h .. 16 h + 15,
// zero padded to 32 x 32), keeps them in shared memory (column-major, leading dimension 36 complex
// so the B-operand / A^H-operand fragment loads are conflict-free) and does every product with the
// Gauss 3-multiplication complex DMMA of the rotation kernels.
// ---------------------------------------------------------------------------------------------
constexpr int S32_LD = 36;
constexpr int S32_MAT = 32 * S32_LD;   // complex elements per padded matrix

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// c[(mt*2 + nt)*2 + j] = sum_l opA[m][l] * B[l][n],  m = mt*8 + g,  n = half*16 + nt*8 + 2t + j
//   AH = true : opA[m][l] = conj(sA[m*LD + l])   (A^H of the column-major array sA)
//   AH = false: opA[m][l] = sA[l*LD + m]
//   B[l][n] = loadB(l, n)
// mtiles: row tiles (of 8) that can be non-zero; ksteps: k-steps (of 4) that can be non-zero.
template <bool AH, typename LoadB>
__device__ __forceinline__ void zgemm32(const cplx* __restrict__ sA, LoadB loadB, int mtiles, int ksteps,
                                        int half, int lane, cplx (&c)[16]) {
    const int g = lane >> 2, t = lane & 3;
    double p1[4][2][2], p2[4][2][2], p3[4][2][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
            for (int e = 0; e < 2; e++) p1[a][b][e] = p2[a][b][e] = p3[a][b][e] = 0.0;
    for (int ks = 0; ks < ksteps; ks++) {
        double qb[2][3];
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            const cplx b = loadB(ks * 4 + t, half * 16 + nt * 8 + g);
            qb[nt][0] = b.x;
            qb[nt][1] = b.y - b.x;
            qb[nt][2] = b.x + b.y;
        }
#pragma unroll
        for (int mt = 0; mt < 4; mt++) {
            if (mt < mtiles) {
                const cplx a = AH ? cconj(sA[(mt * 8 + g) * S32_LD + ks * 4 + t]) : sA[(ks * 4 + t) * S32_LD + mt * 8 + g];
                const double as = a.x + a.y;
#pragma unroll
                for (int nt = 0; nt < 2; nt++) {
                    dmma884(p1[mt][nt][0], p1[mt][nt][1], as, qb[nt][0]);
                    dmma884(p2[mt][nt][0], p2[mt][nt][1], a.x, qb[nt][1]);
                    dmma884(p3[mt][nt][0], p3[mt][nt][1], a.y, qb[nt][2]);
                }
            }
        }
    }
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
#pragma unroll
            for (int j = 0; j < 2; j++)
                c[(mt * 2 + nt) * 2 + j] = make_double2(p1[mt][nt][j] - p3[mt][nt][j], p1[mt][nt][j] + p2[mt][nt][j]);
}

// zero-padded copy of a matrix (nrow x ncol, leading dimension ld) into smem by the 64 threads of its
// warp pair; the 16 loads of a thread are issued together (one exposed latency, not sixteen)
__device__ __forceinline__ void load32(const cplx* __restrict__ g, int nrow, int ncol, int ld, cplx* sm, int t64) {
    cplx v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int e = t64 + 64 * k, m = e & 31, n = e >> 5;
        v[k] = (m < nrow && n < ncol) ? g[n * (long long)ld + m] : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int e = t64 + 64 * k, m = e & 31, n = e >> 5;
        sm[n * S32_LD + m] = v[k];
    }
}

__global__ void __launch_bounds__(256, 1) retract32_kernel(int nrow, int ncol, int ld, long long stride,
                                                           cplx* __restrict__ X, int count, int maxit,
                                                           double tol, int* __restrict__ flag) {
    extern __shared__ cplx sm[];
    __shared__ double red[8][2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp >> 1, half = warp & 1, g = lane >> 2, t = lane & 3;
    const int mat = blockIdx.x * 4 + q;
    const bool valid = mat < count;
    cplx* sX = sm + q * 2 * S32_MAT;
    cplx* sT = sX + S32_MAT;
    cplx* gX = X + (long long)(valid ? mat : 0) * stride;
    load32(gX, valid ? nrow : 0, ncol, ld, sX, tid & 63);
    __syncthreads();
    const int kr = (nrow + 3) / 4, kc = (ncol + 3) / 4, tr = (nrow + 7) / 8, tc = (ncol + 7) / 8;
    bool converged = !valid, bad = false;
    for (int it = 0; it < maxit; it++) {
        cplx c[16];
        // P = X^H X (this warp: columns of its half)
        zgemm32<true>(sX, [&](int l, int n) { return sX[n * S32_LD + l]; }, tc, kr, half, lane, c);
        double e2 = 0.0, f2 = 0.0;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int m = (e >> 2) * 8 + g, n = half * 16 + ((e >> 1) & 1) * 8 + 2 * t + (e & 1);
            if (m < ncol && n < ncol) {
                const double dx = c[e].x - (m == n ? 1.0 : 0.0);
                e2 += dx * dx + c[e].y * c[e].y;
                f2 += c[e].x * c[e].x + c[e].y * c[e].y;
            }
        }
        e2 = warp_sum(e2);
        f2 = warp_sum(f2);
        if (lane == 0) { red[warp][0] = e2; red[warp][1] = f2; }
        __syncthreads();
        const double err = sqrt(red[2 * q][0] + red[2 * q + 1][0]);
        const double fro = sqrt(red[2 * q][1] + red[2 * q + 1][1]);
        if (valid && !(err == err)) bad = true;   // NaN input
        if (valid && err < tol) converged = true;   // the update below takes err -> ~err^2
        // sigma_max^2 <= 1 + err: Newton-Schulz needs sigma^2 < 3; otherwise (first pass only)
        // scale by ||P||_F^(-1/2) so that all singular values are <= 1
        const double s = (it == 0 && err >= 1.9) ? rsqrt(fro) : 1.0;
        const double s3 = 0.5 * s * s * s;
        // T = s (1.5 I - 0.5 s^2 P): this warp's columns, which are exactly the ones its X T needs
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int m = (e >> 2) * 8 + g, n = half * 16 + ((e >> 1) & 1) * 8 + 2 * t + (e & 1);
            cplx v = make_double2(0.0, 0.0);
            if (m < ncol && n < ncol) v = make_double2((m == n ? 1.5 * s : 0.0) - s3 * c[e].x, -s3 * c[e].y);
            sT[n * S32_LD + m] = v;
        }
        __syncwarp();
        zgemm32<false>(sX, [&](int l, int n) { return sT[n * S32_LD + l]; }, tr, kc, half, lane, c);
        // every warp has finished reading the old X; the CTA stops once all four matrices are done
        const int all_done = __syncthreads_and(converged || bad);
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int m = (e >> 2) * 8 + g, n = half * 16 + ((e >> 1) & 1) * 8 + 2 * t + (e & 1);
            sX[n * S32_LD + m] = c[e];
        }
        __syncthreads();
        if (all_done) break;
    }
    if (valid) {
        for (int e = tid & 63; e < nrow * ncol; e += 64) {
            const int m = e % nrow, n = e / nrow;
            gX[n * (long long)ld + m] = sX[n * S32_LD + m];
        }
        if ((!converged || bad) && (tid & 63) == 0) atomicMax(flag, 1);
    }
}

// G <- G - X (X^H G + G^H X)/2
__global__ void __launch_bounds__(256, 1) project32_kernel(int nrow, int ncol, int ld, long long stride,
                                                           const cplx* __restrict__ X, cplx* __restrict__ G,
                                                           int count) {
    extern __shared__ cplx sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp >> 1, half = warp & 1, g = lane >> 2, t = lane & 3;
    const int mat = blockIdx.x * 4 + q;
    const bool valid = mat < count;
    cplx* sX = sm + q * 3 * S32_MAT;
    cplx* sG = sX + S32_MAT;
    cplx* sS = sG + S32_MAT;
    const cplx* gX = X + (long long)(valid ? mat : 0) * stride;
    cplx* gG = G + (long long)(valid ? mat : 0) * stride;
    load32(gX, valid ? nrow : 0, ncol, ld, sX, tid & 63);
    load32(gG, valid ? nrow : 0, ncol, ld, sG, tid & 63);
    __syncthreads();
    const int kr = (nrow + 3) / 4, kc = (ncol + 3) / 4, tr = (nrow + 7) / 8, tc = (ncol + 7) / 8;
    cplx c[16];
    zgemm32<true>(sX, [&](int l, int n) { return sG[n * S32_LD + l]; }, tc, kr, half, lane, c);   // S = X^H G
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const int m = (e >> 2) * 8 + g, n = half * 16 + ((e >> 1) & 1) * 8 + 2 * t + (e & 1);
        sS[n * S32_LD + m] = c[e];
    }
    __syncthreads();   // the Hermitian part needs the other half's columns
    zgemm32<false>(sX,
                   [&](int l, int n) {
                       const cplx a = sS[n * S32_LD + l], b = sS[l * S32_LD + n];
                       return make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
                   },
                   tr, kc, half, lane, c);
    if (valid) {
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int m = (e >> 2) * 8 + g, n = half * 16 + ((e >> 1) & 1) * 8 + 2 * t + (e & 1);
            if (m < nrow && n < ncol) {
                const cplx o = sG[n * S32_LD + m];
                gG[n * (long long)ld + m] = make_double2(o.x - c[e].x, o.y - c[e].y);
            }
        }
    }
}

}  // namespace

bool wb_small_applicable(int nrow, int ncol) { return nrow <= WB_SMALL_MAX && ncol <= WB_SMALL_MAX; }

// DMMA (four matrices per CTA, 32 x 32 padded tiles) or CUDA cores (one CTA per matrix)?  The tensor kernels have a
// latency floor of ~28 us per launch whatever the size (profiles/r4e_launches_valence_maxloc.csv: a 32 x 32 tile
// iterating on 4 x 4 matrices) and win only when there is work to fill them: 1 728 matrices of 32 x 32 take 107 us
// against 306 us on the CUDA cores, but with WB200_STIEFEL_DMMA=0 the optimiser was 6-25 % faster per evaluation at
// every shape of tools/debug_tiny_crossover.py with n <= 24 and up to 729 matrices.  Threshold between those
// measurements: count * nrow * ncol^2 > 5e6 complex multiply-adds per product.  WB200_STIEFEL_DMMA = 0 / 1 forces.
static bool use32(int nrow, int ncol, int count) {
    if (nrow > 32 || ncol > 32) return false;
    if (const char* e = getenv("WB200_STIEFEL_DMMA")) {
        if (e[0] == '0') return false;
        if (e[0] == '1') return true;
    }
    return (double)count * nrow * ncol * ncol > 5e6;
}

// "did every matrix converge?" comes back in h_pinned[8].  Normally the caller waits for it here.  The optimiser
// (ctx->retract_deferred, single-slab runs) does not: the evaluation that follows is enqueued right behind the
// retraction and the flag is read after THAT wait (wb_retract_pending_failed) — one host round trip per trial point
// instead of two; the flag accumulates over the retractions of one trial point (X and Y factors).
static int retract_flag_check(wb200_ctx* ctx, int* flag) {
    int* hflag = reinterpret_cast<int*>(ctx->h_pinned + 8);
    if (ctx->retract_deferred) { ctx->retract_pending = 1; return WB200_OK; }   // delivered with the evaluation's results
    WB_CUDA(ctx, cudaMemcpyAsync(hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*hflag) {
        ctx->err = "retract: Newton-Schulz polar iteration did not converge (singular or NaN input?)";
        return WB200_ERR_NOCONV;
    }
    return WB200_OK;
}
// after a wait on ctx->stream: did a deferred retraction fail?  Clears the pending state.
bool wb_retract_pending_failed(wb200_ctx* ctx) {
    if (!ctx->retract_pending) return false;
    ctx->retract_pending = 0;
    return *reinterpret_cast<int*>(ctx->h_pinned + 8) != 0;
}

int wb_retract_small(wb200_ctx* ctx, cplx* dX, int nrow, int ncol, int ld, long long stride, int count) {
    if (use32(nrow, ncol, count)) {
        const size_t smem32 = sizeof(cplx) * 4 * 2 * S32_MAT;
        if (!(ctx->attr_mask & 16u)) {
            WB_CUDA(ctx, cudaFuncSetAttribute(retract32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32));
            ctx->attr_mask |= 16u;
        }
        int* flag = reinterpret_cast<int*>(ctx->d_scal + 4);
        if (!ctx->retract_pending) WB_CUDA(ctx, cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
        retract32_kernel<<<(count + 3) / 4, 256, smem32, ctx->stream>>>(nrow, ncol, ld, stride, dX, count, 100, 2e-8, flag);
        WB_LAUNCH_CHECK(ctx);
        return retract_flag_check(ctx, flag);
    }
    const size_t smem = sizeof(cplx) * ((size_t)ncol * (nrow + 1) + (size_t)ncol * ncol);
    if (!(ctx->attr_mask & 2u)) {
        WB_CUDA(ctx, cudaFuncSetAttribute(retract_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(sizeof(cplx) * (WB_SMALL_MAX * (WB_SMALL_MAX + 1) + WB_SMALL_MAX * WB_SMALL_MAX))));
        ctx->attr_mask |= 2u;
    }
    int* flag = reinterpret_cast<int*>(ctx->d_scal + 4);
    if (!ctx->retract_pending) WB_CUDA(ctx, cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    retract_small_kernel<<<count, WB_SMALL_THREADS, smem, ctx->stream>>>(nrow, ncol, ld, stride, dX, 100, 2e-8, flag);
    WB_LAUNCH_CHECK(ctx);
    return retract_flag_check(ctx, flag);
}

int wb_project_small(wb200_ctx* ctx, const cplx* dX, cplx* dG, int nrow, int ncol, int ld, long long stride,
                     int count) {
    if (use32(nrow, ncol, count)) {
        const size_t smem32 = sizeof(cplx) * 4 * 3 * S32_MAT;
        if (!(ctx->attr_mask & 32u)) {
            WB_CUDA(ctx, cudaFuncSetAttribute(project32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32));
            ctx->attr_mask |= 32u;
        }
        project32_kernel<<<(count + 3) / 4, 256, smem32, ctx->stream>>>(nrow, ncol, ld, stride, dX, dG, count);
        WB_LAUNCH_CHECK(ctx);
        return WB200_OK;
    }
    const size_t smem = sizeof(cplx) * (2 * (size_t)ncol * (nrow + 1) + (size_t)ncol * ncol);
    if (!(ctx->attr_mask & 4u)) {
        WB_CUDA(ctx, cudaFuncSetAttribute(project_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(sizeof(cplx) * (2 * WB_SMALL_MAX * (WB_SMALL_MAX + 1) + WB_SMALL_MAX * WB_SMALL_MAX))));
        ctx->attr_mask |= 4u;
    }
    project_small_kernel<<<count, WB_SMALL_THREADS, smem, ctx->stream>>>(nrow, ncol, ld, stride, dX, dG);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}
