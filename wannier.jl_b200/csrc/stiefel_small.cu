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

}  // namespace

bool wb_small_applicable(int nrow, int ncol) { return nrow <= WB_SMALL_MAX && ncol <= WB_SMALL_MAX; }

int wb_retract_small(wb200_ctx* ctx, cplx* dX, int nrow, int ncol, int ld, long long stride, int count) {
    const size_t smem = sizeof(cplx) * ((size_t)ncol * (nrow + 1) + (size_t)ncol * ncol);
    static bool attr_done[16] = {false};
    if (!attr_done[ctx->device & 15]) {
        WB_CUDA(ctx, cudaFuncSetAttribute(retract_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(sizeof(cplx) * (WB_SMALL_MAX * (WB_SMALL_MAX + 1) + WB_SMALL_MAX * WB_SMALL_MAX))));
        attr_done[ctx->device & 15] = true;
    }
    int* flag = reinterpret_cast<int*>(ctx->d_scal + 4);
    WB_CUDA(ctx, cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    retract_small_kernel<<<count, WB_SMALL_THREADS, smem, ctx->stream>>>(nrow, ncol, ld, stride, dX, 100, 2e-8, flag);
    WB_LAUNCH_CHECK(ctx);
    int* hflag = reinterpret_cast<int*>(ctx->h_pinned + 8);
    WB_CUDA(ctx, cudaMemcpyAsync(hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*hflag) {
        ctx->err = "retract: Newton-Schulz polar iteration did not converge (singular or NaN input?)";
        return WB200_ERR_NOCONV;
    }
    return WB200_OK;
}

int wb_project_small(wb200_ctx* ctx, const cplx* dX, cplx* dG, int nrow, int ncol, int ld, long long stride,
                     int count) {
    const size_t smem = sizeof(cplx) * (2 * (size_t)ncol * (nrow + 1) + (size_t)ncol * ncol);
    static bool attr_done[16] = {false};
    if (!attr_done[ctx->device & 15]) {
        WB_CUDA(ctx, cudaFuncSetAttribute(project_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(sizeof(cplx) * (2 * WB_SMALL_MAX * (WB_SMALL_MAX + 1) + WB_SMALL_MAX * WB_SMALL_MAX))));
        attr_done[ctx->device & 15] = true;
    }
    project_small_kernel<<<count, WB_SMALL_THREADS, smem, ctx->stream>>>(nrow, ncol, ld, stride, dX, dG);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}
