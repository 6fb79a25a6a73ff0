// stiefel.cu — per-k Stiefel-manifold operations and the (X, Y) disentanglement chain.
//
// Replaces (reference tree):
//   Optim.Stiefel_SVD project_tangent! / retract!  (third-party Optim.jl; call sites
//       src/localization/max_localize.jl:62-63, src/localization/disentangle.jl:687-690)
//   X_Y_to_U!      src/localization/disentangle.jl:351-356
//   XY_to_X_Y!     src/localization/disentangle.jl:424-435   (a view: column ik of XY already IS
//                  [X_k col-major ; Y_k col-major], so no copy is made)
//   GU_to_GX_GY    src/localization/disentangle.jl:481-501
#include "common.cuh"

#include <algorithm>

// ---------------------------------------------------------------------------------------------
// S <- (S + S^H)/2 for a batch of n x n matrices (contiguous)
// ---------------------------------------------------------------------------------------------
__global__ void hermitize_kernel(int n, long long total, cplx* __restrict__ S) {
    const long long nn2 = (long long)n * n;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long mat = e / nn2;
        const int r = (int)(e % nn2);
        const int i = r % n, j = r / n;
        if (i > j) continue;
        cplx* base = S + mat * nn2;
        const cplx a = base[i + (long long)j * n], b = base[j + (long long)i * n];
        const cplx h = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
        base[i + (long long)j * n] = h;
        base[j + (long long)i * n] = cconj(h);
    }
}

static int project_strided(wb200_ctx* ctx, const cplx* dX, cplx* dG, int nrow, int ncol, int ld,
                           long long stride, int count) {
    if (wb_small_applicable(nrow, ncol)) return wb_project_small(ctx, dX, dG, nrow, ncol, ld, stride, count);
    // S = X^H G, S <- (S + S^H)/2, G <- G - X S: two batched GEMMs (FP64 tensor pipe from the plain
    // layouts where the shape allows, the second one accumulating into G) and one small pass over S
    WB_TRY(wb_ensure_tmp(ctx, (size_t)count * ncol * ncol, 0));
    ZgemmArgs a{};
    a.m = ncol; a.n = ncol; a.k = nrow;
    a.A = dX; a.strideA = stride; a.lda = ld; a.transA = 1;
    a.B = dG; a.strideB = stride; a.ldb = ld; a.transB = 0;
    a.C = ctx->d_tmp1; a.strideC = (long long)ncol * ncol; a.ldc = ncol;
    a.alpha = 1.0; a.beta = 0.0; a.batch = count;
    WB_TRY(wb_zgemm(ctx, a));
    const long long total = (long long)count * ncol * ncol;
    const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    hermitize_kernel<<<blocks, 256, 0, ctx->stream>>>(ncol, total, ctx->d_tmp1);
    WB_LAUNCH_CHECK(ctx);
    ZgemmArgs b{};
    b.m = nrow; b.n = ncol; b.k = ncol;
    b.A = dX; b.strideA = stride; b.lda = ld; b.transA = 0;
    b.B = ctx->d_tmp1; b.strideB = (long long)ncol * ncol; b.ldb = ncol; b.transB = 0;
    b.C = dG; b.strideC = stride; b.ldc = ld;
    b.alpha = -1.0; b.beta = 1.0; b.batch = count;
    WB_TRY(wb_zgemm(ctx, b));
    return WB200_OK;
}

int wb_project_tangent_dev(wb200_ctx* ctx, const cplx* dX, cplx* dG, int nrow, int ncol, int count) {
    return project_strided(ctx, dX, dG, nrow, ncol, nrow, (long long)nrow * ncol, count);
}

// ---------------------------------------------------------------------------------------------
// polar factor by Newton-Schulz:  X <- X (3 I - X^H X)/2
// The kernel below turns P = X^H X into T = s (1.5 I - 0.5 s^2 P) in place (s = 1 unless the
// matrix is far from unitary, then s = ||X^H X||_F^(-1/2) so that all singular values are <= 1) and
// records max_k ||P_k - I||_F.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ns_prepare_kernel(int n, cplx* __restrict__ P,
                                                         double* __restrict__ err_max,
                                                         int allow_scale) {
    cplx* p = P + (long long)blockIdx.x * n * n;
    double e2 = 0.0, tr = 0.0;   // ||P - I||_F^2 and ||P||_F^2
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e % n, j = e / n;
        const cplx v = p[e];
        const double dx = v.x - (i == j ? 1.0 : 0.0);
        e2 += dx * dx + v.y * v.y;
        tr += v.x * v.x + v.y * v.y;
    }
    __shared__ double s1[8], s2[8];
    __shared__ double scale;
    e2 = warp_sum(e2);
    tr = warp_sum(tr);
    if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = e2; s2[threadIdx.x >> 5] = tr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; w++) { a += s1[w]; b += s2[w]; }
        const double err = sqrt(a);
        // sigma_max^2 <= 1 + err: Newton-Schulz needs sigma^2 < 3.  Otherwise (first pass only;
        // the map keeps sigma in (0, 1] afterwards) scale by ||P||_F^(-1/2) >= 1/sigma_max.
        scale = (err < 1.9 || !allow_scale) ? 1.0 : rsqrt(sqrt(b));
        atomicMax((unsigned long long*)err_max, (unsigned long long)__double_as_longlong(err));
    }
    __syncthreads();
    const double s = scale, s3 = 0.5 * s * s * s;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e % n, j = e / n;
        cplx v = p[e];
        v.x = (i == j ? 1.5 * s : 0.0) - s3 * v.x;
        v.y = -s3 * v.y;
        p[e] = v;
    }
}

__global__ void copy_strided_kernel(int nrow, int ncol, int ld, long long stride, long long total,
                                    const cplx* __restrict__ src, cplx* __restrict__ dst) {
    const long long per = (long long)nrow * ncol;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long mat = e / per;
        const int r = (int)(e % per);
        dst[mat * stride + (r % nrow) + (long long)(r / nrow) * ld] = src[e];
    }
}

// T <- a I - b P in place for a batch of n x n matrices (first step of the scaled iteration)
__global__ void __launch_bounds__(256) ns_scale_kernel(int n, long long total, double a, double b, cplx* __restrict__ P) {
    const long long nn2 = (long long)n * n;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e % nn2);
        const cplx v = P[e];
        P[e] = make_double2(((r % n) == (r / n) ? a : 0.0) - b * v.x, -b * v.y);
    }
}

// T = a I - b S out of place (first step of a line-search retraction whose X^H X = I + alpha^2 S is known)
__global__ void __launch_bounds__(256) ns_from_gram_kernel(int n, long long total, double a, double b,
                                                           const cplx* __restrict__ S, cplx* __restrict__ T) {
    const long long nn2 = (long long)n * n;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e % nn2);
        const cplx v = S[e];
        T[e] = make_double2(((r % n) == (r / n) ? a : 0.0) - b * v.x, -b * v.y);
    }
}

// Newton-Schulz with the optimal scaling for a known singular-value interval [lo, hi]:
//   X <- c X (3 I - c^2 X^H X) / 2,  c^2 = 3 / (lo^2 + lo hi + hi^2)  makes p(c lo) = p(c hi), the new lower
//   bound, while the maximum of p(x) = x (3 - x^2) / 2, p(1) = 1, is the new upper bound.
// Unscaled, a singular value s << 1 grows by 1.5 per iteration; scaled, the interval [lo, hi] collapses
// like kappa -> ~sqrt(kappa).  The line search of the optimiser takes large trial steps in its first
// iterations (Optim's alphaguess is 1), so this halves the GEMMs of those retractions.
struct NsInterval {
    double lo, hi;
    double c() const { return sqrt(3.0 / (lo * lo + lo * hi + hi * hi)); }
    static double p(double x) { return 0.5 * x * (3.0 - x * x); }
    void tighten(double err) {   // rigorous: |lambda(X^H X) - 1| <= ||X^H X - I||_F
        hi = fmin(hi, sqrt(1.0 + err));
        if (err < 1.0) lo = fmax(lo, sqrt(1.0 - err));
    }
    void step(double cc) {
        const double a = p(cc * lo), b = p(cc * hi);
        const bool spans_one = cc * lo <= 1.0 && cc * hi >= 1.0;
        lo = fmin(a, b);
        hi = spans_one ? 1.0 : fmax(a, b);
        lo = fmax(lo * (1.0 - 1e-12), 0.0);   // guard the analytic bounds against rounding
        hi = hi * (1.0 + 1e-12);
    }
};

// gram != nullptr (tangent steps of one line search): X = X0 + alpha s with X0^H X0 = I and X0^H s skew, hence
// X^H X = I + alpha^2 S with S = s^H s formed ONCE per search direction (wb_tangent_gram): the first pass needs
// neither its GEMM nor a host round trip for the error (alpha^2 max_k ||S_k||_F).
static int retract_strided(wb200_ctx* ctx, cplx* dX, int nrow, int ncol, int ld, long long stride,
                           int count, int* iters_out, bool tangent_step, const WbTangentGram* gram = nullptr) {
    if (wb_small_applicable(nrow, ncol)) {   // fused single-launch kernel, convergence tested on the device
        if (iters_out) *iters_out = -1;
        return wb_retract_small(ctx, dX, nrow, ncol, ld, stride, count);
    }
    const int MAXIT = 100;
    // Quadratic convergence with a known constant: with E = X^H X - I the step X (3 I - X^H X)/2 leaves
    // E' = -(3/4) E^2 + (1/4) E^3, so ||E'||_F <= 0.75 ||E||_F^2 (1 + ||E||/3).  The update applied at a MEASURED
    // err < 1e-6 therefore leaves ||X^H X - I||_F < 7.5e-13 without another X^H X to confirm it (the scaled step
    // only shrinks the bound).  (2e-8 before: ~40 % of the line-search retractions ran one iteration, 2 GEMMs, more
    // just to see 1e-14.)
    const double TOL = 1e-6;
    WB_TRY(wb_ensure_tmp(ctx, (size_t)count * ncol * ncol, (size_t)count * nrow * ncol));
    const bool tensor = !(ctx->flags & WB200_FLAG_NO_TENSOR) && wb_zgemm_tensor_applicable(ncol, ncol, nrow);
    if (tensor) WB_TRY(wb_ensure_zpart(ctx, (size_t)wb_zgemm_tensor_items(ncol, ncol, count) * 8));
    // X ping-pongs between the caller's array and tmp2 (dense), so an iteration is two GEMMs with no
    // copy back.  Tensor path: the P = X^H X GEMM writes the step matrix T = a I - b P straight from
    // its accumulators together with ||P - I||_F^2 partials (no separate pass over P).
    struct Mat { cplx* p; int ld; long long stride; };
    Mat cur{dX, ld, stride}, other{ctx->d_tmp2, nrow, (long long)nrow * ncol};
    const long long ptotal = (long long)count * ncol * ncol;
    const int pblocks = (int)((ptotal + 255) / 256 < 4096 ? (ptotal + 255) / 256 : 4096);
    int it = 0;
    double err = 0.0;
    NsInterval iv{0.0, 0.0};
    bool scaled = false;      // interval known: scaled iteration
    for (; it < MAXIT; it++) {
        ZgemmArgs a{};
        a.m = ncol; a.n = ncol; a.k = nrow;
        a.A = cur.p; a.strideA = cur.stride; a.lda = cur.ld; a.transA = 1;
        a.B = cur.p; a.strideB = cur.stride; a.ldb = cur.ld; a.transB = 0;
        a.C = ctx->d_tmp1; a.strideC = (long long)ncol * ncol; a.ldc = ncol;
        a.alpha = 1.0; a.beta = 0.0; a.batch = count;
        ZgemmArgs b{};
        b.m = nrow; b.n = ncol; b.k = ncol;
        b.A = cur.p; b.strideA = cur.stride; b.lda = cur.ld; b.transA = 0;
        b.B = ctx->d_tmp1; b.strideB = (long long)ncol * ncol; b.ldb = ncol; b.transB = 0;
        b.C = other.p; b.strideC = other.stride; b.ldc = other.ld;
        b.alpha = 1.0; b.beta = 0.0; b.batch = count;
        WB_CUDA(ctx, cudaMemsetAsync(ctx->d_scal, 0, sizeof(double), ctx->stream));
        if (tensor && it == 0 && gram && gram->S) {
            err = gram->alpha * gram->alpha * gram->fro_max;
            if (!(err == err)) { ctx->err = "retract: NaN in s^H s"; return WB200_ERR_NOCONV; }
            iv.hi = sqrt(1.0 + err);
            iv.lo = 1.0 - 1e-9;
            scaled = true;
            const double cc = iv.c(), b3 = 0.5 * cc * cc * cc;
            ns_from_gram_kernel<<<pblocks, 256, 0, ctx->stream>>>(ncol, ptotal, 1.5 * cc - b3, b3 * gram->alpha * gram->alpha,
                                                                gram->S, ctx->d_tmp1);
            WB_LAUNCH_CHECK(ctx);
            iv.step(cc);
            WB_TRY(wb_zgemm(ctx, b));
        } else if (!tensor) {
            WB_TRY(wb_zgemm(ctx, a));   // P = X^H X
            ns_prepare_kernel<<<count, 256, 0, ctx->stream>>>(ncol, ctx->d_tmp1, ctx->d_scal, it == 0);
            WB_LAUNCH_CHECK(ctx);
            WB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_scal, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            WB_TRY(wb_zgemm(ctx, b));   // X T
            WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            err = ctx->h_pinned[0];
        } else if (it == 0) {
            // first pass: P itself and its distance from I; the host then knows the singular-value interval
            WB_TRY(wb_zgemm_tensor(ctx, a, 3, ctx->d_zpart));
            WB_TRY(wb_zgemm_tensor_err(ctx, ncol, ncol, count, ctx->d_zpart, ctx->d_scal));
            WB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_scal, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            err = ctx->h_pinned[0];
            if (err == err) {
                if (tangent_step || err < 0.75) {
                    iv.hi = sqrt(1.0 + err);
                    iv.lo = tangent_step ? fmax(1.0 - 1e-9, err < 1.0 ? sqrt(1.0 - err) : 0.0) : sqrt(1.0 - err);
                    scaled = true;
                    const double cc = iv.c();
                    ns_scale_kernel<<<pblocks, 256, 0, ctx->stream>>>(ncol, ptotal, 1.5 * cc, 0.5 * cc * cc * cc, ctx->d_tmp1);
                    WB_LAUNCH_CHECK(ctx);
                    iv.step(cc);
                } else {
                    // nothing known about the small singular values: Frobenius scaling, then the plain iteration
                    WB_CUDA(ctx, cudaMemsetAsync(ctx->d_scal, 0, sizeof(double), ctx->stream));
                    ns_prepare_kernel<<<count, 256, 0, ctx->stream>>>(ncol, ctx->d_tmp1, ctx->d_scal, 1);
                    WB_LAUNCH_CHECK(ctx);
                }
                WB_TRY(wb_zgemm(ctx, b));
            }
        } else {
            const double cc = scaled ? iv.c() : 1.0;
            ZgemmArgs a2 = a;
            a2.alpha = 1.5 * cc; a2.beta = 0.5 * cc * cc * cc;
            WB_TRY(wb_zgemm_tensor(ctx, a2, 2, ctx->d_zpart));   // T = a I - b P and the error partials
            WB_TRY(wb_zgemm_tensor_err(ctx, ncol, ncol, count, ctx->d_zpart, ctx->d_scal));
            WB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_scal, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            WB_TRY(wb_zgemm(ctx, b));   // X T (enqueued before the host knows err: no bubble)
            WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            err = ctx->h_pinned[0];
            if (scaled && err == err) {   // err belongs to the P of THIS iteration, i.e. to the interval before the step
                iv.tighten(err);
                iv.step(cc);
            }
        }
        if (!(err == err)) {
            ctx->err = "retract: NaN in X^H X";
            return WB200_ERR_NOCONV;
        }
        std::swap(cur, other);
        if (err < TOL) { it++; break; }
    }
    if (cur.p != dX) {
        const long long total = (long long)count * nrow * ncol;
        if (ld == nrow && stride == (long long)nrow * ncol) {
            WB_CUDA(ctx, cudaMemcpyAsync(dX, cur.p, sizeof(cplx) * total, cudaMemcpyDeviceToDevice, ctx->stream));
        } else {
            const int blocks = (int)((total + 255) / 256 < 8192 ? (total + 255) / 256 : 8192);
            copy_strided_kernel<<<blocks, 256, 0, ctx->stream>>>(nrow, ncol, ld, stride, total, cur.p, dX);
            WB_LAUNCH_CHECK(ctx);
        }
    }
    if (iters_out) *iters_out = it;
    {
        static const bool trace = getenv("WB200_TRACE") != nullptr;
        if (trace) fprintf(stderr, "[wb200] retract %d x %d x %d: %d iterations, tangent_step %d, scaled %d, final err %.2e\n",
                           nrow, ncol, count, it, (int)tangent_step, (int)scaled, err);
    }
    if (err >= TOL) {
        ctx->err = "retract: Newton-Schulz polar iteration did not converge (||X^H X - I||_F = " +
                   std::to_string(err) + ")";
        return WB200_ERR_NOCONV;
    }
    return WB200_OK;
}

int wb_retract_dev(wb200_ctx* ctx, cplx* dX, int nrow, int ncol, int count, int* iters, bool tangent_step,
                   const WbTangentGram* gram) {
    return retract_strided(ctx, dX, nrow, ncol, nrow, (long long)nrow * ncol, count, iters, tangent_step, gram);
}

// S_k = s_k^H s_k for a batch of tangent vectors and max_k ||S_k||_F (one synchronisation).  Returns false in
// `ok` when the shape is served by the fused small-shape kernels (which need no such help).
int wb_tangent_gram(wb200_ctx* ctx, const cplx* ds, int nrow, int ncol, int count, cplx* dS, double* fro_max, bool* ok) {
    *ok = false;
    if (wb_small_applicable(nrow, ncol) || (ctx->flags & WB200_FLAG_NO_TENSOR) || !wb_zgemm_tensor_applicable(ncol, ncol, nrow))
        return WB200_OK;
    WB_TRY(wb_ensure_zpart(ctx, (size_t)wb_zgemm_tensor_items(ncol, ncol, count) * 8));
    ZgemmArgs a{};
    a.m = ncol; a.n = ncol; a.k = nrow;
    a.A = ds; a.strideA = (long long)nrow * ncol; a.lda = nrow; a.transA = 1;
    a.B = ds; a.strideB = (long long)nrow * ncol; a.ldb = nrow; a.transB = 0;
    a.C = dS; a.strideC = (long long)ncol * ncol; a.ldc = ncol;
    a.alpha = 1.0; a.beta = 0.0; a.batch = count;
    WB_CUDA(ctx, cudaMemsetAsync(ctx->d_scal, 0, sizeof(double), ctx->stream));
    WB_TRY(wb_zgemm_tensor(ctx, a, 1, ctx->d_zpart));
    WB_TRY(wb_zgemm_tensor_err(ctx, ncol, ncol, count, ctx->d_zpart, ctx->d_scal));
    WB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_scal, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *fro_max = ctx->h_pinned[0];
    *ok = true;
    return WB200_OK;
}

// manifold ops on the packed XY array (product manifold per k)
int wb_retract_xy(wb200_ctx* ctx, cplx* dXY, bool tangent_step, long long st) {
    const int nb = ctx->nb, nw = ctx->nw;
    if (st <= 0) st = (long long)nw * nw + (long long)nb * nw;
    WB_TRY(retract_strided(ctx, dXY, nw, nw, nw, st, ctx->nk, nullptr, tangent_step));
    WB_TRY(retract_strided(ctx, dXY + (long long)nw * nw, nb, nw, nb, st, ctx->nk, nullptr, tangent_step));
    return WB200_OK;
}
int wb_project_xy(wb200_ctx* ctx, const cplx* dXY, cplx* dG, long long st) {
    const int nb = ctx->nb, nw = ctx->nw;
    if (st <= 0) st = (long long)nw * nw + (long long)nb * nw;
    WB_TRY(project_strided(ctx, dXY, dG, nw, nw, nw, st, ctx->nk));
    WB_TRY(project_strided(ctx, dXY + (long long)nw * nw, dG + (long long)nw * nw, nb, nw, nb, st,
                           ctx->nk));
    return WB200_OK;
}

// ---------------------------------------------------------------------------------------------
// (X, Y) chain
// ---------------------------------------------------------------------------------------------
// Small shapes (nb <= 64): one CTA per k-point keeps X_k, Y_k (and G_k) in shared memory and does the
// whole step in one launch.  The batched 32 x 32-tile GEMM kernel below leaves most of its threads idle
// on 18 x 9 matrices and needs three launches per evaluation (0.4 of the 0.52 ms of an (X, Y)
// evaluation at the Cu-like shape were spent there).
#define WB_XY_SMALL_MAX 64

__global__ void __launch_bounds__(128) xy_to_u_small_kernel(int nb, int nw, long long st,
                                                            const cplx* __restrict__ XY, cplx* __restrict__ U) {
    extern __shared__ cplx sxy[];
    cplx* sX = sxy;                 // [nw][nw] col-major
    cplx* sY = sxy + nw * nw;       // [nw][nb + 1] col-major, padded
    const cplx* g = XY + (long long)blockIdx.x * st;
    for (int e = threadIdx.x; e < nw * nw; e += blockDim.x) sX[e] = g[e];
    for (int e = threadIdx.x; e < nb * nw; e += blockDim.x) sY[(e / nb) * (nb + 1) + (e % nb)] = g[nw * nw + e];
    __syncthreads();
    cplx* u = U + (long long)blockIdx.x * nb * nw;
    for (int e = threadIdx.x; e < nb * nw; e += blockDim.x) {
        const int m = e % nb, n = e / nb;
        cplx acc = make_double2(0.0, 0.0);
        for (int l = 0; l < nw; l++) cfma(acc, sY[l * (nb + 1) + m], sX[n * nw + l]);   // U = Y X
        u[e] = acc;
    }
}

// G_X = Y^H G_U, G_Y = G_U X^H, then zero G_Y[frozen rows, :] and G_Y[:, 0:n_frozen)   (disentangle.jl:481-501)
__global__ void __launch_bounds__(128) gu_to_gxy_small_kernel(int nb, int nw, long long st,
                                                              const cplx* __restrict__ XY,
                                                              const cplx* __restrict__ GU,
                                                              const uint8_t* __restrict__ frozen,
                                                              const int32_t* __restrict__ nfrozen,
                                                              cplx* __restrict__ GXY) {
    extern __shared__ cplx sxy[];
    cplx* sX = sxy;                       // [nw][nw + 1]
    cplx* sY = sX + nw * (nw + 1);        // [nw][nb + 1]
    cplx* sG = sY + nw * (nb + 1);        // [nw][nb + 1]
    const int k = blockIdx.x;
    const cplx* g = XY + (long long)k * st;
    const cplx* gu = GU + (long long)k * nb * nw;
    for (int e = threadIdx.x; e < nw * nw; e += blockDim.x) sX[(e / nw) * (nw + 1) + (e % nw)] = g[e];
    for (int e = threadIdx.x; e < nb * nw; e += blockDim.x) {
        sY[(e / nb) * (nb + 1) + (e % nb)] = g[nw * nw + e];
        sG[(e / nb) * (nb + 1) + (e % nb)] = gu[e];
    }
    __syncthreads();
    cplx* out = GXY + (long long)k * st;
    for (int e = threadIdx.x; e < nw * nw; e += blockDim.x) {   // G_X[i][n] = sum_m conj(Y[m][i]) G[m][n]
        const int i = e % nw, n = e / nw;
        cplx acc = make_double2(0.0, 0.0);
        for (int m = 0; m < nb; m++) cfma_conj(acc, sY[i * (nb + 1) + m], sG[n * (nb + 1) + m]);
        out[e] = acc;
    }
    const int nf = frozen ? nfrozen[k] : 0;
    for (int e = threadIdx.x; e < nb * nw; e += blockDim.x) {   // G_Y[m][l] = sum_n G[m][n] conj(X[l][n])
        const int m = e % nb, l = e / nb;
        cplx acc = make_double2(0.0, 0.0);
        if (!(frozen && (frozen[(long long)k * nb + m] || l < nf)))
            for (int n = 0; n < nw; n++) {
                const cplx x = sX[n * (nw + 1) + l];
                cfma(acc, sG[n * (nb + 1) + m], cconj(x));
            }
        out[nw * nw + e] = acc;
    }
}

static int xy_small_attr(wb200_ctx* ctx) {
    if (!(ctx->attr_mask & 8u)) {
        WB_CUDA(ctx, cudaFuncSetAttribute(xy_to_u_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        WB_CUDA(ctx, cudaFuncSetAttribute(gu_to_gxy_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        ctx->attr_mask |= 8u;
    }
    return WB200_OK;
}

int wb_xy_to_u(wb200_ctx* ctx, const cplx* dXY, cplx* dU, long long st) {
    const int nb = ctx->nb, nw = ctx->nw;
    if (st <= 0) st = (long long)nw * nw + (long long)nb * nw;
    if (nb <= WB_XY_SMALL_MAX) {
        WB_TRY(xy_small_attr(ctx));
        const size_t smem = sizeof(cplx) * ((size_t)nw * nw + (size_t)nw * (nb + 1));
        xy_to_u_small_kernel<<<ctx->nk, 128, smem, ctx->stream>>>(nb, nw, st, dXY, dU);
        WB_LAUNCH_CHECK(ctx);
        return WB200_OK;
    }
    ZgemmArgs a{};
    a.m = nb; a.n = nw; a.k = nw;
    a.A = dXY + (long long)nw * nw; a.strideA = st; a.lda = nb; a.transA = 0;   // Y_k
    a.B = dXY; a.strideB = st; a.ldb = nw; a.transB = 0;                        // X_k
    a.C = dU; a.strideC = (long long)nb * nw; a.ldc = nb;
    a.alpha = 1.0; a.beta = 0.0; a.batch = ctx->nk;
    return wb_zgemm(ctx, a);
}

// zero G_Y[frozen rows, :] and G_Y[:, 0:n_frozen)   (disentangle.jl:495-496)
__global__ void mask_gy_kernel(int nb, int nw, long long st, const uint8_t* __restrict__ frozen,
                               const int32_t* __restrict__ nfrozen, cplx* __restrict__ GXY) {
    const int k = blockIdx.x;
    cplx* gy = GXY + (long long)k * st + (long long)nw * nw;
    const int nf = nfrozen[k];
    for (int e = threadIdx.x; e < nb * nw; e += blockDim.x) {
        const int m = e % nb, n = e / nb;
        if (frozen[(long long)k * nb + m] || n < nf) gy[e] = make_double2(0.0, 0.0);
    }
}

int wb_gu_to_gxy(wb200_ctx* ctx, const cplx* dXY, const cplx* dGU, cplx* dGXY, long long st) {
    const int nb = ctx->nb, nw = ctx->nw;
    if (st <= 0) st = (long long)nw * nw + (long long)nb * nw;
    if (nb <= WB_XY_SMALL_MAX) {
        WB_TRY(xy_small_attr(ctx));
        const size_t smem = sizeof(cplx) * ((size_t)nw * (nw + 1) + 2 * (size_t)nw * (nb + 1));
        gu_to_gxy_small_kernel<<<ctx->nk, 128, smem, ctx->stream>>>(nb, nw, st, dXY, dGU, ctx->d_frozen,
                                                                    ctx->d_nfrozen, dGXY);
        WB_LAUNCH_CHECK(ctx);
        return WB200_OK;
    }
    ZgemmArgs a{};  // G_X = Y^H G_U
    a.m = nw; a.n = nw; a.k = nb;
    a.A = dXY + (long long)nw * nw; a.strideA = st; a.lda = nb; a.transA = 1;
    a.B = dGU; a.strideB = (long long)nb * nw; a.ldb = nb; a.transB = 0;
    a.C = dGXY; a.strideC = st; a.ldc = nw;
    a.alpha = 1.0; a.beta = 0.0; a.batch = ctx->nk;
    WB_TRY(wb_zgemm(ctx, a));
    ZgemmArgs b{};  // G_Y = G_U X^H
    b.m = nb; b.n = nw; b.k = nw;
    b.A = dGU; b.strideA = (long long)nb * nw; b.lda = nb; b.transA = 0;
    b.B = dXY; b.strideB = st; b.ldb = nw; b.transB = 1;
    b.C = dGXY + (long long)nw * nw; b.strideC = st; b.ldc = nb;
    b.alpha = 1.0; b.beta = 0.0; b.batch = ctx->nk;
    WB_TRY(wb_zgemm(ctx, b));
    if (ctx->d_frozen) {
        mask_gy_kernel<<<ctx->nk, 256, 0, ctx->stream>>>(nb, nw, st, ctx->d_frozen, ctx->d_nfrozen,
                                                          dGXY);
        WB_LAUNCH_CHECK(ctx);
    }
    return WB200_OK;
}
