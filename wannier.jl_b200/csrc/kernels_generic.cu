// kernels_generic.cu — CUDA-core FP64 kernels valid for ANY (nb, nw): batched complex GEMM, the
// diagonal / spread reductions and the gradient assembly.  The tensor (DMMA) rotation kernel in
// rotate_dmma.cu replaces wb_rotate_generic for large aligned shapes; everything downstream of
// the rotation (pair_reduce -> finalize -> grad) is shared.
//
// Reference functions restated by these kernels (paths under the reference tree):
//   compute_MU_UtMU!  src/spread.jl:164-195      -> wb_rotate_generic, wb_form_N
//   omega! / center!  src/spread.jl:254-360, 565-594 -> pair_reduce_kernel, finalize_kernel
//   omega_grad!       src/spread.jl:398-481      -> grad_kernel
#include "common.cuh"

// ------------------------------------------------------------------------------------------
// generic batched ZGEMM (shared-memory tiled, register blocked)
// ------------------------------------------------------------------------------------------
template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) zgemm_kernel(ZgemmArgs a) {
    constexpr int TX = BM / TM, TY = BN / TN, NT = TX * TY;
    __shared__ cplx As[BK][BM];
    __shared__ cplx Bs[BK][BN];
    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int bz = blockIdx.z;
    const cplx* A = a.A + (long long)(a.idxA ? a.idxA[bz] : bz) * a.strideA;
    const cplx* B = a.B + (long long)(a.idxB ? a.idxB[bz] : bz) * a.strideB;
    cplx* C = a.C + (long long)bz * a.strideC;
    const int row0 = blockIdx.x * BM, col0 = blockIdx.y * BN;

    cplx acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = make_double2(0.0, 0.0);

    for (int k0 = 0; k0 < a.k; k0 += BK) {
        for (int e = tid; e < BM * BK; e += NT) {
            int r, kk;
            if (!a.transA) { r = e % BM; kk = e / BM; } else { kk = e % BK; r = e / BK; }
            const int gr = row0 + r, gk = k0 + kk;
            cplx v = make_double2(0.0, 0.0);
            if (gr < a.m && gk < a.k)
                v = a.transA ? cconj(A[gk + (long long)gr * a.lda]) : A[gr + (long long)gk * a.lda];
            As[kk][r] = v;
        }
        for (int e = tid; e < BN * BK; e += NT) {
            int c, kk;
            if (!a.transB) { kk = e % BK; c = e / BK; } else { c = e % BN; kk = e / BN; }
            const int gc = col0 + c, gk = k0 + kk;
            cplx v = make_double2(0.0, 0.0);
            if (gc < a.n && gk < a.k)
                v = a.transB ? cconj(B[gc + (long long)gk * a.ldb]) : B[gk + (long long)gc * a.ldb];
            Bs[kk][c] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            cplx ar[TM], br[TN];
#pragma unroll
            for (int i = 0; i < TM; i++) ar[i] = As[kk][tx + i * TX];
#pragma unroll
            for (int j = 0; j < TN; j++) br[j] = Bs[kk][ty + j * TY];
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) cfma(acc[i][j], ar[i], br[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < TN; j++) {
        const int gc = col0 + ty + j * TY;
        if (gc >= a.n) continue;
#pragma unroll
        for (int i = 0; i < TM; i++) {
            const int gr = row0 + tx + i * TX;
            if (gr >= a.m) continue;
            cplx* p = C + gr + (long long)gc * a.ldc;
            cplx v = make_double2(a.alpha * acc[i][j].x, a.alpha * acc[i][j].y);
            if (a.beta != 0.0) {
                const cplx o = *p;
                v.x = fma(a.beta, o.x, v.x);
                v.y = fma(a.beta, o.y, v.y);
            }
            *p = v;
        }
    }
}

int wb_zgemm_batched(wb200_ctx* ctx, const ZgemmArgs& a) {
    if (a.batch <= 0 || a.m <= 0 || a.n <= 0) return WB200_OK;
    // blockIdx.z is limited to 65535: split the batch
    const int ZMAX = 65535;
    for (int b0 = 0; b0 < a.batch; b0 += ZMAX) {
        ZgemmArgs s = a;
        s.batch = (a.batch - b0 < ZMAX) ? a.batch - b0 : ZMAX;
        if (a.idxA) s.idxA = a.idxA + b0; else s.A = a.A + (long long)b0 * a.strideA;
        if (a.idxB) s.idxB = a.idxB + b0; else s.B = a.B + (long long)b0 * a.strideB;
        s.C = a.C + (long long)b0 * a.strideC;
        if (a.m <= 16 && a.n <= 16) {
            dim3 g((a.m + 15) / 16, (a.n + 15) / 16, s.batch);
            zgemm_kernel<16, 16, 8, 2, 2><<<g, 64, 0, ctx->stream>>>(s);
        } else if (a.m <= 32 && a.n <= 32) {
            dim3 g((a.m + 31) / 32, (a.n + 31) / 32, s.batch);
            zgemm_kernel<32, 32, 8, 2, 2><<<g, 256, 0, ctx->stream>>>(s);
        } else {
            dim3 g((a.m + 63) / 64, (a.n + 63) / 64, s.batch);
            zgemm_kernel<64, 64, 8, 4, 4><<<g, 256, 0, ctx->stream>>>(s);
        }
        WB_LAUNCH_CHECK(ctx);
    }
    return WB200_OK;
}

// ------------------------------------------------------------------------------------------
// diag N_nn = sum_m conj(U_k[m,n]) MU_kb[m,n]   and   ||MU_kb||_F^2      (one block per pair)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) diag_kernel(int nb, int nw, const cplx* __restrict__ U,
                                                   const int32_t* __restrict__ kself,
                                                   const cplx* __restrict__ MU,
                                                   cplx* __restrict__ dN, double* __restrict__ fro) {
    const int pair = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const cplx* Uk = U + (long long)kself[pair] * nb * nw;
    const cplx* mu = MU + (long long)pair * nb * nw;
    double f = 0.0;
    for (int n = warp; n < nw; n += nwarps) {
        cplx d = make_double2(0.0, 0.0);
        for (int m = lane; m < nb; m += 32) {
            const cplx u = Uk[(long long)n * nb + m], x = mu[(long long)n * nb + m];
            cfma_conj(d, u, x);
            f = fma(x.x, x.x, f);
            f = fma(x.y, x.y, f);
        }
        d.x = warp_sum(d.x);
        d.y = warp_sum(d.y);
        if (lane == 0) dN[(long long)pair * nw + n] = d;
    }
    __shared__ double sf[8];
    f = warp_sum(f);
    if (lane == 0) sf[warp] = f;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < nwarps; w++) t += sf[w];
        fro[pair] = t;
    }
}

// ||N_kb||_F^2 from a dense N (one block per pair)
__global__ void __launch_bounds__(256) fro_kernel(int len, const cplx* __restrict__ N,
                                                  double* __restrict__ fro) {
    const int pair = blockIdx.x;
    const cplx* p = N + (long long)pair * len;
    double f = 0.0;
    for (int e = threadIdx.x; e < len; e += blockDim.x) {
        const cplx x = p[e];
        f = fma(x.x, x.x, f);
        f = fma(x.y, x.y, f);
    }
    __shared__ double sf[8];
    f = warp_sum(f);
    if ((threadIdx.x & 31) == 0) sf[threadIdx.x >> 5] = f;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sf[w];
        fro[pair] = t;
    }
}

int wb_ensure_zpart(wb200_ctx* ctx, size_t elems) {
    if (elems <= ctx->zpart_elems) return WB200_OK;
    if (ctx->d_zpart) wb_dev_free(ctx->d_zpart);
    ctx->d_zpart = nullptr; ctx->zpart_elems = 0;
    WB_CUDA(ctx, wb_dev_alloc(&ctx->d_zpart, sizeof(double) * elems));
    ctx->zpart_elems = elems;
    return WB200_OK;
}

int wb_zgemm(wb200_ctx* ctx, const ZgemmArgs& a) {
    if (!(ctx->flags & WB200_FLAG_NO_TENSOR) && !(a.transA && a.transB) && wb_zgemm_tensor_applicable(a.m, a.n, a.k))
        return wb_zgemm_tensor(ctx, a, 0, nullptr);
    return wb_zgemm_batched(ctx, a);
}

int wb_rotate_generic(wb200_ctx* ctx, const cplx* dU, int ka, int kb) {
    const int nb = ctx->nb, nw = ctx->nw;
    const long long p0 = (long long)ka * ctx->nn;
    const int pairs = (kb - ka) * ctx->nn;
    ZgemmArgs a{};
    a.m = nb; a.n = nw; a.k = nb;
    a.A = ctx->d_M + p0 * nb * nb; a.strideA = (long long)nb * nb; a.lda = nb; a.idxA = nullptr;
    a.B = dU; a.strideB = (long long)nb * nw; a.ldb = nb; a.idxB = ctx->d_kpb + p0;
    a.C = ctx->d_MU + p0 * nb * nw; a.strideC = (long long)nb * nw; a.ldc = nb;
    a.alpha = 1.0; a.beta = 0.0; a.batch = pairs; a.transA = 0; a.transB = 0;
    ProfScope ps(ctx, 1);
    ctx->fro_src = nullptr;
    ctx->dn_src = nullptr;
    WB_TRY(wb_zgemm(ctx, a));     // nb > 256 lands here: tensor pipe from the plain layout of M
    diag_kernel<<<pairs, 256, 0, ctx->stream>>>(nb, nw, dU, ctx->d_kself + p0, ctx->d_MU + p0 * nb * nw,
                                                 ctx->d_dN + p0 * nw, ctx->d_fro + p0);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}

int wb_ensure_N(wb200_ctx* ctx) {
    if (ctx->d_N) return WB200_OK;
    WB_CUDA(ctx, wb_dev_alloc(&ctx->d_N, sizeof(cplx) * (size_t)ctx->nk * ctx->nn * ctx->nw * ctx->nw));
    return WB200_OK;
}

// N_kb = U_k^H MU_kb for all pairs of the slab (second GEMM of compute_MU_UtMU!, spread.jl:174),
// and fro[pair] = ||N_kb||_F^2.  Requires MU from a rotation kernel.
int wb_form_N(wb200_ctx* ctx, const cplx* dU) {
    const int nb = ctx->nb, nw = ctx->nw, pairs = ctx->nk * ctx->nn;
    WB_TRY(wb_ensure_N(ctx));
    ZgemmArgs a{};
    a.m = nw; a.n = nw; a.k = nb;
    a.A = dU; a.strideA = (long long)nb * nw; a.lda = nb; a.idxA = ctx->d_kself; a.transA = 1;
    a.B = ctx->d_MU; a.strideB = (long long)nb * nw; a.ldb = nb; a.idxB = nullptr; a.transB = 0;
    a.C = ctx->d_N; a.strideC = (long long)nw * nw; a.ldc = nw;
    a.alpha = 1.0; a.beta = 0.0; a.batch = pairs;
    ctx->fro_src = nullptr;   // from here on ||N||_F^2 of the exact N
    if (!(ctx->flags & WB200_FLAG_NO_TENSOR) && wb_zgemm_tensor_applicable(a.m, a.n, a.k)) {
        // tensor pipe, ||N||_F^2 from the accumulators (per-warp partials summed in fixed order)
        WB_TRY(wb_ensure_zpart(ctx, (size_t)wb_zgemm_tensor_items(a.m, a.n, pairs) * 8));
        WB_TRY(wb_zgemm_tensor(ctx, a, 1, ctx->d_zpart));
        return wb_zgemm_tensor_fro(ctx, a.m, a.n, pairs, ctx->d_zpart, ctx->d_fro);
    }
    WB_TRY(wb_zgemm_batched(ctx, a));
    fro_kernel<<<pairs, 256, 0, ctx->stream>>>(nw * nw, ctx->d_N, ctx->d_fro);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}

// ------------------------------------------------------------------------------------------
// spread partial sums per k (omega! inner loops, spread.jl:284-322; center!, 565-594)
//   part[j][k], j in [0, 4nw+2):  3n+a: sum_b w_b b_a phi   3nw+n: sum_b w_b (1-|N|^2+phi^2)
//   4nw: sum_b w_b ||N||_F^2       4nw+1: sum_b w_b sum_n |N_nn|^2
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pair_reduce_kernel(int nw, int nn, int nk, int k_off,
                                                          cplx* __restrict__ dN,
                                                          const cplx* __restrict__ dNp, int ndn,
                                                          const double* __restrict__ fro, int nfro,
                                                          const double* __restrict__ bvec,
                                                          const double* __restrict__ wb,
                                                          double* __restrict__ part,
                                                          double* __restrict__ pmin) {
    const int k = blockIdx.x + k_off;
    // phase 1: one (b, n) element per thread — the atan2 calls (the expensive part) run in parallel
    // over all nn * nw diagonal elements of the k-point instead of serially over b
    extern __shared__ double sphi[];            // [nn][nw] phi, then [nn][nw] |N_nn|^2
    double* sa2 = sphi + nn * nw;
    for (int i = threadIdx.x; i < nn * nw; i += blockDim.x) {
        cplx z;
        if (dNp) {   // the tensor rotation leaves diag N as per-warp-row partials: sum them in fixed order, keep the result
            const int b = i / nw, n = i - b * nw;
            const cplx* pp = dNp + (((long long)k * nn + b) * ndn) * nw + n;
            z = make_double2(0.0, 0.0);
            for (int w = 0; w < ndn; w++) { z.x += pp[(long long)w * nw].x; z.y += pp[(long long)w * nw].y; }
            dN[(long long)k * nn * nw + i] = z;
        } else {
            z = dN[(long long)k * nn * nw + i];
        }
        sphi[i] = atan2(z.y, z.x);              // imaglog, src/utils/linalg.jl:15
        sa2[i] = z.x * z.x + z.y * z.y;
    }
    __syncthreads();
    // phase 2: per n, the sums over b in the reference's order
    double sd = 0.0, mn = 1e300;
    for (int n = threadIdx.x; n < nw; n += blockDim.x) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        for (int b = 0; b < nn; b++) {
            const double phi = sphi[b * nw + n];
            const double a2 = sa2[b * nw + n];
            const double w = wb[b];
            const double* bv = bvec + ((long long)k * nn + b) * 3;
            s0 += w * bv[0] * phi;
            s1 += w * bv[1] * phi;
            s2 += w * bv[2] * phi;
            s3 += w * (1.0 - a2 + phi * phi);
            sd += w * a2;
            mn = fmin(mn, a2);
        }
        part[(long long)(3 * n + 0) * nk + k] = s0;
        part[(long long)(3 * n + 1) * nk + k] = s1;
        part[(long long)(3 * n + 2) * nk + k] = s2;
        part[(long long)(3 * nw + n) * nk + k] = s3;
    }
    __shared__ double ssum[4], smin[4];
    sd = warp_sum(sd);
    mn = warp_min(mn);
    if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5] = sd; smin[threadIdx.x >> 5] = mn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0, m = 1e300;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { t += ssum[w]; m = fmin(m, smin[w]); }
        double sf = 0.0;
        for (int b = 0; b < nn; b++) {   // a rotation kernel may leave ||.||^2 of a pair in nfro partials
            double fp = 0.0;
            for (int i = 0; i < nfro; i++) fp += fro[((long long)k * nn + b) * nfro + i];
            sf += wb[b] * fp;
        }
        part[(long long)(4 * nw) * nk + k] = sf;
        part[(long long)(4 * nw + 1) * nk + k] = t;
        pmin[k] = sqrt(m);
    }
}

// sums[j] = sum_k part[j][k] (fixed-order tree: deterministic); block ns reduces the min.
__global__ void __launch_bounds__(256) sum_over_k_kernel(int nk, int ns, const double* __restrict__ part,
                                                         const double* __restrict__ pmin,
                                                         double* __restrict__ sums,
                                                         double* __restrict__ dmin) {
    const int j = blockIdx.x;
    __shared__ double sh[8];
    if (j < ns) {
        double s = 0.0;
        for (int k = threadIdx.x; k < nk; k += blockDim.x) s += part[(long long)j * nk + k];
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; w++) t += sh[w];
            sums[j] = t;
        }
    } else {
        double m = 1e300;
        for (int k = threadIdx.x; k < nk; k += blockDim.x) m = fmin(m, pmin[k]);
        m = warp_min(m);
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 1e300;
            for (int w = 0; w < 8; w++) t = fmin(t, sh[w]);
            *dmin = t;
        }
    }
}

int wb_pair_partial(wb200_ctx* ctx, int ka, int kb) {
    ProfScope ps(ctx, 2);
    const size_t smem = sizeof(double) * 2 * (size_t)ctx->nn * ctx->nw;
    if (smem > 48 * 1024) {
        if (!(ctx->attr_mask & 1u)) {
            WB_CUDA(ctx, cudaFuncSetAttribute(pair_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            ctx->attr_mask |= 1u;
        }
        if (smem > 200 * 1024) {
            ctx->err = "pair_reduce: nn * nw too large for the shared-memory staging";
            return WB200_ERR_ARG;
        }
    }
    pair_reduce_kernel<<<kb - ka, 128, smem, ctx->stream>>>(ctx->nw, ctx->nn, ctx->nk, ka, ctx->d_dN, ctx->dn_src, ctx->dn_parts,
                                                          ctx->fro_src ? ctx->fro_src : ctx->d_fro, ctx->fro_src ? ctx->fro_parts : 1,
                                                          ctx->d_bvec, ctx->d_wb,
                                                          ctx->d_part, ctx->d_pmin);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}

int wb_sum_over_k(wb200_ctx* ctx, double* dsums) {
    ProfScope ps(ctx, 2);
    sum_over_k_kernel<<<ctx->nsums() + 1, 256, 0, ctx->stream>>>(ctx->nk, ctx->nsums(), ctx->d_part,
                                                                  ctx->d_pmin, dsums, ctx->d_min);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}

int wb_pair_reduce(wb200_ctx* ctx, double* dsums) {
    WB_TRY(wb_pair_partial(ctx, 0, ctx->nk));
    return wb_sum_over_k(ctx, dsums);
}

// ------------------------------------------------------------------------------------------
// spread scalars from the (globally reduced) sums  (spread.jl:324-356, 246-252)
// res: {Omega, OmegaI, OmegaOD, OmegaD, OmegaTilde, f, Omega_c, 0, omega_n[nw], r[nw][3]}
// ------------------------------------------------------------------------------------------
__global__ void finalize_kernel(int nw, int nk_total, double wsum, const double* __restrict__ sums,
                                int has_penalty, double lambda, const double* __restrict__ r0,
                                double* __restrict__ res, int split_valid) {
    const double inv = 1.0 / (double)nk_total;
    for (int n = threadIdx.x; n < nw; n += blockDim.x) {
        const double rx = -sums[3 * n + 0] * inv, ry = -sums[3 * n + 1] * inv,
                     rz = -sums[3 * n + 2] * inv;
        const double r2 = sums[3 * nw + n] * inv;
        res[8 + n] = r2 - (rx * rx + ry * ry + rz * rz);
        res[8 + nw + 3 * n + 0] = rx;
        res[8 + nw + 3 * n + 1] = ry;
        res[8 + nw + 3 * n + 2] = rz;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double Om = 0.0, Oc = 0.0;
        for (int n = 0; n < nw; n++) {
            Om += res[8 + n];
            if (has_penalty) {
                const double dx = res[8 + nw + 3 * n + 0] - r0[3 * n + 0];
                const double dy = res[8 + nw + 3 * n + 1] - r0[3 * n + 1];
                const double dz = res[8 + nw + 3 * n + 2] - r0[3 * n + 2];
                Oc += lambda * (dx * dx + dy * dy + dz * dz);
            }
        }
        const double sumF = sums[4 * nw], sumD = sums[4 * nw + 1];
        const double OI = ((double)nw * wsum * (double)nk_total - sumF) * inv;
        const double OOD = (sumF - sumD) * inv;
        const double Ot = Om - OI;
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        res[0] = Om;
        res[1] = split_valid ? OI : nan;
        res[2] = split_valid ? OOD : nan;
        res[3] = split_valid ? Ot - OOD : nan;
        res[4] = split_valid ? Ot : nan;
        res[5] = Om + Oc;
        res[6] = Oc;
        res[7] = 0.0;
    }
}

int wb_finalize(wb200_ctx* ctx, const double* dsums, double* dres) {
    finalize_kernel<<<1, 256, 0, ctx->stream>>>(ctx->nw, ctx->nk_total, ctx->wsum, dsums,
                                                ctx->has_penalty, ctx->lambda, ctx->d_r0, dres, ctx->split_valid);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}

// ------------------------------------------------------------------------------------------
// gradient assembly  G_k = (4/nk) sum_b w_b MU_kb diag(t_n - conj N_nn)   (spread.jl:425-478)
// grid (nk, ceil(nw/NT)); coefficients for the block's NT columns are built once in smem.
// ------------------------------------------------------------------------------------------
#define WB_GRAD_NT 32
__global__ void __launch_bounds__(256) grad_kernel(int nb, int nw, int nn, int nk_total, int k_off,
                                                   const cplx* __restrict__ MU,
                                                   const cplx* __restrict__ dN,
                                                   const double* __restrict__ bvec,
                                                   const double* __restrict__ wb,
                                                   const double* __restrict__ sums, int has_penalty,
                                                   double lambda, const double* __restrict__ r0,
                                                   cplx* __restrict__ G) {
    extern __shared__ cplx coef[];  // [nn][WB_GRAD_NT]
    const int k = blockIdx.x + k_off;
    const int n0 = blockIdx.y * WB_GRAD_NT;
    const int nt = min(WB_GRAD_NT, nw - n0);
    const double inv = 1.0 / (double)nk_total;
    for (int e = threadIdx.x; e < nn * nt; e += blockDim.x) {
        const int b = e / nt, nl = e % nt, n = n0 + nl;
        const cplx z = dN[((long long)k * nn + b) * nw + n];
        double rx = -sums[3 * n + 0] * inv, ry = -sums[3 * n + 1] * inv, rz = -sums[3 * n + 2] * inv;
        if (has_penalty) {  // center_penalty: r - lambda (r - r0[n]), spread.jl:227
            rx -= lambda * (rx - r0[3 * n + 0]);
            ry -= lambda * (ry - r0[3 * n + 1]);
            rz -= lambda * (rz - r0[3 * n + 2]);
        }
        const double* bv = bvec + ((long long)k * nn + b) * 3;
        const double q = atan2(z.y, z.x) + (rx * bv[0] + ry * bv[1] + rz * bv[2]);
        const double a2 = z.x * z.x + z.y * z.y;
        // t = -i q / z = q (-z.y - i z.x) / |z|^2 ;  c = t - conj(z)
        const double s = q / a2;
        const double cx = -s * z.y - z.x, cy = -s * z.x + z.y;
        const double f = 4.0 * wb[b] * inv;
        coef[b * WB_GRAD_NT + nl] = make_double2(f * cx, f * cy);
    }
    __syncthreads();
    const long long mat = (long long)nb * nw;
    const cplx* mu0 = MU + (long long)k * nn * mat + (long long)n0 * nb;
    cplx* g0 = G + (long long)k * mat + (long long)n0 * nb;
    for (int e = threadIdx.x; e < nt * nb; e += blockDim.x) {
        const int nl = e / nb;
        cplx acc = make_double2(0.0, 0.0);
        for (int b = 0; b < nn; b++) {
            const cplx x = mu0[(long long)b * mat + e];
            cfma(acc, x, coef[b * WB_GRAD_NT + nl]);
        }
        g0[e] = acc;
    }
}

int wb_grad(wb200_ctx* ctx, const double* dsums, cplx* dG, int ka, int kb) {
    ProfScope ps(ctx, 3);
    dim3 g(kb - ka, (ctx->nw + WB_GRAD_NT - 1) / WB_GRAD_NT);
    const size_t smem = sizeof(cplx) * (size_t)ctx->nn * WB_GRAD_NT;
    grad_kernel<<<g, 256, smem, ctx->stream>>>(ctx->nb, ctx->nw, ctx->nn, ctx->nk_total, ka, ctx->d_MU,
                                               ctx->d_dN, ctx->d_bvec, ctx->d_wb, dsums,
                                               ctx->has_penalty, ctx->lambda, ctx->d_r0, dG);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}

// copy the rows of `need` that another rank owns from its (peer-mapped) array into the local one
__global__ void __launch_bounds__(256) halo_fetch_kernel(long long mat, int k_begin, int k_end,
                                                         const int32_t* __restrict__ need,
                                                         const int32_t* __restrict__ owner,
                                                         const cplx* const* __restrict__ peerU,
                                                         cplx* __restrict__ U) {
    const long long kg = need[blockIdx.x];
    if (kg >= k_begin && kg < k_end) return;   // own row
    const cplx* src = peerU[owner[blockIdx.x]] + kg * mat;
    cplx* dst = U + kg * mat;
    for (long long e = threadIdx.x; e < mat; e += blockDim.x) dst[e] = src[e];
}

int wb_halo_fetch(wb200_ctx* ctx, cplx* dU) {
    ProfScope ps(ctx, 0);
    halo_fetch_kernel<<<ctx->n_need, 256, 0, ctx->stream>>>((long long)ctx->nb * ctx->nw, ctx->k_begin,
                                                            ctx->k_begin + ctx->nk, ctx->d_need,
                                                            ctx->d_need_owner, ctx->d_peerU, dU);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}

int wb_ensure_tmp(wb200_ctx* ctx, size_t e1, size_t e2) {
    if (e1 > ctx->tmp1_elems) {
        if (ctx->d_tmp1) wb_dev_free(ctx->d_tmp1);
        ctx->d_tmp1 = nullptr; ctx->tmp1_elems = 0;
        WB_CUDA(ctx, wb_dev_alloc(&ctx->d_tmp1, sizeof(cplx) * e1));
        ctx->tmp1_elems = e1;
    }
    if (e2 > ctx->tmp2_elems) {
        if (ctx->d_tmp2) wb_dev_free(ctx->d_tmp2);
        ctx->d_tmp2 = nullptr; ctx->tmp2_elems = 0;
        WB_CUDA(ctx, wb_dev_alloc(&ctx->d_tmp2, sizeof(cplx) * e2));
        ctx->tmp2_elems = e2;
    }
    return WB200_OK;
}
