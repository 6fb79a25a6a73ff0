// extras.cu — by-products of the rotation path (SURVEY 8f "next" rows): the same U_k^H M_kb U_{k+b}
// with a different epilogue.
//   omega_local                       src/spread.jl:522-548
//   get_fg!_rotate (opt_rotate)       src/localization/opt_rotate.jl:12-81   (U_k = W for every k)
//   position_op                       src/spread.jl:674-716
//   compute_berry_connection_kspace   src/spread.jl:747-793
#include "common.cuh"

int wb_fg_device(wb200_ctx* ctx, const cplx* dU, double* dres, cplx* dG, int exact_split);  // api.cu

// loc[k] = sum_b w_b sum_n (1 - |N_nn|^2 + imaglog(N_nn)^2)
__global__ void __launch_bounds__(128) omega_local_kernel(int nw, int nn, const cplx* __restrict__ dN,
                                                          const double* __restrict__ wb,
                                                          double* __restrict__ loc) {
    const int k = blockIdx.x;
    double s = 0.0;
    for (int e = threadIdx.x; e < nn * nw; e += blockDim.x) {
        const cplx z = dN[(long long)k * nn * nw + e];
        const double phi = atan2(z.y, z.x);
        s += wb[e / nw] * (1.0 - (z.x * z.x + z.y * z.y) + phi * phi);
    }
    __shared__ double sh[4];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) loc[k] = sh[0] + sh[1] + sh[2] + sh[3];
}

__global__ void broadcast_gauge_kernel(long long mat, long long total, const cplx* __restrict__ W,
                                       cplx* __restrict__ U) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x)
        U[e] = W[e % mat];
}

// out[e] = scale * sum_k in[k][e]   (complex scale; fixed order over k: deterministic)
__global__ void __launch_bounds__(256) ksum_kernel(long long per, int nk, cplx scale,
                                                   const cplx* __restrict__ in, cplx* __restrict__ out) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= per) return;
    cplx s = make_double2(0.0, 0.0);
    for (int k = 0; k < nk; k++) {
        const cplx v = in[(long long)k * per + e];
        s.x += v.x;
        s.y += v.y;
    }
    out[e] = cmul(scale, s);
}

// A_k[m, n][a] = sum_b w_b b_a f(N_kb[m, n]);  mode 0: i (N - I) with -imaglog on the diagonal,
// mode 1: i (N - I), mode 2: (N - I)  (position_op before its global factor).  Layout [k][n][m][3].
__global__ void __launch_bounds__(256) connection_kernel(int nw, int nn, int mode,
                                                         const cplx* __restrict__ N,
                                                         const double* __restrict__ bvec,
                                                         const double* __restrict__ wb,
                                                         cplx* __restrict__ A) {
    const int k = blockIdx.x;
    for (int e = threadIdx.x; e < nw * nw; e += blockDim.x) {
        const int m = e % nw, n = e / nw;
        cplx acc[3] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
        for (int b = 0; b < nn; b++) {
            const cplx z = N[((long long)k * nn + b) * nw * nw + e];
            cplx v;
            if (mode == 0 && m == n) v = make_double2(-atan2(z.y, z.x), 0.0);
            else {
                const double re = z.x - (m == n ? 1.0 : 0.0), im = z.y;
                v = (mode == 2) ? make_double2(re, im) : make_double2(-im, re);   // i (re + i im)
            }
            const double w = wb[b];
            const double* bv = bvec + ((long long)k * nn + b) * 3;
#pragma unroll
            for (int a = 0; a < 3; a++) {
                acc[a].x = fma(w * bv[a], v.x, acc[a].x);
                acc[a].y = fma(w * bv[a], v.y, acc[a].y);
            }
        }
        cplx* o = A + ((long long)k * nw * nw + e) * 3;
        o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2];
    }
}

// A[m, n][a] <- (A[m, n][a] + conj(A[n, m][a])) / 2   per k
__global__ void __launch_bounds__(256) hermitize3_kernel(int nw, cplx* __restrict__ A) {
    cplx* a = A + (long long)blockIdx.x * nw * nw * 3;
    for (int e = threadIdx.x; e < nw * nw; e += blockDim.x) {
        const int m = e % nw, n = e / nw;
        if (m > n) continue;
        for (int c = 0; c < 3; c++) {
            const cplx x = a[((long long)n * nw + m) * 3 + c], y = a[((long long)m * nw + n) * 3 + c];
            const cplx h = make_double2(0.5 * (x.x + y.x), 0.5 * (x.y - y.y));
            a[((long long)n * nw + m) * 3 + c] = h;
            a[((long long)m * nw + n) * 3 + c] = cconj(h);
        }
    }
}

static int single(wb200_ctx* ctx, const char* who) {
    if (ctx->nk != ctx->nk_total || ctx->k_begin != 0) {
        ctx->err = std::string(who) + ": single-slab ctx only";
        return WB200_ERR_ARG;
    }
    return WB200_OK;
}
static int rotate_all(wb200_ctx* ctx, const cplx* dU) {
    if (ctx->nn == 0) {
        ctx->err = "this context has no stencil / overlaps (wb200_create_stiefel): Stiefel and setup calls only";
        return WB200_ERR_ARG;
    }
    if (ctx->dmma) {
        WB_TRY(wb_pack_dmma(ctx, dU, 0, ctx->n_need));
        return wb_rotate_dmma(ctx, dU, 0, ctx->nk);
    }
    return wb_rotate_generic(ctx, dU, 0, ctx->nk);
}
static int upload(wb200_ctx* ctx, const double* U) {
    WB_CUDA(ctx, cudaMemcpyAsync(ctx->d_U, U, sizeof(cplx) * (size_t)ctx->nk_total * ctx->nb * ctx->nw,
                                 cudaMemcpyHostToDevice, ctx->stream));
    return WB200_OK;
}

extern "C" int wb200_omega_local(wb200_ctx* ctx, const double* U, double* loc) {
    if (!ctx || !U || !loc) return WB200_ERR_ARG;
    cudaSetDevice(ctx->device);
    WB_TRY(single(ctx, "wb200_omega_local"));
    WB_TRY(upload(ctx, U));
    WB_TRY(rotate_all(ctx, ctx->d_U));
    double* d_loc = ctx->d_pmin;   // [nk] scratch
    omega_local_kernel<<<ctx->nk, 128, 0, ctx->stream>>>(ctx->nw, ctx->nn, ctx->d_dN, ctx->d_wb, d_loc);
    WB_LAUNCH_CHECK(ctx);
    WB_CUDA(ctx, cudaMemcpyAsync(loc, d_loc, sizeof(double) * ctx->nk, cudaMemcpyDeviceToHost, ctx->stream));
    WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return WB200_OK;
}

// f and gradient w.r.t. one global rotation W (device pointers): U_k = W for all k, G_W = sum_k G_k
int wb_fg_rotate_device(wb200_ctx* ctx, const cplx* dW, double* dres, cplx* dGW) {
    const long long mat = (long long)ctx->nb * ctx->nw, total = mat * ctx->nk_total;
    broadcast_gauge_kernel<<<(int)((total + 255) / 256 < 2368 ? (total + 255) / 256 : 2368), 256, 0, ctx->stream>>>(
        mat, total, dW, ctx->d_U);
    WB_LAUNCH_CHECK(ctx);
    WB_TRY(wb_fg_device(ctx, ctx->d_U, dres, dGW ? ctx->d_G : nullptr, 0));
    if (dGW) {
        ksum_kernel<<<(int)((mat + 255) / 256), 256, 0, ctx->stream>>>(mat, ctx->nk, make_double2(1.0, 0.0),
                                                                       ctx->d_G, dGW);
        WB_LAUNCH_CHECK(ctx);
    }
    return WB200_OK;
}

extern "C" int wb200_fg_rotate(wb200_ctx* ctx, const double* W, double* f, double* G) {
    if (!ctx || !W) return WB200_ERR_ARG;
    cudaSetDevice(ctx->device);
    WB_TRY(single(ctx, "wb200_fg_rotate"));
    if (ctx->nb != ctx->nw) {
        ctx->err = "n_bands != n_wann, run instead disentanglement?";   // opt_rotate.jl:101
        return WB200_ERR_ARG;
    }
    const size_t mat = (size_t)ctx->nw * ctx->nw;
    WB_TRY(wb_ensure_tmp(ctx, 2 * mat, 0));
    cplx* dW = ctx->d_tmp1;
    cplx* dGW = ctx->d_tmp1 + mat;
    WB_CUDA(ctx, cudaMemcpyAsync(dW, W, sizeof(cplx) * mat, cudaMemcpyHostToDevice, ctx->stream));
    WB_TRY(wb_fg_rotate_device(ctx, dW, ctx->d_res, G ? dGW : nullptr));
    if (G) WB_CUDA(ctx, cudaMemcpyAsync(G, dGW, sizeof(cplx) * mat, cudaMemcpyDeviceToHost, ctx->stream));
    WB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_res, sizeof(double) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    WB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 8, ctx->d_min, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (f) *f = ctx->h_pinned[5];
    if (G && ctx->h_pinned[8] < 1e-10) {   // the guard g! keeps (opt_rotate.jl:60-62)
        ctx->err = "N^kb too small! (|N_nn| < 1e-10)";
        return WB200_ERR_SINGULAR;
    }
    return WB200_OK;
}

static int connection(wb200_ctx* ctx, const double* U, int mode, bool herm, cplx** dA) {
    WB_TRY(upload(ctx, U));
    WB_TRY(rotate_all(ctx, ctx->d_U));
    WB_TRY(wb_form_N(ctx, ctx->d_U));
    const size_t per = (size_t)ctx->nw * ctx->nw * 3;
    WB_TRY(wb_ensure_tmp(ctx, 0, per * ctx->nk + per));
    *dA = ctx->d_tmp2;
    connection_kernel<<<ctx->nk, 256, 0, ctx->stream>>>(ctx->nw, ctx->nn, mode, ctx->d_N, ctx->d_bvec,
                                                        ctx->d_wb, *dA);
    WB_LAUNCH_CHECK(ctx);
    if (herm) {
        hermitize3_kernel<<<ctx->nk, 256, 0, ctx->stream>>>(ctx->nw, *dA);
        WB_LAUNCH_CHECK(ctx);
    }
    return WB200_OK;
}

extern "C" int wb200_berry_connection(wb200_ctx* ctx, const double* U, double* A, int imlog_diag,
                                      int force_hermiticity) {
    if (!ctx || !U || !A) return WB200_ERR_ARG;
    cudaSetDevice(ctx->device);
    WB_TRY(single(ctx, "wb200_berry_connection"));
    cplx* dA = nullptr;
    WB_TRY(connection(ctx, U, imlog_diag ? 0 : 1, force_hermiticity != 0, &dA));
    WB_CUDA(ctx, cudaMemcpyAsync(A, dA, sizeof(cplx) * (size_t)ctx->nk * ctx->nw * ctx->nw * 3,
                                 cudaMemcpyDeviceToHost, ctx->stream));
    WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return WB200_OK;
}

extern "C" int wb200_position_op(wb200_ctx* ctx, const double* U, double* R) {
    if (!ctx || !U || !R) return WB200_ERR_ARG;
    cudaSetDevice(ctx->device);
    WB_TRY(single(ctx, "wb200_position_op"));
    cplx* dA = nullptr;
    WB_TRY(connection(ctx, U, 2, false, &dA));
    const long long per = (long long)ctx->nw * ctx->nw * 3;
    cplx* dR = dA + per * ctx->nk;
    // R = sum / (-i nk) = (i / nk) sum
    ksum_kernel<<<(int)((per + 255) / 256), 256, 0, ctx->stream>>>(per, ctx->nk,
                                                                   make_double2(0.0, 1.0 / ctx->nk_total), dA, dR);
    WB_LAUNCH_CHECK(ctx);
    WB_CUDA(ctx, cudaMemcpyAsync(R, dR, sizeof(cplx) * per, cudaMemcpyDeviceToHost, ctx->stream));
    WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return WB200_OK;
}
