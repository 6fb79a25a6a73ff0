// setup.cu — the initial point of the disentanglement on the device.
//
// Replaces (reference tree, src/localization/disentangle.jl):
//   orthonorm_freeze(U, frozen)   :223-281   Lowdin of the frozen rows Uf, projection of Uf out of the other
//                                            rows Ur, partial isometry of Ur (SVD, singular values -> 1 / 0)
//   U_to_X_Y(U, frozen)           :369-402   Y_k = [I on the frozen rows | an orthonormal basis of range(Ur)],
//                                            X_k = Lowdin(Y_k^H orthonorm_freeze(U_k))
// The reference does this with an SVD, a Hermitian eigendecomposition and two Lowdin steps per k-point on
// the host.  Here ONE kernel (a CTA per k-point) completes the rows of Uf to an orthonormal basis
// [Qf | C] of C^nw (modified Gram-Schmidt with re-orthogonalisation; the complement C is grown from the
// unit vector with the largest remaining residual, so every pivot has norm^2 >= 1/nw) and writes
//     X0_k = [ Uf ; C^H ]                      (nw x nw)
//     Z_k  = [ e_frozen | Ur C on the free rows ]   (nb x nw)
// whose polar factors are exactly the reference's X_k and Y_k up to the basis chosen inside range(Ur):
//   * Ur (I - Uf^H Uf) = (Ur C) C^H, so the partial isometry of the projected Ur is polar(Ur C) C^H and
//     range(Ur) is spanned by polar(Ur C) — the free block of polar(Z_k), whose frozen block is untouched
//     because the two column groups of Z_k live on disjoint rows;
//   * X0 X0^H = blockdiag(Uf Uf^H, I), so polar(X0) = [Lowdin(Uf) ; C^H] = Y_k^H orthonorm_freeze(U_k).
// Both polar factors come from the batched Stiefel retraction the optimiser already uses (wb_retract_xy).
// The reference takes the eigenvectors of the rank-(nw - nf) projector Ur Ur^H for the free block of Y, a
// basis that is only defined up to a unitary mixing inside a degenerate eigenspace (LAPACK-dependent);
// U = Y X = orthonorm_freeze(U), the frozen structure of Y and the unitarity of X are what the optimiser
// needs and what the tests compare.
#include "common.cuh"

#include <cmath>

namespace {

// err[0] is set to 1 (too many frozen bands) or 2 (frozen rows linearly dependent)
__global__ void __launch_bounds__(256) u_to_xy_init_kernel(int nb, int nw, long long st, const cplx* __restrict__ U,
                                                           const uint8_t* __restrict__ frozen, cplx* __restrict__ Bws,
                                                           cplx* __restrict__ XY, int* __restrict__ err) {
    extern __shared__ unsigned char smem_raw[];
    cplx* v = reinterpret_cast<cplx*>(smem_raw);     // [nw]
    cplx* c = v + nw;                                // [nw]
    double* rho = reinterpret_cast<double*>(c + nw); // [nw]
    int* fro = reinterpret_cast<int*>(rho + nw);     // [nb] frozen rows first, then the free rows
    __shared__ double s_red[8];
    __shared__ int s_pick[8];
    __shared__ int s_nf, s_choice;
    __shared__ double s_norm;

    const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const cplx* u = U + (long long)k * nb * nw;     // u[m + n nb]
    cplx* B = Bws + (long long)k * nw * nw;         // B[n + j nw]: column j of the basis
    if (tid == 0) {
        int nf = 0;
        for (int m = 0; m < nb; m++) if (frozen && frozen[(long long)k * nb + m]) fro[nf++] = m;
        int q = nf;
        for (int m = 0; m < nb; m++) if (!(frozen && frozen[(long long)k * nb + m])) fro[q++] = m;
        s_nf = nf;
    }
    for (int n = tid; n < nw; n += blockDim.x) rho[n] = 1.0;
    __syncthreads();
    const int nf = s_nf;
    if (nf > nw) {
        if (tid == 0) atomicMax(err, 1);
        return;
    }
    for (int j = 0; j < nw; j++) {
        if (j < nf) {   // column j of Uf^H
            for (int n = tid; n < nw; n += blockDim.x) v[n] = cconj(u[fro[j] + (long long)n * nb]);
        } else {        // the unit vector least represented so far (ties: smallest index)
            double best = -1.0; int bi = 0;
            for (int n = tid; n < nw; n += blockDim.x) if (rho[n] > best) { best = rho[n]; bi = n; }
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (lane == 0) { s_red[warp] = best; s_pick[warp] = bi; }
            __syncthreads();
            if (tid == 0) {
                double bb = s_red[0]; int ii = s_pick[0];
                for (int w = 1; w < 8; w++)
                    if (s_red[w] > bb || (s_red[w] == bb && s_pick[w] < ii)) { bb = s_red[w]; ii = s_pick[w]; }
                s_choice = ii;
            }
            __syncthreads();
            const int pick = s_choice;
            for (int n = tid; n < nw; n += blockDim.x) v[n] = make_double2(n == pick ? 1.0 : 0.0, 0.0);
        }
        __syncthreads();
        for (int pass = 0; pass < 2; pass++) {   // "twice is enough"
            for (int i = warp; i < j; i += 8) {
                cplx acc = make_double2(0.0, 0.0);
                for (int n = lane; n < nw; n += 32) cfma_conj(acc, B[n + (long long)i * nw], v[n]);
                acc.x = warp_sum(acc.x); acc.y = warp_sum(acc.y);
                if (lane == 0) c[i] = acc;
            }
            __syncthreads();
            for (int n = tid; n < nw; n += blockDim.x) {
                cplx acc = v[n];
                for (int i = 0; i < j; i++) {
                    const cplx b = B[n + (long long)i * nw], ci = c[i];
                    acc.x -= b.x * ci.x - b.y * ci.y;
                    acc.y -= b.x * ci.y + b.y * ci.x;
                }
                v[n] = acc;
            }
            __syncthreads();
        }
        double nrm = 0.0;
        for (int n = tid; n < nw; n += blockDim.x) nrm += v[n].x * v[n].x + v[n].y * v[n].y;
        nrm = warp_sum(nrm);
        if (lane == 0) s_red[warp] = nrm;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; w++) t += s_red[w];
            s_norm = sqrt(t);
        }
        __syncthreads();
        const double nv = s_norm;
        if (!(nv > 1e-10)) {   // the frozen rows of U are linearly dependent (orthonorm_lowdin would divide by ~0)
            if (tid == 0) atomicMax(err, 2);
            return;
        }
        const double inv = 1.0 / nv;
        for (int n = tid; n < nw; n += blockDim.x) {
            const cplx w = make_double2(v[n].x * inv, v[n].y * inv);
            B[n + (long long)j * nw] = w;
            rho[n] -= w.x * w.x + w.y * w.y;
        }
        __threadfence_block();
        __syncthreads();
    }
    // X0 = [Uf ; C^H], nw x nw column-major
    cplx* X0 = XY + (long long)k * st;
    for (int e = tid; e < nw * nw; e += blockDim.x) {
        const int r = e % nw, n = e / nw;
        X0[e] = r < nf ? u[fro[r] + (long long)n * nb] : cconj(B[n + (long long)r * nw]);
    }
    // Z = [e_frozen | Ur C], nb x nw column-major
    cplx* Z = X0 + (long long)nw * nw;
    for (int e = tid; e < nb * nw; e += blockDim.x) Z[e] = make_double2(0.0, 0.0);
    __syncthreads();
    for (int j = tid; j < nf; j += blockDim.x) Z[fro[j] + (long long)j * nb] = make_double2(1.0, 0.0);
    const int nfree = nb - nf, ncomp = nw - nf;
    for (int e = tid; e < nfree * ncomp; e += blockDim.x) {
        const int m = fro[nf + e % nfree], j = nf + e / nfree;
        cplx acc = make_double2(0.0, 0.0);
        for (int n = 0; n < nw; n++) cfma(acc, u[m + (long long)n * nb], B[n + (long long)j * nw]);
        Z[m + (long long)j * nb] = acc;
    }
}

}  // namespace

// XY <- X_Y_to_XY(U_to_X_Y(U, frozen)) for the ctx's k-points; dU [nk][nw][nb], dXY [nk][nw*nw + nb*nw], both device
int wb_u_to_xy_dev(wb200_ctx* ctx, const cplx* dU, cplx* dXY) {
    const int nb = ctx->nb, nw = ctx->nw, nk = ctx->nk;
    const long long st = (long long)nw * nw + (long long)nb * nw;
    WB_TRY(wb_ensure_tmp(ctx, (size_t)nk * nw * nw, 0));
    int* derr = reinterpret_cast<int*>(ctx->d_scal + 8);
    WB_CUDA(ctx, cudaMemsetAsync(derr, 0, sizeof(int), ctx->stream));
    const size_t smem = sizeof(cplx) * 2 * nw + sizeof(double) * nw + sizeof(int) * nb;
    u_to_xy_init_kernel<<<nk, 256, smem, ctx->stream>>>(nb, nw, st, dU, ctx->d_frozen, ctx->d_tmp1, dXY, derr);
    WB_LAUNCH_CHECK(ctx);
    int* herr = reinterpret_cast<int*>(ctx->h_pinned + 8);
    WB_CUDA(ctx, cudaMemcpyAsync(herr, derr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*herr == 1) { ctx->err = "U_to_X_Y: more frozen bands than Wannier functions at some k-point"; return WB200_ERR_ARG; }
    if (*herr == 2) { ctx->err = "U_to_X_Y: the frozen rows of U are linearly dependent at some k-point"; return WB200_ERR_SINGULAR; }
    const int s = wb_retract_xy(ctx, dXY);
    if (s == WB200_ERR_NOCONV)   // orthonorm_freeze's assertion on the singular values of Ur (disentangle.jl:265)
        ctx->err = "U_to_X_Y: the non-frozen rows of U do not span n_wann - n_frozen dimensions (" + ctx->err + ")";
    return s;
}

static int setup_host(wb200_ctx* ctx, const double* U, double* XY, double* U_out) {
    const size_t mat = (size_t)ctx->nb * ctx->nw, st = (size_t)ctx->nw * ctx->nw + mat;
    if (!ctx->d_XY) WB_CUDA(ctx, wb_dev_alloc(&ctx->d_XY, sizeof(cplx) * st * ctx->nk));
    WB_CUDA(ctx, cudaMemcpyAsync(ctx->d_U, U, sizeof(cplx) * mat * ctx->nk, cudaMemcpyHostToDevice, ctx->stream));
    WB_TRY(wb_u_to_xy_dev(ctx, ctx->d_U, ctx->d_XY));
    if (XY)
        WB_CUDA(ctx, cudaMemcpyAsync(XY, ctx->d_XY, sizeof(cplx) * st * ctx->nk, cudaMemcpyDeviceToHost, ctx->stream));
    if (U_out) {
        WB_TRY(wb_xy_to_u(ctx, ctx->d_XY, ctx->d_U));
        WB_CUDA(ctx, cudaMemcpyAsync(U_out, ctx->d_U, sizeof(cplx) * mat * ctx->nk, cudaMemcpyDeviceToHost, ctx->stream));
    }
    WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return WB200_OK;
}

extern "C" int wb200_u_to_xy(wb200_ctx* ctx, const double* U, double* XY) {
    if (!ctx || !U || !XY) return WB200_ERR_ARG;
    cudaSetDevice(ctx->device);
    return setup_host(ctx, U, XY, nullptr);
}

extern "C" int wb200_orthonorm_freeze(wb200_ctx* ctx, double* U_inout) {
    if (!ctx || !U_inout) return WB200_ERR_ARG;
    cudaSetDevice(ctx->device);
    return setup_host(ctx, U_inout, nullptr, U_inout);
}

// A context without a stencil or overlaps: the per-k Stiefel operations and this file's setup step only
// (U_to_X_Y(U, frozen) has no model in the reference either).  Every evaluation entry point rejects it.
extern "C" int wb200_create_stiefel(wb200_ctx** out, int device, int nb, int nw, int nk) {
    if (!out) return WB200_ERR_ARG;
    *out = nullptr;
    if (nb <= 0 || nw <= 0 || nw > nb || nk <= 0) return WB200_ERR_ARG;
    int ndev = 0, cc_major = 0, cc_minor = 0, num_sms = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev ||
        wb_device_info(device, &cc_major, &cc_minor, &num_sms) != WB200_OK || cc_major != 10)
        return WB200_ERR_NODEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return WB200_ERR_CUDA;
    wb200_ctx* ctx = new wb200_ctx();
    ctx->device = device;
    ctx->num_sms = num_sms;
    ctx->use_graphs = 0;
    ctx->nb = nb; ctx->nw = nw; ctx->nk_total = nk; ctx->k_begin = 0; ctx->nk = nk; ctx->nn = 0;
    cudaError_t e = wb_stream_acquire(&ctx->own_stream);
    ctx->stream = ctx->own_stream;
    if (e == cudaSuccess) e = wb_dev_alloc(&ctx->d_U, sizeof(cplx) * (size_t)nk * nb * nw);
    if (e == cudaSuccess) e = wb_dev_alloc(&ctx->d_scal, sizeof(double) * 2048);
    if (e == cudaSuccess) e = wb_dev_alloc(&ctx->d_res, sizeof(double) * ctx->nres());
    if (e == cudaSuccess) e = wb_pinned_alloc(&ctx->h_pinned, sizeof(double) * (ctx->nres() + 64));
    if (e != cudaSuccess) { wb200_destroy(ctx); return WB200_ERR_CUDA; }
    *out = ctx;
    return WB200_OK;
}
