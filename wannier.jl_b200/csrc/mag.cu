// mag.cu — MagModel co-optimisation: spin-up and spin-down models coupled by the overlap
// between up and down Wannier functions.
//
// Replaces (reference tree, src/localization/coopt.jl):
//   overlap_updn        :114-127   Mw = sum_k Uup_k^H Mud_k Udn_k,  |Mw|^2 / nk^2 element-wise
//   omega_updn          :146-152   n_wann - tr(overlap_updn)
//   overlap_updn_grad   :176-201   GUup_k = Mud_k Udn_k diag(conj d) / nk^2, GUdn_k = Mud_k^H Uup_k diag(d) / nk^2,
//                                  d = diag(Mw);  omega_updn_grad = -2 x that (:222-235)
//   omega(::MagModel)   :66-73     up.Omega + dn.Omega + lambda Omega_updn
//   get_fg!_disentangle(::MagModel, lambda)  :252-321  on XY[:, k] = [XYup_k; XYdn_k]
//   disentangle(::MagModel, lambda)          :339-420  L-BFGS on the product of the four Stiefel factors per k
//
// Two ordinary single-slab contexts (up, dn) evaluate their own spreads and gradients; this file adds the
// coupling: two batched GEMMs per evaluation (T_k = Mud_k Udn_k, S_k = Mud_k^H Uup_k, FP64 tensor pipe where
// the shape allows), a deterministic two-stage reduction of the nw diagonal elements d_n, and one pass
// that adds the coupling gradient into the two contexts' gradients.  Only diag(Mw) enters f and g; the
// full nw x nw matrix is formed for overlap_updn alone.
#include "common.cuh"

#include <cmath>

int wb_fg_device(wb200_ctx* ctx, const cplx* dU, double* dres, cplx* dG, int exact_split);  // api.cu

struct wb200_mag {
    wb200_ctx *up = nullptr, *dn = nullptr;
    int nb = 0, nw = 0, nk = 0;
    cplx* d_Mud = nullptr;     // [nk][nb*nb] col-major
    cplx* d_T = nullptr;       // [nk][nb*nw]  Mud_k Udn_k
    cplx* d_S = nullptr;       // [nk][nb*nw]  Mud_k^H Uup_k
    cplx* d_dpart = nullptr;   // [nk][nw]     per-k column dots <Uup_k[:, n], T_k[:, n]>
    cplx* d_d = nullptr;       // [nw]         diag(Mw)
    cplx* d_Mw = nullptr;      // [nk][nw*nw]  (lazy: overlap_updn only)
    cplx* d_XY = nullptr; cplx* d_GXY = nullptr;   // [nk][2 n_inner] staging of the host entry points (lazy)
    double* d_out = nullptr;   // [4]: Omega_updn, f_total, f_up, f_dn
    std::string err;
    size_t n_inner() const { return (size_t)nw * nw + (size_t)nb * nw; }
};

static thread_local std::string g_mag_create_error;

#define MAG_CUDA(m, call)                                                            \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            (m)->err = std::string(#call) + ": " + cudaGetErrorString(e_);           \
            return WB200_ERR_CUDA;                                                   \
        }                                                                            \
    } while (0)
#define MAG_LAUNCH(m, c)                                                             \
    do {                                                                             \
        (c)->launches++;                                                             \
        cudaError_t e_ = cudaGetLastError();                                         \
        if (e_ != cudaSuccess) {                                                     \
            (m)->err = std::string("kernel launch: ") + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + \
                       std::to_string(__LINE__);                                     \
            return WB200_ERR_CUDA;                                                   \
        }                                                                            \
    } while (0)
// a failure inside one of the two contexts: carry its message
#define MAG_TRY(m, c, expr)                                        \
    do {                                                           \
        int s_ = (expr);                                           \
        if (s_ != WB200_OK) { (m)->err = (c)->err; return s_; }    \
    } while (0)

// per-k column dots: dpart[k][n] = sum_m conj(U[k][n][m]) T[k][n][m]; one warp per (k, n)
__global__ void __launch_bounds__(256) mag_coldot_kernel(int nb, int nw, long long total, const cplx* __restrict__ U,
                                                         const cplx* __restrict__ T, cplx* __restrict__ dpart) {
    const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= total) return;
    const cplx* u = U + wid * nb;
    const cplx* t = T + wid * nb;
    cplx acc = make_double2(0.0, 0.0);
    for (int m = lane; m < nb; m += 32) cfma_conj(acc, u[m], t[m]);
    acc.x = warp_sum(acc.x);
    acc.y = warp_sum(acc.y);
    if (lane == 0) dpart[wid] = acc;
}

// stage 1 of the sum over k: chunks of 64 k-points, one block per (chunk, 32 columns), fixed order
__global__ void __launch_bounds__(256) mag_chunk_kernel(int nw, int nk, int chunk, const cplx* __restrict__ dpart,
                                                        cplx* __restrict__ cpart) {
    const int n = blockIdx.y * 32 + (threadIdx.x & 31);
    const int sub = threadIdx.x >> 5;                  // 8 sub-ranges of the chunk
    const int k0 = blockIdx.x * chunk, k1 = min(nk, k0 + chunk);
    __shared__ cplx s[8][32];
    cplx acc = make_double2(0.0, 0.0);
    if (n < nw)
        for (int k = k0 + sub; k < k1; k += 8) {
            const cplx v = dpart[(long long)k * nw + n];
            acc.x += v.x; acc.y += v.y;
        }
    s[sub][threadIdx.x & 31] = acc;
    __syncthreads();
    if (sub == 0 && n < nw) {
        for (int j = 1; j < 8; j++) { acc.x += s[j][threadIdx.x].x; acc.y += s[j][threadIdx.x].y; }
        cpart[(long long)blockIdx.x * nw + n] = acc;
    }
}

// stage 2: d[n] = sum over the `nrows` partial rows (fixed order: deterministic); out[0] = Omega_updn =
// nw - sum_n |d_n|^2 / nk^2 (0 when nrows == 0: lambda == 0 skips the coupling, coopt.jl:267-271); with the
// two contexts' results given, out[1] = f_up + f_dn + lambda Omega_updn, also written to f_dst (where the
// optimiser reads f)
__global__ void __launch_bounds__(256) mag_reduce_kernel(int nw, int nrows, int nk, const cplx* __restrict__ part,
                                                         cplx* __restrict__ d, const double* __restrict__ res_up,
                                                         const double* __restrict__ res_dn, double lambda,
                                                         double* __restrict__ out, double* __restrict__ f_dst) {
    __shared__ double s_tr[8];
    double tr = 0.0;
    if (nrows > 0)
        for (int n = threadIdx.x; n < nw; n += blockDim.x) {
            cplx acc = make_double2(0.0, 0.0);
            for (int k = 0; k < nrows; k++) {
                const cplx v = part[(long long)k * nw + n];
                acc.x += v.x; acc.y += v.y;
            }
            d[n] = acc;
            tr += acc.x * acc.x + acc.y * acc.y;
        }
    tr = warp_sum(tr);
    if ((threadIdx.x & 31) == 0) s_tr[threadIdx.x >> 5] = tr;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += s_tr[w];
        const double o = nrows > 0 ? (double)nw - t / ((double)nk * (double)nk) : 0.0;
        out[0] = o;
        if (res_up && res_dn) {
            const double fu = res_up[5], fd = res_dn[5];
            const double f = fu + fd + lambda * o;
            out[1] = f; out[2] = fu; out[3] = fd;
            if (f_dst) *f_dst = f;
        }
    }
}

// Gup[k] (+)= c T_k diag(conj d),  Gdn[k] (+)= c S_k diag(d)
__global__ void __launch_bounds__(256) mag_grad_add_kernel(int nb, int nw, long long total, double c,
                                                           const cplx* __restrict__ T, const cplx* __restrict__ S,
                                                           const cplx* __restrict__ d, cplx* __restrict__ Gup,
                                                           cplx* __restrict__ Gdn, int accumulate) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int n = (int)((e / nb) % nw);
        const cplx dn = d[n];
        const cplx a = cmul(T[e], cconj(dn)), b = cmul(S[e], dn);
        cplx gu = accumulate ? Gup[e] : make_double2(0.0, 0.0);
        cplx gd = accumulate ? Gdn[e] : make_double2(0.0, 0.0);
        gu.x += c * a.x; gu.y += c * a.y;
        gd.x += c * b.x; gd.y += c * b.y;
        Gup[e] = gu; Gdn[e] = gd;
    }
}

// Mw = sum_k Mwk[k] (fixed order), out = |Mw|^2 / nk^2
__global__ void __launch_bounds__(256) mag_sum_abs2_kernel(int nw2, int nk, const cplx* __restrict__ Mwk, double* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nw2) return;
    cplx acc = make_double2(0.0, 0.0);
    for (int k = 0; k < nk; k++) {
        const cplx v = Mwk[(long long)k * nw2 + e];
        acc.x += v.x; acc.y += v.y;
    }
    out[e] = (acc.x * acc.x + acc.y * acc.y) / ((double)nk * (double)nk);
}

// every launch below goes to the up context's stream; the down context follows it for the duration of a call
struct StreamBind {
    wb200_ctx* dn; cudaStream_t saved;
    explicit StreamBind(wb200_mag* m) : dn(m->dn), saved(m->dn->stream) { dn->stream = m->up->stream; }
    ~StreamBind() { dn->stream = saved; }
};

// T (and S), d and Omega_updn for gauges resident on the device ([nk][nw][nb] each); with res_up / res_dn also
// the total functional.  lambda == 0 skips the coupling as the reference does.
static int mag_coupling(wb200_mag* m, const cplx* dUup, const cplx* dUdn, bool with_S, bool skip, const double* res_up,
                        const double* res_dn, double lambda, double* f_dst) {
    wb200_ctx* c = m->up;
    const int nb = m->nb, nw = m->nw, nk = m->nk;
    const cplx* part = m->d_dpart;
    int nrows = 0;
    if (!skip) {
        ZgemmArgs a{};
        a.m = nb; a.n = nw; a.k = nb;
        a.A = m->d_Mud; a.strideA = (long long)nb * nb; a.lda = nb; a.transA = 0;
        a.B = dUdn; a.strideB = (long long)nb * nw; a.ldb = nb; a.transB = 0;
        a.C = m->d_T; a.strideC = (long long)nb * nw; a.ldc = nb;
        a.alpha = 1.0; a.beta = 0.0; a.batch = nk;
        MAG_TRY(m, c, wb_zgemm(c, a));
        if (with_S) {
            ZgemmArgs b = a;
            b.transA = 1; b.B = dUup; b.C = m->d_S;
            MAG_TRY(m, c, wb_zgemm(c, b));
        }
        const long long warps = (long long)nk * nw;
        mag_coldot_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, c->stream>>>(nb, nw, warps, dUup, m->d_T, m->d_dpart);
        MAG_LAUNCH(m, c);
        nrows = nk;
        if (nk > 64) {   // the tail of d_dpart holds the chunk sums
            const int chunk = 64;
            nrows = (nk + chunk - 1) / chunk;
            cplx* cpart = m->d_dpart + (size_t)nk * nw;
            mag_chunk_kernel<<<dim3(nrows, (nw + 31) / 32), 256, 0, c->stream>>>(nw, nk, chunk, m->d_dpart, cpart);
            MAG_LAUNCH(m, c);
            part = cpart;
        }
    }
    mag_reduce_kernel<<<1, 256, 0, c->stream>>>(nw, nrows, nk, part, m->d_d, res_up, res_dn, lambda, m->d_out, f_dst);
    MAG_LAUNCH(m, c);
    return WB200_OK;
}

static int mag_grad_add(wb200_mag* m, double cfac, cplx* Gup, cplx* Gdn, int accumulate) {
    wb200_ctx* c = m->up;
    const long long total = (long long)m->nk * m->nb * m->nw;
    const int blocks = (int)((total + 255) / 256 < 8192 ? (total + 255) / 256 : 8192);
    mag_grad_add_kernel<<<blocks, 256, 0, c->stream>>>(m->nb, m->nw, total, cfac, m->d_T, m->d_S, m->d_d, Gup, Gdn, accumulate);
    MAG_LAUNCH(m, c);
    return WB200_OK;
}

// f and (optionally) the Euclidean gradient with respect to XY = [XYup_k; XYdn_k] resident on the device;
// f lands in d_out[1] and in up->d_res[5]
static int mag_eval_xy(wb200_mag* m, const cplx* dXY, double lambda, cplx* dGXY) {
    wb200_ctx *up = m->up, *dn = m->dn;
    const long long ni = (long long)m->n_inner(), st2 = 2 * ni;
    MAG_TRY(m, up, wb_xy_to_u(up, dXY, up->d_U, st2));
    MAG_TRY(m, dn, wb_xy_to_u(dn, dXY + ni, dn->d_U, st2));
    MAG_TRY(m, up, wb_fg_device(up, up->d_U, up->d_res, dGXY ? up->d_G : nullptr, 0));
    MAG_TRY(m, dn, wb_fg_device(dn, dn->d_U, dn->d_res, dGXY ? dn->d_G : nullptr, 0));
    WB_TRY(mag_coupling(m, up->d_U, dn->d_U, dGXY != nullptr, lambda == 0.0, up->d_res, dn->d_res, lambda, up->d_res + 5));
    if (dGXY) {
        if (lambda != 0.0)   // omega_updn_grad = -2 overlap_updn_grad (coopt.jl:222-235), times lambda (:300-305)
            WB_TRY(mag_grad_add(m, -2.0 * lambda / ((double)m->nk * (double)m->nk), up->d_G, dn->d_G, 1));
        MAG_TRY(m, up, wb_gu_to_gxy(up, dXY, up->d_G, dGXY, st2));
        MAG_TRY(m, dn, wb_gu_to_gxy(dn, dXY + ni, dn->d_G, dGXY + ni, st2));
    }
    return WB200_OK;
}

extern "C" const char* wb200_mag_last_error(const wb200_mag* m) { return m ? m->err.c_str() : g_mag_create_error.c_str(); }

extern "C" void wb200_mag_destroy(wb200_mag* m) {
    if (!m) return;
    if (m->up) cudaSetDevice(m->up->device);
    void* ptrs[] = {m->d_Mud, m->d_T, m->d_S, m->d_dpart, m->d_d, m->d_Mw, m->d_XY, m->d_GXY, m->d_out};
    for (void* p : ptrs) if (p) wb_dev_free(p);
    delete m;
}

extern "C" int wb200_mag_create(wb200_mag** out, wb200_ctx* up, wb200_ctx* dn, const double* Mud) {
    if (!out) { g_mag_create_error = "wb200_mag_create: out is NULL"; return WB200_ERR_ARG; }
    *out = nullptr;
    if (!up || !dn || !Mud || up == dn) {
        g_mag_create_error = "wb200_mag_create: need two distinct contexts and Mud";
        return WB200_ERR_ARG;
    }
    if (up->nb != dn->nb || up->nw != dn->nw || up->nk != dn->nk || up->device != dn->device) {   // coopt.jl:355-357
        g_mag_create_error = "wb200_mag_create: the up and down contexts must agree in n_bands, n_wann, n_kpts and device";
        return WB200_ERR_ARG;
    }
    if (up->nk != up->nk_total || dn->nk != dn->nk_total || up->allreduce || dn->allreduce) {
        g_mag_create_error = "wb200_mag_create: single-slab contexts only";
        return WB200_ERR_ARG;
    }
    cudaSetDevice(up->device);
    wb200_mag* m = new wb200_mag();
    m->up = up; m->dn = dn; m->nb = up->nb; m->nw = up->nw; m->nk = up->nk;
    const size_t mat = (size_t)m->nb * m->nw;
    const int nchunks = (m->nk + 63) / 64;
    cudaError_t e = wb_dev_alloc(&m->d_Mud, sizeof(cplx) * (size_t)m->nk * m->nb * m->nb);
    if (e == cudaSuccess) e = wb_dev_alloc(&m->d_T, sizeof(cplx) * (size_t)m->nk * mat);
    if (e == cudaSuccess) e = wb_dev_alloc(&m->d_S, sizeof(cplx) * (size_t)m->nk * mat);
    if (e == cudaSuccess) e = wb_dev_alloc(&m->d_dpart, sizeof(cplx) * (size_t)(m->nk + nchunks) * m->nw);
    if (e == cudaSuccess) e = wb_dev_alloc(&m->d_d, sizeof(cplx) * m->nw);
    if (e == cudaSuccess) e = wb_dev_alloc(&m->d_out, sizeof(double) * 4);
    if (e == cudaSuccess) e = cudaMemcpy(m->d_Mud, Mud, sizeof(cplx) * (size_t)m->nk * m->nb * m->nb, cudaMemcpyDefault);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        g_mag_create_error = std::string("wb200_mag_create: ") + cudaGetErrorString(e);
        wb200_mag_destroy(m);
        return WB200_ERR_CUDA;
    }
    *out = m;
    return WB200_OK;
}

static int mag_sync_out(wb200_mag* m, int count, double* host) {
    wb200_ctx* c = m->up;
    MAG_CUDA(m, cudaMemcpyAsync(c->h_pinned + 32, m->d_out, sizeof(double) * count, cudaMemcpyDeviceToHost, c->stream));
    MAG_CUDA(m, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < count; i++) host[i] = c->h_pinned[32 + i];
    return WB200_OK;
}

static int mag_upload_gauges(wb200_mag* m, const double* Uup, const double* Udn) {
    const size_t bytes = sizeof(cplx) * (size_t)m->nk * m->nb * m->nw;
    MAG_CUDA(m, cudaMemcpyAsync(m->up->d_U, Uup, bytes, cudaMemcpyHostToDevice, m->up->stream));
    MAG_CUDA(m, cudaMemcpyAsync(m->dn->d_U, Udn, bytes, cudaMemcpyHostToDevice, m->up->stream));
    return WB200_OK;
}

extern "C" int wb200_mag_overlap(wb200_mag* m, const double* Uup, const double* Udn, double* Mw2, double* omega_updn) {
    if (!m || !Uup || !Udn) return WB200_ERR_ARG;
    cudaSetDevice(m->up->device);
    StreamBind bind(m);
    wb200_ctx* c = m->up;
    const int nb = m->nb, nw = m->nw, nk = m->nk;
    WB_TRY(mag_upload_gauges(m, Uup, Udn));
    WB_TRY(mag_coupling(m, m->up->d_U, m->dn->d_U, false, false, nullptr, nullptr, 0.0, nullptr));
    if (Mw2) {
        if (!m->d_Mw) MAG_CUDA(m, wb_dev_alloc(&m->d_Mw, sizeof(cplx) * (size_t)nk * nw * nw + sizeof(double) * nw * nw));
        ZgemmArgs a{};   // Mw_k = Uup_k^H T_k
        a.m = nw; a.n = nw; a.k = nb;
        a.A = m->up->d_U; a.strideA = (long long)nb * nw; a.lda = nb; a.transA = 1;
        a.B = m->d_T; a.strideB = (long long)nb * nw; a.ldb = nb; a.transB = 0;
        a.C = m->d_Mw; a.strideC = (long long)nw * nw; a.ldc = nw;
        a.alpha = 1.0; a.beta = 0.0; a.batch = nk;
        MAG_TRY(m, c, wb_zgemm(c, a));
        double* dout = (double*)(m->d_Mw + (size_t)nk * nw * nw);
        mag_sum_abs2_kernel<<<(nw * nw + 255) / 256, 256, 0, c->stream>>>(nw * nw, nk, m->d_Mw, dout);
        MAG_LAUNCH(m, c);
        MAG_CUDA(m, cudaMemcpyAsync(Mw2, dout, sizeof(double) * nw * nw, cudaMemcpyDeviceToHost, c->stream));
    }
    double o = 0.0;
    WB_TRY(mag_sync_out(m, 1, &o));
    if (omega_updn) *omega_updn = o;
    return WB200_OK;
}

extern "C" int wb200_mag_omega_updn_grad(wb200_mag* m, const double* Uup, const double* Udn, double* GUup, double* GUdn) {
    if (!m || !Uup || !Udn || !GUup || !GUdn) return WB200_ERR_ARG;
    cudaSetDevice(m->up->device);
    StreamBind bind(m);
    wb200_ctx* c = m->up;
    const size_t bytes = sizeof(cplx) * (size_t)m->nk * m->nb * m->nw;
    WB_TRY(mag_upload_gauges(m, Uup, Udn));
    WB_TRY(mag_coupling(m, m->up->d_U, m->dn->d_U, true, false, nullptr, nullptr, 0.0, nullptr));
    WB_TRY(mag_grad_add(m, -2.0 / ((double)m->nk * (double)m->nk), m->up->d_G, m->dn->d_G, 0));
    MAG_CUDA(m, cudaMemcpyAsync(GUup, m->up->d_G, bytes, cudaMemcpyDeviceToHost, c->stream));
    MAG_CUDA(m, cudaMemcpyAsync(GUdn, m->dn->d_G, bytes, cudaMemcpyDeviceToHost, c->stream));
    MAG_CUDA(m, cudaStreamSynchronize(c->stream));
    return WB200_OK;
}

static int mag_ensure_xy(wb200_mag* m) {
    const size_t bytes = sizeof(cplx) * 2 * m->n_inner() * m->nk;
    if (!m->d_XY) MAG_CUDA(m, wb_dev_alloc(&m->d_XY, bytes));
    if (!m->d_GXY) MAG_CUDA(m, wb_dev_alloc(&m->d_GXY, bytes));
    return WB200_OK;
}

extern "C" int wb200_mag_fg_xy(wb200_mag* m, const double* XY, double lambda, double* f, double* G, double out4[4]) {
    if (!m || !XY) return WB200_ERR_ARG;
    cudaSetDevice(m->up->device);
    StreamBind bind(m);
    wb200_ctx* c = m->up;
    WB_TRY(mag_ensure_xy(m));
    const size_t bytes = sizeof(cplx) * 2 * m->n_inner() * m->nk;
    MAG_CUDA(m, cudaMemcpyAsync(m->d_XY, XY, bytes, cudaMemcpyHostToDevice, c->stream));
    WB_TRY(mag_eval_xy(m, m->d_XY, lambda, G ? m->d_GXY : nullptr));
    if (G) MAG_CUDA(m, cudaMemcpyAsync(G, m->d_GXY, bytes, cudaMemcpyDeviceToHost, c->stream));
    double o[4];
    WB_TRY(mag_sync_out(m, 4, o));
    if (f) *f = o[1];
    if (out4) for (int i = 0; i < 4; i++) out4[i] = o[i];
    return WB200_OK;
}

extern "C" int wb200_mag_disentangle(wb200_mag* m, double* XY_inout, double lambda, double f_tol, double g_tol,
                                     int max_iter, int history, double* f_final, int* iters, int* n_fg) {
    if (!m || !XY_inout) return WB200_ERR_ARG;
    cudaSetDevice(m->up->device);
    StreamBind bind(m);
    wb200_ctx *up = m->up, *dn = m->dn;
    WB_TRY(mag_ensure_xy(m));
    const long long ni = (long long)m->n_inner(), st2 = 2 * ni;
    const size_t bytes = sizeof(cplx) * (size_t)st2 * m->nk;
    MAG_CUDA(m, cudaMemcpyAsync(m->d_XY, XY_inout, bytes, cudaMemcpyHostToDevice, up->stream));
    WbCustomProblem cp;
    cp.ncplx = (size_t)st2 * m->nk;
    cp.retract = [=](cplx* x, bool tangent) -> int {
        int s = wb_retract_xy(up, x, tangent, st2);
        if (s != WB200_OK) return s;
        s = wb_retract_xy(dn, x + ni, tangent, st2);
        if (s != WB200_OK) up->err = dn->err;
        return s;
    };
    cp.project = [=](const cplx* x, cplx* g) -> int {
        int s = wb_project_xy(up, x, g, st2);
        if (s != WB200_OK) return s;
        s = wb_project_xy(dn, x + ni, g + ni, st2);
        if (s != WB200_OK) up->err = dn->err;
        return s;
    };
    cp.eval = [=](const cplx* x, cplx* g) -> int {
        const int s = mag_eval_xy(m, x, lambda, g);
        if (s != WB200_OK) up->err = m->err;
        return s;
    };
    MAG_TRY(m, up, wb_lbfgs_custom(up, cp, m->d_XY, f_tol, g_tol, max_iter, history, f_final, iters, n_fg));
    MAG_CUDA(m, cudaMemcpyAsync(XY_inout, m->d_XY, bytes, cudaMemcpyDeviceToHost, up->stream));
    MAG_CUDA(m, cudaStreamSynchronize(up->stream));
    return WB200_OK;
}
