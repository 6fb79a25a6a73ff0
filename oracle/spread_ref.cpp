// spread_ref.cpp — compiled CPU restatement of the reference's spread / gradient evaluation.
//
// TEST INFRASTRUCTURE AND CPU BASELINE ONLY: tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product (wannier.jl_b200/) never does.
//
// One evaluation = one fg!(F, G, U) of get_fg!_maxloc (src/localization/max_localize.jl:13-23), in the
// reference's own structure and loop order:
//   compute_MU_UtMU!   src/spread.jl:164-195   per k, per b: mul!(MU_kb, M_kb, U_{k+b}); mul!(N_kb, U_k', MU_kb)
//                      (Julia's mul! on Matrix{ComplexF64} is OpenBLAS ZGEMM; the same library is used here,
//                      loaded from numpy's bundled copy, one BLAS thread per caller)
//   center!            src/spread.jl:565-594
//   omega_grad!        src/spread.jl:398-481   (MU is scaled column by column in place, G += 4 w_b MU, G /= n_kpts)
//   omega!             src/spread.jl:254-360   (full N: ts, ts2, diagonal terms; Omega_D = Omega~ - Omega_OD)
//   center_penalty     src/spread.jl:227, omega_center :246-252
// The reference runs these loops on one thread (benchmark/runbenchmarks.jl:5 sets JULIA_NUM_THREADS = 1);
// here the k loop is an OpenMP parallel for, the most favourable parallelisation the structure allows
// (MU, N and G of different k are independent), so that the GPU is compared with all host cores.
// MU and N are stored for every pair, as the reference's Cache does (src/spread.jl:127-162).
//
// Memory layout = the C-ABI of include/wannier_b200.h = Julia's: complex interleaved, column-major,
// M [nk][nn][nb x nb], U [nk][nb x nw], G [nk][nb x nw], kpb 0-based int32, bvec [nk][nn][3].
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <dlfcn.h>
#include <omp.h>

typedef std::complex<double> cd;

// cblas_zgemm with 64-bit integers (scipy_openblas64_: ILP64, the interface Julia's libblastrampoline uses)
typedef void (*cblas_zgemm64_t)(int order, int transA, int transB, int64_t M, int64_t N, int64_t K, const void* alpha,
                                const void* A, int64_t lda, const void* B, int64_t ldb, const void* beta, void* C,
                                int64_t ldc);
typedef void (*set_threads_t)(int);
static cblas_zgemm64_t g_zgemm = nullptr;
static char g_blas_name[512] = "built-in loops (no BLAS loaded)";

// C = op(A) B; op = identity (ta = 0) or conjugate transpose (ta = 1); column-major
static void zgemm_fallback(int ta, int m, int n, int k, const cd* A, int lda, const cd* B, int ldb, cd* C, int ldc) {
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < m; i++) C[i + (size_t)j * ldc] = 0.0;
        if (!ta) {
            for (int l = 0; l < k; l++) {
                const cd b = B[l + (size_t)j * ldb];
                const cd* a = A + (size_t)l * lda;
                cd* c = C + (size_t)j * ldc;
                for (int i = 0; i < m; i++) c[i] += a[i] * b;
            }
        } else {
            for (int i = 0; i < m; i++) {
                const cd* a = A + (size_t)i * lda;
                const cd* b = B + (size_t)j * ldb;
                cd s = 0.0;
                for (int l = 0; l < k; l++) s += std::conj(a[l]) * b[l];
                C[i + (size_t)j * ldc] = s;
            }
        }
    }
}

static inline void zgemm(int ta, int m, int n, int k, const cd* A, int lda, const cd* B, int ldb, cd* C, int ldc) {
    if (g_zgemm) {
        const cd one(1.0, 0.0), zero(0.0, 0.0);
        g_zgemm(102 /*ColMajor*/, ta ? 113 /*ConjTrans*/ : 111 /*NoTrans*/, 111, m, n, k, &one, A, lda, B, ldb, &zero, C, ldc);
    } else {
        zgemm_fallback(ta, m, n, k, A, lda, B, ldb, C, ldc);
    }
}

extern "C" {

// dlopen an ILP64 OpenBLAS (numpy.libs/libscipy_openblas64_*.so) and pin it to one thread per caller
int sr_set_blas(const char* path) {
    void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return -1;
    auto f = (cblas_zgemm64_t)dlsym(h, "scipy_cblas_zgemm64_");
    if (!f) return -2;
    auto st = (set_threads_t)dlsym(h, "scipy_openblas_set_num_threads64_");
    if (st) st(1);
    g_zgemm = f;
    snprintf(g_blas_name, sizeof(g_blas_name), "OpenBLAS cblas_zgemm (ILP64) from %s", path);
    return 0;
}
const char* sr_blas_name() { return g_blas_name; }
int sr_max_threads() { return omp_get_max_threads(); }

static inline double imaglog(cd z) { return std::atan2(z.imag(), z.real()); }   // src/utils/linalg.jl:15

// res: {Omega, OmegaI, OmegaOD, OmegaD, OmegaTilde, f (= Omega_t with a penalty), Omega_c, 0, omega[nw], r[nw][3]}
// MU, N: caller-provided work arrays [nk*nn][nb*nw], [nk*nn][nw*nw] (the reference's Cache); G may be NULL.
// k_sel: optional list of n_sel k indices to evaluate instead of all nk (bounded samples for timing: the
//        sums then run over the sample only and the results are NOT the spread of the model).
int sr_fg(int nb, int nw, int nk, int nn, const int32_t* kpb, const double* bvec, const double* wb, const cd* M,
          const cd* U, const double* r0, double lambda, cd* MU, cd* N, double* res, cd* G, int threads,
          const int32_t* k_sel, int n_sel) {
    if (threads > 0) omp_set_num_threads(threads);
    const size_t mat = (size_t)nb * nw, nmat = (size_t)nw * nw, mm = (size_t)nb * nb;
    const int nloop = k_sel ? n_sel : nk;
    // ---- compute_MU_UtMU! (spread.jl:164-195) ----
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < nloop; i++) {
        const int ik = k_sel ? k_sel[i] : i;
        const cd* Uk = U + (size_t)ik * mat;
        for (int ib = 0; ib < nn; ib++) {
            const size_t p = (size_t)i * nn + ib;
            const int ikpb = kpb[(size_t)ik * nn + ib];
            zgemm(0, nb, nw, nb, M + p * mm, nb, U + (size_t)ikpb * mat, nb, MU + p * mat, nb);
            zgemm(1, nw, nw, nb, Uk, nb, MU + p * mat, nb, N + p * nmat, nw);
        }
    }
    // ---- center! (spread.jl:565-594) ----
    std::vector<double> r(3 * (size_t)nw, 0.0);
    for (int i = 0; i < nloop; i++) {
        const int ik = k_sel ? k_sel[i] : i;
        for (int ib = 0; ib < nn; ib++) {
            const size_t p = (size_t)i * nn + ib;
            const double* b = bvec + ((size_t)ik * nn + ib) * 3;
            const double w = wb[ib];
            const cd* Nb = N + p * nmat;
            for (int n = 0; n < nw; n++) {
                const double fac = w * imaglog(Nb[n + (size_t)n * nw]);
                r[3 * n] -= b[0] * fac; r[3 * n + 1] -= b[1] * fac; r[3 * n + 2] -= b[2] * fac;
            }
        }
    }
    for (auto& v : r) v /= nk;
    // ---- omega_grad! (spread.jl:398-481) ----
    if (G) {
#pragma omp parallel for schedule(dynamic, 1)
        for (int i = 0; i < nloop; i++) {
            const int ik = k_sel ? k_sel[i] : i;
            cd* Gk = G + (size_t)i * mat;
            for (size_t e = 0; e < mat; e++) Gk[e] = 0.0;
            for (int ib = 0; ib < nn; ib++) {
                const size_t p = (size_t)i * nn + ib;
                const double* b = bvec + ((size_t)ik * nn + ib) * 3;
                const double w = wb[ib];
                cd* MUkb = MU + p * mat;
                const cd* Nkb = N + p * nmat;
                for (int n = 0; n < nw; n++) {
                    const cd nnv = Nkb[n + (size_t)n * nw];
                    double rx = r[3 * n], ry = r[3 * n + 1], rz = r[3 * n + 2];
                    if (r0) {   // center_penalty: r - lambda (r - r0[n])  (spread.jl:227)
                        rx -= lambda * (rx - r0[3 * n]); ry -= lambda * (ry - r0[3 * n + 1]); rz -= lambda * (rz - r0[3 * n + 2]);
                    }
                    const double q = imaglog(nnv) + (rx * b[0] + ry * b[1] + rz * b[2]);
                    const cd t = cd(0.0, -1.0) * q / nnv;
                    const cd f = t - std::conj(nnv);
                    cd* col = MUkb + (size_t)n * nb;
                    for (int m = 0; m < nb; m++) col[m] *= f;
                }
                const double w4 = 4.0 * w;
                for (size_t e = 0; e < mat; e++) Gk[e] += w4 * MUkb[e];
            }
            for (size_t e = 0; e < mat; e++) Gk[e] /= (double)nk;
        }
    }
    // ---- omega! (spread.jl:254-360) ----
    if (res) {
        std::vector<double> r2(nw, 0.0);
        double OI = 0.0, OOD = 0.0;
        std::vector<double> rr(3 * (size_t)nw, 0.0);
#pragma omp parallel
        {
            std::vector<double> r2_l(nw, 0.0), rr_l(3 * (size_t)nw, 0.0);
            double OI_l = 0.0, OOD_l = 0.0;
#pragma omp for schedule(static) nowait
            for (int i = 0; i < nloop; i++) {
                const int ik = k_sel ? k_sel[i] : i;
                for (int ib = 0; ib < nn; ib++) {
                    const size_t p = (size_t)i * nn + ib;
                    const double* b = bvec + ((size_t)ik * nn + ib) * 3;
                    const double w = wb[ib];
                    const cd* Nkb = N + p * nmat;
                    double ts = 0.0, ts2 = 0.0;
                    for (int ii = 0; ii < nw; ii++)
                        for (int j = 0; j < nw; j++) {
                            const cd nt = Nkb[j + (size_t)ii * nw];
                            const double a2 = std::norm(nt);
                            ts += a2;
                            if (ii == j) {
                                const double il = imaglog(nt);
                                rr_l[3 * ii] -= il * w * b[0]; rr_l[3 * ii + 1] -= il * w * b[1]; rr_l[3 * ii + 2] -= il * w * b[2];
                                r2_l[ii] += w * (1.0 - a2 + il * il);
                            } else {
                                ts2 += a2;
                            }
                        }
                    OI_l += w * (nw - ts);
                    OOD_l += w * ts2;
                }
            }
#pragma omp critical
            {
                for (int n = 0; n < nw; n++) r2[n] += r2_l[n];
                for (size_t e = 0; e < rr.size(); e++) rr[e] += rr_l[e];
                OI += OI_l; OOD += OOD_l;
            }
        }
        double Om = 0.0, Oc = 0.0;
        for (int n = 0; n < nw; n++) {
            const double x = rr[3 * n] / nk, y = rr[3 * n + 1] / nk, z = rr[3 * n + 2] / nk;
            const double om = r2[n] / nk - (x * x + y * y + z * z);
            res[8 + n] = om;
            res[8 + nw + 3 * n] = x; res[8 + nw + 3 * n + 1] = y; res[8 + nw + 3 * n + 2] = z;
            Om += om;
            if (r0) {
                const double dx = x - r0[3 * n], dy = y - r0[3 * n + 1], dz = z - r0[3 * n + 2];
                Oc += lambda * (dx * dx + dy * dy + dz * dz);
            }
        }
        OI /= nk; OOD /= nk;
        const double Ot = Om - OI;
        res[0] = Om; res[1] = OI; res[2] = OOD; res[3] = Ot - OOD; res[4] = Ot; res[5] = Om + Oc; res[6] = Oc; res[7] = 0.0;
    }
    return 0;
}

// M_kb = V_k^H V_{k+b} for the synthetic inputs (input generation on the host, parallel over k):
// V [nk][nH x nb] column-major (the smooth subspace family of wannier.jl_b200/synthetic.py)
int sr_overlaps_from_subspaces(int nH, int nb, int nk, int nn, const int32_t* kpb, const cd* V, cd* M, int threads) {
    if (threads > 0) omp_set_num_threads(threads);
    const size_t vm = (size_t)nH * nb, mm = (size_t)nb * nb;
#pragma omp parallel for schedule(dynamic, 1)
    for (int ik = 0; ik < nk; ik++)
        for (int ib = 0; ib < nn; ib++)
            zgemm(1, nb, nb, nH, V + (size_t)ik * vm, nH, V + (size_t)kpb[(size_t)ik * nn + ib] * vm, nH,
                  M + ((size_t)ik * nn + ib) * mm, nb);
    return 0;
}

}  // extern "C"
