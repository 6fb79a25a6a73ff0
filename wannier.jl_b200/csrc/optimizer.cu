// optimizer.cu — device-resident Riemannian L-BFGS for max_localize / disentangle.
//
// Replaces the Optim.optimize(only_fg!(fg!), x0, LBFGS(manifold, HagerZhang, m)) calls at
// src/localization/max_localize.jl:73-84 and src/localization/disentangle.jl:702-713.  Optim.jl
// is third-party code that is not in the reference tree; its exact trajectory cannot be pinned
// (SURVEY 8c).  This driver keeps Optim's stopping rules (|f - f_prev| <= f_tol |f| or
// max|g| <= g_tol on the projected gradient, max_iter) and is the same algorithm, step for step,
// as oracle/spread_oracle.py::lbfgs_stiefel, which the parity tests compare it with.
// x, g, the search direction and the 2m history vectors never leave HBM; per iteration only a
// handful of scalars (f, dot products) cross PCIe.
#include "common.cuh"

#include <cmath>
#include <deque>

int wb_fg_device(wb200_ctx* ctx, const cplx* dU, double* dres, cplx* dG, int exact_split);  // api.cu
int wb_fg_rotate_device(wb200_ctx* ctx, const cplx* dW, double* dres, cplx* dGW);           // extras.cu

namespace {

struct Problem {
    wb200_ctx* ctx;
    int mode;     // 0: U on (Stiefel nb x nw)^nk ; 1: XY on (Stiefel nw x nw  x  Stiefel nb x nw)^nk ;
                  // 2: one global W on Stiefel nw x nw (opt_rotate)
    size_t n;     // doubles per vector
    int n_fg = 0;

    int retract(cplx* x) {
        if (mode == 2) return wb_retract_dev(ctx, x, ctx->nw, ctx->nw, 1, nullptr);
        if (mode == 0) return wb_retract_dev(ctx, x, ctx->nb, ctx->nw, ctx->nk, nullptr);
        return wb_retract_xy(ctx, x);
    }
    int project(const cplx* x, cplx* g) {
        if (mode == 2) return wb_project_tangent_dev(ctx, x, g, ctx->nw, ctx->nw, 1);
        if (mode == 0) return wb_project_tangent_dev(ctx, x, g, ctx->nb, ctx->nw, ctx->nk);
        return wb_project_xy(ctx, x, g);
    }
    // f and the PROJECTED gradient at x
    int fg(const cplx* x, double* f, cplx* g) {
        n_fg++;
        if (mode == 0) {
            WB_TRY(wb_fg_device(ctx, x, ctx->d_res, g, 0));
        } else if (mode == 2) {
            WB_TRY(wb_fg_rotate_device(ctx, x, ctx->d_res, g));
        } else {
            WB_TRY(wb_xy_to_u(ctx, x, ctx->d_U));
            WB_TRY(wb_fg_device(ctx, ctx->d_U, ctx->d_res, ctx->d_G, 0));
            WB_TRY(wb_gu_to_gxy(ctx, x, ctx->d_G, g));
        }
        WB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 16, ctx->d_res + 5, sizeof(double),
                                     cudaMemcpyDeviceToHost, ctx->stream));
        WB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 17, ctx->d_min, sizeof(double),
                                     cudaMemcpyDeviceToHost, ctx->stream));
        WB_TRY(project(x, g));
        WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *f = ctx->h_pinned[16];
        if (mode == 2 && !ctx->allreduce && ctx->h_pinned[17] < 1e-10) {   // opt_rotate.jl:60-62
            ctx->err = "N^kb too small! (|N_nn| < 1e-10)";
            return WB200_ERR_SINGULAR;
        }
        return WB200_OK;
    }
};

}  // namespace

int wb_lbfgs(wb200_ctx* ctx, int mode, cplx* dx, double f_tol, double g_tol, int max_iter,
             int history, double* f_final, int* iters_out, int* n_fg_out) {
    Problem P{ctx, mode, 0};
    const size_t ncplx = (mode == 2) ? (size_t)ctx->nw * ctx->nw
                         : (mode == 0) ? (size_t)ctx->nk * ctx->nb * ctx->nw
                                       : (size_t)ctx->nk * ((size_t)ctx->nw * ctx->nw + (size_t)ctx->nb * ctx->nw);
    P.n = 2 * ncplx;
    const size_t n = P.n;
    const int m = history > 0 ? history : 3;

    // work vectors: g, d, q, xn, gn, td, xb, gb + 2m history
    const int nvec = 8 + 2 * (m + 1);  // m history pairs + one candidate pair
    double* pool = nullptr;
    WB_CUDA(ctx, cudaMalloc(&pool, sizeof(double) * n * nvec));
    struct Free { double* p; ~Free() { cudaFree(p); } } guard{pool};
    auto vec = [&](int i) { return pool + (size_t)i * n; };
    double *g = vec(0), *d = vec(1), *q = vec(2), *xn = vec(3), *gn = vec(4), *td = vec(5),
           *xb = vec(6), *gb = vec(7);
    std::deque<int> slots;            // history slot ids, oldest first
    std::vector<double> rho(m + 1, 0.0);
    auto S = [&](int s) { return vec(8 + 2 * s); };
    auto Y = [&](int s) { return vec(8 + 2 * s + 1); };
    std::vector<int> free_slots;
    for (int i = m; i >= 0; i--) free_slots.push_back(i);

    double* x = (double*)dx;
    WB_TRY(P.retract(dx));
    double f = 0.0;
    WB_TRY(P.fg(dx, &f, (cplx*)g));
    int it = 0;
    int status = WB200_OK;
    while (it < max_iter) {
        double gmax = 0.0;
        WB_TRY(wb_absmax(ctx, g, n, &gmax));
        if (gmax <= g_tol) break;
        // two-loop recursion
        WB_TRY(wb_axpbyz(ctx, 1.0, g, 0.0, g, q, n));
        std::vector<double> al;
        for (auto itS = slots.rbegin(); itS != slots.rend(); ++itS) {
            double sq = 0.0;
            WB_TRY(wb_dot(ctx, S(*itS), q, n, &sq));
            const double a = rho[*itS] * sq;
            al.push_back(a);
            WB_TRY(wb_axpby(ctx, -a, Y(*itS), 1.0, q, n));
        }
        if (!slots.empty()) {
            const int last = slots.back();
            double sy = 0.0, yy = 0.0;
            WB_TRY(wb_dot(ctx, S(last), Y(last), n, &sy));
            WB_TRY(wb_dot(ctx, Y(last), Y(last), n, &yy));
            WB_TRY(wb_axpby(ctx, 0.0, q, sy / yy, q, n));
        }
        {
            int j = (int)al.size() - 1;
            for (auto itS = slots.begin(); itS != slots.end(); ++itS, --j) {
                double yq = 0.0;
                WB_TRY(wb_dot(ctx, Y(*itS), q, n, &yq));
                const double b = rho[*itS] * yq;
                WB_TRY(wb_axpby(ctx, al[j] - b, S(*itS), 1.0, q, n));
            }
        }
        WB_TRY(wb_axpbyz(ctx, -1.0, q, 0.0, q, d, n));
        WB_TRY(P.project(dx, (cplx*)d));
        double dphi0 = 0.0, gg = 0.0;
        WB_TRY(wb_dot(ctx, g, d, n, &dphi0));
        if (dphi0 >= 0.0) {
            WB_TRY(wb_axpbyz(ctx, -1.0, g, 0.0, g, d, n));
            WB_TRY(wb_dot(ctx, g, d, n, &dphi0));
            while (!slots.empty()) { free_slots.push_back(slots.front()); slots.pop_front(); }
        }
        double alpha = 1.0;
        if (slots.empty()) {
            WB_TRY(wb_dot(ctx, g, g, n, &gg));
            alpha = fmin(1.0, 1.0 / fmax(sqrt(gg), 1e-300));
        }
        // strong-Wolfe bracketing line search
        const double c1 = 1e-4, c2 = 0.9;
        double lo = 0.0, hi = -1.0, f_lo = f;
        bool have_best = false;
        double f_best = 0.0, a_best = 0.0;
        // the best trial point so far lives in (xb, gb); (xn, gn) is the current trial
        for (int ls = 0; ls < 30; ls++) {
            // xn = retract(x + alpha d)
            WB_TRY(wb_axpbyz(ctx, 1.0, x, alpha, d, xn, n));
            WB_TRY(P.retract((cplx*)xn));
            double fn = 0.0;
            WB_TRY(P.fg((cplx*)xn, &fn, (cplx*)gn));
            WB_TRY(wb_axpbyz(ctx, 1.0, d, 0.0, d, td, n));
            WB_TRY(P.project((cplx*)xn, (cplx*)td));
            double dn = 0.0;
            WB_TRY(wb_dot(ctx, gn, td, n, &dn));
            if (!(fn == fn)) {  // NaN: shrink
                hi = alpha;
            } else if (fn > f + c1 * alpha * dphi0 || (have_best && fn >= f_lo && lo > 0.0)) {
                hi = alpha;
            } else if (fabs(dn) <= -c2 * dphi0) {
                std::swap(xn, xb); std::swap(gn, gb);
                f_best = fn; a_best = alpha; have_best = true;
                break;
            } else if (dn >= 0.0) {
                hi = alpha;
            } else {
                lo = alpha; f_lo = fn;
                std::swap(xn, xb); std::swap(gn, gb);
                f_best = fn; a_best = alpha; have_best = true;
            }
            alpha = (hi < 0.0) ? 2.0 * alpha : 0.5 * (lo + hi);
        }
        if (!have_best) break;
        // secant pair as Optim.jl's LBFGS forms it on a manifold (update_state! / update_h!):
        // dx = alpha * s (the step before retraction), dg = g_new - g_previous, each gradient
        // projected at its own point; the history is not transported.
        const int slot = free_slots.back();
        free_slots.pop_back();
        WB_TRY(wb_axpbyz(ctx, a_best, d, 0.0, d, S(slot), n));
        WB_TRY(wb_axpbyz(ctx, 1.0, gb, -1.0, g, Y(slot), n));
        double sy = 0.0, ss = 0.0, yy = 0.0;
        WB_TRY(wb_dot(ctx, S(slot), Y(slot), n, &sy));
        WB_TRY(wb_dot(ctx, S(slot), S(slot), n, &ss));
        WB_TRY(wb_dot(ctx, Y(slot), Y(slot), n, &yy));
        if (sy > 1e-14 * sqrt(ss * yy)) {
            rho[slot] = 1.0 / sy;
            slots.push_back(slot);
            if ((int)slots.size() > m) { free_slots.push_back(slots.front()); slots.pop_front(); }
        } else {
            free_slots.push_back(slot);
        }
        const double df = fabs(f_best - f);
        // x <- xb, g <- gb
        WB_CUDA(ctx, cudaMemcpyAsync(x, xb, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        WB_CUDA(ctx, cudaMemcpyAsync(g, gb, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        f = f_best;
        it++;
        if (df <= f_tol * fabs(f)) break;
    }
    WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (f_final) *f_final = f;
    if (iters_out) *iters_out = it;
    if (n_fg_out) *n_fg_out = P.n_fg;
    return status;
}
