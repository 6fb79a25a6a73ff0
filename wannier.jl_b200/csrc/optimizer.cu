// optimizer.cu — device-resident Riemannian L-BFGS for max_localize / disentangle / opt_rotate.
//
// Replaces the Optim.optimize(only_fg!(fg!), x0, LBFGS(manifold, HagerZhang, m)) calls at
// src/localization/max_localize.jl:73-84, src/localization/disentangle.jl:702-713 and
// src/localization/opt_rotate.jl:118-141.  Optim.jl is third-party code that is not in the
// reference tree; its exact trajectory cannot be pinned.  This driver keeps Optim's
// stopping rules (|f - f_prev| <= f_tol |f| or max|g| <= g_tol on the projected gradient,
// max_iter) and is the same algorithm as oracle/spread_oracle.py::lbfgs_stiefel, which the parity
// tests compare it with.
//
// x, g, the search direction and the 2m history vectors never leave HBM.  The two-loop recursion is
// done in COEFFICIENT space: the Gram matrix of {g, s_i, y_i} is kept on the host (the rows of g
// and of the newest secant pair are refreshed by ONE fused pass over memory per iteration,
// wb_cross_launch), the recursion runs on those 7 x 7 numbers, and the search direction is ONE
// linear-combination pass (wb_lincomb).  Per iteration that is 3 host round trips and ~27 vector
// reads/writes instead of 13 round trips and ~64 (it was one kernel + one synchronisation per
// inner product and per axpy).
#include "common.cuh"

#include <algorithm>
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
    // launches f and the PROJECTED gradient at x, and (optionally) <g, d> in the same submission;
    // one synchronisation returns f, min |N_nn| and the inner product
    int fg(const cplx* x, double* f, cplx* g, const double* d, double* gd) {
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
        if (d) {
            WbCrossArgs c{};
            c.a[0] = (const double*)g; c.na = 1;
            c.b[0] = d; c.nb = 1;
            WB_TRY(wb_cross_launch(ctx, c, n));
        }
        WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *f = ctx->h_pinned[16];
        if (d) *gd = ctx->h_pinned[WB_PIN_CROSS];
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

    // work vectors: g, d, xn, gn + two spare (x, g) buffers for the best trial point + (m + 1)
    // history pairs (m live + one candidate)
    const int nfixed = 6;
    const int nvec = nfixed + 2 * (m + 1);
    double* pool = nullptr;
    WB_CUDA(ctx, cudaMalloc(&pool, sizeof(double) * n * nvec));
    struct Free { double* p; ~Free() { cudaFree(p); } } guard{pool};
    auto vec = [&](int i) { return pool + (size_t)i * n; };
    // x lives in the caller's array or in one of the spares (pointer swaps instead of copies)
    double *x = (double*)dx, *g = vec(0), *d = vec(1), *xn = vec(2), *gn = vec(3), *xb = vec(4),
           *gb = vec(5);
    auto S = [&](int s) { return vec(nfixed + 2 * s); };
    auto Y = [&](int s) { return vec(nfixed + 2 * s + 1); };
    std::deque<int> slots;            // live history slots, oldest first
    std::vector<int> free_slots;
    for (int i = m; i >= 0; i--) free_slots.push_back(i);

    // Gram matrix of the vectors {g, S(0..m), Y(0..m)} by id: 0 = g, 1 + s = S(s), 2 + m + s = Y(s)
    const int nid = 1 + 2 * (m + 1);
    std::vector<double> GM((size_t)nid * nid, 0.0);
    auto idS = [&](int s) { return 1 + s; };
    auto idY = [&](int s) { return 2 + m + s; };
    auto ptr_of = [&](int id) -> const double* {
        return id == 0 ? g : (id <= m + 1 ? S(id - 1) : Y(id - 2 - m));
    };
    // refresh the Gram rows of the first `na` ids of `ids` (against all of `ids`); returns max |g|
    auto refresh = [&](const std::vector<int>& ids, int na, double* gmax) -> int {
        for (size_t off = 0; off < ids.size(); off += WB_CROSS_NB) {
            WbCrossArgs c{};
            c.na = na;
            for (int i = 0; i < na; i++) c.a[i] = ptr_of(ids[i]);
            c.nb = (int)std::min<size_t>(WB_CROSS_NB, ids.size() - off);
            for (int j = 0; j < c.nb; j++) c.b[j] = ptr_of(ids[off + j]);
            c.a_in_b = off == 0;
            WB_TRY(wb_cross_launch(ctx, c, n));
            WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            const double* h = ctx->h_pinned + WB_PIN_CROSS;
            for (int i = 0; i < na; i++)
                for (int j = 0; j < c.nb; j++) {
                    const int a = ids[i], b = ids[off + j];
                    GM[(size_t)a * nid + b] = GM[(size_t)b * nid + a] = h[i * WB_CROSS_NB + j];
                }
            if (off == 0 && gmax) *gmax = sqrt(h[WB_CROSS_NOUT - 1]);
        }
        return WB200_OK;
    };

    WB_TRY(P.retract(dx));
    double f = 0.0;
    WB_TRY(P.fg(dx, &f, (cplx*)g, nullptr, nullptr));
    int it = 0;
    int pending = -1;   // slot of the secant pair formed by the last iteration, not yet admitted
    while (it < max_iter) {
        // ---- Gram rows of g (and of the new pair), max |g| : one pass over memory ----
        // (a full history evicts its oldest pair when the new one is admitted, so that pair is left
        // out of the pass: g + new pair + the m - 1 newest pairs = 7 vectors at m = 3)
        const bool skip_oldest = pending >= 0 && (int)slots.size() == m;
        std::vector<int> ids{0};
        if (pending >= 0) { ids.push_back(idS(pending)); ids.push_back(idY(pending)); }
        for (size_t i = skip_oldest ? 1 : 0; i < slots.size(); i++) ids.push_back(idS(slots[i]));
        for (size_t i = skip_oldest ? 1 : 0; i < slots.size(); i++) ids.push_back(idY(slots[i]));
        double gmax = 0.0;
        WB_TRY(refresh(ids, pending >= 0 ? 3 : 1, &gmax));
        if (pending >= 0) {   // admit the pair if it has positive curvature
            const double sy = GM[(size_t)idS(pending) * nid + idY(pending)];
            const double ss = GM[(size_t)idS(pending) * nid + idS(pending)];
            const double yy = GM[(size_t)idY(pending) * nid + idY(pending)];
            if (sy > 1e-14 * sqrt(ss * yy)) {
                slots.push_back(pending);
                if ((int)slots.size() > m) { free_slots.push_back(slots.front()); slots.pop_front(); }
            } else {
                free_slots.push_back(pending);
                if (skip_oldest)   // the oldest pair stays: its products with the new g are still missing
                    WB_TRY(refresh({0, idS(slots.front()), idY(slots.front())}, 1, nullptr));
            }
            pending = -1;
        }
        if (gmax <= g_tol) break;

        // ---- two-loop recursion on the coefficients of q = sum_id c[id] v_id ----
        std::vector<double> c(nid, 0.0);
        c[0] = 1.0;
        auto dotq = [&](int id) {
            double t = 0.0;
            for (int j = 0; j < nid; j++) if (c[j] != 0.0) t += c[j] * GM[(size_t)id * nid + j];
            return t;
        };
        auto rho = [&](int s) { return 1.0 / GM[(size_t)idS(s) * nid + idY(s)]; };
        std::vector<double> al;
        for (auto itS = slots.rbegin(); itS != slots.rend(); ++itS) {
            const double a = rho(*itS) * dotq(idS(*itS));
            al.push_back(a);
            c[idY(*itS)] -= a;
        }
        if (!slots.empty()) {
            const int last = slots.back();
            const double gamma = GM[(size_t)idS(last) * nid + idY(last)] / GM[(size_t)idY(last) * nid + idY(last)];
            for (double& v : c) v *= gamma;
        }
        {
            int j = (int)al.size() - 1;
            for (auto itS = slots.begin(); itS != slots.end(); ++itS, --j) {
                const double b = rho(*itS) * dotq(idY(*itS));
                c[idS(*itS)] += al[j] - b;
            }
        }
        // d = -q in one pass (chunks of 7 vectors)
        {
            std::vector<int> used;
            for (int id = 0; id < nid; id++) if (c[id] != 0.0) used.push_back(id);
            if (used.empty()) used.push_back(0);
            for (size_t off = 0; off < used.size(); off += WB_CROSS_NB) {
                WbLinArgs l{};
                l.nv = (int)std::min<size_t>(WB_CROSS_NB, used.size() - off);
                for (int j = 0; j < l.nv; j++) { l.v[j] = ptr_of(used[off + j]); l.c[j] = -c[used[off + j]]; }
                l.beta = off == 0 ? 0.0 : 1.0;
                WB_TRY(wb_lincomb(ctx, l, d, n));
            }
        }
        WB_TRY(P.project((const cplx*)x, (cplx*)d));
        const double gg = GM[0];
        double dphi0 = 0.0;
        {
            WbCrossArgs cr{};
            cr.a[0] = g; cr.na = 1; cr.b[0] = d; cr.nb = 1;
            WB_TRY(wb_cross_launch(ctx, cr, n));
            WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            dphi0 = ctx->h_pinned[WB_PIN_CROSS];
        }
        if (dphi0 >= 0.0) {   // not a descent direction: steepest descent, forget the history
            WbLinArgs l{};
            l.v[0] = g; l.c[0] = -1.0; l.nv = 1;
            WB_TRY(wb_lincomb(ctx, l, d, n));
            dphi0 = -gg;
            while (!slots.empty()) { free_slots.push_back(slots.front()); slots.pop_front(); }
        }
        double alpha = slots.empty() ? fmin(1.0, 1.0 / fmax(sqrt(gg), 1e-300)) : 1.0;

        // ---- strong-Wolfe bracketing line search ----
        // phi'(alpha) = <g(x_alpha), d> with g projected at x_alpha (the tangent projection is
        // self-adjoint, so projecting d as well would not change the product)
        const double c1 = 1e-4, c2 = 0.9;
        double lo = 0.0, hi = -1.0, f_lo = f;
        bool have_best = false;
        double f_best = 0.0, a_best = 0.0;
        // the best trial point so far lives in (xb, gb); (xn, gn) is the current trial
        for (int ls = 0; ls < 30; ls++) {
            WbLinArgs l{};
            l.v[0] = x; l.c[0] = 1.0; l.v[1] = d; l.c[1] = alpha; l.nv = 2;
            WB_TRY(wb_lincomb(ctx, l, xn, n));
            WB_TRY(P.retract((cplx*)xn));
            double fn = 0.0, dn = 0.0;
            WB_TRY(P.fg((cplx*)xn, &fn, (cplx*)gn, d, &dn));
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
        // projected at its own point; the history is not transported.  Its inner products are
        // taken by the next iteration's Gram pass, which also decides whether it is admitted.
        pending = free_slots.back();
        free_slots.pop_back();
        WB_TRY(wb_secant(ctx, a_best, d, gb, g, S(pending), Y(pending), n));
        const double df = fabs(f_best - f);
        std::swap(x, xb);   // x <- best trial point, g <- its gradient (the old ones become spares)
        std::swap(g, gb);
        f = f_best;
        it++;
        if (df <= f_tol * fabs(f)) break;
    }
    if (x != (double*)dx)
        WB_CUDA(ctx, cudaMemcpyAsync(dx, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (f_final) *f_final = f;
    if (iters_out) *iters_out = it;
    if (n_fg_out) *n_fg_out = P.n_fg;
    return WB200_OK;
}
