// optimizer.cu — device-resident Riemannian L-BFGS for max_localize / disentangle / opt_rotate.
//
// Replaces the Optim.optimize(only_fg!(fg!), x0, LBFGS(manifold, HagerZhang, m), Options(...)) calls
// at src/localization/max_localize.jl:73-84, src/localization/disentangle.jl:702-713 and
// src/localization/opt_rotate.jl:118-141.  Optim.jl / LineSearches.jl are third-party code that is
// not in the reference tree; this driver follows their published algorithms step for step, exactly as
// oracle/optim_lbfgs.py restates them (that restatement reproduces the reference's trajectory-
// dependent max_iter = 4 goldens to 1e-14, tests/test_oracle_golden.py):
//   * optimize loop: initial |g|_inf test, then update_state! -> update_g! -> assess_convergence ->
//     update_h!; f_tol is relative (Options.f_reltol) and must hold on successive_f_tol + 1 = 2
//     consecutive iterations, g_tol is absolute on the projected gradient, a zero step ends the run;
//   * LBFGS on a manifold: two-loop recursion with the Nocedal-Wright (7.20) scaling, direction
//     projected at x, pairs dx = alpha s (before retraction), dg = g_new - g_previous (each projected
//     at its own point, not transported), rejected only when <dx, dg> = 0; a non-descent direction
//     resets to steepest descent;
//   * alphaguess InitialStatic (alpha = 1), line search LineSearches.HagerZhang with its defaults;
//   * phi(a) = f(retract(x + a s)), phi'(a) = <project(grad f, retract(x + a s)), s>.
//
// x, g, the search direction and the 2m history vectors never leave HBM.  The two-loop recursion is
// done in COEFFICIENT space: the Gram matrix of {g, dx_i, dg_i} is kept on the host (the rows of g
// and of the newest pair are refreshed by ONE fused pass over memory per iteration, wb_cross_launch),
// the recursion runs on those 7 x 7 numbers, and the search direction is ONE linear-combination pass
// (wb_lincomb).  In k-sharded runs every scalar the control flow depends on is all-reduced, so all
// ranks take the same branches; a rank-local failure of the retraction poisons that evaluation's
// sums with NaN (f = NaN on every rank), which the line search treats as a rejected trial point.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <deque>
#include <functional>
#include <limits>

int wb_fg_device(wb200_ctx* ctx, const cplx* dU, double* dres, cplx* dG, int exact_split);  // api.cu
int wb_fg_rotate_device(wb200_ctx* ctx, const cplx* dW, double* dres, cplx* dGW);           // extras.cu

namespace {

// WB200_TRACE: wall-clock share of each phase of the driver (the stream is drained at the end of every scope,
// so this mode is for attribution only, not for timing the driver)
struct Trace {
    bool on = getenv("WB200_TRACE") != nullptr && strcmp(getenv("WB200_TRACE"), "coarse") != 0;
    double t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long n[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    static double now() {
        timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec + 1e-9 * ts.tv_nsec;
    }
    void report() const {
        if (!on) return;
        static const char* names[8] = {"retract", "fg (rotation, spread, gradient)", "project g", "inner product + sync",
                                       "direction (lincomb)", "project s", "trial point (lincomb)", "secant + Gram refresh"};
        for (int i = 0; i < 8; i++)
            if (n[i]) fprintf(stderr, "[wb200] %-34s %6ld calls %9.3f ms total %8.3f ms each\n", names[i], n[i], 1e3 * t[i], 1e3 * t[i] / n[i]);
    }
};
// WB200_TRACE=coarse: host wall clock of the three parts of an iteration as they are (no added waits)
struct Coarse {
    bool on = getenv("WB200_TRACE") && !strcmp(getenv("WB200_TRACE"), "coarse");
    double t[4] = {0, 0, 0, 0}; long n[4] = {0, 0, 0, 0};
    void add(int i, double t0) { if (on) { t[i] += Trace::now() - t0; n[i]++; } }
    void report() const {
        if (!on) return;
        static const char* names[4] = {"direction + project + dphi0 (1 wait)", "trial point (1 wait)", "secant + Gram refresh (1 wait)", "whole run"};
        for (int i = 0; i < 4; i++)
            if (n[i]) fprintf(stderr, "[wb200] %-40s %6ld x %9.3f us\n", names[i], n[i], 1e6 * t[i] / n[i]);
    }
};
struct TraceScope {
    Trace& tr; int id; cudaStream_t st; double t0 = 0.0;
    TraceScope(Trace& t, int i, cudaStream_t s) : tr(t), id(i), st(s) { if (tr.on) { cudaStreamSynchronize(st); t0 = Trace::now(); } }
    ~TraceScope() { if (tr.on) { cudaStreamSynchronize(st); tr.t[id] += Trace::now() - t0; tr.n[id]++; } }
};

// ---------------------------------------------------------------------------------------------
// LineSearches.HagerZhang (Hager & Zhang, ACM TOMS 32 (2006), Algorithm 851) with the package
// defaults; indices are 0-based.  phidphi(a, phi, dphi) returns a WB200 status.
// ---------------------------------------------------------------------------------------------
struct HagerZhang {
    double delta = 0.1, sigma = 0.9, alphamax = INFINITY, rho = 5.0, epsilon = 1e-6, gamma = 0.66, psi3 = 0.1;
    int linesearchmax = 50;
    std::function<int(double, double&, double&)> phidphi;
    std::vector<double> alphas, values, slopes;
    double phi_lim = 0.0;
    std::string msg;

    enum { LS_OK = 0, LS_EXCEPTION = 1 };   // LS_EXCEPTION: LineSearchException (alpha still returned)

    static bool fin(double a, double b) { return std::isfinite(a) && std::isfinite(b); }
    static double eps_of(double x) { return std::nextafter(std::fabs(x), INFINITY) - std::fabs(x); }
    static double nextfloat(double x) { return std::nextafter(x, INFINITY); }
    bool wolfe(double c, double phi_c, double dphi_c, double phi_0, double dphi_0) const {
        const bool w1 = delta * dphi_0 >= (phi_c - phi_0) / c && dphi_c >= sigma * dphi_0;
        const bool w2 = (2 * delta - 1) * dphi_0 >= dphi_c && dphi_c >= sigma * dphi_0 && phi_c <= phi_lim;
        return w1 || w2;
    }
    int eval(double c, double& p, double& d) { return phidphi(c, p, d); }
    void push(double c, double p, double d) { alphas.push_back(c); values.push_back(p); slopes.push_back(d); }

    // status < 0: WB200 error; LS_OK / LS_EXCEPTION otherwise
    int run(double c, double phi_0, double dphi_0, double& alpha_out, double& phi_out) {
        const double eps64 = std::numeric_limits<double>::epsilon();
        alpha_out = 0.0; phi_out = phi_0;
        if (!fin(phi_0, dphi_0)) { msg = "Value and slope at step length = 0 must be finite."; return LS_EXCEPTION; }
        if (dphi_0 >= eps64 * std::fabs(phi_0)) { msg = "Search direction is not a direction of descent."; return LS_EXCEPTION; }
        if (dphi_0 >= 0.0) return LS_OK;
        double amax = alphamax;
        const int iterfinitemax = (int)std::ceil(-std::log2(eps64));
        alphas = {0.0}; values = {phi_0}; slopes = {dphi_0};
        phi_lim = phi_0 + epsilon * std::fabs(phi_0);
        if (c <= eps64) return LS_OK;
        double phi_c, dphi_c;
        WB_TRY(eval(c, phi_c, dphi_c));
        int iterfinite = 1;
        while (!fin(phi_c, dphi_c) && iterfinite < iterfinitemax) {
            iterfinite++;
            c *= psi3;
            WB_TRY(eval(c, phi_c, dphi_c));
        }
        if (!fin(phi_c, dphi_c)) return LS_OK;   // "Failed to achieve finite new evaluation point, using alpha=0"
        push(c, phi_c, dphi_c);
        // (mayterminate is never set by InitialStatic: the first point is not tested for Wolfe here)
        bool isbracketed = false;
        int ia = 0, ib = 1, iter = 1;
        double cold = -1.0, phi_cold = phi_0;
        while (!isbracketed && iter < linesearchmax) {
            if (dphi_c >= 0.0) {
                ib = (int)alphas.size() - 1;
                for (int i = ib - 1; i >= 0; i--)
                    if (values[i] <= phi_lim) { ia = i; break; }
                isbracketed = true;
            } else if (values.back() > phi_lim) {
                ib = (int)alphas.size() - 1;
                ia = 0;
                WB_TRY(bisect(ia, ib));
                isbracketed = true;
            } else {
                cold = c; phi_cold = phi_c;
                if (nextfloat(cold) >= amax) { alpha_out = cold; phi_out = phi_cold; return LS_OK; }
                c *= rho;
                if (c > amax) c = amax;
                WB_TRY(eval(c, phi_c, dphi_c));
                iterfinite = 1;
                while (!fin(phi_c, dphi_c) && c > nextfloat(cold) && iterfinite < iterfinitemax) {
                    amax = c;
                    iterfinite++;
                    c = (cold + c) / 2;
                    WB_TRY(eval(c, phi_c, dphi_c));
                }
                if (!fin(phi_c, dphi_c)) { alpha_out = cold; phi_out = phi_cold; return LS_OK; }
                push(c, phi_c, dphi_c);
            }
            iter++;
        }
        while (iter < linesearchmax) {
            const double a = alphas[ia], b = alphas[ib];
            if (!(b > a)) { msg = "HagerZhang: bracket lost (b <= a)"; return WB200_ERR_NOCONV; }
            if (b - a <= eps_of(b)) { alpha_out = a; phi_out = values[ia]; return LS_OK; }
            bool iswolfe = false;
            int iA = ia, iB = ib;
            WB_TRY(secant2(ia, ib, phi_0, dphi_0, iswolfe, iA, iB));
            if (iswolfe) { alpha_out = alphas[iA]; phi_out = values[iA]; return LS_OK; }
            const double A = alphas[iA], B = alphas[iB];
            if (!(B > A)) { msg = "HagerZhang: bracket lost (B <= A)"; return WB200_ERR_NOCONV; }
            if (B - A < gamma * (b - a)) {
                if (nextfloat(values[ia]) >= values[ib] && nextfloat(values[iA]) >= values[iB]) {
                    alpha_out = A; phi_out = values[iA];   // so flat that secant did nothing useful
                    return LS_OK;
                }
                ia = iA; ib = iB;
            } else {
                c = (A + B) / 2;
                WB_TRY(eval(c, phi_c, dphi_c));
                if (!fin(phi_c, dphi_c)) { msg = "HagerZhang: non-finite value inside the bracket"; return WB200_ERR_NOCONV; }
                push(c, phi_c, dphi_c);
                ia = iA; ib = iB;
                WB_TRY(update(ia, ib, (int)alphas.size() - 1));
            }
            iter++;
        }
        msg = "Linesearch failed to converge, reached maximum iterations.";
        alpha_out = alphas[ia]; phi_out = values[ia];
        return LS_EXCEPTION;
    }

    int secant2(int ia, int ib, double phi_0, double dphi_0, bool& iswolfe, int& iA, int& iB) {
        const double a = alphas[ia], b = alphas[ib], da = slopes[ia], db = slopes[ib];
        iswolfe = false;
        if (!(da < 0.0 && db >= 0.0)) { msg = "HagerZhang: search direction is not a direction of descent"; return WB200_ERR_NOCONV; }
        double c = (a * db - b * da) / (db - da);
        double phi_c, dphi_c;
        WB_TRY(eval(c, phi_c, dphi_c));
        if (!fin(phi_c, dphi_c)) { msg = "HagerZhang: non-finite value in secant2"; return WB200_ERR_NOCONV; }
        push(c, phi_c, dphi_c);
        int ic = (int)alphas.size() - 1;
        if (wolfe(c, phi_c, dphi_c, phi_0, dphi_0)) { iswolfe = true; iA = iB = ic; return WB200_OK; }
        iA = ia; iB = ib;
        WB_TRY(update(iA, iB, ic));
        const double a2 = alphas[iA], b2 = alphas[iB];
        if (iB == ic) c = (alphas[ib] * slopes[iB] - alphas[iB] * slopes[ib]) / (slopes[iB] - slopes[ib]);
        else if (iA == ic) c = (alphas[ia] * slopes[iA] - alphas[iA] * slopes[ia]) / (slopes[iA] - slopes[ia]);
        if ((iA == ic || iB == ic) && a2 <= c && c <= b2) {
            WB_TRY(eval(c, phi_c, dphi_c));
            if (!fin(phi_c, dphi_c)) { msg = "HagerZhang: non-finite value in secant2"; return WB200_ERR_NOCONV; }
            push(c, phi_c, dphi_c);
            ic = (int)alphas.size() - 1;
            if (wolfe(c, phi_c, dphi_c, phi_0, dphi_0)) { iswolfe = true; iA = iB = ic; return WB200_OK; }
            WB_TRY(update(iA, iB, ic));
        }
        return WB200_OK;
    }

    // (ia, ib) in/out
    int update(int& ia, int& ib, int ic) {
        const double a = alphas[ia], b = alphas[ib], c = alphas[ic];
        if (c < a || c > b) return WB200_OK;
        if (slopes[ic] >= 0.0) { ib = ic; return WB200_OK; }
        if (values[ic] <= phi_lim) { ia = ic; return WB200_OK; }
        ib = ic;
        return bisect(ia, ib);
    }

    int bisect(int& ia, int& ib) {
        double a = alphas[ia], b = alphas[ib];
        while (b - a > eps_of(b)) {
            const double d = (a + b) / 2;
            double phi_d, g;
            WB_TRY(eval(d, phi_d, g));
            if (!fin(phi_d, g)) { msg = "HagerZhang: non-finite value in bisect"; return WB200_ERR_NOCONV; }
            push(d, phi_d, g);
            const int id = (int)alphas.size() - 1;
            if (g >= 0.0) { ib = id; return WB200_OK; }
            if (phi_d <= phi_lim) { a = d; ia = id; } else { b = d; ib = id; }
        }
        return WB200_OK;
    }
};

struct Problem {
    wb200_ctx* ctx;
    int mode;     // 0: U on (Stiefel nb x nw)^nk ; 1: XY on (Stiefel nw x nw  x  Stiefel nb x nw)^nk ;
                  // 2: one global W on Stiefel nw x nw (opt_rotate)
    size_t n;     // doubles per vector
    int n_fg = 0;
    const WbCustomProblem* custom = nullptr;   // mode 3
    Trace trace;
    // mode 0, large n: S = s^H s of the current search direction (every trial point of the line search has
    // X^H X = I + alpha^2 S, which saves the first GEMM and the first host round trip of each retraction)
    cplx* dS = nullptr;
    WbTangentGram gram{nullptr, 0.0, 0.0};
    int prepare_direction(const cplx* s) {
        gram.S = nullptr;
        if (mode != 0) return WB200_OK;
        if (!dS) {
            if (wb_small_applicable(ctx->nb, ctx->nw) || (ctx->flags & WB200_FLAG_NO_TENSOR) ||
                !wb_zgemm_tensor_applicable(ctx->nw, ctx->nw, ctx->nb)) return WB200_OK;
            WB_CUDA(ctx, wb_dev_alloc(&dS, sizeof(cplx) * (size_t)ctx->nk * ctx->nw * ctx->nw));
        }
        bool ok = false;
        double fro = 0.0;
        WB_TRY(wb_tangent_gram(ctx, s, ctx->nb, ctx->nw, ctx->nk, dS, &fro, &ok));
        if (ok) { gram.S = dS; gram.fro_max = fro; }
        return WB200_OK;
    }
    ~Problem() { if (dS) wb_dev_free(dS); }

    // tangent_step: x = x0 + a s with x0 on the manifold and s tangent at x0 (every trial point of a line search)
    // alpha > 0: x = x0 + alpha s with s the direction prepare_direction() saw
    int retract(cplx* x, bool tangent_step = false, double alpha = 0.0) {
        if (mode == 3) return custom->retract(x, tangent_step);
        if (mode == 2) return wb_retract_dev(ctx, x, ctx->nw, ctx->nw, 1, nullptr, tangent_step);
        if (mode == 0) {
            gram.alpha = alpha;
            return wb_retract_dev(ctx, x, ctx->nb, ctx->nw, ctx->nk, nullptr, tangent_step,
                                  (tangent_step && alpha > 0.0 && gram.S) ? &gram : nullptr);
        }
        return wb_retract_xy(ctx, x, tangent_step);
    }
    int project(const cplx* x, cplx* g) {
        if (mode == 3) return custom->project(x, g);
        if (mode == 2) return wb_project_tangent_dev(ctx, x, g, ctx->nw, ctx->nw, 1);
        if (mode == 0) return wb_project_tangent_dev(ctx, x, g, ctx->nb, ctx->nw, ctx->nk);
        return wb_project_xy(ctx, x, g);
    }
    // launches f and the PROJECTED gradient at x, and (optionally) <g, d> in the same submission;
    // one synchronisation (fg_finish) returns f, min |N_nn| and the inner product
    int fg_enqueue(const cplx* x, cplx* g, const double* d) {
        n_fg++;
        {
        TraceScope ts_(trace, 1, ctx->stream);
        if (mode == 3) {
            WB_TRY(custom->eval(x, g));
        } else if (mode == 0) {
            WB_TRY(wb_fg_device(ctx, x, ctx->d_res, g, 0));
        } else if (mode == 2) {
            WB_TRY(wb_fg_rotate_device(ctx, x, ctx->d_res, g));
        } else {
            WB_TRY(wb_xy_to_u(ctx, x, ctx->d_U));
            WB_TRY(wb_fg_device(ctx, ctx->d_U, ctx->d_res, ctx->d_G, 0));
            WB_TRY(wb_gu_to_gxy(ctx, x, ctx->d_G, g));
        }
        }
        // f, min |N_nn| and the deferred retraction flag reach the host with the inner product's results (zero-copy
        // stores of the last kernel, vecops.cu) when this evaluation has one and the run is a single slab; by copies
        // otherwise
        const bool published = d && !ctx->allreduce;
        if (!published) {
            WB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 16, ctx->d_res + 5, sizeof(double),
                                         cudaMemcpyDeviceToHost, ctx->stream));
            WB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 17, ctx->d_min, sizeof(double),
                                         cudaMemcpyDeviceToHost, ctx->stream));
            if (ctx->retract_pending)
                WB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 8, ctx->d_scal + 4, sizeof(int),
                                             cudaMemcpyDeviceToHost, ctx->stream));
        }
        {
            TraceScope ts_(trace, 2, ctx->stream);
            WB_TRY(project(x, g));
        }
        TraceScope ts_(trace, 3, ctx->stream);
        if (d) {
            WbCrossArgs c{};
            c.a[0] = (const double*)g; c.na = 1;
            c.b[0] = d; c.nb = 1;
            ctx->publish_eval = published ? 1 : 0;
            const int st_ = wb_cross_launch(ctx, c, n);
            ctx->publish_eval = 0;
            WB_TRY(st_);
        }
        return WB200_OK;
    }
    int fg_finish(double* f, const double* d, double* gd) {
        WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *f = ctx->h_pinned[16];
        if (d) *gd = ctx->h_pinned[WB_PIN_CROSS];
        if (wb_retract_pending_failed(ctx)) {   // the deferred check of the retraction in front of this evaluation:
            *f = NAN;                           // a point that is not on the manifold is a rejected trial point
            if (d) *gd = NAN;
        }
        if (mode == 2 && !ctx->allreduce && ctx->h_pinned[17] < 1e-10) {   // opt_rotate.jl:60-62
            ctx->err = "N^kb too small! (|N_nn| < 1e-10)";
            return WB200_ERR_SINGULAR;
        }
        return WB200_OK;
    }
    int fg(const cplx* x, double* f, cplx* g, const double* d, double* gd) {
        WB_TRY(fg_enqueue(x, g, d));
        return fg_finish(f, d, gd);
    }

    // ---- one trial point of the line search: xn = retract(x + a s), gn = the projected gradient there,
    //      phi = f(xn), dphi = <gn, s> ----
    // Small problems (the reference's own benchmark sizes) are pure launch latency: ~12 dependent kernels of a few
    // microseconds each.  On a single slab with the fused small-shape retraction the whole chain — step length
    // included, which a 8-byte copy node brings from page-locked memory — is captured ONCE per (x, s, xn, gn)
    // buffer combination into a CUDA graph and replayed: one submission and one wait per trial point.
    bool trial_graphs_usable() const {
        return ctx->use_graphs && !trace.on && !ctx->profiling && !ctx->allreduce && ctx->retract_deferred && mode != 3 &&
               ctx->stream != nullptr && !gram.S && wb_small_applicable(ctx->nb, ctx->nw);
    }
    int trial_enqueue(const double* x, const double* s, double* xn, double* gn, double a, bool alpha_from_device) {
        WbLinArgs l{};
        l.v[0] = x; l.c[0] = 1.0; l.v[1] = s; l.c[1] = a; l.nv = 2;
        if (alpha_from_device) {
            WB_CUDA(ctx, cudaMemcpyAsync(ctx->d_scal + 16, ctx->h_pinned + 20, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            l.cdev = ctx->d_scal + 16; l.cdev_idx = 1;
        }
        {
            TraceScope ts_(trace, 6, ctx->stream);
            WB_TRY(wb_lincomb(ctx, l, xn, n));
        }
        {
            TraceScope ts_(trace, 0, ctx->stream);
            int st = retract((cplx*)xn, true, a);
            if (st == WB200_ERR_NOCONV) ctx->poison_next = 1;   // this trial point is rejected on every rank
            else WB_TRY(st);
        }
        return fg_enqueue((const cplx*)xn, (cplx*)gn, s);
    }
    // tiny problems: the whole trial point is one cooperative launch (tiny.cu)
    bool tiny_usable() { return !trace.on && mode != 3 && wb_tiny_usable(ctx, mode); }
    int trial(const double* x, const double* s, double* xn, double* gn, double a, double* phi, double* dphi) {
        if (tiny_usable()) {
            n_fg++;
            WB_TRY(wb_tiny_trial(ctx, mode, (const cplx*)x, (const cplx*)s, a, (cplx*)xn, (cplx*)gn));
            const int st_ = fg_finish(phi, s, dphi);
#ifdef WB_TINY_TIMING
            fprintf(stderr, "[tiny] a=%g ns:", a);
            for (int i = 0; i < 7; i++) fprintf(stderr, " %6.0f", ctx->h_pinned[48 + i]);
            fprintf(stderr, "\n");
#endif
            return st_;
        }
        if (!trial_graphs_usable()) {
            WB_TRY(trial_enqueue(x, s, xn, gn, a, false));
            return fg_finish(phi, s, dphi);
        }
        wb200_ctx::TrialGraph* slot = nullptr;
        for (auto& g : ctx->tgraphs)
            if (g.seen && g.x == x && g.s == s && g.xn == xn && g.gn == gn && g.mode == mode && g.has_penalty == ctx->has_penalty &&
                g.lambda == ctx->lambda && g.frozen == ctx->d_frozen) { slot = &g; break; }
        ctx->h_pinned[20] = a;
        if (slot && slot->exec) {
            n_fg++;
            ctx->retract_pending = 1;           // the graph contains a deferred retraction; its flag is published
            WB_CUDA(ctx, cudaGraphLaunch(slot->exec, ctx->stream));
            ctx->launches += slot->nlaunch;
            slot->stamp = ++ctx->graph_clock;
            return fg_finish(phi, s, dphi);
        }
        if (!slot) {   // first sighting: run plainly (lazy allocations, function attributes), remember the key
            slot = &ctx->tgraphs[0];
            for (auto& g : ctx->tgraphs) if (g.stamp < slot->stamp) slot = &g;
            if (slot->exec) { cudaGraphExecDestroy(slot->exec); slot->exec = nullptr; }
            *slot = wb200_ctx::TrialGraph{};
            slot->x = x; slot->s = s; slot->xn = xn; slot->gn = gn; slot->mode = mode; slot->has_penalty = ctx->has_penalty;
            slot->lambda = ctx->lambda; slot->frozen = ctx->d_frozen; slot->seen = 1; slot->stamp = ++ctx->graph_clock;
            WB_TRY(trial_enqueue(x, s, xn, gn, a, true));
            return fg_finish(phi, s, dphi);
        }
        // second sighting: capture, instantiate, launch
        const int saved_graphs = ctx->use_graphs;
        const int saved_nfg = n_fg;
        const long long l0 = ctx->launches;
        ctx->use_graphs = 0;                    // the evaluation inside is captured as plain launches
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed);
        int st = WB200_OK;
        if (ce == cudaSuccess) {
            st = trial_enqueue(x, s, xn, gn, a, true);
            ce = cudaStreamEndCapture(ctx->stream, &graph);
        }
        ctx->use_graphs = saved_graphs;
        n_fg = saved_nfg;
        ctx->retract_pending = 0;
        if (ce == cudaSuccess && st == WB200_OK && graph &&
            cudaGraphInstantiate(&slot->exec, graph, 0) == cudaSuccess) {
            slot->nlaunch = ctx->launches - l0;
            ctx->launches = l0;
        } else {
            slot->exec = nullptr;
            cudaGetLastError();
            ctx->use_graphs = 0;                // graphs are not usable here: plain launches from now on
        }
        if (graph) cudaGraphDestroy(graph);
        if (st != WB200_OK) return st;
        return trial(x, s, xn, gn, a, phi, dphi);   // replays the new graph (or runs plainly if it could not be built)
    }
};

}  // namespace

static int lbfgs_run(wb200_ctx* ctx, int mode, const WbCustomProblem* custom, cplx* dx, double f_tol, double g_tol,
                     int max_iter, int history, double* f_final, int* iters_out, int* n_fg_out);

int wb_lbfgs(wb200_ctx* ctx, int mode, cplx* dx, double f_tol, double g_tol, int max_iter,
             int history, double* f_final, int* iters_out, int* n_fg_out) {
    return lbfgs_run(ctx, mode, nullptr, dx, f_tol, g_tol, max_iter, history, f_final, iters_out, n_fg_out);
}
int wb_lbfgs_custom(wb200_ctx* ctx, const WbCustomProblem& cp, cplx* dx, double f_tol, double g_tol,
                    int max_iter, int history, double* f_final, int* iters_out, int* n_fg_out) {
    return lbfgs_run(ctx, 3, &cp, dx, f_tol, g_tol, max_iter, history, f_final, iters_out, n_fg_out);
}

static int lbfgs_run(wb200_ctx* ctx, int mode, const WbCustomProblem* custom, cplx* dx, double f_tol, double g_tol,
                     int max_iter, int history, double* f_final, int* iters_out, int* n_fg_out) {
    Problem P{ctx, mode, 0};
    P.custom = custom;
    Coarse coarse;
    const double t_run0 = Trace::now();
    // single slab: the convergence flag of a small-shape retraction is read after the evaluation's wait (one host round
    // trip per trial point); k-sharded runs keep the immediate check, every rank must poison the same evaluation
    struct Deferred {
        wb200_ctx* c;
        ~Deferred() { c->retract_deferred = 0; c->retract_pending = 0; }
    } deferred_guard{ctx};
    ctx->retract_deferred = (mode != 3 && !ctx->allreduce) ? 1 : 0;
    ctx->retract_pending = 0;
    const size_t ncplx = (mode == 3) ? custom->ncplx : (mode == 2) ? (size_t)ctx->nw * ctx->nw
                         : (mode == 0) ? (size_t)ctx->nk * ctx->nb * ctx->nw
                                       : (size_t)ctx->nk * ((size_t)ctx->nw * ctx->nw + (size_t)ctx->nb * ctx->nw);
    P.n = 2 * ncplx;
    const size_t n = P.n;
    const int m = history > 0 ? history : 3;
    const int successive_f_tol = 1;   // Optim.Options default

    // work vectors: g, s, xn, gn + m history pairs
    const int nfixed = 4;
    const int nvec = nfixed + 2 * m;
    double* pool = nullptr;
    WB_CUDA(ctx, wb_dev_alloc(&pool, sizeof(double) * n * nvec));
    struct Free { double* p; ~Free() { wb_dev_free(p); } } guard{pool};
    auto vec = [&](int i) { return pool + (size_t)i * n; };
    // x lives in the caller's array or in a spare (pointer swaps instead of copies)
    double *x = (double*)dx, *g = vec(0), *s = vec(1), *xn = vec(2), *gn = vec(3);
    auto DX = [&](int sl) { return vec(nfixed + 2 * sl); };
    auto DG = [&](int sl) { return vec(nfixed + 2 * sl + 1); };
    std::deque<int> slots;            // live history slots, oldest first (Optim's ring mod1(pseudo, m))
    std::vector<int> free_slots;
    for (int i = m - 1; i >= 0; i--) free_slots.push_back(i);

    // Gram matrix of the vectors {g, DX(0..m-1), DG(0..m-1)} by id: 0 = g, 1 + sl = DX(sl), 1 + m + sl = DG(sl)
    const int nid = 1 + 2 * m;
    std::vector<double> GM((size_t)nid * nid, 0.0);
    auto idX = [&](int sl) { return 1 + sl; };
    auto idG = [&](int sl) { return 1 + m + sl; };
    auto ptr_of = [&](int id) -> const double* {
        return id == 0 ? g : (id <= m ? DX(id - 1) : DG(id - 1 - m));
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

    // ---- initial_state: retract!, value_gradient!!, project_tangent! ----
    double f = 0.0;
    if (P.tiny_usable()) {   // retraction, evaluation and projection in one launch
        P.n_fg++;
        WB_TRY(wb_tiny_trial(ctx, mode, dx, nullptr, 0.0, dx, (cplx*)g));
        WB_TRY(P.fg_finish(&f, nullptr, nullptr));
    } else {
        {
            TraceScope ts_(P.trace, 0, ctx->stream);
            int st = P.retract(dx);
            if (st == WB200_ERR_NOCONV && ctx->allreduce) ctx->poison_next = 1;   // peers must see the failure too
            else WB_TRY(st);
        }
        WB_TRY(P.fg(dx, &f, (cplx*)g, nullptr, nullptr));
    }
    if (!std::isfinite(f)) {
        ctx->err = "initial point: spread is not finite (retraction failed or singular overlaps)";
        return WB200_ERR_NOCONV;
    }
    double gmax = 0.0;
    WB_TRY(refresh({0}, 1, &gmax));
    bool converged = gmax <= g_tol;   // initial_convergence
    int it = 0, counter_f_tol = 0;
    bool hist_reset = false;
    while (!converged && it < max_iter) {
        it++;
        const double t_dir0 = Trace::now();
        // ---- update_state!: two-loop recursion on the coefficients of q = sum_id c[id] v_id ----
        std::vector<double> c(nid, 0.0);
        c[0] = 1.0;
        auto dotq = [&](int id) {
            double t = 0.0;
            for (int j = 0; j < nid; j++) if (c[j] != 0.0) t += c[j] * GM[(size_t)id * nid + j];
            return t;
        };
        auto rho = [&](int sl) { return 1.0 / GM[(size_t)idX(sl) * nid + idG(sl)]; };
        std::vector<double> al;
        for (auto itS = slots.rbegin(); itS != slots.rend(); ++itS) {
            const double a = rho(*itS) * dotq(idX(*itS));
            al.push_back(a);
            c[idG(*itS)] -= a;
        }
        if (!slots.empty()) {   // Nocedal & Wright (7.20) with the newest pair
            const int last = slots.back();
            const double scaling = GM[(size_t)idX(last) * nid + idG(last)] / GM[(size_t)idG(last) * nid + idG(last)];
            for (double& v : c) v *= scaling;
        }
        {
            int j = (int)al.size() - 1;
            for (auto itS = slots.begin(); itS != slots.end(); ++itS, --j) {
                const double b = rho(*itS) * dotq(idG(*itS));
                c[idX(*itS)] += al[j] - b;
            }
        }
        {   // s = -q in one pass (chunks of 7 vectors), then project_tangent!(s, x)
            TraceScope ts_(P.trace, 4, ctx->stream);
            std::vector<int> used;
            for (int id = 0; id < nid; id++) if (c[id] != 0.0) used.push_back(id);
            if (used.empty()) used.push_back(0);
            for (size_t off = 0; off < used.size(); off += WB_CROSS_NB) {
                WbLinArgs l{};
                l.nv = (int)std::min<size_t>(WB_CROSS_NB, used.size() - off);
                for (int j = 0; j < l.nv; j++) { l.v[j] = ptr_of(used[off + j]); l.c[j] = -c[used[off + j]]; }
                l.beta = off == 0 ? 0.0 : 1.0;
                WB_TRY(wb_lincomb(ctx, l, s, n));
            }
        }
        {
            TraceScope ts_(P.trace, 5, ctx->stream);
            WB_TRY(P.project((const cplx*)x, (cplx*)s));
        }
        // ---- perform_linesearch! ----
        const double gg = GM[0];
        double dphi0 = 0.0;
        {
            WbCrossArgs cr{};
            cr.a[0] = g; cr.na = 1; cr.b[0] = s; cr.nb = 1;
            WB_TRY(wb_cross_launch(ctx, cr, n));
            WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                dphi0 = ctx->h_pinned[WB_PIN_CROSS];
        }
        hist_reset = false;
        if (dphi0 >= 0.0) {   // reset_search_direction!: steepest descent, pseudo_iteration = 1
            WbLinArgs l{};
            l.v[0] = g; l.c[0] = -1.0; l.nv = 1;
            WB_TRY(wb_lincomb(ctx, l, s, n));
            dphi0 = -gg;
            hist_reset = true;
        }
        WB_TRY(P.prepare_direction((const cplx*)s));
        coarse.add(0, t_dir0);
        const double f_prev = f;
        double last_alpha = -1.0, last_f = 0.0;
        HagerZhang hz;
        hz.phidphi = [&](double a, double& phi, double& dphi) -> int {
            double fv = 0.0, dv = 0.0;
            const double t_tr0 = Trace::now();
            WB_TRY(P.trial(x, s, xn, gn, a, &fv, &dv));
            coarse.add(1, t_tr0);
            phi = fv; dphi = dv;
            last_alpha = a; last_f = fv;
            return WB200_OK;
        };
        double alpha = 0.0, phi_a = f;
        const int ls = hz.run(1.0 /* InitialStatic */, f, dphi0, alpha, phi_a);
        if (ls < 0) {
            if (ctx->err.empty() || ls == WB200_ERR_NOCONV) ctx->err = hz.msg.empty() ? ctx->err : hz.msg;
            return ls;
        }
        if (alpha == 0.0) {   // x_converged: the iterate did not move (Optim counts the iteration)
            converged = true;
            break;
        }
        if (alpha != last_alpha) {   // update_g! re-evaluates when the accepted point is not the last trial
            double pa, da;
            WB_TRY(hz.phidphi(alpha, pa, da));
            phi_a = pa;
            if (!std::isfinite(pa)) {
                ctx->err = "line search returned a point with a non-finite spread";
                return WB200_ERR_NOCONV;
            }
        } else {
            phi_a = last_f;
        }
        // ---- update_h!: dx = alpha s, dg = g_new - g_previous into the slot mod1(pseudo_iteration, m) ----
        if (hist_reset) while (!slots.empty()) { free_slots.push_back(slots.front()); slots.pop_front(); }
        int sl;
        if (free_slots.empty()) { sl = slots.front(); slots.pop_front(); }
        else { sl = free_slots.back(); free_slots.pop_back(); }
        TraceScope ts_secant(P.trace, 7, ctx->stream);
        const double t_sec0 = Trace::now();
        WB_TRY(wb_secant(ctx, alpha, s, gn, g, DX(sl), DG(sl), n));
        std::swap(x, xn);   // x <- accepted point, g <- its projected gradient (the old ones become the trial buffers)
        std::swap(g, gn);
        f = phi_a;
        slots.push_back(sl);
        {   // Gram rows of g and of the new pair against everything live: one pass over memory
            std::vector<int> ids{0, idX(sl), idG(sl)};
            for (size_t i = 0; i + 1 < slots.size(); i++) ids.push_back(idX(slots[i]));
            for (size_t i = 0; i + 1 < slots.size(); i++) ids.push_back(idG(slots[i]));
            WB_TRY(refresh(ids, 3, &gmax));
        }
        coarse.add(2, t_sec0);
        const double sy = GM[(size_t)idX(sl) * nid + idG(sl)];
        if (sy == 0.0 || !std::isfinite(1.0 / sy)) {   // isinf(rho): pseudo_iteration = 0, history dropped
            while (!slots.empty()) { free_slots.push_back(slots.front()); slots.pop_front(); }
        }
        // ---- assess_convergence ----
        const bool f_converged = fabs(f - f_prev) <= f_tol * fabs(f);
        const bool g_converged = gmax <= g_tol;
        counter_f_tol = f_converged ? counter_f_tol + 1 : 0;
        converged = g_converged || counter_f_tol > successive_f_tol;
        if (ls == HagerZhang::LS_EXCEPTION) break;   // Optim leaves the loop when the line search failed
    }
    if (x != (double*)dx)
        WB_CUDA(ctx, cudaMemcpyAsync(dx, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    P.trace.report();
    coarse.add(3, t_run0);
    coarse.report();
    if (f_final) *f_final = f;
    if (iters_out) *iters_out = it;
    if (n_fg_out) *n_fg_out = P.n_fg;
    return WB200_OK;
}
