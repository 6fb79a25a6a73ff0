"""CPU restatement of the third-party optimiser the reference calls on this path (TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module).

The reference drives max_localize / disentangle / opt_rotate with

    Optim.optimize(Optim.only_fg!(fg!), x0,
                   Optim.LBFGS(; manifold, linesearch=Optim.HagerZhang(), m=history_size),
                   Optim.Options(; iterations, f_tol, g_tol, allow_f_increases=true))

(src/localization/max_localize.jl:62-84, disentangle.jl:687-713, opt_rotate.jl:118-141).  Optim.jl
(compat "1.7") and LineSearches.jl (7.x) are NOT in /root/reference; this file restates their
PUBLISHED algorithms:

  * Optim `optimize` loop for first-order methods: initial_convergence on |g|_inf, then per iteration
    update_state! -> update_g! -> assess_convergence -> update_h!; `f_tol` is Options.f_reltol,
    `g_tol` is g_abstol, x tolerances are 0, and f-convergence must be hit
    `successive_f_tol + 1 = 2` times in a row (Optim >= 1.3);
  * Optim LBFGS on a manifold: gradient projected at the iterate, two-loop recursion with the
    Nocedal-Wright (7.20) scaling, direction projected, history pairs dx = alpha*s (before
    retraction) and dg = g_new - g_previous, not transported; a pair is rejected only when
    1/<dx, dg> is infinite; non-descent directions reset to steepest descent;
  * alphaguess = LineSearches.InitialStatic(alpha = 1, scaled = false) (LBFGS's default);
  * LineSearches.HagerZhang with its defaults (delta 0.1, sigma 0.9, rho 5, epsilon 1e-6, gamma 0.66,
    linesearchmax 50, psi3 0.1): bracket (B0-B3), secant2 (S1-S4), update (U0-U3), bisect (U3a-c),
    Wolfe / approximate Wolfe tests, Hager & Zhang, ACM TOMS 32 (2006) Algorithm 851;
  * the line-search objective of a ManifoldObjective: phi(a) = f(retract(x + a s)),
    phi'(a) = Re<project_tangent(grad f, retract(x + a s)), s>.

Pinning: the reference's own tests hold TRAJECTORY-DEPENDENT goldens produced by this optimiser at
max_iter = 4 (test/localization/disentangle.jl:56-97, constrain_center/max_localize.jl:46-96,
constrain_center/disentangle.jl:53-120); tests/test_oracle_golden.py reproduces them to 1e-7 with
this restatement, which pins the restatement (a different line search or history rule lands
elsewhere after 4 iterations).
"""
from __future__ import annotations

import math

import numpy as np


class LineSearchException(Exception):
    def __init__(self, msg, alpha):
        super().__init__(msg)
        self.alpha = alpha


def _rdot(a, b):
    return float(np.real(np.vdot(a, b)))


def _eps(x):
    return float(np.spacing(abs(x))) if x != 0 else float(np.finfo(np.float64).tiny)


def _nextfloat(x):
    return float(np.nextafter(x, np.inf))


# --------------------------------------------------------------------------------------------
# LineSearches.HagerZhang
# --------------------------------------------------------------------------------------------
class HagerZhang:
    def __init__(self, delta=0.1, sigma=0.9, alphamax=math.inf, rho=5.0, epsilon=1e-6, gamma=0.66,
                 linesearchmax=50, psi3=0.1):
        self.delta, self.sigma, self.alphamax, self.rho = delta, sigma, alphamax, rho
        self.epsilon, self.gamma, self.linesearchmax, self.psi3 = epsilon, gamma, linesearchmax, psi3
        self.mayterminate = False       # only InitialHagerZhang sets it; InitialStatic never does

    def satisfies_wolfe(self, c, phi_c, dphi_c, phi_0, dphi_0, phi_lim):
        wolfe1 = self.delta * dphi_0 >= (phi_c - phi_0) / c and dphi_c >= self.sigma * dphi_0
        wolfe2 = ((2 * self.delta - 1) * dphi_0 >= dphi_c >= self.sigma * dphi_0) and phi_c <= phi_lim
        return wolfe1 or wolfe2

    def __call__(self, phidphi, c, phi_0, dphi_0):
        """returns (alpha, phi(alpha)); `phidphi(a) -> (phi, dphi)`.  Indices below are 0-based
        (LineSearches' are 1-based)."""
        eps64 = float(np.finfo(np.float64).eps)
        if not (math.isfinite(phi_0) and math.isfinite(dphi_0)):
            raise LineSearchException("Value and slope at step length = 0 must be finite.", 0.0)
        if dphi_0 >= eps64 * abs(phi_0):
            raise LineSearchException("Search direction is not a direction of descent.", 0.0)
        elif dphi_0 >= 0:
            return 0.0, phi_0
        alphamax = self.alphamax
        iterfinitemax = int(math.ceil(-math.log2(eps64)))
        alphas, values, slopes = [0.0], [phi_0], [dphi_0]
        phi_lim = phi_0 + self.epsilon * abs(phi_0)
        assert c >= 0
        if c <= eps64:
            return 0.0, phi_0
        assert math.isfinite(c) and c <= alphamax
        phi_c, dphi_c = phidphi(c)
        iterfinite = 1
        while not (math.isfinite(phi_c) and math.isfinite(dphi_c)) and iterfinite < iterfinitemax:
            self.mayterminate = False
            iterfinite += 1
            c *= self.psi3
            phi_c, dphi_c = phidphi(c)
        if not (math.isfinite(phi_c) and math.isfinite(dphi_c)):
            self.mayterminate = False
            return 0.0, phi_0
        alphas.append(c); values.append(phi_c); slopes.append(dphi_c)
        if self.mayterminate and self.satisfies_wolfe(c, phi_c, dphi_c, phi_0, dphi_0, phi_lim):
            self.mayterminate = False
            return c, phi_c
        # initial bracketing (HZ stages B0-B3)
        isbracketed = False
        ia, ib = 0, 1
        it = 1
        cold = -1.0
        phi_cold = phi_0
        while not isbracketed and it < self.linesearchmax:
            if dphi_c >= 0.0:
                ib = len(alphas) - 1
                for i in range(ib - 1, -1, -1):
                    if values[i] <= phi_lim:
                        ia = i
                        break
                isbracketed = True
            elif values[-1] > phi_lim:
                ib = len(alphas) - 1
                ia = 0
                ia, ib = self.bisect(phidphi, alphas, values, slopes, ia, ib, phi_lim)
                isbracketed = True
            else:
                cold = c
                phi_cold = phi_c
                if _nextfloat(cold) >= alphamax:
                    self.mayterminate = False
                    return cold, phi_cold
                c *= self.rho
                if c > alphamax:
                    c = alphamax
                phi_c, dphi_c = phidphi(c)
                iterfinite = 1
                while (not (math.isfinite(phi_c) and math.isfinite(dphi_c)) and c > _nextfloat(cold)
                       and iterfinite < iterfinitemax):
                    alphamax = c
                    iterfinite += 1
                    c = (cold + c) / 2
                    phi_c, dphi_c = phidphi(c)
                if not (math.isfinite(phi_c) and math.isfinite(dphi_c)):
                    return cold, phi_cold
                alphas.append(c); values.append(phi_c); slopes.append(dphi_c)
            it += 1
        while it < self.linesearchmax:
            a, b = alphas[ia], alphas[ib]
            assert b > a
            if b - a <= _eps(b):
                self.mayterminate = False
                return a, values[ia]
            iswolfe, iA, iB = self.secant2(phidphi, alphas, values, slopes, ia, ib, phi_lim)
            if iswolfe:
                self.mayterminate = False
                return alphas[iA], values[iA]
            A, B = alphas[iA], alphas[iB]
            assert B > A
            if B - A < self.gamma * (b - a):
                if _nextfloat(values[ia]) >= values[ib] and _nextfloat(values[iA]) >= values[iB]:
                    self.mayterminate = False      # so flat that secant did nothing useful
                    return A, values[iA]
                ia, ib = iA, iB
            else:
                c = (A + B) / 2
                phi_c, dphi_c = phidphi(c)
                assert math.isfinite(phi_c) and math.isfinite(dphi_c)
                alphas.append(c); values.append(phi_c); slopes.append(dphi_c)
                ia, ib = self.update(phidphi, alphas, values, slopes, iA, iB, len(alphas) - 1, phi_lim)
            it += 1
        raise LineSearchException("Linesearch failed to converge, reached maximum iterations %d." %
                                  self.linesearchmax, alphas[ia])

    def secant2(self, phidphi, alphas, values, slopes, ia, ib, phi_lim):
        phi_0, dphi_0 = values[0], slopes[0]
        a, b = alphas[ia], alphas[ib]
        dphi_a, dphi_b = slopes[ia], slopes[ib]
        if not (dphi_a < 0.0 and dphi_b >= 0.0):
            raise RuntimeError("Search direction is not a direction of descent; (dphi_a = %f; dphi_b = %f)"
                               % (dphi_a, dphi_b))
        c = (a * dphi_b - b * dphi_a) / (dphi_b - dphi_a)
        assert math.isfinite(c)
        phi_c, dphi_c = phidphi(c)
        assert math.isfinite(phi_c) and math.isfinite(dphi_c)
        alphas.append(c); values.append(phi_c); slopes.append(dphi_c)
        ic = len(alphas) - 1
        if self.satisfies_wolfe(c, phi_c, dphi_c, phi_0, dphi_0, phi_lim):
            return True, ic, ic
        iA, iB = self.update(phidphi, alphas, values, slopes, ia, ib, ic, phi_lim)
        a, b = alphas[iA], alphas[iB]
        if iB == ic:      # b was updated: make sure a is updated too
            c = (alphas[ib] * slopes[iB] - alphas[iB] * slopes[ib]) / (slopes[iB] - slopes[ib])
        elif iA == ic:    # a was updated: do it for b too
            c = (alphas[ia] * slopes[iA] - alphas[iA] * slopes[ia]) / (slopes[iA] - slopes[ia])
        if (iA == ic or iB == ic) and a <= c <= b:
            phi_c, dphi_c = phidphi(c)
            assert math.isfinite(phi_c) and math.isfinite(dphi_c)
            alphas.append(c); values.append(phi_c); slopes.append(dphi_c)
            ic = len(alphas) - 1
            if self.satisfies_wolfe(c, phi_c, dphi_c, phi_0, dphi_0, phi_lim):
                return True, ic, ic
            iA, iB = self.update(phidphi, alphas, values, slopes, iA, iB, ic, phi_lim)
        return False, iA, iB

    def update(self, phidphi, alphas, values, slopes, ia, ib, ic, phi_lim):
        a, b, c = alphas[ia], alphas[ib], alphas[ic]
        assert slopes[ia] < 0.0 and values[ia] <= phi_lim and slopes[ib] >= 0.0 and b > a
        phi_c, dphi_c = values[ic], slopes[ic]
        if c < a or c > b:
            return ia, ib
        if dphi_c >= 0.0:
            return ia, ic
        if phi_c <= phi_lim:
            return ic, ib
        return self.bisect(phidphi, alphas, values, slopes, ia, ic, phi_lim)

    def bisect(self, phidphi, alphas, values, slopes, ia, ib, phi_lim):
        a, b = alphas[ia], alphas[ib]
        assert slopes[ia] < 0.0 and values[ia] <= phi_lim and slopes[ib] < 0.0 and values[ib] > phi_lim and b > a
        while b - a > _eps(b):
            d = (a + b) / 2
            phi_d, gphi = phidphi(d)
            assert math.isfinite(phi_d) and math.isfinite(gphi)
            alphas.append(d); values.append(phi_d); slopes.append(gphi)
            idd = len(alphas) - 1
            if gphi >= 0.0:
                return ia, idd
            if phi_d <= phi_lim:
                a, ia = d, idd
            else:
                b, ib = d, idd
        return ia, ib


# --------------------------------------------------------------------------------------------
# Optim.optimize with LBFGS(manifold, HagerZhang, m) and Options(f_tol, g_tol, iterations)
# --------------------------------------------------------------------------------------------
def optim_lbfgs(fg, x0, retract, project, f_tol=1e-7, g_tol=1e-5, max_iter=200, m=3,
                successive_f_tol=1, verbose=False, trace=None):
    """fg(x) -> (f, grad) Euclidean; retract(x) -> point on the manifold; project(g, x) -> tangent
    part of g at x.  Returns (x, f, iterations, info)."""
    n_fg = 0

    def value_gradient(xa):          # ManifoldObjective: value_gradient!(inner, retract(x)); project
        nonlocal n_fg
        n_fg += 1
        xin = retract(xa)
        f_, g_ = fg(xin)
        return f_, project(g_, xin), xin

    x = retract(np.array(x0, dtype=np.complex128, copy=True))       # initial_state
    f, g = fg(x)
    n_fg += 1
    g = project(g, x)
    dx_hist = [None] * m
    dg_hist = [None] * m
    rho = [math.nan] * m
    pseudo = 0
    it = 0
    counter_f_tol = 0
    converged = float(np.max(np.abs(g))) <= g_tol                   # initial_convergence
    ls_failed = False
    f_prev = math.nan
    stop_reason = "g_tol (initial)" if converged else "iterations"
    while not converged and it < max_iter:
        it += 1
        # ---- update_state! ----
        pseudo += 1
        g = project(g, x)
        # twoloop!
        lower, upper = pseudo - m, pseudo - 1
        q = g.copy()
        alpha_tl = [0.0] * m
        for index in range(upper, lower - 1, -1):
            if index < 1:
                continue
            i = (index - 1) % m
            alpha_tl[i] = rho[i] * _rdot(dx_hist[i], q)
            q = q - alpha_tl[i] * dg_hist[i]
        if pseudo > 1:
            i = (upper - 1) % m
            scaling = _rdot(dx_hist[i], dg_hist[i]) / float(np.sum(np.abs(dg_hist[i]) ** 2))
            s = scaling * q
        else:
            s = q.copy()
        for index in range(lower, upper + 1):
            if index < 1:
                continue
            i = (index - 1) % m
            beta = rho[i] * _rdot(dg_hist[i], s)
            s = s + dx_hist[i] * (alpha_tl[i] - beta)
        s = -s
        s = project(s, x)
        g_previous = g.copy()
        # perform_linesearch!
        dphi_0 = _rdot(g, s)
        if dphi_0 >= 0.0:                       # reset_search_direction!
            pseudo = 1
            s = -g
            dphi_0 = _rdot(g, s)
        phi_0 = f
        alpha0 = 1.0                            # InitialStatic(alpha = 1.0, scaled = false)
        f_prev = phi_0
        x_previous = x
        cache = {}

        def phidphi(a):
            f_, g_, xin = value_gradient(x + a * s)
            cache["last"] = (a, f_, g_, xin)
            return f_, _rdot(g_, s)

        ls = HagerZhang()
        try:
            alpha, _ = ls(phidphi, alpha0, phi_0, dphi_0)
            ls_ok = True
        except LineSearchException as ex:
            alpha = ex.alpha
            ls_ok = False
        dx = alpha * s
        x_new = retract(x + dx)
        # ---- update_g! (NLSolversBase caches the last evaluation point) ----
        if "last" in cache and cache["last"][0] == alpha:
            _, f_new, g_new, _ = cache["last"]
        else:
            f_new, g_new = fg(x_new)
            n_fg += 1
            g_new = project(g_new, x_new)
        x, f, g = x_new, f_new, g_new
        if not ls_ok:
            ls_failed = True
            stop_reason = "line search failed"
            break
        # ---- assess_convergence ----
        x_converged = float(np.max(np.abs(x - x_previous))) <= 0.0
        f_converged = abs(f - f_prev) <= f_tol * abs(f)
        g_converged = float(np.max(np.abs(g))) <= g_tol
        counter_f_tol = counter_f_tol + 1 if f_converged else 0
        converged = x_converged or g_converged or counter_f_tol > successive_f_tol
        if converged:
            stop_reason = "x" if x_converged else ("g_tol" if g_converged else "f_tol")
        # ---- update_h! ----
        dg = g - g_previous
        sy = _rdot(dx, dg)
        rho_it = math.inf if sy == 0.0 else 1.0 / sy
        if math.isinf(rho_it):
            pseudo = 0
        else:
            i = (pseudo - 1) % m
            dx_hist[i], dg_hist[i], rho[i] = dx, dg, rho_it
        if trace is not None:
            trace.append(dict(it=it, f=f, gmax=float(np.max(np.abs(g))), alpha=alpha, n_fg=n_fg))
        if verbose:
            print(f"iter {it:4d} f={f:.12f} |g|inf={np.max(np.abs(g)):.3e} alpha={alpha:.4g} n_fg={n_fg}")
    return x, f, it, dict(n_fg=n_fg, converged=converged, ls_failed=ls_failed, stop=stop_reason)
