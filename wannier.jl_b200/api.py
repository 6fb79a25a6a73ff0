"""Host-side mirror of the reference's Julia API for the spread / gradient hot path.

Julia is not available in the build or GPU images, so the host layer above the C-ABI is this
Python module; names, argument meaning and error behaviour follow the reference
(paths below are under the reference tree):

    KspaceStencil, Model                      src/kpoints/kstencil.jl:19-60, src/model.jl:26-66
    Spread, SpreadCenter                      src/spread.jl:30-54, 62-97
    SpreadPenalty, CenterSpreadPenalty        src/spread.jl:203-227
    Cache, compute_MU_UtMU                    src/spread.jl:127-198
    omega, omega_grad, center, omega_center   src/spread.jl:241-252, 370-381, 496-510, 560-564
    transform_gauge                           src/localization/gauge.jl:83-98
    get_fg_maxloc, max_localize               src/localization/max_localize.jl:10-101
    X_Y_to_XY, XY_to_X_Y, X_Y_to_U, U_to_X_Y  src/localization/disentangle.jl:351-466
    get_fg_disentangle, disentangle           src/localization/disentangle.jl:586-728
    MagModel, SpreadMag, overlap_updn, omega_updn, omega_updn_grad,
    omega / get_fg_disentangle / disentangle on a MagModel      src/localization/coopt.jl

Arrays use math indexing in numpy C order: M[k, b, m, l], U[k, m, n], G[k, m, n]; they are
converted to the C-ABI's column-major (Julia) memory image at this boundary.  All arithmetic of
the hot path runs in libwannier_b200.so on the GPU; nothing here falls back to the CPU.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import lib as _lib


# ------------------------------------------------------------------------------------------
# data model
# ------------------------------------------------------------------------------------------
@dataclass
class KspaceStencil:
    """src/kpoints/kstencil.jl:19-60.  recip_lattice columns are the reciprocal vectors (1/A);
    kpoints fractional [nk, 3]; bweights [nn]; kpb_k [nk, nn] 0-based; kpb_G [nk, nn, 3]."""
    recip_lattice: np.ndarray
    kpoints: np.ndarray
    bweights: np.ndarray
    kpb_k: np.ndarray
    kpb_G: np.ndarray

    @property
    def n_kpoints(self):
        return self.kpb_k.shape[0]

    @property
    def n_bvectors(self):
        return self.kpb_k.shape[1]

    def bvectors_cart(self):
        """b = recip_lattice * (kpoints[kpb_k] + kpb_G - kpoints[k])   (src/spread.jl:293)"""
        kp = np.asarray(self.kpoints, dtype=np.float64)
        d = kp[self.kpb_k] + np.asarray(self.kpb_G, dtype=np.float64) - kp[:, None, :]
        return d @ np.asarray(self.recip_lattice, dtype=np.float64).T


@dataclass
class Model:
    """src/model.jl:26-66 (the fields the hot path reads)."""
    lattice: np.ndarray
    kstencil: KspaceStencil
    overlaps: np.ndarray                 # M[k, b, m, l]
    gauges: np.ndarray                   # U[k, m, n]
    frozen_bands: np.ndarray | None = None   # [nk, nb] bool
    eigenvalues: np.ndarray | None = None

    def __post_init__(self):
        nk, nn, nb, nb2 = self.overlaps.shape
        if nb != nb2:
            raise ValueError("overlaps must be n_bands x n_bands")
        if self.gauges.shape[0] != nk or self.gauges.shape[1] != nb:
            raise ValueError("gauges must be [n_kpts, n_bands, n_wann]")
        if self.kstencil.kpb_k.shape != (nk, nn):
            raise ValueError("kstencil does not match overlaps")
        if self.gauges.shape[2] > nb:
            raise ValueError("n_wann > n_bands")

    n_bands = property(lambda self: self.overlaps.shape[2])
    n_wannier = property(lambda self: self.gauges.shape[2])
    n_kpoints = property(lambda self: self.overlaps.shape[0])
    n_bvectors = property(lambda self: self.overlaps.shape[1])


@dataclass
class Spread:
    """src/spread.jl:30-54 (ASCII field names; unicode aliases below)."""
    Omega: float
    OmegaI: float
    OmegaOD: float
    OmegaD: float
    OmegaTilde: float
    omega: np.ndarray
    r: np.ndarray
    min_abs_Nnn: float = float("nan")

    Ω = property(lambda s: s.Omega)
    ΩI = property(lambda s: s.OmegaI)
    ΩOD = property(lambda s: s.OmegaOD)
    ΩD = property(lambda s: s.OmegaD)
    Ω̃ = property(lambda s: s.OmegaTilde)
    ω = property(lambda s: s.omega)


@dataclass
class SpreadCenter(Spread):
    """src/spread.jl:62-97"""
    Omegac: float = 0.0
    Omegat: float = 0.0
    omegac: np.ndarray = field(default_factory=lambda: np.zeros(0))
    omegat: np.ndarray = field(default_factory=lambda: np.zeros(0))


class SpreadPenalty:
    """src/spread.jl:203"""


@dataclass
class CenterSpreadPenalty:
    """src/spread.jl:214-217: r0 [nw, 3] (Cartesian A), lam."""
    r0: np.ndarray
    lam: float


# ------------------------------------------------------------------------------------------
# layout conversion at the C-ABI boundary
# ------------------------------------------------------------------------------------------
def to_abi_U(U):
    """U[k, m, n] -> [k][n][m] contiguous (Julia nb x nw col-major per k)."""
    return np.ascontiguousarray(np.asarray(U, dtype=np.complex128).transpose(0, 2, 1))


def from_abi_U(Uc):
    return np.ascontiguousarray(Uc.transpose(0, 2, 1))


def to_abi_M(M):
    return np.ascontiguousarray(np.asarray(M, dtype=np.complex128).transpose(0, 1, 3, 2))


# ------------------------------------------------------------------------------------------
# Cache = device context for one (stencil, M)
# ------------------------------------------------------------------------------------------
class Cache:
    """Cache(model) / Cache(stencil, M, U) (src/spread.jl:139-162): here it owns the device
    context — M resident in HBM, all work buffers — instead of host MU/UtMU arrays."""

    def __init__(self, kstencil, M, n_wann, *, device=0, devices=None, flags=_lib.FLAG_DEFAULT):
        """devices = [d0, d1, ...]: shard the k-points over several GPUs behind the same calls
        (omega, omega_grad, center, the fg closures, max_localize, disentangle)."""
        nk, nn, nb, _ = M.shape
        self.nb, self.nw, self.nk, self.nn = nb, n_wann, nk, nn
        if devices is not None and len(devices) > 1:
            self.ctx = _lib.MultiContext(nb, n_wann, nk, nn, kstencil.kpb_k, kstencil.bvectors_cart(),
                                         kstencil.bweights, to_abi_M(M), devices=devices, flags=flags)
        else:
            # M[k, b, m, l] in C order is the row-major image of every M_kb: the library transposes it on the
            # device (WB200_FLAG_M_ROW_MAJOR) instead of numpy on the host (0.08 of the 0.3 ms of a one-shot call)
            self.ctx = _lib.Context(nb, n_wann, nk, nn, kstencil.kpb_k, kstencil.bvectors_cart(),
                                    kstencil.bweights, np.ascontiguousarray(M, dtype=np.complex128),
                                    device=devices[0] if devices else device, flags=flags | _lib.FLAG_M_ROW_MAJOR)
        self._penalty = None

    @classmethod
    def from_model(cls, model, **kw):
        c = cls(model.kstencil, model.overlaps, model.n_wannier, **kw)
        if model.frozen_bands is not None:
            c.ctx.set_frozen(np.asarray(model.frozen_bands, dtype=np.uint8))
        return c

    def set_penalty(self, p):
        if p is None or isinstance(p, SpreadPenalty):
            self.ctx.set_penalty(None, 0.0)
            self._penalty = None
        elif isinstance(p, CenterSpreadPenalty):
            self.ctx.set_penalty(np.asarray(p.r0, dtype=np.float64), p.lam)
            self._penalty = p
        else:
            raise TypeError("penalty must be SpreadPenalty or CenterSpreadPenalty")

    def close(self):
        self.ctx.close()


def _split_args(args):
    """([penalty,] model[, U])  or  ([penalty,] stencil, M, U)"""
    args = list(args)
    p = None
    if args and isinstance(args[0], (SpreadPenalty, CenterSpreadPenalty)):
        p = args.pop(0)
    if isinstance(args[0], Model):
        model = args[0]
        U = args[1] if len(args) > 1 else model.gauges
        return p, model.kstencil, model.overlaps, U
    st, M, U = args
    return p, st, M, U


def _with_cache(args, cache):
    """In the reference the penalty travels with every call; here it is state of the device context.
    A caller-owned cache keeps whatever penalty it has (e.g. the one its fg closure set) unless this
    call names one."""
    p, st, M, U = _split_args(args)
    own = cache is None
    if own:
        cache = Cache(st, M, U.shape[2])
    if own or p is not None:
        cache.set_penalty(p)
    return p, cache, U, own


def omega(*args, cache=None):
    """omega([penalty,] model[, U]) / omega([penalty,] stencil, M, U)  (src/spread.jl:370-381;
    with CenterSpreadPenalty -> omega_center, :221).  omega(model::MagModel, [Uup, Udn,] lambda) -> SpreadMag
    (coopt.jl:66-77)."""
    if args and isinstance(args[0], MagModel):
        if len(args) == 2:
            return omega_mag(args[0], lam=args[1], cache=cache)
        return omega_mag(args[0], args[1], args[2], args[3], cache=cache)
    p, cache, U, own = _with_cache(args, cache)
    try:
        out5, om, r, mn = cache.ctx.spread(to_abi_U(U))
    finally:
        if own:
            cache.close()
    sp = Spread(*[float(x) for x in out5], om, r, mn)
    if isinstance(p, CenterSpreadPenalty):
        return omega_center(sp, p.r0, p.lam)
    return sp


def omega_center(sp, r0, lam):
    """src/spread.jl:246-252"""
    omegac = lam * ((sp.r - np.asarray(r0)) ** 2).sum(axis=1)
    return SpreadCenter(sp.Omega, sp.OmegaI, sp.OmegaOD, sp.OmegaD, sp.OmegaTilde, sp.omega, sp.r,
                        sp.min_abs_Nnn, float(omegac.sum()), float(sp.Omega + omegac.sum()), omegac,
                        sp.omega + omegac)


def omega_grad(*args, cache=None):
    """omega_grad([penalty,] stencil, M, U) -> G[k, m, n]   (src/spread.jl:496-510)"""
    p, cache, U, own = _with_cache(args, cache)
    try:
        G = np.empty((cache.nk, cache.nw, cache.nb), dtype=np.complex128)
        cache.ctx.fg(to_abi_U(U), want_f=False, G=G)
    finally:
        if own:
            cache.close()
    return from_abi_U(G)


def center(*args, cache=None):
    """center(model[, U]) / center(stencil, M, U) -> r[nw, 3]   (src/spread.jl:560-564)"""
    p, cache, U, own = _with_cache(args, cache)
    try:
        if not hasattr(cache.ctx, "center"):      # several GPUs: the centres come with the spread
            return cache.ctx.spread(to_abi_U(U))[2]
        return cache.ctx.center(to_abi_U(U))
    finally:
        if own:
            cache.close()


def compute_MU_UtMU(cache, U):
    """compute_MU_UtMU!(cache, stencil, M, U) (src/spread.jl:164-198): returns (MU, UtMU) with
    MU[k, b, m, n], UtMU[k, b, j, n]."""
    N, MU = cache.ctx.rotate_overlaps(to_abi_U(U), want_N=True, want_MU=True)
    return MU.transpose(0, 1, 3, 2), N.transpose(0, 1, 3, 2)


def transform_gauge(M, kstencil_or_kpb, U, cache=None):
    """transform_gauge(M, kpb_k, U) -> N[k, b, j, n] = U_k^H M_kb U_{k+b}  (gauge.jl:83-98)"""
    own = cache is None
    if own:
        if isinstance(kstencil_or_kpb, KspaceStencil):
            st = kstencil_or_kpb
        else:  # only kpb_k matters for the rotation
            kpb = np.asarray(kstencil_or_kpb)
            st = KspaceStencil(np.eye(3), np.zeros((kpb.shape[0], 3)), np.ones(kpb.shape[1]), kpb,
                               np.zeros(kpb.shape + (3,), dtype=np.int64))
        cache = Cache(st, M, U.shape[2])
    try:
        N, _ = cache.ctx.rotate_overlaps(to_abi_U(U), want_N=True, want_MU=False)
    finally:
        if own:
            cache.close()
    return np.ascontiguousarray(N.transpose(0, 1, 3, 2))


def omega_local(*args, cache=None):
    """omega_local(stencil, M, U) -> loc[nk]   (src/spread.jl:522-548)"""
    p, cache, U, own = _with_cache(args, cache)
    try:
        return cache.ctx.omega_local(to_abi_U(U))
    finally:
        if own:
            cache.close()


def position_op(*args, cache=None):
    """position_op(model[, U]) / position_op(stencil, M, U) -> R[m, n, alpha]  (src/spread.jl:674-736)"""
    p, cache, U, own = _with_cache(args, cache)
    try:
        return np.ascontiguousarray(cache.ctx.position_op(to_abi_U(U)).transpose(1, 0, 2))
    finally:
        if own:
            cache.close()


def compute_berry_connection_kspace(*args, imlog_diag=True, force_hermiticity=True, cache=None):
    """compute_berry_connection_kspace(model[, gauges]; imlog_diag, force_hermiticity)
    -> A[k, m, n, alpha]   (src/spread.jl:747-801)"""
    p, cache, U, own = _with_cache(args, cache)
    try:
        A = cache.ctx.berry_connection(to_abi_U(U), imlog_diag, force_hermiticity)
        return np.ascontiguousarray(A.transpose(0, 2, 1, 3))
    finally:
        if own:
            cache.close()


# ------------------------------------------------------------------------------------------
# opt_rotate: one global rotation W
# ------------------------------------------------------------------------------------------
def merge_gauge(U, W):
    """U_k W for every k"""
    return np.asarray(U) @ np.asarray(W)


def _rotated_cache(model):
    """model2 of opt_rotate.jl:113-116: overlaps pre-rotated by the model's gauges, identity gauges."""
    if model.n_bands != model.n_wannier:
        raise ValueError("n_bands != n_wann, run instead disentanglement?")      # opt_rotate.jl:101
    N = transform_gauge(model.overlaps, model.kstencil, model.gauges)
    return Cache(model.kstencil, N, model.n_wannier)


def get_fg_rotate(model):
    """get_fg!_rotate(model) -> (f, g!)  (src/localization/opt_rotate.jl:12-81); f(W) -> Omega,
    g(G, W) writes G[m, n] in place."""
    cache = _rotated_cache(model)
    nw = model.n_wannier

    def f(W):
        return cache.ctx.fg_rotate(np.ascontiguousarray(np.asarray(W, dtype=np.complex128).T))

    def g(G, W):
        Gb = np.empty((nw, nw), dtype=np.complex128)
        cache.ctx.fg_rotate(np.ascontiguousarray(np.asarray(W, dtype=np.complex128).T), want_f=False, G=Gb)
        G[...] = Gb.T

    f.cache = cache
    f.close = g.close = cache.close      # the pair owns a device context (tiled M, MU, U): release it when done
    return f, g


def opt_rotate(model, f_tol=1e-7, g_tol=1e-5, max_iter=200, history_size=3, return_info=False):
    """opt_rotate(model; ...) -> Wmin[m, n]   (src/localization/opt_rotate.jl:97-154)"""
    cache = _rotated_cache(model)
    try:
        W = np.eye(model.n_wannier, dtype=np.complex128)
        f, iters, nfg = cache.ctx.opt_rotate(W, f_tol, g_tol, max_iter, history_size)
    finally:
        cache.close()
    Wm = np.ascontiguousarray(W.T)
    return (Wm, dict(f=f, iterations=iters, n_fg=nfg)) if return_info else Wm


# ------------------------------------------------------------------------------------------
# max_localize
# ------------------------------------------------------------------------------------------
def get_fg_maxloc(p, model, cache=None):
    """get_fg!_maxloc (src/localization/max_localize.jl:10-28).  Returns fg(F, G, U) with
    Optim's only_fg! convention: F / G may be None to skip; G[k, m, n] is written in place;
    returns Omega (or None)."""
    cache = cache or Cache.from_model(model)
    cache.set_penalty(p)
    Gbuf = np.empty((cache.nk, cache.nw, cache.nb), dtype=np.complex128)

    def fg(F, G, U):
        cache.set_penalty(p)      # the closure's penalty, whatever else used the cache in between
        f = cache.ctx.fg(to_abi_U(U), want_f=F is not None, G=Gbuf if G is not None else None)
        if G is not None:
            G[...] = Gbuf.transpose(0, 2, 1)
        return f

    fg.cache = cache
    return fg


def max_localize(*args, f_tol=1e-7, g_tol=1e-5, max_iter=200, history_size=3, cache=None,
                 return_info=False, devices=None):
    """max_localize([p,] model; f_tol, g_tol, max_iter, history_size) -> gauges U[k, m, n]
    (src/localization/max_localize.jl:44-101).  The whole optimiser loop is device-resident."""
    args = list(args)
    p = SpreadPenalty()
    if isinstance(args[0], (SpreadPenalty, CenterSpreadPenalty)):
        p = args.pop(0)
    model = args[0]
    if model.n_bands != model.n_wannier:
        raise ValueError("n_bands != n_wann, run instead disentanglement?")   # max_localize.jl:52-53
    own = cache is None
    cache = cache or Cache.from_model(model, devices=devices)
    cache.set_penalty(p)
    try:
        Uc = to_abi_U(model.gauges)
        f, iters, nfg = cache.ctx.max_localize(Uc, f_tol, g_tol, max_iter, history_size)
    finally:
        if own:
            cache.close()
    U = from_abi_U(Uc)
    return (U, dict(f=f, iterations=iters, n_fg=nfg)) if return_info else U


# ------------------------------------------------------------------------------------------
# disentangle: (X, Y) parametrisation
# ------------------------------------------------------------------------------------------
def X_Y_to_U(X, Y):
    """U_k = Y_k X_k   (disentangle.jl:351-356).  Host helper (layout/setup only)."""
    return Y @ X


def X_Y_to_XY(X, Y):
    """row k of the result = Julia column k = [vec(X_k); vec(Y_k)]  (disentangle.jl:454-466)"""
    nk = X.shape[0]
    return np.ascontiguousarray(np.concatenate(
        [X.transpose(0, 2, 1).reshape(nk, -1), Y.transpose(0, 2, 1).reshape(nk, -1)], axis=1))


def XY_to_X_Y(XY, nb, nw):
    """disentangle.jl:424-435"""
    nk = XY.shape[0]
    X = XY[:, :nw * nw].reshape(nk, nw, nw).transpose(0, 2, 1)
    Y = XY[:, nw * nw:].reshape(nk, nw, nb).transpose(0, 2, 1)
    return np.ascontiguousarray(X), np.ascontiguousarray(Y)


def _stiefel_ctx(U, frozen, cache):
    nk, nb, nw = U.shape
    if cache is not None:
        return cache.ctx, False
    ctx = _lib.StiefelContext(nb, nw, nk)
    ctx.set_frozen(None if frozen is None else np.asarray(frozen, dtype=np.uint8))
    return ctx, True


def orthonorm_freeze(U, frozen, cache=None):
    """orthonorm_freeze(U_k, frozen_k) for every k (disentangle.jl:223-281): U[k, m, n], frozen[k, m].
    Runs on the GPU (wb200_orthonorm_freeze); `cache` supplies a context whose frozen bands are already set."""
    U = np.asarray(U, dtype=np.complex128)
    single = U.ndim == 2
    if single:
        U, frozen = U[None], np.asarray(frozen)[None]
    ctx, own = _stiefel_ctx(U, frozen, cache)
    try:
        Uc = to_abi_U(U)
        ctx.orthonorm_freeze(Uc)
    finally:
        if own:
            ctx.close()
    V = from_abi_U(Uc)
    return V[0] if single else V


def U_to_X_Y(U, frozen, cache=None):
    """U_to_X_Y(U, frozen) -> (X[k, :, :], Y[k, :, :])   (disentangle.jl:369-402), on the GPU (wb200_u_to_xy).
    The free block of Y_k is an orthonormal basis of the non-frozen range; the reference's is LAPACK's
    eigenbasis of a degenerate projector, i.e. the same up to a unitary mixing that X_k absorbs."""
    U = np.asarray(U, dtype=np.complex128)
    ctx, own = _stiefel_ctx(U, frozen, cache)
    try:
        XY = ctx.u_to_xy(to_abi_U(U))
    finally:
        if own:
            ctx.close()
    return XY_to_X_Y(XY, U.shape[1], U.shape[2])


def get_fg_disentangle(p, model=None, cache=None):
    """get_fg!_disentangle (disentangle.jl:586-614): fg(F, G, XY) with XY [nk, nw*nw + nb*nw].
    get_fg_disentangle(model::MagModel, lambda) -> (f, g!) (coopt.jl:252-321)."""
    if isinstance(p, MagModel):
        return get_fg_disentangle_mag(p, 1.0 if model is None else model, cache=cache)
    cache = cache or Cache.from_model(model)
    cache.set_penalty(p)

    def fg(F, G, XY):
        cache.set_penalty(p)
        XY = np.ascontiguousarray(XY, dtype=np.complex128)
        Gb = np.empty_like(XY) if G is not None else None
        f = cache.ctx.fg_xy(XY, want_f=F is not None, G=Gb)
        if G is not None:
            G[...] = Gb
        return f

    fg.cache = cache
    return fg


def disentangle(*args, f_tol=1e-7, g_tol=1e-5, max_iter=200, history_size=3, cache=None,
                return_info=False, devices=None):
    """disentangle([p,] model; ...) -> gauges U[k, m, n]   (disentangle.jl:631-728);
    disentangle(model::MagModel[, lambda]; ...) -> (Uup, Udn)   (coopt.jl:339-420)"""
    args = list(args)
    if isinstance(args[0], MagModel):
        return disentangle_mag(args[0], args[1] if len(args) > 1 else 1.0, f_tol=f_tol, g_tol=g_tol,
                               max_iter=max_iter, history_size=history_size, cache=cache, return_info=return_info)
    p = SpreadPenalty()
    if isinstance(args[0], (SpreadPenalty, CenterSpreadPenalty)):
        p = args.pop(0)
    model = args[0]
    if model.frozen_bands is None:
        raise ValueError("disentangle needs model.frozen_bands")
    own = cache is None
    cache = cache or Cache.from_model(model, devices=devices)
    cache.set_penalty(p)
    try:
        if hasattr(cache.ctx, "u_to_xy"):
            XY = cache.ctx.u_to_xy(to_abi_U(model.gauges))       # disentangle.jl:666-670, on the device
        else:                                                   # several GPUs: set up on the first one
            XY = X_Y_to_XY(*U_to_X_Y(model.gauges, model.frozen_bands))
        f, iters, nfg = cache.ctx.disentangle(XY, f_tol, g_tol, max_iter, history_size)
    finally:
        if own:
            cache.close()
    X, Y = XY_to_X_Y(XY, model.n_bands, model.n_wannier)
    U = X_Y_to_U(X, Y)
    return (U, dict(f=f, iterations=iters, n_fg=nfg)) if return_info else U


# ------------------------------------------------------------------------------------------
# MagModel: spin-up / spin-down co-optimisation (src/localization/coopt.jl)
# ------------------------------------------------------------------------------------------
@dataclass
class MagModel:
    """coopt.jl:8-17: the two spin channels and M[k, m, l] = <u_mk^up | u_lk^dn>."""
    up: Model
    dn: Model
    M: np.ndarray


@dataclass
class SpreadMag:
    """coopt.jl:46-64"""
    up: Spread
    dn: Spread
    Omegaupdn: float
    Omegat: float
    M: np.ndarray
    lam: float


class MagCache:
    """Device state of a MagModel: one Cache per spin channel and the coupling context."""

    def __init__(self, model: MagModel, device=0):
        if (model.up.n_bands, model.up.n_wannier, model.up.n_kpoints) != \
                (model.dn.n_bands, model.dn.n_wannier, model.dn.n_kpoints):
            raise ValueError("MagModel: up and down models differ in n_bands / n_wann / n_kpts")   # coopt.jl:355-357
        self.up = Cache.from_model(model.up, device=device)
        self.dn = Cache.from_model(model.dn, device=device)
        Mud = np.ascontiguousarray(np.asarray(model.M, dtype=np.complex128).transpose(0, 2, 1))
        self.mag = _lib.MagContext(self.up.ctx, self.dn.ctx, Mud)
        self.nb, self.nw, self.nk = self.up.nb, self.up.nw, self.up.nk

    def close(self):
        self.mag.close()
        self.up.close()
        self.dn.close()


def _mag_call(model, cache, fn):
    own = cache is None
    cache = cache or MagCache(model)
    try:
        return fn(cache)
    finally:
        if own:
            cache.close()


def overlap_updn(model, Uup=None, Udn=None, cache=None):
    """overlap_updn(model[, Uup, Udn]) -> |sum_k Uup_k^H M_k Udn_k|^2 / nk^2  [nw, nw]  (coopt.jl:114-135)"""
    Uup = model.up.gauges if Uup is None else Uup
    Udn = model.dn.gauges if Udn is None else Udn
    return _mag_call(model, cache, lambda c: np.ascontiguousarray(c.mag.overlap(to_abi_U(Uup), to_abi_U(Udn))[0].T))


def omega_updn(model, Uup=None, Udn=None, cache=None):
    """omega_updn(model[, Uup, Udn]) = n_wann - tr(overlap_updn)   (coopt.jl:146-160)"""
    Uup = model.up.gauges if Uup is None else Uup
    Udn = model.dn.gauges if Udn is None else Udn
    return _mag_call(model, cache, lambda c: c.mag.overlap(to_abi_U(Uup), to_abi_U(Udn), want_M=False)[1])


def omega_updn_grad(model, Uup, Udn, cache=None):
    """omega_updn_grad(model, Uup, Udn) -> (GUup, GUdn)[k, m, n]   (coopt.jl:222-229)"""
    def run(c):
        Gu, Gd = c.mag.omega_updn_grad(to_abi_U(Uup), to_abi_U(Udn))
        return from_abi_U(Gu), from_abi_U(Gd)
    return _mag_call(model, cache, run)


def omega_mag(model, Uup=None, Udn=None, lam=1.0, cache=None):
    """omega(model::MagModel, [Uup, Udn,] lambda) -> SpreadMag   (coopt.jl:66-77)"""
    Uup = model.up.gauges if Uup is None else Uup
    Udn = model.dn.gauges if Udn is None else Udn

    def run(c):
        up = omega(model.up, Uup, cache=c.up)
        dn = omega(model.dn, Udn, cache=c.dn)
        Mw2, o = c.mag.overlap(to_abi_U(Uup), to_abi_U(Udn))
        return SpreadMag(up, dn, o, up.Omega + dn.Omega + lam * o, np.ascontiguousarray(Mw2.T), lam)
    return _mag_call(model, cache, run)


def mag_XY(model):
    """XY0 of disentangle(::MagModel) (coopt.jl:369-374): row k = [XYup_k; XYdn_k]."""
    XYu = X_Y_to_XY(*U_to_X_Y(model.up.gauges, model.up.frozen_bands))      # coopt.jl:369-370
    XYd = X_Y_to_XY(*U_to_X_Y(model.dn.gauges, model.dn.frozen_bands))
    return np.ascontiguousarray(np.concatenate([XYu, XYd], axis=1))


def get_fg_disentangle_mag(model, lam=1.0, cache=None):
    """get_fg!_disentangle(model::MagModel, lambda) -> (f, g!)   (coopt.jl:252-321); XY [nk, 2 n_inner]."""
    cache = cache or MagCache(model)

    def f(XY):
        return cache.mag.fg_xy(np.ascontiguousarray(XY, dtype=np.complex128), lam)[0]

    def g(G, XY):
        Gb = np.empty((cache.nk, 2 * cache.mag.n_inner), dtype=np.complex128)
        cache.mag.fg_xy(np.ascontiguousarray(XY, dtype=np.complex128), lam, want_f=False, G=Gb)
        G[...] = Gb

    f.cache = cache
    f.close = g.close = cache.close
    return f, g


def disentangle_mag(model, lam=1.0, f_tol=1e-7, g_tol=1e-5, max_iter=200, history_size=3, cache=None,
                    return_info=False):
    """disentangle(model::MagModel, lambda; ...) -> (Uup, Udn)   (coopt.jl:339-420)"""
    if model.up.frozen_bands is None or model.dn.frozen_bands is None:
        raise ValueError("disentangle needs frozen_bands on both spin channels")
    nb, nw = model.up.n_bands, model.up.n_wannier
    n_inner = nw * nw + nb * nw

    def run(c):
        XY = mag_XY(model)
        f, iters, nfg = c.mag.disentangle(XY, lam, f_tol, g_tol, max_iter, history_size)
        Uup = X_Y_to_U(*XY_to_X_Y(np.ascontiguousarray(XY[:, :n_inner]), nb, nw))
        Udn = X_Y_to_U(*XY_to_X_Y(np.ascontiguousarray(XY[:, n_inner:]), nb, nw))
        return (Uup, Udn, dict(f=f, iterations=iters, n_fg=nfg)) if return_info else (Uup, Udn)
    return _mag_call(model, cache, run)
