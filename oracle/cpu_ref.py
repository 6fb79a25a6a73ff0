"""ctypes wrapper of oracle/spread_ref.cpp, the compiled (C++17 / OpenMP / OpenBLAS ZGEMM) CPU
restatement of the reference's evaluation in the reference's own loop order.

TEST INFRASTRUCTURE AND CPU BASELINE ONLY (tests/, smoke(), bench.py's cpu_baseline and
--impl reference legs); nothing under wannier.jl_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "spread_ref.cpp")
LIB = os.path.join(HERE, "_build", "libspread_ref.so")
CXXFLAGS = ["-O3", "-march=x86-64-v3", "-fopenmp", "-std=c++17", "-shared", "-fPIC"]

_lib = None


def build(force=False):
    """g++ -O3 -fopenmp (AVX2 baseline so the .so runs on any box; the flops are in OpenBLAS, which
    dispatches on the CPU at run time)."""
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["g++", *CXXFLAGS, "-o", LIB, SRC, "-ldl"])
    return LIB


def _find_openblas():
    import numpy
    pats = [os.path.join(os.path.dirname(numpy.__file__), "..", "numpy.libs", "libscipy_openblas64_*.so")]
    for p in pats:
        hits = sorted(glob.glob(p))
        if hits:
            return os.path.abspath(hits[0])
    return None


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(LIB)
    lib.sr_set_blas.argtypes = [C.c_char_p]
    lib.sr_blas_name.restype = C.c_char_p
    vp = C.c_void_p
    lib.sr_fg.argtypes = [C.c_int] * 4 + [vp] * 6 + [C.c_double] + [vp] * 4 + [C.c_int, vp, C.c_int]
    lib.sr_overlaps_from_subspaces.argtypes = [C.c_int] * 4 + [vp] * 3 + [C.c_int]
    ob = _find_openblas()
    if ob:
        lib.sr_set_blas(ob.encode())
    _lib = lib
    return lib


def blas_name():
    return load().sr_blas_name().decode()


def max_threads():
    return load().sr_max_threads()


class Workspace:
    """The reference's Cache (src/spread.jl:127-162): MU and UtMU for every pair."""

    def __init__(self, nb, nw, n_k, nn):
        self.MU = np.empty((n_k * nn, nw, nb), dtype=np.complex128)
        self.N = np.empty((n_k * nn, nw, nw), dtype=np.complex128)


def fg_abi(M, kpb, bvec, wb, U, nk, *, r0=None, lam=0.0, want_G=True, threads=0, k_sel=None, ws=None):
    """One evaluation on C-ABI (Julia-layout) arrays: M [n_k][nn][nb][nb], U [nk][nw][nb] (both stored
    column-major per matrix, i.e. index order [.., col, row]).  Returns (res, G) with res as the GPU
    library's `res` vector.  k_sel: global k indices of the rows of M (bounded sample)."""
    lib = load()
    n_k, nn, nb, _ = M.shape
    nw = U.shape[1]
    assert M.flags.c_contiguous and U.flags.c_contiguous and M.dtype == np.complex128 and U.dtype == np.complex128
    kpb = np.ascontiguousarray(kpb, dtype=np.int32)
    bvec = np.ascontiguousarray(bvec, dtype=np.float64)
    wb = np.ascontiguousarray(wb, dtype=np.float64)
    ws = ws or Workspace(nb, nw, n_k, nn)
    res = np.zeros(8 + 4 * nw)
    G = np.empty((n_k, nw, nb), dtype=np.complex128) if want_G else None
    r0a = np.ascontiguousarray(r0, dtype=np.float64) if r0 is not None else None
    ks = np.ascontiguousarray(k_sel, dtype=np.int32) if k_sel is not None else None
    a = lambda x: x.ctypes.data if x is not None else None
    st = lib.sr_fg(nb, nw, nk, nn, a(kpb), a(bvec), a(wb), a(M), a(U), a(r0a), float(lam), a(ws.MU), a(ws.N),
                   a(res), a(G), int(threads), a(ks), 0 if ks is None else len(ks))
    assert st == 0
    return res, G


def fg(M, kpb_k, bvec, wb, U, r0=None, lam=0.0, threads=0):
    """Same call on oracle-layout arrays (M[k, b, m, l], U[k, m, n]); returns (spread dict, G[k, m, n])."""
    Mc = np.ascontiguousarray(np.asarray(M, dtype=np.complex128).transpose(0, 1, 3, 2))
    Uc = np.ascontiguousarray(np.asarray(U, dtype=np.complex128).transpose(0, 2, 1))
    nw = Uc.shape[1]
    res, G = fg_abi(Mc, kpb_k, bvec, wb, Uc, len(Uc), r0=r0, lam=lam, threads=threads)
    sp = dict(Omega=res[0], OmegaI=res[1], OmegaOD=res[2], OmegaD=res[3], OmegaTilde=res[4], f=res[5],
              omega=res[8:8 + nw].copy(), r=res[8 + nw:].reshape(nw, 3).copy())
    return sp, np.ascontiguousarray(G.transpose(0, 2, 1))


def overlaps_from_subspaces(V, kpb, threads=0):
    """M[k][b] = V_k^H V_{kpb[k, b]} in the C-ABI layout (M [nk][nn][l][m]) from V[j, h, n] (math indexing)
    for the len(kpb) first rows k of V (kpb may point at any row of V)."""
    lib = load()
    _, nH, nb = V.shape
    nk = kpb.shape[0]
    Vc = np.ascontiguousarray(np.asarray(V, dtype=np.complex128).transpose(0, 2, 1))    # column-major nH x nb
    nn = kpb.shape[1]
    kpb = np.ascontiguousarray(kpb, dtype=np.int32)
    M = np.empty((nk, nn, nb, nb), dtype=np.complex128)
    lib.sr_overlaps_from_subspaces(nH, nb, nk, nn, kpb.ctypes.data, Vc.ctypes.data, M.ctypes.data, int(threads))
    return M
