"""Multithreaded CPU evaluation of the reference algorithm on a k-point SAMPLE, for timing only.

TEST / MEASUREMENT INFRASTRUCTURE (see oracle/spread_oracle.py header): used by bench.py's
`cpu_baseline` leg and `--impl reference` arm, never by the product path.

The reference (Julia) cannot run in this image; this is the port it is timed by.  Per pair it does
what the reference does (paths under the reference tree): both ZGEMMs of compute_MU_UtMU!
(src/spread.jl:173-174) through OpenBLAS, the |N|^2 / imaglog sweeps of omega! (spread.jl:284-322),
center! (565-594) and the column scaling + accumulation of omega_grad! (448-474).  The reference
runs this loop on one thread (benchmark/runbenchmarks.jl:5); here k-points are spread over all
host cores (one OpenBLAS thread each), which is the most favourable reading of "multithreaded
CPU path".
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import spread_oracle as so

try:
    from threadpoolctl import threadpool_limits
except Exception:  # pragma: no cover
    threadpool_limits = None


def sample_inputs(w, st, nb, nw, ks):
    """Overlaps for the sampled k-points and gauges for them and their neighbours, generated with
    the product's synthetic recipe restricted to the k-points that are needed.
    Returns (M [ns, nn, nb, nb], (Ucompact, umap))."""
    syn = w.synthetic
    nk = len(st["kpoints"])
    need = np.unique(np.concatenate([ks, st["kpb_k"][ks].ravel()]))
    umap = -np.ones(nk, dtype=np.int64)
    umap[need] = np.arange(len(need))
    Rs, c = syn._coeffs(nb, syn.SEED_M)
    phase = np.exp(2j * np.pi * (st["kpoints"][need] @ Rs.T))
    B = np.einsum("kr,rhn->khn", phase, c)
    B[:, :nb, :] += np.eye(nb)
    V = syn._polar_np(B)
    Vh = np.conj(V[umap[ks]].transpose(0, 2, 1))
    M = np.ascontiguousarray(Vh[:, None] @ V[umap[st["kpb_k"][ks]]])
    U = syn.make_gauge(len(need), nb, nw)
    return M, (U, umap)


def fg_sample(M, kpb, bvec, wb, U, ks, nk_total, threads=1, single_blas=False):
    """One spread+gradient evaluation restricted to the sampled k-points.  U is either the full
    [nk_total, nb, nw] array or (compact, umap).  Returns (Omega_sample, G_sample).
    threads = 1 with single_blas = True is the reference's own benchmark setting (one Julia thread,
    one OpenBLAS thread: benchmark/runbenchmarks.jl:3-7)."""
    if isinstance(U, tuple):
        U, umap = U
        ik, ikpb = umap[ks], umap[kpb]
    else:
        ik, ikpb = ks, kpb
    ns = len(ks)
    nchunk = max(1, min(ns, threads * 2))
    chunks = np.array_split(np.arange(ns), nchunk)
    MU = [None] * nchunk
    N = [None] * nchunk

    def rotate(ci):
        sel = chunks[ci]
        MU[ci], N[ci] = so.compute_MU_UtMU_gathered(M[sel], U[ik[sel]], U[ikpb[sel]])
        return so.partial_sums_from_N(N[ci], bvec[sel], wb)

    def grad(ci, r):
        sel = chunks[ci]
        return so.grad_from_MU_N_r(MU[ci], N[ci], bvec[sel], wb, r, ns)

    ctxmgr = threadpool_limits(limits=1) if (threadpool_limits and (threads > 1 or single_blas)) else None
    try:
        if ctxmgr is not None:
            ctxmgr.__enter__()
        with ThreadPoolExecutor(max_workers=threads) as ex:
            parts = list(ex.map(rotate, range(nchunk)))
            tot = {k: sum(p[k] for p in parts) for k in parts[0]}
            r = -tot["wbphi"] / ns
            r2 = tot["r2"] / ns
            omega = r2 - (r ** 2).sum(axis=1)
            G = list(ex.map(lambda ci: grad(ci, r), range(nchunk)))
    finally:
        if ctxmgr is not None:
            ctxmgr.__exit__(None, None, None)
    return float(omega.sum()), np.concatenate(G, axis=0)
