"""Ingest of Wannier90 text files into the hot path's data model (the setup step
in front of the path).

The reference delegates file parsing to WannierIO.jl (not vendored) and assembles a `Model` in
`read_w90` (src/io/w90/model.jl:16-94): lattice / k-points from `.win`, overlaps and the b-vector
stencil from `.mmn` (src/kpoints/kstencil.jl:68-85), B1 weights (kstencil.jl:132-170), gauges =
Lowdin-orthonormalised `.amn` projections (src/io/w90/amn.jl:12-17), frozen window from
`dis_froz_min/max` and `.eig` (src/localization/disentangle.jl:20-22).  The formats are restated
from the files themselves.  Parsing is host work; the Lowdin orthonormalisation of the projections
runs on the GPU (`wb200_retract` = polar factor = U V^H of the SVD).
"""
from __future__ import annotations

import numpy as np

from . import api as _api
from . import lib as _lib

BOHR = 0.52917721092   # WannierIO's constant (CODATA-2010); pinned through Omega_I of the valence fixture


def _strip(line):
    for c in "!#":
        if c in line:
            line = line[:line.index(c)]
    return line.strip()


def read_win(path):
    """dict(lattice [3,3] columns = lattice vectors in Angstrom, kpoints [nk,3], mp_grid, num_wann,
    num_bands, dis_froz_min/max when present)."""
    lines = [s for s in (_strip(l) for l in open(path)) if s]
    blocks, keys, i = {}, {}, 0
    while i < len(lines):
        low = lines[i].lower()
        if low.startswith("begin"):
            name = low.split()[1]
            j = i + 1
            body = []
            while not lines[j].lower().startswith("end"):
                body.append(lines[j])
                j += 1
            blocks[name] = body
            i = j + 1
            continue
        parts = lines[i].replace("=", " ").replace(":", " ").split()
        if parts:
            keys[parts[0].lower()] = parts[1:]
        i += 1
    cell = blocks["unit_cell_cart"]
    scale = 1.0
    first = cell[0].lower()
    if first.startswith("bohr"):
        scale, cell = BOHR, cell[1:]
    elif first.startswith("ang"):
        cell = cell[1:]
    rows = np.array([[float(x) for x in r.split()[:3]] for r in cell[:3]]) * scale
    out = dict(lattice=rows.T.copy(),
               kpoints=np.array([[float(x) for x in r.split()[:3]] for r in blocks["kpoints"]]),
               mp_grid=[int(x) for x in keys["mp_grid"][:3]],
               num_wann=int(keys["num_wann"][0]))
    out["num_bands"] = int(keys["num_bands"][0]) if "num_bands" in keys else out["num_wann"]
    for k in ("dis_froz_max", "dis_froz_min"):
        if k in keys:
            out[k] = float(keys[k][0])
    return out


def read_mmn(path):
    """(M[k, b, m, l], kpb_k[k, b] 0-based, kpb_G[k, b, 3]).  After the header `nb nk nn` every pair
    is a line `ik ikpb G1 G2 G3` (1-based) followed by nb*nb lines `re im`, m fastest; the order
    of appearance of the pairs of a k-point defines b.  Parsed by the native multithreaded reader
    (wb200_mmn_read), which fills the C-ABI image [k][b][l][m]; the view returned here is M[k, b, m, l]."""
    Mc, kpb_k, kpb_G = _lib.read_mmn_native(path)
    return Mc.transpose(0, 1, 3, 2), kpb_k, kpb_G


def read_amn(path):
    """A[k, m, n] from lines `m n ik re im` (1-based); native reader (wb200_amn_read)."""
    return _lib.read_amn_native(path).transpose(0, 2, 1)


def read_eig(path, nb=None, nk=None):
    """E[k, m] from lines `m ik eps`; native reader (wb200_eig_read).  Without nb / nk the sizes are taken
    from the last line (Wannier90 writes m fastest, k slowest)."""
    if nb is None or nk is None:
        with open(path, "rb") as f:
            f.seek(0, 2)
            f.seek(max(0, f.tell() - 256))
            last = [l for l in f.read().decode().splitlines() if l.strip()][-1].split()
        nb, nk = int(last[0]), int(last[1])
    return _lib.read_eig_native(path, nb, nk)


def compute_bweights(bvectors, atol=1e-6, sigma_atol=1e-5):
    """B1 weights for the b-vectors of one k-point (kstencil.jl:132-170, kstencil_shell.jl:341-391):
    shells of equal |b| by increasing norm; add a shell, solve sum_s w_s sum_{b in s} b b^T = I in the
    least-squares sense (SVD), drop the shell if it makes the system singular, stop when satisfied."""
    b = np.asarray(bvectors, dtype=np.float64)
    norms = np.linalg.norm(b, axis=1)
    order = np.argsort(norms, kind="stable")
    shells, i = [], 0
    while i < len(order):
        j = i
        while j < len(order) and abs(norms[order[j]] - norms[order[i]]) <= atol:
            j += 1
        shells.append(order[i:j])
        i = j
    pick = [(0, 0), (0, 1), (1, 1), (0, 2), (1, 2), (2, 2)]
    tri = lambda m: np.array([m[r, c] for r, c in pick])
    target = tri(np.eye(3))
    cols, keep, w_keep = {}, [], None
    for s, idx in enumerate(shells):
        cols[s] = tri(b[idx].T @ b[idx])
        keep.append(s)
        Bm = np.stack([cols[t] for t in keep], axis=1)
        Um, S, Vh = np.linalg.svd(Bm, full_matrices=False)
        if np.all(S > sigma_atol):
            w = Vh.T @ ((Um.T @ target) / S)
            if np.allclose(Bm @ w, target, rtol=0, atol=atol):
                w_keep = w
                break
        else:
            keep.pop()
    if w_keep is None:
        raise ValueError("not enough shells to satisfy the B1 condition")
    out = np.zeros(len(b))
    for t, s in enumerate(keep):
        out[shells[s]] = w_keep[t]
    return out


def lowdin_gpu(A, device=0):
    """Lowdin orthonormalisation A_k -> A_k (A_k^H A_k)^(-1/2) of a stack of projections, on the GPU
    (read_amn_ortho, src/io/w90/amn.jl:12-17; src/utils/linalg.jl:25-29)."""
    nk, nb, nw = A.shape
    lib = _lib.load()
    import ctypes as C
    # the retraction only needs a device and scratch space: a minimal context carries them
    kpb = np.zeros((1, 1), dtype=np.int32)
    ctx = _lib.Context(nb, nw, 1, 1, kpb, np.zeros((1, 1, 3)), np.ones(1), np.zeros((1, 1, nb, nb), dtype=np.complex128),
                       device=device, flags=_lib.FLAG_NO_TENSOR)
    try:
        Ac = _api.to_abi_U(A)
        ctx.retract(Ac, nb, nw)
    finally:
        ctx.close()
    return _api.from_abi_U(Ac)


def read_w90(prefix, ortho_amn=True, device=0):
    """read_w90(prefix) -> Model   (src/io/w90/model.jl:16-94, with the b-vectors taken from the .mmn)."""
    win = read_win(prefix + ".win")
    M, kpb_k, kpb_G = read_mmn(prefix + ".mmn")
    nk, nn, nb, _ = M.shape
    if nk != len(win["kpoints"]) or nb != win["num_bands"]:
        raise ValueError("win and mmn disagree on the number of k-points / bands")
    recip = 2.0 * np.pi * np.linalg.inv(win["lattice"]).T
    M = np.ascontiguousarray(M)     # (Cache re-lays it out for the C-ABI; the native reader's image is a view)
    st = _api.KspaceStencil(recip, win["kpoints"], np.ones(nn), kpb_k, kpb_G)
    st.bweights = compute_bweights(st.bvectors_cart()[0])
    A = read_amn(prefix + ".amn")
    U = lowdin_gpu(A, device) if ortho_amn else A
    E, frozen = None, None
    try:
        E = read_eig(prefix + ".eig")
    except OSError:
        pass
    if nb != win["num_wann"] and E is not None and "dis_froz_max" in win:
        frozen = frozen_bands(E, win["dis_froz_max"], win.get("dis_froz_min", -np.inf))
    return _api.Model(win["lattice"], st, M, U, frozen, E)


def frozen_bands(E, dis_froz_max, dis_froz_min=-np.inf):
    """states inside the frozen window (src/localization/disentangle.jl:20-22)"""
    return (E >= dis_froz_min) & (E <= dis_froz_max)
