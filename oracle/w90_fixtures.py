"""Test-side parsers for the plain-text Wannier90 fixture files of the reference
(test/fixtures/**).  TEST INFRASTRUCTURE ONLY (see oracle/spread_oracle.py header).

The reference delegates parsing to WannierIO.jl 0.2.5 (not vendored); the formats are
restated here from the files themselves and from how src/io/w90/model.jl:16-94 assembles
a Model.  Used only by tests/golden/make_golden.py, in the build container where
/root/reference exists; the GPU box reads the committed .npz files instead.
"""
from __future__ import annotations

import re
import numpy as np

from . import spread_oracle as so

# WannierIO's Bohr (CODATA-2010); pinned through OmegaI = 5.812709709242578
# (test/localization/max_localize.jl:53), see SURVEY 8c.
BOHR = 0.52917721092


def read_win(path):
    txt = open(path).read().splitlines()
    clean = []
    for ln in txt:
        ln = re.split(r"[!#]", ln, 1)[0].strip()
        if ln:
            clean.append(ln)
    out = {}
    i = 0
    while i < len(clean):
        ln = clean[i]
        low = ln.lower()
        if low.startswith("begin"):
            name = low.split()[1]
            j = i + 1
            block = []
            while not clean[j].lower().startswith("end"):
                block.append(clean[j])
                j += 1
            out["block_" + name] = block
            i = j + 1
            continue
        mt = re.match(r"([A-Za-z_0-9]+)\s*[=:]?\s*(.*)", ln)
        if mt:
            out[mt.group(1).lower()] = mt.group(2).strip()
        i += 1
    cell = out["block_unit_cell_cart"]
    scale = 1.0
    if cell[0].lower().startswith("bohr"):
        scale = BOHR
        cell = cell[1:]
    elif cell[0].lower().startswith("ang"):
        cell = cell[1:]
    rows = np.array([[float(x) for x in r.split()] for r in cell[:3]]) * scale
    lattice = rows.T  # columns = lattice vectors
    kpts = np.array([[float(x) for x in r.split()[:3]] for r in out["block_kpoints"]])
    res = dict(lattice=lattice, kpoints=kpts,
               mp_grid=[int(x) for x in out["mp_grid"].split()],
               num_wann=int(out["num_wann"]),
               num_bands=int(out.get("num_bands", out["num_wann"])))
    for key in ("dis_froz_max", "dis_froz_min"):
        if key in out:
            res[key] = float(out[key])
    return res


def read_mmn(path):
    """Returns M[k,b,m,l], kpb_k[k,b] (0-based), kpb_G[k,b,3].  Pair order in the file
    defines b; per pair nb^2 lines 're im', column-major (m fastest)."""
    with open(path) as f:
        f.readline()
        nb, nk, nn = [int(x) for x in f.readline().split()]
        M = np.zeros((nk, nn, nb, nb), dtype=np.complex128)
        kpb_k = np.zeros((nk, nn), dtype=np.int64)
        kpb_G = np.zeros((nk, nn, 3), dtype=np.int64)
        cnt = np.zeros(nk, dtype=np.int64)
        for _ in range(nk * nn):
            h = [int(x) for x in f.readline().split()]
            ik = h[0] - 1
            ib = cnt[ik]
            cnt[ik] += 1
            kpb_k[ik, ib] = h[1] - 1
            kpb_G[ik, ib] = h[2:5]
            vals = np.array([f.readline().split() for _ in range(nb * nb)], dtype=np.float64)
            z = vals[:, 0] + 1j * vals[:, 1]
            M[ik, ib] = z.reshape(nb, nb).T  # column-major file -> [m, l]
    return M, kpb_k, kpb_G


def read_amn(path):
    """Returns A[k,m,n]."""
    with open(path) as f:
        f.readline()
        nb, nk, nw = [int(x) for x in f.readline().split()]
        data = np.loadtxt(f)
    A = np.zeros((nk, nb, nw), dtype=np.complex128)
    m = data[:, 0].astype(int) - 1
    n = data[:, 1].astype(int) - 1
    k = data[:, 2].astype(int) - 1
    A[k, m, n] = data[:, 3] + 1j * data[:, 4]
    return A


def read_eig(path):
    data = np.loadtxt(path)
    nb = int(data[:, 0].max())
    nk = int(data[:, 1].max())
    E = np.zeros((nk, nb))
    E[data[:, 1].astype(int) - 1, data[:, 0].astype(int) - 1] = data[:, 2]
    return E


def read_chk_fmt(path):
    """Formatted W90 checkpoint (layout in SURVEY 8c item 1)."""
    with open(path) as f:
        lines = f.read().splitlines()
    it = iter(lines)
    next(it)
    nb = int(next(it))
    nexcl = int(next(it))
    if nexcl > 0:
        for _ in range(nexcl):
            next(it)
    lattice = np.array([float(x) for x in next(it).split()]).reshape(3, 3)
    recip = np.array([float(x) for x in next(it).split()]).reshape(3, 3)
    nk = int(next(it))
    grid = [int(x) for x in next(it).split()]
    kpts = np.array([[float(x) for x in next(it).split()] for _ in range(nk)])
    nn = int(next(it))
    nw = int(next(it))
    next(it)  # checkpoint name
    have_dis = int(next(it)) != 0

    def cplx(n):
        a = np.array([next(it).split() for _ in range(n)], dtype=np.float64)
        return a[:, 0] + 1j * a[:, 1]

    out = dict(nb=nb, nk=nk, nn=nn, nw=nw, kpoints=kpts, grid=grid, lattice=lattice, recip=recip)
    if have_dis:
        out["OmegaI"] = float(next(it))
        lw = np.array([int(next(it)) for _ in range(nb * nk)]).reshape(nk, nb)
        out["lwindow"] = lw != 0
        out["ndimwin"] = np.array([int(next(it)) for _ in range(nk)])
        uo = cplx(nb * nw * nk).reshape(nk, nw, nb).transpose(0, 2, 1)  # [k, m(band), n]
        out["u_matrix_opt"] = uo
    um = cplx(nw * nw * nk).reshape(nk, nw, nw).transpose(0, 2, 1)
    out["u_matrix"] = um
    mm = cplx(nw * nw * nn * nk).reshape(nk, nn, nw, nw).transpose(0, 1, 3, 2)
    out["m_matrix"] = mm
    cen = np.array([[float(x) for x in next(it).split()] for _ in range(nw)])
    spr = np.array([float(next(it)) for _ in range(nw)])
    out["centres"] = cen
    out["spreads"] = spr
    return out


def chk_gauges(chk):
    """U = U_opt * U_rot (all bands inside the window in these fixtures)."""
    if "u_matrix_opt" in chk:
        return chk["u_matrix_opt"] @ chk["u_matrix"]
    return chk["u_matrix"]


def read_model(prefix, ortho_amn=True, amn=None):
    """Restatement of read_w90 (src/io/w90/model.jl:16-94) with use_mmn_bvecs=true."""
    win = read_win(prefix + ".win")
    recip = so.reciprocal_lattice(win["lattice"])
    M, kpb_k, kpb_G = read_mmn(prefix + ".mmn")
    nk, nn, nb, _ = M.shape
    assert nk == len(win["kpoints"]) and nb == win["num_bands"]
    bvec = so.bvectors_cart(recip, win["kpoints"], kpb_k, kpb_G)
    wb = so.compute_bweights(bvec[0])           # kstencil.jl:68-85: weights from k-point 1
    A = read_amn(amn if amn else prefix + ".amn")
    U = so.orthonorm_lowdin(A) if ortho_amn else A
    model = dict(lattice=win["lattice"], recip=recip, kpoints=win["kpoints"], M=M, kpb_k=kpb_k,
                 kpb_G=kpb_G, bvec=bvec, wb=wb, U=U, nb=nb, nw=win["num_wann"], nk=nk, nn=nn)
    try:
        E = read_eig(prefix + ".eig")
        model["E"] = E
    except OSError:
        E = None
    frozen = np.zeros((nk, nb), dtype=bool)
    if nb != win["num_wann"] and E is not None and "dis_froz_max" in win:
        lo = win.get("dis_froz_min", -np.inf)
        frozen = (E >= lo) & (E <= win["dis_froz_max"])  # disentangle.jl:20-22
    model["frozen"] = frozen
    return model
