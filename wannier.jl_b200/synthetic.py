"""Deterministic synthetic inputs of the BASELINE shapes (host setup, not on the iterated path).

Stencil: Monkhorst-Pack grid, l fastest (src/kpoints/kpoint.jl:115-140), b-vector shells found
by brute force over neighbour offsets and weighted by the B1 condition (the algorithm of
src/kpoints/kstencil.jl:132-170 / kstencil_shell.jl:341-391: add shells by increasing |b|, drop
a shell that makes the system singular, stop when sum_b w_b b b^T = I).

Overlaps: a smooth family of nb-dimensional subspaces of C^nH,
    B(k) = [I; 0] + s * sum_R c_R exp(2 pi i k.R)   (|R|_inf <= 1, c_R complex Gaussian nH x nb),
    V_k = polar(B(k)),   M_kb = V_k^H V_{k+b}
so that M_{k+b,-b} = M_kb^H, ||M||_2 <= 1 (non-unitary: Omega_I > 0) and diag(N) stays away
from 0.  Gauge: U_k = polar([I; 0] + eps Z_k), Z_k complex Gaussian.
The same recipe exists in numpy (tests, small shapes, shared with the CPU oracle) and in torch
on the GPU (bench at the 13 GB north-star shape); both are seeded.
"""
from __future__ import annotations

import itertools

import numpy as np

SEED_M = 20240601
SEED_U = 20240602


def reciprocal_lattice(lattice):
    """columns of `lattice` are the lattice vectors; returns columns = reciprocal vectors."""
    return 2.0 * np.pi * np.linalg.inv(np.asarray(lattice, dtype=np.float64)).T


LATTICES = {
    # real-space bcc (reciprocal fcc): 12 nearest neighbours in one shell
    "bcc": lambda a=5.0: a / 2 * np.array([[-1.0, 1, 1], [1, -1, 1], [1, 1, -1]]).T,
    # real-space fcc (reciprocal bcc): 8 nearest neighbours
    "fcc": lambda a=5.0: a / 2 * np.array([[0.0, 1, 1], [1, 0, 1], [1, 1, 0]]).T,
    "cubic": lambda a=5.0: a * np.eye(3),
    "tetragonal": lambda a=5.0, c=20.0: np.diag([a, a, c]),
}


def kgrid(n1, n2, n3):
    i, j, l = np.meshgrid(np.arange(n1), np.arange(n2), np.arange(n3), indexing="ij")
    return np.stack([i.ravel() / n1, j.ravel() / n2, l.ravel() / n3], axis=1)


def _b1_weights(shell_vectors, atol=1e-6, sigma_atol=1e-5):
    """shell_vectors: list of [n_s, 3] arrays by increasing |b|.  Returns (kept shell ids, w)."""
    iu = [(0, 0), (0, 1), (1, 1), (0, 2), (1, 2), (2, 2)]
    tri = lambda m: np.array([m[a, b] for a, b in iu])
    target = tri(np.eye(3))
    cols, keep = [], []
    for s, bv in enumerate(shell_vectors):
        cols.append(tri(bv.T @ bv))
        keep.append(s)
        Bm = np.stack([cols[t] for t in keep], axis=1)
        Um, S, Vh = np.linalg.svd(Bm, full_matrices=False)
        if np.all(S > sigma_atol):
            w = Vh.T @ ((Um.T @ target) / S)
            if np.allclose(Bm @ w, target, rtol=0, atol=atol):
                return keep, w
        else:
            keep.pop()
    raise ValueError("not enough shells to satisfy the B1 condition")


def make_stencil(lattice, grid, search=2):
    """Returns dict(recip, kpoints, kpb_k [nk, nn] int32, kpb_G [nk, nn, 3] int32, wb [nn],
    bvec [nk, nn, 3])."""
    n = np.array(grid, dtype=np.int64)
    recip = reciprocal_lattice(lattice)
    offs = [d for d in itertools.product(range(-search, search + 1), repeat=3) if any(d)]
    # with one k-point along an axis every offset along it is a pure G shift; keep |d| <= 1 there
    offs = [d for d in offs if all(abs(d[a]) <= 1 or n[a] > 1 for a in range(3))]
    offs = np.array(offs, dtype=np.int64)
    bcart = (offs / n) @ recip.T
    norms = np.linalg.norm(bcart, axis=1)
    order = np.argsort(norms, kind="stable")
    shells, i = [], 0
    while i < len(order):
        j = i
        while j < len(order) and abs(norms[order[j]] - norms[order[i]]) <= 1e-6:
            j += 1
        shells.append(order[i:j])
        i = j
    keep, w = _b1_weights([bcart[s] for s in shells[:12]])
    sel = np.concatenate([shells[s] for s in keep])
    wb = np.concatenate([np.full(len(shells[s]), w[t]) for t, s in enumerate(keep)])
    d = offs[sel]                                           # [nn, 3]
    idx = np.stack(np.meshgrid(*[np.arange(x) for x in n], indexing="ij"), axis=-1).reshape(-1, 3)
    tgt = idx[:, None, :] + d[None, :, :]                   # [nk, nn, 3]
    G = np.floor_divide(tgt, n)
    wrapped = tgt - G * n
    kpb_k = (wrapped[..., 0] * n[1] + wrapped[..., 1]) * n[2] + wrapped[..., 2]
    kpoints = idx / n
    bvec = np.broadcast_to(bcart[sel], (len(idx), len(sel), 3)).copy()
    return dict(recip=recip, kpoints=kpoints, kpb_k=kpb_k.astype(np.int32), kpb_G=G.astype(np.int32),
                wb=wb, bvec=bvec, lattice=np.asarray(lattice, dtype=np.float64))


def _n_hilbert(nb):
    return nb + max(8, nb // 4)


def _coeffs(nb, seed):
    """c_R for |R|_inf <= 1 and the k-independent part; returns (Rs [27,3], c [27, nH, nb])."""
    nH = _n_hilbert(nb)
    rng = np.random.default_rng(seed)
    Rs = np.array(list(itertools.product((-1, 0, 1), repeat=3)), dtype=np.float64)
    c = rng.standard_normal((27, nH, nb)) + 1j * rng.standard_normal((27, nH, nb))
    sig = 0.35 / np.sqrt(nH) / (1.0 + np.abs(Rs).sum(axis=1))
    return Rs, c * sig[:, None, None]


def _polar_np(A):
    W, _, Vh = np.linalg.svd(A, full_matrices=False)
    return W @ Vh


def _frames(st, nb, Rs, c):
    """orthonormal frames V[k] (nH x nb) of the smooth subspace: polar factor of sum_R c_R e^{2 pi i k.R} + [I; 0]"""
    phase = np.exp(2j * np.pi * (st["kpoints"] @ Rs.T))            # [nk, 27]
    B = np.einsum("kr,rhn->khn", phase, c)
    B[:, :nb, :] += np.eye(nb)
    return _polar_np(B)                                            # [nk, nH, nb]


def _overlaps_of(st, V):
    Vh = np.conj(V.transpose(0, 2, 1))
    return np.ascontiguousarray(Vh[:, None] @ V[st["kpb_k"]])       # [nk, nn, nb, nb]


def make_overlaps(st, nb, seed=SEED_M):
    """numpy: M[k, b, m, l] complex128."""
    Rs, c = _coeffs(nb, seed)
    return _overlaps_of(st, _frames(st, nb, Rs, c))


def make_gauge(nk, nb, nw, eps=0.05, seed=SEED_U):
    rng = np.random.default_rng(seed)
    Z = rng.standard_normal((nk, nb, nw)) + 1j * rng.standard_normal((nk, nb, nw))
    A = eps * Z
    A[:, :nw, :] += np.eye(nw)
    return np.ascontiguousarray(_polar_np(A))


def make_frozen(nk, nb, n_frozen):
    fr = np.zeros((nk, nb), dtype=bool)
    fr[:, :n_frozen] = True
    return fr


CONFIGS = {
    # name: (lattice name, grid, nb, nw)  — BASELINE.json configs
    "cfg1_si_valence_shape": ("fcc", (4, 4, 4), 4, 4),
    "cfg2_cu_shape": ("fcc", (9, 9, 9), 18, 9),
    "cfg3_nw32": ("bcc", (12, 12, 12), 32, 32),
    "cfg4_nw128": ("bcc", (16, 16, 16), 128, 128),
    "cfg5_2d_nw256": ("tetragonal", (10, 10, 1), 256, 256),
}


def make_model_arrays(lattice_name, grid, nb, nw, eps=0.05):
    """Everything a Context needs, as numpy arrays with math indexing."""
    st = make_stencil(LATTICES[lattice_name](), grid)
    M = make_overlaps(st, nb)
    U = make_gauge(len(st["kpoints"]), nb, nw, eps)
    return st, M, U


def make_mag_arrays(lattice_name, grid, nb, nw, eps=0.05, mix=0.3):
    """A synthetic spin-polarised pair for the MagModel path (src/localization/coopt.jl): the down frames are a
    smooth perturbation of the up frames, so Mud[k, m, l] = <u_mk^up | u_lk^dn> is close to the identity, as
    for a weakly polarised material.  Returns (st, M_up, M_dn, Mud, U_up, U_dn)."""
    st = make_stencil(LATTICES[lattice_name](), grid)
    Rs, c_up = _coeffs(nb, SEED_M)
    _, c_2 = _coeffs(nb, SEED_M + 1)
    V_up = _frames(st, nb, Rs, c_up)
    V_dn = _frames(st, nb, Rs, c_up + mix * c_2)
    nk = len(st["kpoints"])
    Mud = np.ascontiguousarray(np.conj(V_up.transpose(0, 2, 1)) @ V_dn)
    return (st, _overlaps_of(st, V_up), _overlaps_of(st, V_dn), Mud,
            make_gauge(nk, nb, nw, eps, SEED_U), make_gauge(nk, nb, nw, eps, SEED_U + 1))


# ---- torch / GPU generation for the large bench shapes (plumbing only; same recipe) -----------
def _polar_torch(A, iters=60):
    """polar factor by Newton-Schulz on the GPU (matmul only); A is close to an isometry."""
    import torch
    eye = torch.eye(A.shape[-1], dtype=A.dtype, device=A.device)
    # scale so that every singular value is <= 1 (sigma_max^2 <= ||A^H A||_F): Newton-Schulz then
    # converges monotonically for any full-rank A
    P = A.mH @ A
    X = A * torch.linalg.matrix_norm(P).pow(-0.5)[..., None, None].to(A.dtype)
    for _ in range(iters):
        P = X.mH @ X
        err = (P - eye).abs().amax().item()
        X = X @ (1.5 * eye - 0.5 * P)
        if err < 1e-8:
            break
    return X


def make_overlaps_torch(st, nb, device, seed=SEED_M, k_begin=0, nk=None, abi_layout=True):
    """M for the k-slab [k_begin, k_begin+nk) on `device`; with abi_layout the result is already
    in the C-ABI memory image [nk][nn][l][m] (column-major nb x nb per pair)."""
    import torch
    Rs, c = _coeffs(nb, seed)
    kp = torch.as_tensor(st["kpoints"], device=device)
    phase = torch.exp(2j * np.pi * (kp @ torch.as_tensor(Rs, device=device).T))
    ct = torch.as_tensor(c, device=device)
    B = torch.einsum("kr,rhn->khn", phase, ct)
    B[:, :nb, :] += torch.eye(nb, dtype=B.dtype, device=device)
    V = _polar_torch(B)
    del B
    nk_total = V.shape[0]
    nk = nk_total if nk is None else nk
    kpb = torch.as_tensor(st["kpb_k"][k_begin:k_begin + nk].astype(np.int64), device=device)
    nn = kpb.shape[1]
    M = torch.empty((nk, nn, nb, nb), dtype=torch.complex128, device=device)
    Vk = V[k_begin:k_begin + nk]
    for b in range(nn):
        Mb = Vk.mH @ V[kpb[:, b]]                      # M[k, m, l]
        M[:, b] = Mb.transpose(1, 2) if abi_layout else Mb
    return M


def make_gauge_torch(nk, nb, nw, device, eps=0.05, seed=SEED_U, abi_layout=True):
    """U on `device`; with abi_layout in the C-ABI image [nk][n][m]."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    Z = torch.randn((nk, nb, nw, 2), generator=g, dtype=torch.float64)
    A = (eps * torch.view_as_complex(Z)).to(device)
    A[:, :nw, :] += torch.eye(nw, dtype=A.dtype, device=device)
    U = _polar_torch(A)
    return U.transpose(1, 2).contiguous() if abi_layout else U
