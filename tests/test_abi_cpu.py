"""CPU-only checks: the C-ABI library builds/loads and exports every symbol the header declares,
the host-side layout helpers agree with the oracle, the synthetic generators are well formed.
No compute call is made (there is no GPU here and the library has no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import spread_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(w):
    import __graft_entry__ as ge
    if not os.path.exists(w.lib.LIB_PATH):
        ge.build()
    hdr = open(os.path.join(ROOT, "include", "wannier_b200.h")).read()
    declared = set(re.findall(r"\b(wb200_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(w.lib.SYMBOLS)
    lib = ctypes.CDLL(w.lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert b"sm_100a" in w.lib.load().wb200_version()


def test_flag_constants_match_the_header(w):
    """the ctypes side must pass the bit values include/wannier_b200.h defines"""
    hdr = open(os.path.join(ROOT, "include", "wannier_b200.h")).read()
    flags = dict(re.findall(r"#define\s+WB200_FLAG_([A-Z_]+)\s+(\d+)u", hdr))
    assert set(flags) >= {"DEFAULT", "NO_TENSOR", "FORCE_TENSOR", "NO_FUSED_TRIAL", "M_ROW_MAJOR"}
    for name, val in flags.items():
        assert getattr(w.lib, "FLAG_" + name) == int(val), name
    vals = [int(v) for k, v in flags.items() if k != "DEFAULT"]
    assert len(set(vals)) == len(vals) and all(v & (v - 1) == 0 for v in vals)      # distinct single bits


def test_no_cpu_fallback_create_fails_without_device(w):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    st, M, U = w.synthetic.make_model_arrays("cubic", (2, 2, 2), 4, 4)
    with pytest.raises(w.WB200Error) as e:
        w.Context(4, 4, 8, 6, st["kpb_k"], st["bvec"], st["wb"], w.api.to_abi_M(M))
    assert e.value.status == -3


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "wannier.jl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


@pytest.mark.parametrize("lat,grid,nn", [("bcc", (4, 4, 4), 12), ("fcc", (3, 3, 3), 8),
                                         ("cubic", (3, 2, 2), 6), ("tetragonal", (10, 10, 1), 6)])
def test_synthetic_stencil(w, lat, grid, nn):
    st = w.synthetic.make_stencil(w.synthetic.LATTICES[lat](), grid)
    assert st["kpb_k"].shape == (np.prod(grid), nn)
    b = st["bvec"][0]
    assert np.allclose(np.einsum("b,bx,by->xy", st["wb"], b, b), np.eye(3), atol=1e-10)   # B1
    # same b through the reference formula (src/spread.jl:293)
    b2 = so.bvectors_cart(st["recip"], st["kpoints"], st["kpb_k"], st["kpb_G"])
    assert np.allclose(b2, st["bvec"], atol=1e-12)
    # every b has its -b partner at the neighbour: kpb(kpb(k, b), -b) == k
    for ib in range(nn):
        jb = int(np.argmin(np.abs(b + b[ib]).sum(axis=1)))
        assert np.allclose(b[jb], -b[ib])
        assert np.array_equal(st["kpb_k"][st["kpb_k"][:, ib], jb], np.arange(len(st["kpoints"])))
    if lat == "tetragonal":
        assert (st["kpb_k"][:, 4:] == np.arange(100)[:, None]).all()      # self-neighbours, G != 0
        assert np.abs(st["kpb_G"][:, 4:, 2]).min() == 1


def test_synthetic_overlaps_are_consistent(w):
    st, M, U = w.synthetic.make_model_arrays("bcc", (3, 3, 3), 6, 6)
    b = st["bvec"][0]
    for ib in range(12):
        jb = int(np.argmin(np.abs(b + b[ib]).sum(axis=1)))
        assert np.allclose(M[st["kpb_k"][:, ib], jb], np.conj(M[:, ib].transpose(0, 2, 1)), atol=1e-12)
    assert np.linalg.norm(M, ord=2, axis=(-2, -1)).max() <= 1 + 1e-12
    assert np.abs(np.conj(U.transpose(0, 2, 1)) @ U - np.eye(6)).max() < 1e-12
    sp = so.omega(M, st["kpb_k"], st["bvec"], st["wb"], U)
    assert sp["OmegaI"] > 1e-3 and np.isfinite(sp["Omega"])


def test_host_layout_helpers_match_oracle(w, silicon):
    nb, nw = 12, 8
    fro = silicon["frozen"]
    X, Y = so.U_to_X_Y(silicon["U"], fro)      # (the product's U_to_X_Y runs on the GPU: tests/test_gpu_parity.py)
    assert np.array_equal(w.api.X_Y_to_U(X, Y), so.X_Y_to_U(X, Y))
    XY = w.api.X_Y_to_XY(X, Y)
    assert np.array_equal(XY, so.X_Y_to_XY(X, Y))
    X2, Y2 = w.api.XY_to_X_Y(XY, nb, nw)
    assert np.array_equal(X, X2) and np.array_equal(Y, Y2)
    U = silicon["U"]
    assert np.array_equal(w.api.from_abi_U(w.api.to_abi_U(U)), U)
    # Julia column-major image: element (m, n) of U_k at n*nb + m
    assert w.api.to_abi_U(U).reshape(len(U), -1)[3, 2 * nb + 5] == U[3, 5, 2]
    st = w.KspaceStencil(silicon["recip"], silicon["kpoints"], silicon["wb"], silicon["kpb_k"], silicon["kpb_G"])
    assert np.allclose(st.bvectors_cart(), silicon["bvec"], atol=1e-12)
    m = w.Model(silicon["lattice"], st, silicon["M"], U, fro)
    assert (m.n_bands, m.n_wannier, m.n_kpoints, m.n_bvectors) == (12, 8, 64, 8)


def test_w90_ingest_matches_goldens(w, valence, silicon):
    """Product-side parsers (wannier.jl_b200/io_w90.py) against the goldens the oracle-side parsers
    produced from the same reference fixtures: two independent implementations of the formats."""
    import os
    gold = os.path.join(ROOT, "tests", "golden")
    io = w.io_w90
    win = io.read_win(os.path.join(gold, "w90_valence", "silicon.win"))
    assert np.allclose(win["lattice"], valence["lattice"]) and np.allclose(win["kpoints"], valence["kpoints"])
    assert (win["num_wann"], win["num_bands"], win["mp_grid"]) == (4, 4, [4, 4, 4])
    M, kpb_k, kpb_G = io.read_mmn(os.path.join(gold, "w90_valence", "silicon.mmn"))
    assert np.array_equal(M, valence["M"]) and np.array_equal(kpb_k, valence["kpb_k"]) and np.array_equal(kpb_G, valence["kpb_G"])
    recip = 2 * np.pi * np.linalg.inv(win["lattice"]).T
    st = w.KspaceStencil(recip, win["kpoints"], np.ones(8), kpb_k, kpb_G)
    assert np.allclose(st.bvectors_cart(), valence["bvec"], atol=1e-13)
    assert np.allclose(io.compute_bweights(st.bvectors_cart()[0]), valence["wb"], rtol=1e-12)
    A = io.read_amn(os.path.join(gold, "w90_valence", "silicon.amn"))
    assert np.abs(so.orthonorm_lowdin(A) - valence["U"]).max() < 1e-13
    # frozen window of the 12 -> 8 silicon model (dis_froz_max = 10 eV)
    win2 = io.read_win(os.path.join(gold, "w90_silicon", "silicon.win"))
    E = io.read_eig(os.path.join(gold, "w90_silicon", "silicon.eig"))
    assert np.array_equal(E, silicon["E"])
    assert np.array_equal(io.frozen_bands(E, win2["dis_froz_max"], win2.get("dis_froz_min", -np.inf)), silicon["frozen"])
    assert (win2["num_wann"], win2["num_bands"]) == (8, 12)


def test_native_ingest_edge_cases(w, tmp_path):
    """wb200_mmn_read / wb200_amn_read / wb200_eig_read: pairs not grouped by k-point, Fortran 'D' exponents,
    blank lines, every thread count, malformed input (WannierIO.jl's read_mmn / read_amn / read_eig semantics:
    b = order of appearance of a k-point's pairs, 1-based indices in the file)."""
    rng = np.random.default_rng(5)
    nb, nk, nn = 3, 4, 2
    M = rng.standard_normal((nk, nn, nb, nb)) + 1j * rng.standard_normal((nk, nn, nb, nb))
    kpb = rng.integers(0, nk, size=(nk, nn))
    G = rng.integers(-1, 2, size=(nk, nn, 3))
    order = [(k, b) for b in range(nn) for k in range(nk)]        # b-major: NOT grouped by k-point
    lines = ["generated by the test", f"  {nb}  {nk}  {nn}"]
    for k, b in order:
        lines.append(f"  {k + 1}  {kpb[k, b] + 1}  {G[k, b, 0]}  {G[k, b, 1]}  {G[k, b, 2]}")
        for l in range(nb):
            for m in range(nb):
                z = M[k, b, m, l]
                re = f"{z.real:.17e}".replace("e", "D") if (l + m) % 2 else f"{z.real:+.17e}"
                lines.append(f"  {re}   {z.imag:.17e}")
        lines.append("")                                            # blank lines are skipped
    path = tmp_path / "t.mmn"
    path.write_text("\n".join(lines) + "\n")
    for nthreads in (1, 2, 3, 7, 0):
        Mc, kk, GG = w.lib.read_mmn_native(str(path), nthreads)
        assert np.array_equal(Mc.transpose(0, 1, 3, 2), M) and np.array_equal(kk, kpb) and np.array_equal(GG, G)
    M2, k2, G2 = w.io_w90.read_mmn(str(path))
    assert np.array_equal(M2, M)
    # .amn / .eig
    nw = 2
    A = rng.standard_normal((nk, nb, nw)) + 1j * rng.standard_normal((nk, nb, nw))
    al = ["amn", f"{nb} {nk} {nw}"] + [f"{m + 1} {n + 1} {k + 1} {A[k, m, n].real:.17e} {A[k, m, n].imag:.17e}"
                                        for k in range(nk) for n in range(nw) for m in range(nb)]
    (tmp_path / "t.amn").write_text("\n".join(al) + "\n")
    assert np.array_equal(w.io_w90.read_amn(str(tmp_path / "t.amn")), A)
    E = rng.standard_normal((nk, nb))
    (tmp_path / "t.eig").write_text("\n".join(f"{m + 1} {k + 1} {E[k, m]:.17e}" for k in range(nk) for m in range(nb)) + "\n")
    assert np.array_equal(w.io_w90.read_eig(str(tmp_path / "t.eig")), E)
    # truncated file, malformed number, missing file
    (tmp_path / "short.mmn").write_text("\n".join(lines[:20]) + "\n")
    with pytest.raises(w.WB200Error, match="ends before"):
        w.lib.read_mmn_native(str(tmp_path / "short.mmn"))
    bad = list(lines); bad[5] = "  0.1  abc"
    (tmp_path / "bad.mmn").write_text("\n".join(bad) + "\n")
    with pytest.raises(w.WB200Error, match="malformed"):
        w.lib.read_mmn_native(str(tmp_path / "bad.mmn"))
    with pytest.raises(w.WB200Error, match="cannot open"):
        w.lib.read_mmn_native(str(tmp_path / "nope.mmn"))


def test_julia_wrapper_binds_only_exported_symbols(w):
    """julia/WannierB200Ext.jl cannot be executed here (no Julia runtime): at least every symbol it `ccall`s must be
    exported by the library, and the reference functions north_star names must each have a binding."""
    import os, re, ctypes
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "julia", "WannierB200Ext.jl")).read()
    syms = set(re.findall(r"\(:(wb200_[a-z0-9_]+), LIB\)", src))
    lib = ctypes.CDLL(w.lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), s
    for needed in ("wb200_fg", "wb200_spread", "wb200_center", "wb200_max_localize", "wb200_disentangle",
                   "wb200_rotate_overlaps", "wb200_opt_rotate", "wb200_position_op", "wb200_berry_connection",
                   "wb200_omega_local", "wb200_u_to_xy", "wb200_multi_create", "wb200_multi_max_localize",
                   "wb200_mag_disentangle", "wb200_host_register", "wb200_retract", "wb200_project_tangent"):
        assert needed in syms, needed
    for fn in ("omega(", "omega_grad(", "center(", "max_localize(", "disentangle(", "get_fg!_maxloc(", "transform_gauge(",
               "compute_MU_UtMU!(", "omega!(", "omega_grad!(", "opt_rotate(", "position_op(", "U_to_X_Y("):
        assert fn in src, fn
