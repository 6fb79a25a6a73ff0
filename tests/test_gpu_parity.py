"""GPU parity tests (-m gpu): every call goes through the C-ABI of libwannier_b200.so and is
compared with the CPU oracle on identical inputs.  Tolerances: 1e-10 relative for FP64 results
(BASELINE north_star); OmegaD, a difference of near-equal numbers, uses 1e-10 * Omega absolute;
max_localize must reach the reference optimum within 1e-8 A^2."""
import numpy as np
import pytest

from oracle import spread_oracle as so
from conftest import make_ctx, fixture_stencil
import optim_goldens as og

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def check_spread(got, ref):
    out5, om, r, _ = got
    scale = max(abs(ref["Omega"]), 1.0)   # degenerate stencils give Omega ~ 1e-14: absolute floor
    for i, key in enumerate(["Omega", "OmegaI", "OmegaOD", "OmegaD", "OmegaTilde"]):
        assert abs(out5[i] - ref[key]) <= RTOL * scale, (key, out5[i], ref[key])
    assert np.abs(om - ref["omega"]).max() <= RTOL * max(np.abs(ref["omega"]).max(), 1.0)
    assert np.abs(r - ref["r"]).max() <= RTOL * max(1.0, np.abs(ref["r"]).max())


def oracle_args(d):
    return d["M"], d["kpb_k"], d["bvec"], d["wb"]


# ---------------------------------------------------------------------------------------------
# reference fixtures (BASELINE config 1 and the 12 -> 8 silicon model)
# ---------------------------------------------------------------------------------------------
def test_valence_spread_and_known_answers(w, valence):
    ctx = make_ctx(w, fixture_stencil(valence), valence["M"], 4)
    assert ctx.rotation_kernel == "dmma-fp64"      # nb <= 32: warp-pair-per-pair tensor kernel (zero padded)
    ctx_cc = make_ctx(w, fixture_stencil(valence), valence["M"], 4, flags=w.lib.FLAG_NO_TENSOR)
    assert ctx_cc.rotation_kernel == "cuda-core-fp64"
    check_spread(ctx_cc.spread(w.api.to_abi_U(valence["U"])), so.omega(*oracle_args(valence), valence["U"]))
    ctx_cc.close()
    for U in (valence["U"], valence["U_ptg"], valence["U"] @ valence["W1"]):
        got = ctx.spread(w.api.to_abi_U(U))
        check_spread(got, so.omega(*oracle_args(valence), U))
    # reference constant: test/localization/max_localize.jl:53 (OmegaI is gauge invariant)
    assert abs(ctx.spread(w.api.to_abi_U(valence["U_ptg"]))[0][1] - valence["ref_maxloc"][1]) < 1e-11
    r = ctx.center(w.api.to_abi_U(valence["U"]))
    assert np.abs(r - so.center(*oracle_args(valence), valence["U"])).max() < 1e-10
    ctx.close()


def test_valence_fg_matches_oracle_and_finite_differences(w, valence):
    # test/localization/max_localize.jl:11-43
    ctx = make_ctx(w, fixture_stencil(valence), valence["M"], 4)
    rng = np.random.default_rng(1)
    for U in (valence["U"], valence["U"] @ valence["W1"]):
        G = np.empty((64, 4, 4), dtype=np.complex128)
        f = ctx.fg(w.api.to_abi_U(U), G=G)
        G = w.api.from_abi_U(G)
        f_ref, G_ref = so.fg_maxloc(*oracle_args(valence), U)
        assert abs(f - f_ref) < RTOL * abs(f_ref)
        assert rel(G, G_ref) < RTOL
        d = rng.standard_normal(U.shape) + 1j * rng.standard_normal(U.shape)
        h = 1e-5
        fd = (ctx.fg(w.api.to_abi_U(U + h * d)) - ctx.fg(w.api.to_abi_U(U - h * d))) / (2 * h)
        assert abs(fd - np.real(np.vdot(G, d))) < 1e-6 * max(1.0, abs(fd))
    # only_fg! convention: F or G may be skipped
    assert ctx.fg(w.api.to_abi_U(valence["U"]), want_f=False, G=np.empty((64, 4, 4), dtype=np.complex128)) is None
    ctx.close()


def test_center_penalty(w, valence):
    # test/localization/constrain_center/max_localize.jl: CenterSpreadPenalty(r0, lambda = 10)
    ctx = make_ctx(w, fixture_stencil(valence), valence["M"], 4)
    rng = np.random.default_rng(5)
    r0 = rng.standard_normal((4, 3))
    lam = 10.0
    ctx.set_penalty(r0, lam)
    U = valence["U"]
    G = np.empty((64, 4, 4), dtype=np.complex128)
    f = ctx.fg(w.api.to_abi_U(U), G=G)
    f_ref, G_ref = so.fg_maxloc(*oracle_args(valence), U, r0, lam)
    assert abs(f - f_ref) < RTOL * abs(f_ref)
    assert rel(w.api.from_abi_U(G), G_ref) < RTOL
    ctx.set_penalty(None, 0.0)
    assert abs(ctx.fg(w.api.to_abi_U(U)) - so.fg_maxloc(*oracle_args(valence), U)[0]) < 1e-9
    ctx.close()
    # host mirror: omega(CenterSpreadPenalty, ...) -> SpreadCenter (src/spread.jl:221, 246-252)
    st = w.KspaceStencil(valence["recip"], valence["kpoints"], valence["wb"], valence["kpb_k"], valence["kpb_G"])
    sc = w.omega(w.CenterSpreadPenalty(r0, lam), st, valence["M"], U)
    ref = so.omega_center(so.omega(*oracle_args(valence), U), r0, lam)
    assert abs(sc.Omegat - ref["Omegat"]) < RTOL * abs(ref["Omegat"])
    assert rel(sc.omegat, ref["omegat"]) < RTOL


def test_silicon_rotation_matches_w90_chk(w, silicon):
    # test/io/model.jl:38-75 analogue: U^H M U == chk.m_matrix (rectangular U: 12 -> 8)
    ctx = make_ctx(w, fixture_stencil(silicon), silicon["M"], 8)
    N, MU = ctx.rotate_overlaps(w.api.to_abi_U(silicon["U_chk"]), want_N=True, want_MU=True)
    N = N.transpose(0, 1, 3, 2)
    assert np.abs(N - silicon["chk_m_matrix"]).max() < 1e-13
    MU_ref, N_ref = so.compute_MU_UtMU(silicon["M"], silicon["kpb_k"], silicon["U_chk"])
    assert rel(MU.transpose(0, 1, 3, 2), MU_ref) < 1e-13
    assert rel(N, N_ref) < 1e-13
    got = ctx.spread(w.api.to_abi_U(silicon["U_chk"]))
    check_spread(got, so.omega(*oracle_args(silicon), silicon["U_chk"]))
    ref = silicon["wout_final"]       # W90's own numbers (slightly different Bohr): 3e-8 relative
    assert np.allclose(got[0][:4], ref, rtol=3e-8, atol=0)
    ctx.close()


def test_split_overlaps(w, silicon, silicon_split):
    # test/localization/split.jl:17-47
    for V, ref in ((silicon_split["Vv"], silicon_split["Mv_ref"]), (silicon_split["Vc"], silicon_split["Mc_ref"])):
        N = w.transform_gauge(silicon["M"], silicon["kpb_k"], silicon["U_chk"] @ V)
        assert np.abs(N - ref).max() < 1e-7


def test_silicon_rectangular_fg_and_disentangle_chain(w, silicon):
    # test/localization/disentangle.jl:27-54
    nb, nw = 12, 8
    ctx = make_ctx(w, fixture_stencil(silicon), silicon["M"], nw)
    U = silicon["U"]
    G = np.empty((64, nw, nb), dtype=np.complex128)
    f = ctx.fg(w.api.to_abi_U(U), G=G)
    f_ref, G_ref = so.fg_maxloc(*oracle_args(silicon), U)
    assert abs(f - f_ref) < RTOL * abs(f_ref)
    assert rel(w.api.from_abi_U(G), G_ref) < RTOL
    fro = silicon["frozen"]
    ctx.set_frozen(fro)
    X, Y = so.U_to_X_Y(U, fro)
    XY = so.X_Y_to_XY(X, Y)
    Gxy = np.empty_like(XY)
    f = ctx.fg_xy(np.ascontiguousarray(XY), G=Gxy)
    f_ref, G_ref = so.fg_disentangle(*oracle_args(silicon), XY, nb, nw, fro)
    assert abs(f - f_ref) < RTOL * abs(f_ref)
    assert rel(Gxy, G_ref) < RTOL
    ctx.close()


@pytest.mark.parametrize("lat,grid,nb,nw,nfroz", [
    ("fcc", (3, 3, 3), 18, 9, 5),      # config 2 shape: fused single-launch (X, Y) kernels (nb <= 64)
    ("cubic", (2, 2, 2), 40, 24, 7),   # fused kernels, CUDA-core Stiefel kernels (32 < n <= 64)
    ("fcc", (2, 2, 2), 72, 40, 9),     # nb > 64: batched-GEMM (X, Y) chain
])
def test_synthetic_xy_chain_and_disentangle(w, lat, grid, nb, nw, nfroz):
    # disentangle.jl:351-356, 424-466, 481-548: U = Y X, G_X = Y^H G_U, G_Y = G_U X^H with the frozen masks
    st, M, U = w.synthetic.make_model_arrays(lat, grid, nb, nw)
    nk = len(U)
    frozen = w.synthetic.make_frozen(nk, nb, nfroz)
    ctx = make_ctx(w, st, M, nw)
    ctx.set_frozen(frozen)
    X, Y = so.U_to_X_Y(U, frozen)
    XY = np.ascontiguousarray(so.X_Y_to_XY(X, Y))
    Gxy = np.empty_like(XY)
    f = ctx.fg_xy(XY, G=Gxy)
    f_ref, G_ref = so.fg_disentangle(M, st["kpb_k"], st["bvec"], st["wb"], XY, nb, nw, frozen)
    assert abs(f - f_ref) < RTOL * abs(f_ref)
    assert rel(Gxy, G_ref) < RTOL
    # a few steps of the device driver: the spread goes down, U stays semi-unitary, frozen states stay in
    XYd = XY.copy()
    f1, iters, nfg = ctx.disentangle(XYd, 1e-7, 1e-5, 5, 3)
    assert f1 < f_ref and iters >= 1
    Xd, Yd = so.XY_to_X_Y(XYd, nb, nw)
    Ud = so.X_Y_to_U(Xd, Yd)
    assert abs(so.omega(M, st["kpb_k"], st["bvec"], st["wb"], Ud)["Omega"] - f1) < 1e-9 * abs(f1)
    assert np.abs(np.conj(Ud.transpose(0, 2, 1)) @ Ud - np.eye(nw)).max() < 1e-11
    assert np.abs((np.abs(Ud) ** 2).sum(axis=2)[frozen] - 1).max() < 1e-10
    ctx.close()


@pytest.mark.parametrize("dmma", [None, "0", "1"])
def test_stiefel_project_and_retract(w, valence, silicon, dmma):
    # n <= 32: tensor-pipe or CUDA-core kernels, chosen by the amount of work; WB200_STIEFEL_DMMA = 0 / 1 forces one
    import os
    os.environ.pop("WB200_STIEFEL_DMMA", None)
    if dmma is not None:
        os.environ["WB200_STIEFEL_DMMA"] = dmma
    try:
        _stiefel_project_and_retract(w, valence, silicon)
    finally:
        os.environ.pop("WB200_STIEFEL_DMMA", None)


def _stiefel_project_and_retract(w, valence, silicon):
    ctx = make_ctx(w, fixture_stencil(valence), valence["M"], 4)
    rng = np.random.default_rng(7)
    # shapes above 64 run on the plain-layout tensor GEMM (zgemm_dmma.cu): ragged m / n / k, one and
    # several row / column tiles, both CTA shapes (m <= 64: 64 x 64 tiles; above: 128 x 32), A^H and beta = 1
    for (nrow, ncol, count) in ((4, 4, 64), (12, 8, 17), (128, 128, 3), (40, 24, 5), (32, 32, 9), (31, 17, 6),
                               (9, 9, 5), (18, 9, 7), (1, 1, 3), (33, 32, 2), (72, 40, 5), (100, 100, 7),
                               (130, 70, 3), (257, 129, 2), (65, 65, 11), (200, 33, 4), (256, 256, 2),
                               (96, 64, 301), (32, 32, 200), (24, 20, 600)):
        X = so.stiefel_retract(rng.standard_normal((count, nrow, ncol)) + 1j * rng.standard_normal((count, nrow, ncol)))
        G = rng.standard_normal(X.shape) + 1j * rng.standard_normal(X.shape)
        Gc = w.api.to_abi_U(G)
        ctx.project_tangent(w.api.to_abi_U(X), Gc, nrow, ncol)
        assert rel(w.api.from_abi_U(Gc), so.stiefel_project_tangent(G, X)) < 1e-12
        Xp = X + 0.1 * G
        Xc = w.api.to_abi_U(Xp)
        ctx.retract(Xc, nrow, ncol)
        assert np.abs(w.api.from_abi_U(Xc) - so.stiefel_retract(Xp)).max() < 1e-12
    # far-from-unitary input (projections from an .amn): polar factor == Lowdin (src/utils/linalg.jl:25-29)
    A = silicon["U"] * 3.0 + 0.3 * (rng.standard_normal(silicon["U"].shape))
    Ac = w.api.to_abi_U(A)
    ctx.retract(Ac, 12, 8)
    assert np.abs(w.api.from_abi_U(Ac) - so.orthonorm_lowdin(A)).max() < 1e-11
    ctx.close()


def _model(w, d, U=None, frozen=False):
    st = w.KspaceStencil(d["recip"], d["kpoints"], d["wb"], d["kpb_k"], d["kpb_G"])
    return w.Model(d["lattice"], st, d["M"], d["U"] if U is None else U, d["frozen"] if frozen else None)


def _as_dict(sp):
    return dict(Omega=sp.Omega, OmegaI=sp.OmegaI, OmegaOD=sp.OmegaOD, OmegaD=sp.OmegaD,
                OmegaTilde=sp.OmegaTilde, omega=sp.omega, r=sp.r)


def test_max_localize_reaches_reference_optimum(w, valence):
    # test/localization/max_localize.jl:45-56 verbatim: max_localize(p, model) with the DEFAULT
    # tolerances (f_tol 1e-7, g_tol 1e-5), then the reference's own assertions at atol 1e-7;
    # north_star: with tight tolerances within 1e-8 A^2
    ref = valence["ref_maxloc"]
    model = _model(w, valence, valence["U_ptg"])
    U, info = w.max_localize(model, return_info=True)             # reference default tolerances
    sp = w.omega(model, U)
    og.check_spread(_as_dict(sp), og.MAXLOC_DEFAULT, atol=1e-7)
    # the device driver is Optim's LBFGS + HagerZhang step for step: same iterations, evaluations and
    # spread as the oracle's restatement (which the reference's max_iter = 4 goldens pin)
    Uo, fo, ito, infoo = so.max_localize(*oracle_args(valence), valence["U_ptg"])
    assert info["iterations"] == ito and info["n_fg"] == infoo["n_fg"]
    assert abs(info["f"] - fo) < 1e-9
    assert np.abs(U - Uo).max() < 1e-6
    U, info = w.max_localize(model, f_tol=1e-13, g_tol=1e-8, max_iter=500, return_info=True)
    sp = so.omega(*oracle_args(valence), U)
    assert abs(sp["Omega"] - ref[0]) < 1e-8
    assert abs(sp["OmegaI"] - ref[1]) < 1e-8
    assert abs(sp["OmegaOD"] - ref[2]) < 1e-8
    assert abs(info["f"] - sp["Omega"]) < 1e-10
    assert np.abs(np.conj(U.transpose(0, 2, 1)) @ U - np.eye(4)).max() < 1e-12
    # other history sizes (the Gram pass is chunked above 3 pairs): same optimum
    for hist in (1, 5, 9):
        U, info = w.max_localize(model, f_tol=1e-13, g_tol=1e-8, max_iter=500, history_size=hist, return_info=True)
        assert abs(info["f"] - ref[0]) < 1e-8, hist
    with pytest.raises(ValueError):   # max_localize.jl:52-53
        st = model.kstencil
        w.max_localize(w.Model(valence["lattice"], st, np.zeros((64, 8, 5, 5), complex), np.zeros((64, 5, 4), complex)))


def test_optimiser_trajectory_goldens(w, valence, silicon):
    """The reference's trajectory-dependent known answers (Optim LBFGS + HagerZhang, max_iter = 4),
    through the public API and the device-resident driver, at the reference's own atol 1e-7."""
    # test/localization/disentangle.jl:56-97
    model = _model(w, silicon, frozen=True)
    U, info = w.disentangle(model, max_iter=4, return_info=True)
    assert info["iterations"] == 4
    og.check_spread(_as_dict(w.omega(model, U)), og.DISENTANGLE_IT4)
    # test/localization/constrain_center/disentangle.jl:43-120
    ref = og.CC_DISENTANGLE_IT4
    p = w.CenterSpreadPenalty(ref["r0"], ref["lam"])
    U, info = w.disentangle(p, model, max_iter=4, return_info=True)
    spc = w.omega(p, model, U)
    og.check_spread(_as_dict(spc), ref)
    assert abs(spc.Omegac - ref["Omegac"]) < 1e-7 and abs(spc.Omegat - ref["Omegat"]) < 1e-7
    assert abs(info["f"] - ref["Omegat"]) < 1e-7
    # test/localization/constrain_center/max_localize.jl:46-96
    ref = og.CC_MAXLOC_IT4
    model = _model(w, valence, valence["U_ptg"])
    p = w.CenterSpreadPenalty(ref["r0"], ref["lam"])
    U, info = w.max_localize(p, model, max_iter=4, return_info=True)
    spc = w.omega(p, model, U)
    og.check_spread(_as_dict(spc), ref)
    assert abs(spc.Omegac - ref["Omegac"]) < 1e-7 and abs(spc.Omegat - ref["Omegat"]) < 1e-7
    assert np.abs(spc.omegac - ref["omegac"]).max() < 1e-7


def test_disentangle_driver(w, silicon):
    # the device driver against the oracle's driver on the 12 -> 8 silicon model (same iterations, same
    # evaluations, same spread), semi-unitarity and frozen states kept
    model = _model(w, silicon, frozen=True)
    X0, Y0 = so.U_to_X_Y(silicon["U"], silicon["frozen"])
    f0 = so.omega(*oracle_args(silicon), so.X_Y_to_U(X0, Y0))["Omega"]
    U, info = w.disentangle(model, max_iter=30, return_info=True)
    sp = so.omega(*oracle_args(silicon), U)
    assert abs(sp["Omega"] - info["f"]) < 1e-9
    assert info["f"] < f0 - 1.0
    Uo, xyo, fo, ito, infoo = so.disentangle(*oracle_args(silicon), silicon["U"], silicon["frozen"], max_iter=30)
    assert info["iterations"] == ito and info["n_fg"] == infoo["n_fg"]
    assert abs(info["f"] - fo) < 1e-8
    assert np.abs(np.conj(U.transpose(0, 2, 1)) @ U - np.eye(8)).max() < 1e-11
    # frozen states stay inside the span of U:  sum_n |U[m, n]|^2 == 1 for frozen m
    w2 = (np.abs(U) ** 2).sum(axis=2)
    assert np.abs(w2[silicon["frozen"]] - 1).max() < 1e-10


@pytest.mark.parametrize("lat,grid,nb,nw,nfroz", [
    ("fcc", (4, 4, 4), 4, 4, 0),       # the Si2_valence shape of benchmark/bench_maxloc.jl
    ("bcc", (3, 3, 3), 11, 11, 0),     # 12 neighbours
    ("cubic", (3, 2, 2), 7, 7, 0),     # odd sizes
    ("fcc", (4, 4, 4), 12, 8, 4),      # the Si2 shape of benchmark/bench_disentangle.jl: (X, Y) chain, frozen bands
    ("tetragonal", (5, 5, 1), 16, 5, 3),   # self-neighbours with G != 0, few Wannier functions, 16 bands
    ("cubic", (2, 2, 2), 9, 9, 9),     # every band frozen at every k (the n_froz == n_wann branch, disentangle.jl:390)
    ("cubic", (6, 6, 6), 8, 8, 0),     # 216 k-points: more CTAs than SMs, the global sums formed once (two-level)
    ("cubic", (10, 8, 8), 6, 6, 0),    # 640 k-points: two k-points per CTA by the plan itself
    ("fcc", (5, 5, 5), 18, 9, 5),      # the band counts of BASELINE config 2 (Cu-like), n_bands > 16
])
def test_tiny_trial_kernel_matches_general_path(w, lat, grid, nb, nw, nfroz):
    """tiny.cu (one cooperative launch per trial point, small n_bands and little arithmetic) against the chain of general kernels
    (WB200_FLAG_NO_FUSED_TRIAL): the same L-BFGS / HagerZhang run must take the same iterations and evaluations and
    arrive at the same point; with and without the centre penalty; each also checked against the oracle's spread."""
    st, M, U = w.synthetic.make_model_arrays(lat, grid, nb, nw)
    nk = len(U)
    rng = np.random.default_rng(11)
    r0 = rng.standard_normal((nw, 3))
    import os
    for penalty in (None, (r0, 0.3)):
        out = []
        # fused kernel with one k-point per CTA, the general kernels, the fused kernel with (at least) three k-points per CTA
        for flags, kper in ((w.lib.FLAG_DEFAULT, None), (w.lib.FLAG_NO_FUSED_TRIAL, None), (w.lib.FLAG_DEFAULT, "3")):
            os.environ.pop("WB200_TINY_KPER", None)
            if kper:
                os.environ["WB200_TINY_KPER"] = kper
            ctx = make_ctx(w, st, M, nw, flags=flags)
            if penalty:
                ctx.set_penalty(penalty[0], penalty[1])
            if nfroz:
                frozen = w.synthetic.make_frozen(nk, nb, nfroz)
                ctx.set_frozen(frozen)
                X, Y = so.U_to_X_Y(U, frozen)
                x = np.ascontiguousarray(so.X_Y_to_XY(X, Y))
                f, it, nfg = ctx.disentangle(x, 1e-7, 1e-5, 6, 3)
                Xd, Yd = so.XY_to_X_Y(x, nb, nw)
                Ud = so.X_Y_to_U(Xd, Yd)
                assert np.abs((np.abs(Ud) ** 2).sum(axis=2)[frozen] - 1).max() < 1e-10
            else:
                x = w.api.to_abi_U(U).copy()
                f, it, nfg = ctx.max_localize(x, 1e-7, 1e-5, 6, 3)
                Ud = w.api.from_abi_U(x)
            launches = ctx.launch_count
            ctx.close()
            os.environ.pop("WB200_TINY_KPER", None)
            sp = so.omega(M, st["kpb_k"], st["bvec"], st["wb"], Ud)
            fo = sp["Omega"] + (penalty[1] * ((sp["r"] - penalty[0]) ** 2).sum() if penalty else 0.0)
            assert abs(fo - f) < 1e-9 * max(1.0, abs(f))
            assert np.abs(np.conj(Ud.transpose(0, 2, 1)) @ Ud - np.eye(nw)).max() < 1e-11
            out.append((f, it, nfg, x, launches))
        (f1, it1, n1, x1, l1), (f2, it2, n2, x2, l2), (f3, it3, n3, x3, l3) = out
        assert (it1, n1) == (it2, n2) == (it3, n3)
        assert abs(f1 - f2) < 1e-9 * max(1.0, abs(f2))
        assert np.abs(x1 - x2).max() < 1e-7
        assert l1 < l2 / 2               # the fused path really ran: one launch per trial point instead of ~13
        # one or several k-points per CTA: the same arithmetic in the same order (a context whose three slots do not
        # fit into shared memory falls back to the general kernels)
        assert l3 in (l1, l2)
        ref = (f1, x1) if l3 == l1 else (f2, x2)
        assert f3 == ref[0] and np.array_equal(x3, ref[1])


# ---------------------------------------------------------------------------------------------
# synthetic shapes: both rotation kernels against the oracle, padding and edge cases
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lat,grid,nb,nw", [
    ("bcc", (2, 2, 2), 128, 128),     # north-star tile shape (DMMA 128 x 32 CTA tiles)
    ("bcc", (2, 2, 1), 128, 96),      # rectangular, nw not a multiple of... 32 (it is) but < nb
    ("cubic", (2, 2, 2), 100, 70),    # padded rows / columns / k-chunks
    ("fcc", (2, 2, 2), 64, 64),
    ("fcc", (2, 1, 2), 40, 33),
    ("tetragonal", (3, 3, 1), 256, 256),   # self-neighbours with G != 0 (config 5 shape)
    ("cubic", (3, 1, 2), 200, 130),
    ("cubic", (2, 1, 1), 33, 33),     # degenerate: every neighbour is a periodic image, Omega ~ 0
    ("fcc", (3, 3, 3), 18, 9),        # config 2 shape (CUDA-core kernel)
    ("bcc", (3, 3, 3), 32, 32),       # config 3 shape
    ("cubic", (1, 1, 1), 5, 3),       # single k-point, all neighbours are the point itself
    ("cubic", (2, 2, 2), 260, 20),    # nb > 256: CUDA-core kernel only
    # >= 1184 pairs with nb <= 32: the persistent ring kernel (rotate_small_ring_kernel); pair counts
    # that do not divide by the CTA / warp-pair split, padded rows and columns, rectangular U
    ("fcc", (6, 6, 6), 18, 9),
    ("bcc", (5, 5, 5), 32, 32),
    ("cubic", (8, 8, 5), 7, 7),
    ("bcc", (5, 4, 5), 31, 17),
])
def test_synthetic_both_kernels(w, lat, grid, nb, nw):
    st, M, U = w.synthetic.make_model_arrays(lat, grid, nb, nw)
    nk = len(U)
    f_ref, G_ref = so.fg_maxloc(M, st["kpb_k"], st["bvec"], st["wb"], U)
    sp_ref = so.omega(M, st["kpb_k"], st["bvec"], st["wb"], U)
    kernels = set()
    for flags in (w.lib.FLAG_DEFAULT, w.lib.FLAG_NO_TENSOR):
        ctx = make_ctx(w, st, M, nw, flags=flags)
        kernels.add(ctx.rotation_kernel)
        G = np.empty((nk, nw, nb), dtype=np.complex128)
        f = ctx.fg(w.api.to_abi_U(U), G=G)
        assert abs(f - f_ref) < RTOL * max(abs(f_ref), 1.0), ctx.rotation_kernel
        assert np.abs(w.api.from_abi_U(G) - G_ref).max() < RTOL * max(np.abs(G_ref).max(), 1.0), ctx.rotation_kernel
        check_spread(ctx.spread(w.api.to_abi_U(U)), sp_ref)
        N, MU = ctx.rotate_overlaps(w.api.to_abi_U(U), want_N=True, want_MU=True)
        MU_ref, N_ref = so.compute_MU_UtMU(M, st["kpb_k"], U)
        assert rel(MU.transpose(0, 1, 3, 2), MU_ref) < 1e-12
        assert rel(N.transpose(0, 1, 3, 2), N_ref) < 1e-12
        # determinism: bitwise identical on a second call
        G2 = np.empty_like(G)
        f2 = ctx.fg(w.api.to_abi_U(U), G=G2)
        assert f2 == f and np.array_equal(G, G2)
        ctx.close()
    assert kernels == ({"cuda-core-fp64", "dmma-fp64"} if nb <= 256 else {"cuda-core-fp64"})


@pytest.mark.parametrize("name", ["cfg2_cu_shape", "cfg3_nw32", "cfg5_2d_nw256"])
def test_baseline_configs_at_their_exact_sizes(w, name):
    """BASELINE.json configs 2, 3 and 5 at their full k-grids (9^3 / 12^3 / 10 x 10 x 1), not only as tile
    shapes: Omega, the five-way split, r, omega_n and G against the oracle at 1e-10; config 2 through the
    (X, Y) chain with the five lowest bands frozen (its path in the reference, disentangle.jl:586-614);
    config 3 additionally two iterations of the device max_localize against the oracle's Optim restatement."""
    lat, grid, nb, nw = w.synthetic.CONFIGS[name]
    st, M, U = w.synthetic.make_model_arrays(lat, grid, nb, nw)
    nk = len(U)
    args = (M, st["kpb_k"], st["bvec"], st["wb"])
    ctx = make_ctx(w, st, M, nw)
    assert ctx.rotation_kernel == "dmma-fp64"
    f_ref, G_ref = so.fg_maxloc(*args, U)
    G = np.empty((nk, nw, nb), dtype=np.complex128)
    f = ctx.fg(w.api.to_abi_U(U), G=G)
    assert abs(f - f_ref) < RTOL * abs(f_ref)
    assert rel(w.api.from_abi_U(G), G_ref) < RTOL
    check_spread(ctx.spread(w.api.to_abi_U(U)), so.omega(*args, U))
    if name == "cfg2_cu_shape":
        frozen = w.synthetic.make_frozen(nk, nb, 5)
        ctx.set_frozen(frozen)
        X, Y = so.U_to_X_Y(U, frozen)
        XY = np.ascontiguousarray(so.X_Y_to_XY(X, Y))
        Gxy = np.empty_like(XY)
        fx = ctx.fg_xy(XY, G=Gxy)
        fx_ref, Gxy_ref = so.fg_disentangle(*args, XY, nb, nw, frozen)
        assert abs(fx - fx_ref) < RTOL * abs(fx_ref)
        assert rel(Gxy, Gxy_ref) < RTOL
    if name == "cfg3_nw32":
        Ud = np.ascontiguousarray(w.api.to_abi_U(U))
        fd, itd, nfgd = ctx.max_localize(Ud, max_iter=2)
        Uo, fo, ito, infoo = so.max_localize(*args, U, max_iter=2)
        assert (itd, nfgd) == (ito, infoo["n_fg"])
        assert abs(fd - fo) < 1e-9 * abs(fo)
        assert np.abs(w.api.from_abi_U(Ud) - Uo).max() < 1e-8
    ctx.close()


def test_unitary_shortcut_vs_exact_split(w):
    # fg (one GEMM, ||N||_F = ||MU||_F for square unitary U) and spread (full N) agree on all
    # five scalars when nb == nw
    import torch
    st, M, U = w.synthetic.make_model_arrays("bcc", (2, 2, 2), 64, 64)
    ctx = make_ctx(w, st, M, 64)
    dU = torch.from_numpy(w.api.to_abi_U(U)).cuda()
    res = torch.zeros(ctx.nres, dtype=torch.float64, device="cuda")
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.fg_dev(dU.data_ptr(), res.data_ptr(), None, exact_split=False)
    a = res.cpu().numpy().copy()
    ctx.fg_dev(dU.data_ptr(), res.data_ptr(), None, exact_split=True)
    b = res.cpu().numpy()
    assert np.abs(a[:5] - b[:5]).max() < 1e-11 * abs(b[0])
    sp = so.omega(M, st["kpb_k"], st["bvec"], st["wb"], U)
    assert abs(a[1] - sp["OmegaI"]) < 1e-10 * sp["Omega"] and abs(a[2] - sp["OmegaOD"]) < 1e-10 * sp["Omega"]
    ctx.close()
    # both epilogues of the shortcut (the gauge-invariant norm ||M_kb||_F computed once, and ||M U||_F per
    # evaluation) give the same split; the small ring kernel likewise (nb = nw = 32, enough pairs for it)
    import os
    for lat, grid, n in [("bcc", (2, 2, 2), 64), ("cubic", (6, 6, 6), 32)]:
        st, M, U = w.synthetic.make_model_arrays(lat, grid, n, n)
        sp = so.omega(M, st["kpb_k"], st["bvec"], st["wb"], U)
        for env in ("1", "0"):
            os.environ["WB200_CONST_NORM"] = env
            try:
                ctx = make_ctx(w, st, M, n)
            finally:
                os.environ.pop("WB200_CONST_NORM", None)
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)
            dU = torch.from_numpy(w.api.to_abi_U(U)).cuda()
            ctx.fg_dev(dU.data_ptr(), res[:ctx.nres].data_ptr(), None, exact_split=False)
            a = res.cpu().numpy()
            for i, key in enumerate(["Omega", "OmegaI", "OmegaOD", "OmegaD", "OmegaTilde"]):
                assert abs(a[i] - sp[key]) < 1e-10 * abs(sp["Omega"]), (n, env, key)
            ctx.close()


def test_rectangular_shortcut_reports_no_split(w):
    # exact_split = 0 equates ||N||_F with ||M U||_F, which only holds for square gauges: with nb > nw the spread
    # and f are right and the OmegaI / OmegaOD / OmegaD / OmegaTilde slots are NaN, not plausible wrong numbers
    import torch
    st, M, U = w.synthetic.make_model_arrays("cubic", (3, 2, 2), 72, 40)
    ctx = make_ctx(w, st, M, 40)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    dU = torch.from_numpy(w.api.to_abi_U(U)).cuda()
    res = torch.zeros(ctx.nres, dtype=torch.float64, device="cuda")
    ctx.fg_dev(dU.data_ptr(), res.data_ptr(), None, exact_split=False)
    a = res.cpu().numpy().copy()
    sp = so.omega(M, st["kpb_k"], st["bvec"], st["wb"], U)
    assert abs(a[0] - sp["Omega"]) < 1e-10 * abs(sp["Omega"]) and abs(a[5] - sp["Omega"]) < 1e-10 * abs(sp["Omega"])
    assert np.isnan(a[1:5]).all()
    ctx.fg_dev(dU.data_ptr(), res.data_ptr(), None, exact_split=True)
    b = res.cpu().numpy()
    for i, key in enumerate(["Omega", "OmegaI", "OmegaOD", "OmegaD", "OmegaTilde"]):
        assert abs(b[i] - sp[key]) < 1e-10 * abs(sp["Omega"]), key
    ctx.close()


def test_k_slab_sharding_on_one_gpu(w):
    """Multi-GPU path emulated on one device: two slab contexts, phase1 each, sums added (what the
    NCCL all-reduce does), phase2 each; equals the single-slab result."""
    import torch
    st, M, U = w.synthetic.make_model_arrays("bcc", (4, 2, 2), 64, 48)
    nk, nb, nw = len(U), 64, 48
    f_ref, G_ref = so.fg_maxloc(M, st["kpb_k"], st["bvec"], st["wb"], U)
    dU = torch.from_numpy(w.api.to_abi_U(U)).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    ctxs, sums = [], []
    for (k0, k1) in ((0, 6), (6, 16)):
        # anisotropic grid: 18 neighbours in 3 shells, two of them with ~0 weight
        c = w.Context(nb, nw, nk, st["kpb_k"].shape[1], st["kpb_k"][k0:k1], st["bvec"][k0:k1], st["wb"],
                      w.api.to_abi_M(M[k0:k1]), k_begin=k0, nk=k1 - k0)
        c.set_stream(stream)
        s = torch.zeros(c.nsums, dtype=torch.float64, device="cuda")
        c.fg_phase1_dev(dU.data_ptr(), s.data_ptr())
        ctxs.append((c, k0, k1))
        sums.append(s)
    total = sums[0] + sums[1]
    G = torch.zeros((nk, nw, nb), dtype=torch.complex128, device="cuda")
    res = torch.zeros(ctxs[0][0].nres, dtype=torch.float64, device="cuda")
    for c, k0, k1 in ctxs:
        c.fg_phase2_dev(total.data_ptr(), res.data_ptr(), G[k0:k1].data_ptr())
    torch.cuda.synchronize()
    assert abs(res[5].item() - f_ref) < RTOL * abs(f_ref)
    assert rel(G.cpu().numpy().transpose(0, 2, 1), G_ref) < RTOL
    for c, _, _ in ctxs:
        c.close()


def test_error_behaviour(w, valence):
    st = fixture_stencil(valence)
    bad = st["kpb_k"].copy()
    bad[3, 2] = 64
    with pytest.raises(w.WB200Error) as e:
        w.Context(4, 4, 64, 8, bad, st["bvec"], st["wb"], w.api.to_abi_M(valence["M"]))
    assert e.value.status == -1
    with pytest.raises(w.WB200Error):
        w.Context(4, 5, 64, 8, st["kpb_k"], st["bvec"], st["wb"], w.api.to_abi_M(valence["M"]))
    ctx = make_ctx(w, st, valence["M"], 4)
    with pytest.raises(ValueError):
        ctx.fg(np.zeros((3, 4, 4), dtype=np.complex128))
    # a singular gauge gives |N_nn| ~ 0: reported through min_abs_Nnn instead of silent NaN
    U = valence["U"].copy()
    U[:, :, 0] = 0
    mn = ctx.spread(w.api.to_abi_U(U))[3]
    assert mn < 1e-10
    ctx.close()


def test_caching_allocator_and_one_shot_contexts(w, valence):
    """mempool.cu: a context per call (what the reference's one-shot API does with its Cache, spread.jl:378, 497) must
    not cost driver allocations the second time, must give the same answers from recycled blocks, and the kept blocks
    must go back to the driver on request."""
    import os
    if os.environ.get("WB200_POOL") == "0":
        pytest.skip("caching allocator disabled by WB200_POOL=0")
    st = fixture_stencil(valence)
    Uc = w.api.to_abi_U(valence["U"])
    ref = None
    w.lib.trim_pools()
    assert w.lib.pool_stats(0)[0] == 0 and w.lib.pool_stats(-1)[0] == 0
    for rep in range(4):
        _, h0, m0 = w.lib.pool_stats(0)
        ctx = make_ctx(w, st, valence["M"], 4)
        G = np.empty_like(Uc)
        f = ctx.fg(Uc, G=G)
        out5 = ctx.spread(Uc)[0]
        ctx.close()
        kept, h1, m1 = w.lib.pool_stats(0)
        assert kept > 0 and w.lib.pool_stats(-1)[0] > 0
        if rep == 0:
            ref = (f, G.copy(), np.array(out5))
            assert m1 > m0                       # first context: everything from the driver
        else:
            assert m1 == m0 and h1 > h0          # later contexts: nothing from the driver
            assert f == ref[0] and np.array_equal(G, ref[1]) and np.array_equal(np.array(out5), ref[2])   # bitwise
    kept_before = w.lib.pool_stats(0)[0]
    for rep in range(20):                        # no growth: the same blocks go round
        make_ctx(w, st, valence["M"], 4).close()
    assert w.lib.pool_stats(0)[0] == kept_before
    w.lib.trim_pools()
    assert w.lib.pool_stats(0)[0] == 0 and w.lib.pool_stats(-1)[0] == 0
    ctx = make_ctx(w, st, valence["M"], 4)       # and the library still works from the driver again
    assert ctx.fg(Uc) == ref[0]
    ctx.close()
    # numpy-order overlaps transposed on the device (WB200_FLAG_M_ROW_MAJOR) = the caller transposing them
    nk, nn, nb, _ = valence["M"].shape
    ctx = w.Context(nb, 4, nk, nn, st["kpb_k"], st["bvec"], st["wb"], np.ascontiguousarray(valence["M"]),
                    flags=w.lib.FLAG_M_ROW_MAJOR)
    assert ctx.fg(Uc) == ref[0]
    ctx.close()


def test_multi_wave_determinism_and_parity(w):
    """864 CTAs of the DMMA kernel (about 6 waves on 148 SMs): results must be bitwise identical
    from call to call and match the oracle.  Guards the shared-memory ring of the rotation kernel
    (a missing cross-proxy fence once let a stage be overwritten under a slow warp, visible only
    with more CTAs than SMs)."""
    import torch
    st = w.synthetic.make_stencil(w.synthetic.LATTICES["bcc"](), (6, 6, 6))
    nk, nn = st["kpb_k"].shape
    nb = nw = 128
    dev = torch.device("cuda", 0)
    M = w.synthetic.make_overlaps_torch(st, nb, dev)
    dU = w.synthetic.make_gauge_torch(nk, nb, nw, dev)
    ctx = w.Context(nb, nw, nk, nn, st["kpb_k"], st["bvec"], st["wb"], M.data_ptr())
    assert ctx.rotation_kernel == "dmma-fp64"
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    res = torch.zeros(ctx.nres, dtype=torch.float64, device=dev)
    G = torch.zeros((nk, nw, nb), dtype=torch.complex128, device=dev)
    outs = []
    for _ in range(6):
        ctx.fg_dev(dU.data_ptr(), res.data_ptr(), G.data_ptr())
        torch.cuda.synchronize()
        outs.append((res.cpu().numpy().copy(), G.cpu().numpy().copy()))
    for r, g in outs[1:]:
        assert np.array_equal(r, outs[0][0]) and np.array_equal(g, outs[0][1])
    Mh = M.transpose(2, 3).cpu().numpy()
    Uh = dU.cpu().numpy().transpose(0, 2, 1)
    f_ref, G_ref = so.fg_maxloc(Mh, st["kpb_k"], st["bvec"], st["wb"], Uh)
    assert abs(outs[0][0][5] - f_ref) < RTOL * abs(f_ref)
    assert rel(outs[0][1].transpose(0, 2, 1), G_ref) < RTOL
    _, MU = ctx.rotate_overlaps(dU.cpu().numpy(), want_N=False, want_MU=True)
    MU_ref, _ = so.compute_MU_UtMU(Mh, st["kpb_k"], Uh)
    assert np.abs(MU.transpose(0, 1, 3, 2) - MU_ref).max() < 1e-13
    ctx.close()


@pytest.mark.parametrize("nb,nw", [(64, 48), (12, 8)])
def test_peer_gauge_reads_fused_into_phase1(w, nb, nw):
    """k-sharded run where each rank's gauge array holds ONLY its own rows: phase 1 must fetch the
    neighbour rows from the owning rank's array (peer pointers; two arrays on one GPU stand in for
    two GPUs, the kernels cannot tell).  Covers the fused read in the tensor path's pack kernel
    (nb = 64) and the halo-fetch kernel of the CUDA-core path (nb = 12)."""
    import torch
    st, M, U = w.synthetic.make_model_arrays("bcc", (4, 2, 2), nb, nw)
    nk, nn = st["kpb_k"].shape
    f_ref, G_ref = so.fg_maxloc(M, st["kpb_k"], st["bvec"], st["wb"], U)
    bounds = [0, 6, 16]
    Uabi = torch.from_numpy(w.api.to_abi_U(U)).cuda()
    arrays = []
    for r in range(2):
        a = torch.full_like(Uabi, complex(float("nan"), float("nan")))
        a[bounds[r]:bounds[r + 1]] = Uabi[bounds[r]:bounds[r + 1]]
        arrays.append(a)
    stream = torch.cuda.current_stream().cuda_stream
    ctxs, sums = [], []
    for r in range(2):
        k0, k1 = bounds[r], bounds[r + 1]
        c = w.Context(nb, nw, nk, nn, st["kpb_k"][k0:k1], st["bvec"][k0:k1], st["wb"],
                      w.api.to_abi_M(M[k0:k1]), k_begin=k0, nk=k1 - k0)
        c.set_stream(stream)
        c.set_peer_gauges([a.data_ptr() for a in arrays], bounds)
        s = torch.zeros(c.nsums, dtype=torch.float64, device="cuda")
        c.fg_phase1_dev(arrays[r].data_ptr(), s.data_ptr())
        ctxs.append(c)
        sums.append(s)
    total = sums[0] + sums[1]
    G = torch.zeros((nk, nw, nb), dtype=torch.complex128, device="cuda")
    res = torch.zeros(ctxs[0].nres, dtype=torch.float64, device="cuda")
    for r, c in enumerate(ctxs):
        c.fg_phase2_dev(total.data_ptr(), res.data_ptr(), G[bounds[r]:bounds[r + 1]].data_ptr())
    torch.cuda.synchronize()
    assert abs(res[5].item() - f_ref) < RTOL * abs(f_ref)
    assert rel(G.cpu().numpy().transpose(0, 2, 1), G_ref) < RTOL
    for c in ctxs:
        c.set_peer_gauges(None, None)
        c.close()


def test_full_size_north_star_properties(w):
    """BASELINE config 4 at full size (n_wann = 128, 16^3 k, 12 b: 49 152 pairs, 13 GB of overlaps).
    The oracle cannot run this in seconds, so parity is checked through size-independent
    properties plus an oracle spot check on a few k-points."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < 90e9:
        pytest.skip("needs ~80 GB of free HBM")
    lat, grid, nb, nw = w.synthetic.CONFIGS["cfg4_nw128"]
    st = w.synthetic.make_stencil(w.synthetic.LATTICES[lat](), grid)
    nk, nn = st["kpb_k"].shape
    dev = torch.device("cuda", 0)
    M = w.synthetic.make_overlaps_torch(st, nb, dev)
    dU = w.synthetic.make_gauge_torch(nk, nb, nw, dev)
    spot = np.array([0, 1377, 4095])
    M_spot = M[torch.as_tensor(spot, device=dev)].transpose(2, 3).cpu().numpy()      # [3, nn, m, l]
    ctx = w.Context(nb, nw, nk, nn, st["kpb_k"], st["bvec"], st["wb"], M.data_ptr())
    del M
    torch.cuda.empty_cache()
    assert ctx.rotation_kernel == "dmma-fp64"
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    res = torch.zeros(ctx.nres, dtype=torch.float64, device=dev)
    G = torch.zeros((nk, nw, nb), dtype=torch.complex128, device=dev)

    def fg(U, want_G=True, exact=False):
        ctx.fg_dev(U.data_ptr(), res.data_ptr(), G.data_ptr() if want_G else None, exact_split=exact)
        torch.cuda.synchronize()
        return res.cpu().numpy().copy()

    a = fg(dU)
    G0 = G.clone()
    Om, OI, OOD, OD, Ot = a[:5]
    omega_n, r = a[8:8 + nw], a[8 + nw:8 + 4 * nw].reshape(nw, 3)
    # (1) the decomposition closes and is consistent
    assert abs(omega_n.sum() - Om) < 1e-10 * Om
    assert abs(OI + OOD + OD - Om) < 1e-10 * Om and abs(Ot - (Om - OI)) < 1e-10 * Om
    assert OI > 0 and OOD > 0
    # (2) one-GEMM shortcut (||N||_F = ||MU||_F for unitary U) == full N (second GEMM)
    b = fg(dU, want_G=False, exact=True)
    assert np.abs(a[:5] - b[:5]).max() < 1e-10 * Om
    # (3) gauge covariance under a global unitary W: Omega_I invariant; under diagonal phases D
    #     Omega, omega_n, r invariant and G(U D) = G(U) D
    g = torch.Generator(device="cpu").manual_seed(7)
    Wm = torch.linalg.qr(torch.view_as_complex(torch.randn((nw, nw, 2), generator=g, dtype=torch.float64)))[0].to(dev)
    # abi image of U_k is [n][m] = U_k^T, so (U W)^T = W^T U^T
    c = fg(torch.matmul(Wm.T.contiguous(), dU).contiguous(), want_G=False, exact=True)
    assert abs(c[1] - b[1]) < 1e-10 * Om
    ph = torch.exp(2j * np.pi * torch.rand(nw, generator=g, dtype=torch.float64)).to(dev)
    d = fg((dU * ph[None, :, None]).contiguous())
    assert abs(d[0] - Om) < 1e-10 * Om and np.abs(d[8:8 + 4 * nw] - a[8:8 + 4 * nw]).max() < 1e-9
    assert (G - G0 * ph[None, :, None]).abs().max().item() < 1e-10 * G0.abs().max().item()
    # (4) the gradient is the derivative: central difference along a random direction
    Z = torch.view_as_complex(torch.randn((nk, nw, nb, 2), generator=torch.Generator(device="cpu").manual_seed(8),
                                          dtype=torch.float64)).to(dev)
    h = 1e-6
    fp = fg(dU + h * Z, want_G=False)[5]
    fm = fg(dU - h * Z, want_G=False)[5]
    an = torch.sum(G0.real * Z.real + G0.imag * Z.imag).item()
    diag = dict(f0=Om, fp=fp, fm=fm, an=an, fp_again=fg(dU + h * Z, want_G=False)[5],
                f0_again=fg(dU, want_G=False)[5], f0_copy=fg(dU + 0.0 * Z, want_G=False)[5])
    assert abs((fp - fm) / (2 * h) - an) < 1e-6 * abs(an), diag
    # (5) oracle spot check on three k-points: MU, diag N and G_k (with the global r of the GPU run)
    fg(dU)
    Uh = dU.cpu().numpy().transpose(0, 2, 1)
    kpb = st["kpb_k"][spot]
    MU_ref, N_ref = so.compute_MU_UtMU_gathered(M_spot, Uh[spot], Uh[kpb])
    G_ref = so.grad_from_MU_N_r(MU_ref, N_ref, st["bvec"][spot], st["wb"], r, nk)
    Gh = G0[torch.as_tensor(spot, device=dev)].cpu().numpy().transpose(0, 2, 1)
    assert np.abs(Gh - G_ref).max() < 1e-10 * np.abs(G_ref).max()
    ctx.close()


def test_next_rows_omega_local_position_berry_opt_rotate(w, valence, silicon):
    """SURVEY 8f 'next' rows: by-products of the same rotation, each against the oracle."""
    from test_oracle_golden import W_OPTROT_REF
    for d, nw in ((valence, 4), (silicon, 8)):
        st = w.KspaceStencil(d["recip"], d["kpoints"], d["wb"], d["kpb_k"], d["kpb_G"])
        U = d["U"]
        args = oracle_args(d)
        assert rel(w.omega_local(st, d["M"], U), so.omega_local(d["M"], d["kpb_k"], d["wb"], U)) < 1e-12
        assert rel(w.position_op(st, d["M"], U), so.position_op(*args, U)) < 1e-12
        for imlog in (True, False):
            for herm in (True, False):
                A = w.compute_berry_connection_kspace(st, d["M"], U, imlog_diag=imlog, force_hermiticity=herm)
                assert rel(A, so.berry_connection_kspace(*args, U, imlog, herm)) < 1e-12
    # opt_rotate (test/localization/opt_rotate.jl): gradient vs oracle, optimum vs the reference's Wmin
    st = w.KspaceStencil(valence["recip"], valence["kpoints"], valence["wb"], valence["kpb_k"], valence["kpb_G"])
    args = oracle_args(valence)
    model = w.Model(valence["lattice"], st, valence["M"], valence["U"])
    f, g = w.get_fg_rotate(model)
    N = so.transform_gauge(valence["M"], valence["kpb_k"], valence["U"])
    for W0 in (np.eye(4, dtype=complex), valence["W1"]):
        G = np.zeros((4, 4), dtype=complex)
        g(G, W0)
        f_ref, G_ref = so.fg_rotate(N, *args[1:], W0)
        assert abs(f(W0) - f_ref) < RTOL * abs(f_ref) and rel(G, G_ref) < RTOL
    f.cache.close()
    model = w.Model(valence["lattice"], st, valence["M"], valence["U_ptg"])
    W, info = w.opt_rotate(model, f_tol=1e-12, g_tol=1e-8, max_iter=300, return_info=True)
    f_ref = so.omega(*args, valence["U_ptg"] @ so.stiefel_retract(W_OPTROT_REF))["Omega"]
    assert abs(info["f"] - f_ref) < 1e-6 and info["f"] <= f_ref + 1e-9
    assert abs(so.omega(*args, valence["U_ptg"] @ W)["Omega"] - info["f"]) < 1e-10
    with pytest.raises(ValueError):
        w.opt_rotate(w.Model(silicon["lattice"], w.KspaceStencil(silicon["recip"], silicon["kpoints"], silicon["wb"],
                                                                silicon["kpb_k"], silicon["kpb_G"]), silicon["M"], silicon["U"]))


def test_read_w90_to_max_localize(w, valence):
    """The whole drop-in flow from Wannier90 files: parse, Lowdin on the GPU, spread, max_localize
    (test/localization/max_localize.jl builds its model with read_w90 on this same fixture)."""
    import os
    prefix = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "w90_valence", "silicon")
    model = w.read_w90(prefix)
    assert np.abs(model.gauges - valence["U"]).max() < 1e-12            # device Lowdin == SVD Lowdin
    sp = w.omega(model)
    ref = so.omega(*oracle_args(valence), valence["U"])
    assert abs(sp.Omega - ref["Omega"]) < RTOL * ref["Omega"]
    assert abs(sp.OmegaI - valence["ref_maxloc"][1]) < 1e-10             # reference constant
    U = w.max_localize(model, f_tol=1e-13, g_tol=1e-8, max_iter=500)
    assert abs(so.omega(*oracle_args(valence), U)["Omega"] - valence["ref_maxloc"][0]) < 1e-8


def test_opt_rotate_singular_guard(w, valence):
    # opt_rotate.jl:60-62: g! raises "N^kb too small" when a diagonal element vanishes
    st = w.KspaceStencil(valence["recip"], valence["kpoints"], valence["wb"], valence["kpb_k"], valence["kpb_G"])
    f, g = w.get_fg_rotate(w.Model(valence["lattice"], st, valence["M"], valence["U"]))
    W = np.eye(4, dtype=complex)
    W[:, 0] = 0
    with pytest.raises(w.WB200Error) as e:
        g(np.zeros((4, 4), dtype=complex), W)
    assert e.value.status == -5
    f.cache.close()


@pytest.mark.parametrize("lat,grid,n", [("bcc", (2, 2, 2), 128), ("cubic", (3, 3, 1), 72), ("tetragonal", (3, 3, 1), 256)])
def test_stiefel_ops_on_the_tensor_pipe(w, lat, grid, n):
    """n x n Stiefel GEMMs (X^H X, X T, X^H G, X S) routed through the DMMA rotation kernel when the
    matrices have the ctx's own shape (max_localize at large n_wann): projection, retraction and a few
    L-BFGS iterations against the oracle."""
    st, M, U = w.synthetic.make_model_arrays(lat, grid, n, n, eps=0.02)
    nk = len(U)
    ctx = make_ctx(w, st, M, n)
    assert ctx.rotation_kernel == "dmma-fp64"
    rng = np.random.default_rng(11)
    G = rng.standard_normal(U.shape) + 1j * rng.standard_normal(U.shape)
    Gc = w.api.to_abi_U(G)
    ctx.project_tangent(w.api.to_abi_U(U), Gc, n, n)
    assert rel(w.api.from_abi_U(Gc), so.stiefel_project_tangent(G, U)) < 1e-12
    Xp = U + 0.05 * G / np.sqrt(n)
    Xc = w.api.to_abi_U(Xp)
    ctx.retract(Xc, n, n)
    assert np.abs(w.api.from_abi_U(Xc) - so.stiefel_retract(Xp)).max() < 1e-12
    # a short optimisation: same iterates as the oracle's driver
    Uc = w.api.to_abi_U(U)
    f, it, nfg = ctx.max_localize(Uc, f_tol=1e-12, g_tol=1e-9, max_iter=3)
    Uo, fo, ito, _ = so.max_localize(M, st["kpb_k"], st["bvec"], st["wb"], U, f_tol=1e-12, g_tol=1e-9, max_iter=3)
    assert it == ito and abs(f - fo) < 1e-9 * max(abs(fo), 1.0)   # degenerate stencils give Omega ~ 0
    assert np.abs(w.api.from_abi_U(Uc) - Uo).max() < 1e-8
    ctx.close()


# ---------------------------------------------------------------------------------------------
# MagModel co-optimisation (src/localization/coopt.jl)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lat,grid,nb,nw,nfroz", [
    ("fcc", (2, 2, 2), 6, 4, 2),        # fused small-shape kernels
    ("cubic", (5, 4, 5), 18, 9, 5),     # > 64 k-points: two-stage reduction of diag(Mw)
    ("fcc", (2, 2, 2), 72, 40, 9),      # batched GEMMs on the FP64 tensor pipe
])
def test_magmodel_coupling_and_disentangle(w, lat, grid, nb, nw, nfroz):
    st, Mu, Md, Mud, Uu, Ud = w.synthetic.make_mag_arrays(lat, grid, nb, nw)
    nk = len(Uu)
    frozen = w.synthetic.make_frozen(nk, nb, nfroz)
    stencil = w.KspaceStencil(st["recip"], st["kpoints"], st["wb"], st["kpb_k"], st["kpb_G"])
    model = w.MagModel(w.Model(st["lattice"], stencil, Mu, Uu, frozen_bands=frozen),
                       w.Model(st["lattice"], stencil, Md, Ud, frozen_bands=frozen), Mud)
    cache = w.MagCache(model)
    up = (Mu, st["kpb_k"], st["bvec"], st["wb"])
    dn = (Md, st["kpb_k"], st["bvec"], st["wb"])
    # overlap_updn, omega_updn, omega_updn_grad (coopt.jl:114-235)
    Mw = w.overlap_updn(model, cache=cache)
    assert rel(Mw, so.overlap_updn(Mud, Uu, Ud)) < RTOL
    assert abs(w.omega_updn(model, cache=cache) - so.omega_updn(Mud, Uu, Ud)) < RTOL * nw
    Gu, Gd = w.omega_updn_grad(model, Uu, Ud, cache=cache)
    Gu_ref, Gd_ref = so.omega_updn_grad(Mud, Uu, Ud)
    assert rel(Gu, Gu_ref) < RTOL and rel(Gd, Gd_ref) < RTOL
    # omega(model::MagModel, lambda) (coopt.jl:66-77)
    for lam in (1.0, 0.35):
        sp = w.omega(model, lam, cache=cache)
        ref = so.omega_mag(up, dn, Mud, Uu, Ud, lam)
        assert abs(sp.Omegat - ref["Omegat"]) < RTOL * abs(ref["Omegat"])
        assert abs(sp.up.Omega - ref["up"]["Omega"]) < RTOL * abs(ref["Omegat"])
        assert abs(sp.Omegaupdn - ref["Omegaupdn"]) < RTOL * nw
    # f / g! over XY = [XYup; XYdn] (coopt.jl:252-321), lambda = 0 included (the coupling is skipped)
    XY = np.concatenate([so.X_Y_to_XY(*so.U_to_X_Y(Uu, frozen)), so.X_Y_to_XY(*so.U_to_X_Y(Ud, frozen))], axis=1)
    ni = nw * nw + nb * nw
    XYd = w.api.mag_XY(model)                 # device U_to_X_Y of both spins: same U = Y X per spin
    for a, Uref in ((XYd[:, :ni], Uu), (XYd[:, ni:], Ud)):
        Xa, Ya = so.XY_to_X_Y(np.ascontiguousarray(a), nb, nw)
        Af = np.stack([so.orthonorm_freeze(Uref[k], frozen[k]) for k in range(nk)])
        assert np.abs(so.X_Y_to_U(Xa, Ya) - Af).max() < 1e-10
    for lam in (1.0, 0.0, 2.5):
        f, g = w.get_fg_disentangle(model, lam, cache=cache)
        G = np.empty_like(XY)
        g(G, XY)
        f_ref, G_ref = so.fg_mag_disentangle(up, dn, Mud, XY, nb, nw, frozen, frozen, lam)
        assert abs(f(XY) - f_ref) < RTOL * abs(f_ref), lam
        assert rel(G, G_ref) < RTOL, lam
    # the device driver against the oracle's driver (same Optim restatement): same path, same spread
    XYrun = XY.copy()
    f1, it1, nfg1 = cache.mag.disentangle(XYrun, 1.0, max_iter=6)
    Uo, Do, xyo, fo, ito, infoo = so.disentangle_mag(up, dn, Mud, Uu, Ud, frozen, frozen, lam=1.0, max_iter=6)
    assert (it1, nfg1) == (ito, infoo["n_fg"])
    assert np.abs(XYrun - xyo).max() < 1e-7          # same starting point: the very same iterates
    # through the public API the start comes from the device U_to_X_Y (another basis of the free range: the
    # trajectory is the same up to that unitary mixing, the gauges U = Y X are the same)
    Uup, Udn, info = w.disentangle(model, 1.0, max_iter=6, cache=cache, return_info=True)
    assert (info["iterations"], info["n_fg"]) == (ito, infoo["n_fg"])
    assert abs(info["f"] - fo) < 1e-9 * abs(fo)
    assert np.abs(Uup - Uo).max() < 1e-7 and np.abs(Udn - Do).max() < 1e-7
    ref = so.omega_mag(up, dn, Mud, Uup, Udn, 1.0)
    assert abs(ref["Omegat"] - info["f"]) < 1e-9 * abs(info["f"])
    for U in (Uup, Udn):
        assert np.abs(np.conj(U.transpose(0, 2, 1)) @ U - np.eye(nw)).max() < 1e-11
        assert np.abs((np.abs(U) ** 2).sum(axis=2)[frozen] - 1).max() < 1e-10
    cache.close()
    # error behaviour: mismatched spin channels (coopt.jl:355-357)
    with pytest.raises(ValueError):
        w.MagCache(w.MagModel(model.up, w.Model(st["lattice"], stencil, Md, Ud[:, :, :nw - 1], frozen_bands=frozen), Mud))


# ---------------------------------------------------------------------------------------------
# setup step on the device: orthonorm_freeze / U_to_X_Y, disentangle.jl:223-281, 369-402
# ---------------------------------------------------------------------------------------------
def _check_u_to_xy(w, U, frozen, ctx=None):
    nk, nb, nw = U.shape
    X, Y = w.U_to_X_Y(U, frozen, cache=ctx)
    Af = np.stack([so.orthonorm_freeze(U[k], frozen[k]) for k in range(nk)])
    assert np.abs(so.X_Y_to_U(X, Y) - Af).max() < 1e-10                      # Y X = orthonorm_freeze(U)
    eye = np.eye(nw)
    assert np.abs(np.conj(X.transpose(0, 2, 1)) @ X - eye).max() < 1e-11
    assert np.abs(np.conj(Y.transpose(0, 2, 1)) @ Y - eye).max() < 1e-11
    Xo, Yo = so.U_to_X_Y(U, frozen)
    for k in range(nk):
        f = frozen[k]
        nf = int(f.sum())
        # frozen block of Y exactly as in the reference (disentangle.jl:386): identity on the frozen rows
        assert np.array_equal(Y[k][f, :nf], np.eye(nf)) and not Y[k][f, nf:].any() and not Y[k][~f, :nf].any()
        # free block: the same subspace as the reference's eigenvectors (projectors agree)
        P, Po = Y[k][~f, nf:] @ np.conj(Y[k][~f, nf:].T), Yo[k][~f, nf:] @ np.conj(Yo[k][~f, nf:].T)
        assert np.abs(P - Po).max() < 1e-10
    V = w.orthonorm_freeze(U, frozen, cache=ctx)
    assert np.abs(V - Af).max() < 1e-10
    return X, Y


def test_device_u_to_xy_and_orthonorm_freeze(w, silicon):
    # the reference's 12 -> 8 silicon model: 4 ... 8 frozen bands per k-point, n_frozen == n_wann included
    _check_u_to_xy(w, silicon["U"], silicon["frozen"])
    assert silicon["frozen"].sum(axis=1).max() == 8
    # U <-> (X, Y) round trip of the reference's own test (test/localization/disentangle.jl:8-25)
    X, Y = w.U_to_X_Y(silicon["U"], silicon["frozen"])
    X2, Y2 = w.U_to_X_Y(so.X_Y_to_U(X, Y), silicon["frozen"])
    assert np.abs(so.X_Y_to_U(X2, Y2) - so.X_Y_to_U(X, Y)).max() < 1e-10
    # synthetic shapes: fused small kernels, batched-GEMM retraction (nb > 64), nothing frozen, ragged frozen sets
    rng = np.random.default_rng(11)
    for lat, grid, nb, nw, nfroz in (("fcc", (3, 3, 3), 18, 9, 5), ("fcc", (2, 2, 2), 72, 40, 9),
                                     ("cubic", (2, 2, 2), 40, 24, 0), ("bcc", (2, 2, 2), 128, 96, 17)):
        st, M, U = w.synthetic.make_model_arrays(lat, grid, nb, nw, eps=0.2)
        frozen = np.zeros((len(U), nb), dtype=bool)
        for k in range(len(U)):          # a different frozen set at every k-point, not the lowest bands
            frozen[k, rng.choice(nb, size=rng.integers(0, nfroz + 1), replace=False)] = True
        ctx = make_ctx(w, st, M, nw)
        ctx.set_frozen(frozen)

        class _C:        # api-level cache stand-in: only .ctx is read
            pass
        c = _C(); c.ctx = ctx
        _check_u_to_xy(w, U, frozen, ctx=c)
        ctx.close()
    # more frozen bands than Wannier functions
    st, M, U = w.synthetic.make_model_arrays("fcc", (2, 2, 2), 6, 4)
    with pytest.raises(w.WB200Error):
        w.U_to_X_Y(U, w.synthetic.make_frozen(len(U), 6, 5))
    # a Stiefel-only context refuses evaluations
    sc = w.StiefelContext(6, 4, len(U))
    with pytest.raises(w.WB200Error):
        sc.fg(w.api.to_abi_U(U))
    sc.close()
