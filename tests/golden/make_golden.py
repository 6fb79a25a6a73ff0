"""Generate tests/golden/*.npz from the reference's in-tree fixtures.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The .npz files hold (a) the INPUT arrays of each fixture model (overlaps, stencil tables,
gauges) exactly as the reference's read_w90 would assemble them, and (b) the reference's own
known answers that go with them (W90 m_matrix from the .chk.fmt, .wout final state, literal
constants copied from the reference's test files with their file:line).  Oracle OUTPUTS are
not stored as goldens, except where labelled ``oracle_*`` (regression values, unpinned).
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import w90_fixtures as wf  # noqa: E402
from oracle import spread_oracle as so  # noqa: E402

FIX = "/root/reference/test/fixtures"
OUT = os.path.join(ROOT, "tests", "golden")


def pack_model(m):
    return dict(M=m["M"], kpb_k=m["kpb_k"].astype(np.int32), kpb_G=m["kpb_G"].astype(np.int32),
                bvec=m["bvec"], wb=m["wb"], U=m["U"], kpoints=m["kpoints"], recip=m["recip"],
                lattice=m["lattice"], frozen=m["frozen"])


def main():
    # ---- valence silicon: 4 bands -> 4 WF, 4x4x4 k, 8 b  (BASELINE config 1) ----
    val = wf.read_model(f"{FIX}/valence/silicon")
    ptg = so.orthonorm_lowdin(wf.read_amn(f"{FIX}/valence/silicon.ptg.amn"))
    # W1 literal: test/localization/max_localize.jl:31-36
    W1 = np.array([
        [0.497264+0.129534j, -0.281405-0.540402j, 0.0625545+0.237438j, -0.516272-0.194678j],
        [0.113245-0.532338j, 0.47272+0.0873106j, -0.546504+0.19511j, -0.36623+0.042955j],
        [-0.0956088+0.205024j, 0.468778+0.234124j, 0.563928-0.141122j, -0.572565+0.0921779j],
        [0.561868-0.269944j, 0.350754-0.00970417j, 0.451343+0.247661j, 0.471665-0.0282528j]])
    d = pack_model(val)
    d.update(U_ptg=ptg, W1=W1,
             # test/localization/max_localize.jl:52-55 (converged max_localize from ptg gauge)
             ref_maxloc=np.array([6.374823673444644, 5.812709709242578,
                                  0.5621139641912166, 0.5621139642020658]))
    np.savez_compressed(f"{OUT}/valence_silicon.npz", **d)

    # ---- silicon 12 bands -> 8 WF, frozen window dis_froz_max = 10 (disentangle) ----
    sil = wf.read_model(f"{FIX}/silicon/silicon")
    chk = wf.read_chk_fmt(f"{FIX}/silicon/silicon.chk.fmt")
    d = pack_model(sil)
    d.update(U_chk=wf.chk_gauges(chk), chk_m_matrix=chk["m_matrix"], chk_OmegaI=chk["OmegaI"],
             chk_centres=chk["centres"], chk_spreads=chk["spreads"],
             # test/fixtures/silicon/silicon.wout:3393-3407 (Final State)
             wout_final=np.array([20.491096007, 12.599083622, 7.719386591, 0.172625794]),
             wout_centres=np.array([[1.349402] * 3] * 4 + [[0.0] * 3] * 4),
             wout_spreads=np.array([1.68869719, 2.85228361, 2.85228360, 2.85228362,
                                    1.68869719, 2.85228361, 2.85228360, 2.85228360]),
             E=sil["E"])
    np.savez_compressed(f"{OUT}/silicon_12to8.npz", **d)

    # ---- split valence/conduction overlaps: test/localization/split.jl:17-47 ----
    Vv = wf.read_amn(f"{FIX}/valence/silicon.vmn")
    Vc = wf.read_amn(f"{FIX}/conduction/silicon.vmn")
    Mv, _, _ = wf.read_mmn(f"{FIX}/valence/silicon.mmn")
    Mc, _, _ = wf.read_mmn(f"{FIX}/conduction/silicon.mmn")
    np.savez_compressed(f"{OUT}/silicon_split.npz", Vv=Vv, Vc=Vc, Mv_ref=Mv, Mc_ref=Mc)

    # ---- sanity print: pin the oracle (also asserted in tests/test_oracle_golden.py) ----
    N = so.transform_gauge(sil["M"], sil["kpb_k"], wf.chk_gauges(chk))
    print("max|N - chk.m_matrix| =", np.abs(N - chk["m_matrix"]).max())
    sp = so.omega(sil["M"], sil["kpb_k"], sil["bvec"], sil["wb"], wf.chk_gauges(chk))
    print("silicon chk gauge: Omega, OmegaI, OmegaOD, OmegaD =", sp["Omega"], sp["OmegaI"],
          sp["OmegaOD"], sp["OmegaD"], " chk OmegaI", chk["OmegaI"])
    sp = so.omega(val["M"], val["kpb_k"], val["bvec"], val["wb"], ptg)
    print("valence ptg gauge: Omega, OmegaI =", sp["Omega"], sp["OmegaI"])
    print("wb valence", val["wb"], "b0", val["bvec"][0, 0])
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
