"""GPU: where the fused trial-point kernel (tiny.cu) stops paying: device max_localize / disentangle per evaluation
with and without WB200_FLAG_NO_FUSED_TRIAL on a ladder of shapes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
w = ge.load_package()
shapes = [("fcc", (4, 4, 4), 4, 4, 0), ("fcc", (4, 4, 4), 8, 8, 0), ("fcc", (4, 4, 4), 12, 8, 4), ("fcc", (4, 4, 4), 12, 12, 0),
          ("fcc", (4, 4, 4), 16, 16, 0), ("bcc", (4, 4, 4), 16, 16, 0), ("fcc", (6, 6, 6), 8, 8, 0), ("fcc", (8, 8, 8), 8, 8, 0),
          ("fcc", (6, 6, 6), 12, 12, 0), ("fcc", (8, 8, 8), 12, 8, 4), ("fcc", (6, 6, 6), 18, 9, 5), ("fcc", (9, 9, 9), 18, 9, 5),
          ("fcc", (4, 4, 4), 24, 24, 0), ("fcc", (4, 4, 4), 20, 12, 6), ("fcc", (10, 10, 10), 12, 8, 4),
          ("cubic", (12, 12, 10), 8, 8, 0), ("fcc", (12, 12, 12), 6, 6, 0), ("fcc", (10, 10, 10), 4, 4, 0)]
if os.environ.get("WB_LADDER") == "big":
    shapes = shapes[-6:]
for lat, grid, nb, nw, nfroz in shapes:
    st, M, U = w.synthetic.make_model_arrays(lat, grid, nb, nw)
    nk, nn = st["kpb_k"].shape
    row = []
    for flags in (w.lib.FLAG_DEFAULT, w.lib.FLAG_NO_FUSED_TRIAL):
        ctx = w.Context(nb, nw, nk, nn, st["kpb_k"], st["bvec"], st["wb"], w.api.to_abi_M(M), flags=flags)
        if nfroz:
            frozen = w.synthetic.make_frozen(nk, nb, nfroz)
            ctx.set_frozen(frozen)
            X, Y = w.api.U_to_X_Y(U, frozen)          # the product's own setup step (wb200_u_to_xy)
            x0 = np.ascontiguousarray(w.api.X_Y_to_XY(X, Y))
            run = lambda x: ctx.disentangle(x, 1e-7, 1e-5, 15, 3)
        else:
            x0 = w.api.to_abi_U(U).copy()
            run = lambda x: ctx.max_localize(x, 1e-7, 1e-5, 15, 3)
        best = 1e9
        for rep in range(4):
            x = x0.copy()
            l0 = ctx.launch_count
            t0 = time.perf_counter(); f, it, nfg = run(x); dt = time.perf_counter() - t0
            best = min(best, dt / nfg)
            launches = (ctx.launch_count - l0) / nfg
        ctx.close()
        row.append((1e6 * best, launches))
    cm = nn * nb * nb * nw
    print("%-5s %-12s nb %2d nw %2d nk %4d nn %2d  CMAC/k %6d  fused %7.1f us/eval (%.1f launches)  general %7.1f us/eval (%.1f launches)  ratio %.2f" % (
        lat, grid, nb, nw, nk, nn, cm, row[0][0], row[0][1], row[1][0], row[1][1], row[1][0] / row[0][0]), flush=True)
