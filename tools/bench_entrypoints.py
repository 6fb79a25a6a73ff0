"""GPU: timings of the user-facing entry points beyond the headline fg (round 2):
  * one evaluation per BASELINE config with U resident, CUDA-graph replay on / off (WB200_GRAPH);
  * the exact-split evaluation behind omega(model) (second GEMM U_k^H MU on the tensor pipe);
  * host-pointer one-shots: wb200_spread, wb200_rotate_overlaps, wb200_fg_xy;
  * device-resident max_localize / disentangle: ms per evaluation.
Wall clock around n calls with a stream synchronisation on both sides (the ctx works on its own stream).
    python tools/bench_entrypoints.py [out.json]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

w = ge.load_package()
dev = torch.device("cuda", 0)
dmma_peak, dfma_peak = w.lib.probe_fp64_peak(0)
out = {"fp64_peak_tflops": {"dmma": dmma_peak, "dfma": dfma_peak}}


def timed(fn, sync, min_s=0.3, nmin=5, nmax=3000):
    fn(); sync()
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < 0.3 or n < nmin:      # warm-up (clock ramp)
        fn(); n += 1
    sync()
    est = (time.perf_counter() - t0) / n
    n = int(min(nmax, max(nmin, min_s / est)))
    sync(); t0 = time.perf_counter()
    for _ in range(n):
        fn()
    sync()
    return 1e3 * (time.perf_counter() - t0) / n


def make(name, graph=True):
    lat, grid, nb, nw = w.synthetic.CONFIGS[name]
    st = w.synthetic.make_stencil(w.synthetic.LATTICES[lat](), grid)
    nk, nn = st["kpb_k"].shape
    M = w.synthetic.make_overlaps_torch(st, nb, dev)
    dU = w.synthetic.make_gauge_torch(nk, nb, nw, dev)
    torch.cuda.synchronize()
    os.environ["WB200_GRAPH"] = "1" if graph else "0"
    ctx = w.Context(nb, nw, nk, nn, st["kpb_k"], st["bvec"], st["wb"], M.data_ptr())
    del M
    torch.cuda.empty_cache()
    return ctx, st, dU, (nb, nw, nk, nn)


_all = ["cfg1_si_valence_shape", "cfg2_cu_shape", "cfg3_nw32", "cfg5_2d_nw256", "cfg4_nw128"]
_sel = os.environ.get("WB_CONFIGS")      # comma-separated subset
for name in ([c for c in _all if c in _sel.split(",")] if _sel else _all):
    r = {}
    for graph in (True, False):
        ctx, st, dU, (nb, nw, nk, nn) = make(name, graph)
        res = torch.zeros(ctx.nres, dtype=torch.float64, device=dev)
        G = torch.zeros((nk, nw, nb), dtype=torch.complex128, device=dev)
        torch.cuda.synchronize()
        l0 = ctx.launch_count
        ms = timed(lambda: ctx.fg_dev(dU.data_ptr(), res.data_ptr(), G.data_ptr()), ctx.synchronize)
        key = "graph" if graph else "plain"
        r[key + "_ms"] = ms
        if graph:
            flops = 8.0 * nk * nn * nb * nb * nw
            r.update(nb=nb, nw=nw, nk=nk, nn=nn, kernel=ctx.rotation_kernel, omega=res[0].item(),
                     tflops_algorithmic=flops / ms / 1e9, frac_of_dmma_peak=flops / ms / 1e9 / dmma_peak)
            # exact split: MU GEMM + N = U_k^H MU GEMM (what omega(model) / transform_gauge run)
            ms_x = timed(lambda: ctx.fg_dev(dU.data_ptr(), res.data_ptr(), 0, exact_split=True), ctx.synchronize)
            fl2 = flops + 8.0 * nk * nn * nb * nw * nw
            r["exact_split_ms"] = ms_x
            r["exact_split_tflops"] = fl2 / ms_x / 1e9
            r["exact_split_frac_of_dmma_peak"] = fl2 / ms_x / 1e9 / dmma_peak
            if name in ("cfg2_cu_shape", "cfg3_nw32", "cfg4_nw128"):
                hU = dU.cpu().numpy().copy()
                t0 = time.perf_counter()
                out5, om, rr, mn = ctx.spread(hU)
                r["spread_host_first_ms"] = 1e3 * (time.perf_counter() - t0)
                r["spread_host_ms"] = timed(lambda: ctx.spread(hU), lambda: None, min_s=0.2, nmin=3)
                if nb == nw:
                    hU2 = hU.copy()
                    w.lib.host_register(hU2)       # what the Julia wrapper does with the caller's Array
                    for rep in range(2):           # the first call pays lazy kernel loading and one-off allocations
                        hU2[...] = hU
                        t0 = time.perf_counter()
                        f, it, nfg = ctx.max_localize(hU2, max_iter=200 if name != "cfg4_nw128" else 20)
                        dt = time.perf_counter() - t0
                    w.lib.host_unregister(hU2)
                    r["max_localize"] = dict(f=f, iterations=it, n_fg=nfg, seconds=dt, ms_per_fg=1e3 * dt / nfg)
                if name != "cfg4_nw128":
                    N, _ = ctx.rotate_overlaps(hU)
                    r["rotate_overlaps_host_ms"] = timed(lambda: ctx.rotate_overlaps(hU), lambda: None, min_s=0.2, nmin=3)
            if nb > nw:    # (X, Y) chain and the device disentangle driver
                frozen = w.synthetic.make_frozen(nk, nb, 5)
                ctx.set_frozen(frozen)
                U = np.ascontiguousarray(dU.cpu().numpy().transpose(0, 2, 1))
                X, Y = w.api.U_to_X_Y(U, frozen)          # the product's own setup step (wb200_u_to_xy)
                XY = np.ascontiguousarray(w.api.X_Y_to_XY(X, Y))
                Gxy = np.empty_like(XY)
                r["fg_xy_host_ms"] = timed(lambda: ctx.fg_xy(XY, G=Gxy), lambda: None, min_s=0.2, nmin=3)
                for rep in range(2):               # first call: lazy kernel loading, one-off allocations
                    XYc = XY.copy()
                    t0 = time.perf_counter()
                    f, it, nfg = ctx.disentangle(XYc, max_iter=50)
                    dt = time.perf_counter() - t0
                r["disentangle"] = dict(f=f, iterations=it, n_fg=nfg, seconds=dt, ms_per_fg=1e3 * dt / nfg)
        ctx.close()
        del dU, G, res
        torch.cuda.empty_cache()
    out[name] = r
    print(name, json.dumps(r), flush=True)
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2_entrypoints.json", "w"), indent=1)
