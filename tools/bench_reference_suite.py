"""GPU: the reference's OWN benchmark suite (benchmark/bench_spread.jl, bench_grad.jl, bench_maxloc.jl,
bench_disentangle.jl — BenchmarkTools on the Si2 / Si2_valence datasets, one Julia thread), run through
this package's public API the way the reference's suite calls its own: ONE-SHOT calls with host arrays,
so every call pays what a user pays (context creation, upload of M, the computation, the download,
context destruction).  Next to each: the same call on a caller-held Cache (what a loop over gauges costs),
the creation / destruction cost alone, and the compiled CPU port of the oracle on one host thread (the
reference's benchmark setting, benchmark/runbenchmarks.jl:5) as the baseline.

Datasets: tests/golden/silicon_12to8.npz (= Si2: 12 bands -> 8 WFs, 4x4x4 k, 8 b, frozen window) and
tests/golden/valence_silicon.npz (= Si2_valence: 4 WFs), both made from the reference's fixture files.

    python tools/bench_reference_suite.py [out.json] [--no-cpu] [--quick]"""
import json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _load(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name))
    return {k: z[k] for k in z.files}


def bench(fn, min_s=0.25, nmin=5, nmax=2000):
    """BenchmarkTools-style: warm up, then the MINIMUM and the median of n timed calls (ms)."""
    fn(); fn()
    t0 = time.perf_counter(); fn(); est = time.perf_counter() - t0
    n = int(min(nmax, max(nmin, min_s / max(est, 1e-6))))
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    ts.sort()
    return {"min_ms": 1e3 * ts[0], "median_ms": 1e3 * ts[len(ts) // 2], "n": n}


def run(cpu=True, min_s=0.25):
    """Returns the result dictionary (GPU: the public API of the package; cpu=True adds the compiled CPU port on one
    thread as the baseline — bench.py's cpu_baseline leg and this tool are the only callers)."""
    import __graft_entry__ as ge
    w = ge.load_package()
    api = w.api

    def model_of(d, U, frozen=None):
        st = api.KspaceStencil(d["recip"], d["kpoints"], d["wb"], d["kpb_k"], d["kpb_G"])
        return api.Model(d["lattice"], st, d["M"], U, frozen)

    B = lambda fn, **kw: bench(fn, min_s=min_s, **kw)
    si2 = _load("silicon_12to8.npz")
    val = _load("valence_silicon.npz")
    m_si2 = model_of(si2, si2["U"], si2["frozen"])
    m_val = model_of(val, val["U"])
    out = {"datasets": {"Si2": "12 bands -> 8 WFs, 64 k, 8 b (tests/golden/silicon_12to8.npz)",
                        "Si2_valence": "4 WFs, 64 k, 8 b (tests/golden/valence_silicon.npz)"},
           "timer": "time.perf_counter around the public call (host numpy arrays in, numpy out); minimum and median of n calls",
           "what": "the reference's own benchmark suite (benchmark/bench_spread.jl, bench_grad.jl, bench_maxloc.jl, "
                   "bench_disentangle.jl) as ONE-SHOT calls: context creation, upload of M, computation, download and "
                   "context destruction are all inside every timed call; `cache=c` rows reuse a caller-held Cache"}
    st, M, U = m_si2.kstencil, m_si2.overlaps, m_si2.gauges
    r = {}
    # --- bench_spread.jl / bench_grad.jl: omega(bvectors, M, U), omega_grad(bvectors, M, U) on Si2 ---
    r["omega(stencil, M, U) one-shot"] = B(lambda: api.omega(st, M, U))
    r["omega_grad(stencil, M, U) one-shot"] = B(lambda: api.omega_grad(st, M, U))
    c = api.Cache(st, M, U.shape[2])
    r["omega(..., cache=c)"] = B(lambda: api.omega(st, M, U, cache=c))
    r["omega_grad(..., cache=c)"] = B(lambda: api.omega_grad(st, M, U, cache=c))
    c.close()

    def create_close():
        cc = api.Cache(st, M, U.shape[2]); cc.close()

    r["Cache(stencil, M) create + close"] = B(create_close)
    # --- bench_disentangle.jl ---
    r["U_to_X_Y(U, frozen) one-shot"] = B(lambda: api.U_to_X_Y(U, m_si2.frozen_bands))
    info = {}

    def dis():
        info["dis"] = api.disentangle(m_si2, max_iter=10, return_info=True)[1]

    r["disentangle(model, max_iter=10) one-shot"] = B(dis, nmin=5)
    r["disentangle(model, max_iter=10) one-shot"]["info"] = info["dis"]
    c = api.Cache.from_model(m_si2)
    r["disentangle(model, max_iter=10, cache=c)"] = B(lambda: api.disentangle(m_si2, max_iter=10, cache=c), nmin=5)
    c.close()
    out["Si2"] = r

    # --- bench_maxloc.jl: max_localize(model, max_iter=10) on Si2_valence ---
    r = {}

    def mx():
        info["mx"] = api.max_localize(m_val, max_iter=10, return_info=True)[1]

    r["max_localize(model, max_iter=10) one-shot"] = B(mx, nmin=5)
    r["max_localize(model, max_iter=10) one-shot"]["info"] = info["mx"]
    c = api.Cache.from_model(m_val)
    r["max_localize(model, max_iter=10, cache=c)"] = B(lambda: api.max_localize(m_val, max_iter=10, cache=c), nmin=5)
    r["omega(model, cache=c)"] = B(lambda: api.omega(m_val, cache=c))
    c.close()
    r["omega(model) one-shot"] = B(lambda: api.omega(m_val))
    out["Si2_valence"] = r
    kept, hits, misses = w.lib.pool_stats(0)
    out["caching_allocator"] = {"device_bytes_kept": kept, "hits": hits, "misses": misses}

    if cpu:
        # the CPU side: the compiled port of the oracle on ONE thread (the reference's benchmark setting,
        # benchmark/runbenchmarks.jl:5), its Cache (MU / UtMU work arrays) reused across calls; baseline only
        from oracle import cpu_ref
        cpud = {}
        for name, d, Uc, nev in (("Si2", si2, si2["U"], info["dis"]["n_fg"]), ("Si2_valence", val, val["U"], info["mx"]["n_fg"])):
            Mc, Ua = api.to_abi_M(d["M"]), api.to_abi_U(Uc)
            ws = cpu_ref.Workspace(Mc.shape[2], Ua.shape[1], Mc.shape[0], Mc.shape[1])
            kw = dict(threads=1, ws=ws)
            one = B(lambda: cpu_ref.fg_abi(Mc, d["kpb_k"], d["bvec"], d["wb"], Ua, len(Ua), **kw))
            cpud[name] = {
                "one evaluation (both ZGEMMs per pair + omega! + omega_grad!), 1 thread": one,
                "the same without the gradient (= omega)":
                    B(lambda: cpu_ref.fg_abi(Mc, d["kpb_k"], d["bvec"], d["wb"], Ua, len(Ua), want_G=False, **kw)),
                "lower bound of the 10-iteration driver: its %d evaluations alone, ms" % nev: nev * one["min_ms"],
                "blas": cpu_ref.blas_name()}
        out["cpu_port"] = cpud
    return out


if __name__ == "__main__":
    out = run(cpu="--no-cpu" not in sys.argv, min_s=0.12 if "--quick" in sys.argv else 0.25)
    for k, v in out.items():
        print(k, json.dumps(v, default=float), flush=True)
    dst = next((a for a in sys.argv[1:] if not a.startswith("--")), os.path.join(ROOT, "gpurun_out", "ref_suite.json"))
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    json.dump(out, open(dst, "w"), indent=1, default=float)
