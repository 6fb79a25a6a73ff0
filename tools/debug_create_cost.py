"""GPU: where a one-shot call's fixed cost goes (WB200_TRACE_CREATE=1 prints the library's own phases)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
w = ge.load_package(); api = w.api
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "silicon_12to8.npz")); d = {k: z[k] for k in z.files}
st = api.KspaceStencil(d["recip"], d["kpoints"], d["wb"], d["kpb_k"], d["kpb_G"])
M, U = d["M"], d["U"]
def t(fn, n=20):
    fn(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return 1e3 * min(ts)
print("to_abi_M            %.3f ms" % t(lambda: api.to_abi_M(M)))
print("to_abi_U            %.3f ms" % t(lambda: api.to_abi_U(U)))
print("bvectors_cart       %.3f ms" % t(lambda: st.bvectors_cart()))
Mc = api.to_abi_M(M); bv = st.bvectors_cart()
for rep in range(3):
    print("--- rep", rep, flush=True)
    t0 = time.perf_counter()
    c = w.Context(12, 8, 64, 8, st.kpb_k, bv, st.bweights, Mc)
    t1 = time.perf_counter()
    out = c.spread(api.to_abi_U(U))
    t2 = time.perf_counter()
    c.close()
    t3 = time.perf_counter()
    print("Context() %.3f ms, spread %.3f ms, close %.3f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)), flush=True)
