import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
import __graft_entry__ as ge
w = ge.load_package(); api = w.api
z = np.load("tests/golden/valence_silicon.npz"); d = {k: z[k] for k in z.files}
st = api.KspaceStencil(d["recip"], d["kpoints"], d["wb"], d["kpb_k"], d["kpb_G"])
m = api.Model(d["lattice"], st, d["M"], d["U"])
c = api.Cache.from_model(m)
for rep in range(3):
    t0 = time.perf_counter()
    U, info = api.max_localize(m, max_iter=10, cache=c, return_info=True)
    print("rep", rep, info, 1e3 * (time.perf_counter() - t0), "ms", flush=True)
c.close()
