import sys, os, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
w = ge.load_package()
dev = torch.device("cuda", 0)
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg3_nw32"
lat, grid, nb, nw = w.synthetic.CONFIGS[cfg]
st = w.synthetic.make_stencil(w.synthetic.LATTICES[lat](), grid)
nk, nn = st["kpb_k"].shape
M = w.synthetic.make_overlaps_torch(st, nb, dev)
dU = w.synthetic.make_gauge_torch(nk, nb, nw, dev)
ctx = w.Context(nb, nw, nk, nn, st["kpb_k"], st["bvec"], st["wb"], M.data_ptr())
del M
torch.cuda.empty_cache()
hU = dU.cpu().numpy().copy()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
t0 = time.perf_counter()
f, it, nfg = ctx.max_localize(hU, max_iter=n)
print("iters", it, "nfg", nfg, "f", f, "s", time.perf_counter() - t0, "launches", ctx.launch_count)
