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
hU0 = dU.cpu().numpy().copy()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
hU = hU0.copy()
if "--pageable" not in sys.argv:
    w.lib.host_register(hU)          # what the Julia wrapper does with the caller's Array (wb200_host_register)
for rep in range(3 if "--once" not in sys.argv else 1):      # the first call pays the one-off allocations
    hU[...] = hU0
    l0 = ctx.launch_count
    t0 = time.perf_counter()
    f, it, nfg = ctx.max_localize(hU, max_iter=n)
    dt = time.perf_counter() - t0
    print("iters", it, "nfg", nfg, "f", f, "s", dt, "ms/fg", 1e3 * dt / nfg, "launches", ctx.launch_count - l0, flush=True)
