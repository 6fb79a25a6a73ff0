"""GPU: the ONE-process multi-GPU boundary (wb200_multi_*: what a single Julia process calling max_localize(model)
uses) at the north-star shape — end-to-end fg!(F, G, U) with host pointers and the device max_localize driver.

    python tools/bench_multi.py [ndev] [config] [max_iter] > out.json

One host thread per k-slab inside the library; with one slab per device every collective runs over peer memory
(peer.cu).  Host arrays are page-locked (wb200_host_register), as the Julia wrapper does."""
import json, os, sys, time
import numpy as np, torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

w = ge.load_package()
ndev = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg4_nw128"
max_iter = int(sys.argv[3]) if len(sys.argv) > 3 else 10
lat, grid, nb, nw = w.synthetic.CONFIGS[cfg]
st = w.synthetic.make_stencil(w.synthetic.LATTICES[lat](), grid)
nk, nn = st["kpb_k"].shape
dev = torch.device("cuda", 0)
# the full model on the host, as a caller holds it (synthetic inputs generated on device 0, slab by slab)
M = np.empty((nk, nn, nb, nb), dtype=np.complex128)
step = max(1, nk // 8)
for k0 in range(0, nk, step):
    k1 = min(nk, k0 + step)
    M[k0:k1] = w.synthetic.make_overlaps_torch(st, nb, dev, k_begin=k0, nk=k1 - k0).cpu().numpy().reshape(k1 - k0, nn, nb, nb)
U0 = w.synthetic.make_gauge_torch(nk, nb, nw, dev).cpu().numpy()
torch.cuda.empty_cache()
t0 = time.perf_counter()
m = w.MultiContext(nb, nw, nk, nn, st["kpb_k"], st["bvec"], st["wb"], M, devices=list(range(ndev)))
t_create = time.perf_counter() - t0
del M
U = U0.copy(); G = np.empty_like(U0)
w.lib.host_register(U); w.lib.host_register(G)
f = m.fg(U, G=G)
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    f = m.fg(U, G=G)
    ts.append(time.perf_counter() - t0)
out = dict(config=cfg, n_gpus=ndev, create_seconds=t_create, f=f, g_check=float((np.abs(G) ** 2).sum()),
           e2e_fg=dict(ms=1e3 * float(np.mean(ts)), evals_per_s=1.0 / float(np.mean(ts)),
                       api="wb200_multi_fg: full U in, f and full G out, page-locked host arrays"))
if nb == nw:
    for rep in range(2):
        U[...] = U0
        t0 = time.perf_counter()
        fm, it, nfg = m.max_localize(U, max_iter=max_iter)
        dt = time.perf_counter() - t0
    out["max_localize"] = dict(max_iter=max_iter, iterations=it, n_evaluations=nfg, seconds=dt,
                               ms_per_evaluation=1e3 * dt / max(nfg, 1), f_end=fm,
                               api="wb200_multi_max_localize: full U in and out")
print(json.dumps(out), flush=True)
w.lib.host_unregister(U); w.lib.host_unregister(G)
m.close()
