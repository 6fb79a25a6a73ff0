"""GPU: config 2 shape (18 bands -> 9 WFs, 9^3 k, 8 b, 5 frozen bands): one fg!(Omega, G, XY) of the
disentanglement through the host entry point, and the device disentangle driver."""
import sys, os, time, json, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
w = ge.load_package()
lat, grid, nb, nw = w.synthetic.CONFIGS["cfg2_cu_shape"]
st, M, U = w.synthetic.make_model_arrays(lat, grid, nb, nw)
nk, nn = st["kpb_k"].shape
frozen = w.synthetic.make_frozen(nk, nb, 5)
kst = w.KspaceStencil(st["recip"], st["kpoints"], st["wb"], st["kpb_k"], st["kpb_G"])
model = w.Model(st["lattice"], kst, M, U, frozen)
cache = w.Cache.from_model(model)
X, Y = w.U_to_X_Y(U, frozen)
XY = w.X_Y_to_XY(X, Y)
G = np.empty_like(XY)
for _ in range(5):
    f = cache.ctx.fg_xy(XY, G=G)
n = 50
t0 = time.perf_counter()
for _ in range(n):
    f = cache.ctx.fg_xy(XY, G=G)
dt = (time.perf_counter() - t0) / n
out = dict(fg_xy_ms=1e3 * dt, f=f, kernel=cache.ctx.rotation_kernel)
t0 = time.perf_counter()
Umin, info = w.disentangle(model, max_iter=50, return_info=True, cache=cache)
out["disentangle"] = dict(seconds=time.perf_counter() - t0, **info)
# the device driver alone (warm; the call above also pays U_to_X_Y on the host and one-off allocations)
reps = []
for _ in range(3):
    XYc = XY.copy()
    l0 = cache.ctx.launch_count
    t0 = time.perf_counter()
    f, iters, nfg = cache.ctx.disentangle(XYc, 1e-7, 1e-5, 50, 3)
    dt = time.perf_counter() - t0
    reps.append(dict(seconds=dt, f=f, iterations=iters, n_fg=nfg, ms_per_fg=1e3 * dt / nfg,
                     launches=cache.ctx.launch_count - l0))
out["disentangle_driver_warm"] = reps
print(json.dumps(out))
json.dump(out, open("gpurun_out/bench_disentangle.json", "w"), indent=1)
