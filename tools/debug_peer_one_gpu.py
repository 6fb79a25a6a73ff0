"""debug: peer-memory collectives between slabs of one device (see tests/test_gpu_multislab.py)"""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
w = ge.load_package()
case = sys.argv[1]
lat, grid, nb, nw, ndev = {"big": ("cubic", (4, 2, 2), 72, 72, 2), "small": ("bcc", (4, 2, 2), 16, 16, 2)}[case]
st, M, U = w.synthetic.make_model_arrays(lat, grid, nb, nw, eps=0.02)
nk, nn = st["kpb_k"].shape
m = w.MultiContext(nb, nw, nk, nn, st["kpb_k"], st["bvec"], st["wb"], w.api.to_abi_M(M), devices=[0] * ndev)
Ua = w.api.to_abi_U(U)
G = np.empty((nk, nw, nb), dtype=np.complex128)
if "pin" in sys.argv:
    w.lib.host_register(Ua); w.lib.host_register(G)
mode = sys.argv[2]
print("start", case, mode, flush=True)
reps = 3 if "rep" in sys.argv else 1
for r in range(reps):
    if mode == "spread":
        print(m.spread(Ua)[0], flush=True)
    elif mode == "f":
        print(m.fg(Ua, G=None), flush=True)
    else:
        print(m.fg(Ua, G=G), flush=True)
print("done", flush=True)
