"""GPU debugging aid: multi-wave determinism / parity of the DMMA rotation kernel."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
w = ge.load_package()
from oracle import spread_oracle as so
grid = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (6, 6, 6)
nb = nw = 128
st = w.synthetic.make_stencil(w.synthetic.LATTICES["bcc"](), grid)
nk, nn = st["kpb_k"].shape
dev = torch.device("cuda", 0)
M = w.synthetic.make_overlaps_torch(st, nb, dev)
dU = w.synthetic.make_gauge_torch(nk, nb, nw, dev)
ctx = w.Context(nb, nw, nk, nn, st["kpb_k"], st["bvec"], st["wb"], M.data_ptr())
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
res = torch.zeros(ctx.nres, dtype=torch.float64, device=dev)
G = torch.zeros((nk, nw, nb), dtype=torch.complex128, device=dev)
outs = []
for i in range(4):
    ctx.fg_dev(dU.data_ptr(), res.data_ptr(), G.data_ptr())
    torch.cuda.synchronize()
    outs.append((res.cpu().numpy().copy(), G.cpu().numpy().copy()))
for i in range(1, 4):
    print("call", i, "f diff", outs[i][0][0] - outs[0][0][0], "G maxdiff", np.abs(outs[i][1] - outs[0][1]).max(),
          "n differing k", int((np.abs(outs[i][1] - outs[0][1]).reshape(nk, -1).max(axis=1) > 0).sum()))
hU = dU.cpu().numpy()
hG = np.empty((nk, nw, nb), dtype=np.complex128)
f = ctx.fg(hU, G=hG)
print("host path f diff", f - outs[0][0][5], "G maxdiff", np.abs(hG - outs[0][1]).max())
if nk <= 512:
    Mh = M.transpose(2, 3).cpu().numpy(); Uh = hU.transpose(0, 2, 1)
    f_ref, G_ref = so.fg_maxloc(Mh, st["kpb_k"], st["bvec"], st["wb"], Uh)
    print("oracle rel f", abs(outs[0][0][5] - f_ref) / abs(f_ref), "G", np.abs(outs[0][1].transpose(0, 2, 1) - G_ref).max() / np.abs(G_ref).max())
    N, MU = ctx.rotate_overlaps(hU, want_N=False, want_MU=True)
    MU_ref, _ = so.compute_MU_UtMU(Mh, st["kpb_k"], Uh)
    d = np.abs(MU.transpose(0, 1, 3, 2) - MU_ref)
    print("MU max abs err", d.max(), "pairs with err>1e-12:", int((d.reshape(nk * nn, -1).max(axis=1) > 1e-12).sum()))
    bad = np.nonzero(d.reshape(nk * nn, -1).max(axis=1) > 1e-12)[0]
    for p in bad[:4]:
        e = d.reshape(nk * nn, nb, nw)[p]
        rows = np.nonzero(e.max(axis=1) > 1e-12)[0]; cols = np.nonzero(e.max(axis=0) > 1e-12)[0]
        print("bad pair", p, "k", p // nn, "b", p % nn, "rows", rows.min(), rows.max(), len(rows), "cols", cols.min(), cols.max(), len(cols))
        # which l-range explains the error?  err = sum over a chunk range of M[:, l] U[l, :]
        k, b = p // nn, p % nn
        Uk = Uh[st["kpb_k"][k, b]]
        diff = (MU.transpose(0, 1, 3, 2)[k, b] - MU_ref[k, b])
        for c in range(0, nb, 8):
            contrib = Mh[k, b][:, c:c + 8] @ Uk[c:c + 8, :]
            sub = np.ix_(rows, cols)
            r1 = np.abs(diff[sub] + contrib[sub]).max(); r2 = np.abs(diff[sub] - contrib[sub]).max()
            if min(r1, r2) < 1e-3 * np.abs(diff[sub]).max():
                print("   explained by chunk l0 =", c, "missing" if r1 < r2 else "doubled")
