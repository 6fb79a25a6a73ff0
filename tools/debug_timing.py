"""GPU debugging aid: per-warp cycle accounting of rotate_dmma_kernel (needs the -DWB_TIMING build)."""
import sys, os, ctypes, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
w = ge.load_package()
grid = (16, 16, 16)
nb = nw = 128
st = w.synthetic.make_stencil(w.synthetic.LATTICES["bcc"](), grid)
nk, nn = st["kpb_k"].shape
dev = torch.device("cuda", 0)
M = w.synthetic.make_overlaps_torch(st, nb, dev)
dU = w.synthetic.make_gauge_torch(nk, nb, nw, dev)
ctx = w.Context(nb, nw, nk, nn, st["kpb_k"], st["bvec"], st["wb"], M.data_ptr())
del M
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
res = torch.zeros(ctx.nres, dtype=torch.float64, device=dev)
G = torch.zeros((nk, nw, nb), dtype=torch.complex128, device=dev)
nct = 4
dbg = torch.zeros((nk * nct, 8, 8), dtype=torch.int64, device=dev)
lib = w.lib.load()
lib.wb200_debug_set_timing_buffer.argtypes = [ctypes.c_void_p]
for i in range(3):
    ctx.fg_dev(dU.data_ptr(), res.data_ptr(), G.data_ptr())
torch.cuda.synchronize()
assert lib.wb200_debug_set_timing_buffer(dbg.data_ptr()) == 0
ctx.fg_dev(dU.data_ptr(), res.data_ptr(), G.data_ptr())
torch.cuda.synchronize()
d = dbg.cpu().numpy().astype(np.float64)
tot, wait, epi, first, g0, g1, smid, setup = [d[..., i] for i in range(8)]
print("per-warp cycles: total %.0f  wait %.0f (%.1f%%)  epilogue %.0f (%.1f%%)  first-data %.0f  setup(to setmaxnreg) %.0f" % (
    tot.mean(), wait.mean(), 100 * wait.mean() / tot.mean(), epi.mean(), 100 * epi.mean() / tot.mean(), first.mean(), setup.mean()))
print("  compute(main loop minus wait) %.0f (%.1f%%); ideal DMMA pipe cycles per warp if alone on SMSP %.0f, shared by two %.0f" % (
    (tot - wait - epi).mean(), 100 * (tot - wait - epi).mean() / tot.mean(), 12 * 16 * 2 * 24 * 16.0, 12 * 16 * 2 * 24 * 32.0))
print("  per-pair epilogue cycles %.0f ; per-chunk wait cycles %.1f" % (epi.mean() / nn, wait.mean() / (nn * 16)))
for wg in (slice(0, 4), slice(4, 8)):
    print("  warps", wg, "total %.0f wait %.0f epi %.0f" % (tot[:, wg].mean(), wait[:, wg].mean(), epi[:, wg].mean()))
# inter-CTA gaps per SM from globaltimer (ns)
cta_start = g0.min(axis=1); cta_end = g1.max(axis=1); sm = smid[:, 0]
gaps = []; durs = cta_end - cta_start
for s_ in np.unique(sm):
    idx = np.nonzero(sm == s_)[0]
    o = np.argsort(cta_start[idx])
    a, b = cta_start[idx][o], cta_end[idx][o]
    gaps.extend((a[1:] - b[:-1]).tolist())
print("CTA duration (consumer start..end) mean %.1f us; gap between consecutive CTAs on an SM mean %.2f us (median %.2f); CTAs/SM %.1f" % (
    durs.mean() / 1e3, np.mean(gaps) / 1e3, np.median(gaps) / 1e3, len(sm) / len(np.unique(sm))))
print("kernel span %.2f ms; sum of (dur+gap) per SM %.2f ms" % ((cta_end.max() - cta_start.min()) / 1e6, (durs.mean() + np.mean(gaps)) * len(sm) / len(np.unique(sm)) / 1e6))
