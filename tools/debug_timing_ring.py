"""GPU debugging aid: per-warp cycle accounting of rotate_small_ring_kernel at cfg3 (needs the -DWB_TIMING build,
WB200_LIB pointing at it)."""
import sys, os, ctypes, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
w = ge.load_package()
grid = (12, 12, 12)
nb = nw = 32
st = w.synthetic.make_stencil(w.synthetic.LATTICES["bcc"](), grid)
nk, nn = st["kpb_k"].shape
dev = torch.device("cuda", 0)
M = w.synthetic.make_overlaps_torch(st, nb, dev)
dU = w.synthetic.make_gauge_torch(nk, nb, nw, dev)
os.environ["WB200_GRAPH"] = "0"
ctx = w.Context(nb, nw, nk, nn, st["kpb_k"], st["bvec"], st["wb"], M.data_ptr())
del M
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
res = torch.zeros(ctx.nres, dtype=torch.float64, device=dev)
G = torch.zeros((nk, nw, nb), dtype=torch.complex128, device=dev)
dbg = torch.zeros((148, 8, 8), dtype=torch.int64, device=dev)
lib = w.lib.load()
lib.wb200_debug_set_timing_buffer.argtypes = [ctypes.c_void_p]
for i in range(5):
    ctx.fg_dev(dU.data_ptr(), res.data_ptr(), G.data_ptr())
torch.cuda.synchronize()
assert lib.wb200_debug_set_timing_buffer(dbg.data_ptr()) == 0
ctx.fg_dev(dU.data_ptr(), res.data_ptr(), G.data_ptr())
torch.cuda.synchronize()
d = dbg.cpu().numpy().astype(np.float64)
tot, wait, mma, epi, npairs = [d[..., i] for i in range(5)]
ok = tot > 0
print("warps reporting: %d; pairs per warp pair: %.1f" % (ok.sum(), npairs[ok].mean()))
print("per-warp cycles: total %.0f  wait %.0f (%.1f%%)  k-steps %.0f (%.1f%%)  epilogue %.0f (%.1f%%)  other %.0f" % (
    tot[ok].mean(), wait[ok].mean(), 100 * wait[ok].mean() / tot[ok].mean(), mma[ok].mean(), 100 * mma[ok].mean() / tot[ok].mean(),
    epi[ok].mean(), 100 * epi[ok].mean() / tot[ok].mean(), (tot - wait - mma - epi)[ok].mean()))
pp = npairs[ok].mean()
print("per pair per warp: wait %.0f  k-steps %.0f  epilogue %.0f  (DMMA pipe alone: %d cycles for 192 DMMA; shared by two warps: %d)" % (
    wait[ok].mean() / pp, mma[ok].mean() / pp, epi[ok].mean() / pp, 192 * 16, 192 * 32))
print("slowest / fastest warp total: %.0f / %.0f" % (tot[ok].max(), tot[ok].min()))
