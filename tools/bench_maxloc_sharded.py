"""GPU: device max_localize, k-sharded over the ranks of a torchrun launch (or one GPU without it).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_maxloc_sharded.py [config] [max_iter]

Every rank owns a contiguous k-slab: overlaps, L-BFGS vectors and gradients never move; neighbour
gauges are read from peer memory (CUDA IPC over NVLink) inside the pack kernel and the optimiser's
scalars go through the all-reduce callback (torch.distributed / NCCL).  Prints one JSON line."""
import json, os, sys, time
import numpy as np, torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

w = ge.load_package()
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4_nw128"
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 6
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)

lat, grid, nb, nw = w.synthetic.CONFIGS[cfg]
st = w.synthetic.make_stencil(w.synthetic.LATTICES[lat](), grid)
nk, nn = st["kpb_k"].shape
bounds = w.sharding.slab_bounds(nk, world)
k0, k1 = bounds[rank], bounds[rank + 1]
M = w.synthetic.make_overlaps_torch(st, nb, dev, k_begin=k0, nk=k1 - k0)
ctx = w.Context(nb, nw, nk, nn, st["kpb_k"][k0:k1], st["bvec"][k0:k1], st["wb"], M.data_ptr(),
                device=local, k_begin=k0, nk=k1 - k0)
del M
torch.cuda.empty_cache()
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
Uslab0 = w.synthetic.make_gauge_torch(nk, nb, nw, dev)[k0:k1].cpu().numpy().copy()
shared = None
if world > 1:
    comm = os.environ.get("WB200_PEER_COMM", "1") != "0"
    shared = w.sharding.SharedGauges(ctx, dist, local, rank, world, nk, nb, nw, bounds, comm=comm)
    if not comm:
        ctx.set_allreduce(w.sharding.torch_allreduce(dist, dev))

res = []
for rep in range(2):          # the first call pays one-off allocations
    U = Uslab0.copy()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    f, it, nfg = ctx.max_localize(U, f_tol=1e-7, g_tol=1e-5, max_iter=max_iter)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    res.append(time.perf_counter() - t0)
if rank == 0:
    print(json.dumps(dict(config=cfg, n_gpus=world, max_iter=max_iter, iterations=it, n_fg=nfg, f=f,
                          seconds=res, ms_per_fg=1e3 * res[-1] / nfg,
                          note="includes H2D of the U slab and D2H of the result")), flush=True)
if world > 1:
    ctx.set_allreduce(None)
    shared.close()
ctx.close()
if world > 1:
    dist.destroy_process_group()
