"""Multi-GPU tests (-m gpu; skipped with fewer than 2 GPUs): one process per GPU over NCCL, gauges
shared through CUDA-IPC peer memory, the L-BFGS driver running k-sharded through the caller-supplied
all-reduce.  The sharded run must reach the same spread as the single-GPU run."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, out):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    w = ge.load_package()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        lat, grid, nb = case[:3]
        nw = case[3] if len(case) > 3 else nb
        st, M, U = w.synthetic.make_model_arrays(lat, grid, nb, nw, eps=0.02)
        nk, nn = st["kpb_k"].shape
        bounds = w.sharding.slab_bounds(nk, world)
        k0, k1 = bounds[rank], bounds[rank + 1]
        ctx = w.Context(nb, nw, nk, nn, st["kpb_k"][k0:k1], st["bvec"][k0:k1], st["wb"],
                        w.api.to_abi_M(M[k0:k1]), device=rank, k_begin=k0, nk=k1 - k0)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        # the library's own peer-memory collectives, except one case that keeps the caller-supplied NCCL all-reduce
        comm = nb != 16
        shared = w.sharding.SharedGauges(ctx, dist, rank, rank, world, nk, nb, nw, bounds, comm=comm)
        if not comm:
            ctx.set_allreduce(w.sharding.torch_allreduce(dist, dev))
        if len(case) > 3:      # disentangle: (X, Y) slab, frozen bands of the slab
            from oracle import spread_oracle as so
            frozen = w.synthetic.make_frozen(nk, nb, case[4])
            ctx.set_frozen(frozen[k0:k1])
            X, Y = so.U_to_X_Y(U, frozen)
            Uslab = np.ascontiguousarray(so.X_Y_to_XY(X, Y)[k0:k1])
            f, it, nfg = ctx.disentangle(Uslab, 1e-10, 1e-7, 10, 3)
        else:
            Uslab = w.api.to_abi_U(U[k0:k1])
            f, it, nfg = ctx.max_localize(Uslab, f_tol=1e-10, g_tol=1e-7, max_iter=10)
        torch.cuda.synchronize()
        dist.barrier()
        out[rank] = (f, it, nfg, Uslab.copy())
        ctx.set_allreduce(None)
        shared.close()
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", [("bcc", (4, 2, 2), 16), ("cubic", (4, 2, 2), 64)])
def test_sharded_max_localize_matches_single_gpu(w, case):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import spread_oracle as so
    lat, grid, nb = case
    st, M, U = w.synthetic.make_model_arrays(lat, grid, nb, nb, eps=0.02)
    nk, nn = st["kpb_k"].shape
    ctx = w.Context(nb, nb, nk, nn, st["kpb_k"], st["bvec"], st["wb"], w.api.to_abi_M(M))
    U1 = w.api.to_abi_U(U)
    f1, it1, nfg1 = ctx.max_localize(U1, f_tol=1e-10, g_tol=1e-7, max_iter=10)
    ctx.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), case, out), nprocs=2, join=True)
    (fa, ita, nfga, Ua), (fb, itb, nfgb, Ub) = out[0], out[1]
    assert fa == fb and ita == itb and nfga == nfgb            # same control flow on every rank
    assert ita == it1 and abs(fa - f1) < 1e-9 * abs(f1)        # same trajectory as one GPU
    Ush = np.concatenate([Ua, Ub], axis=0)
    assert np.abs(Ush - U1).max() < 1e-8
    f_check = so.omega(M, st["kpb_k"], st["bvec"], st["wb"], w.api.from_abi_U(Ush))["Omega"]
    assert abs(f_check - fa) < 1e-9 * abs(fa)


def test_sharded_disentangle_matches_single_gpu(w):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import spread_oracle as so
    case = ("fcc", (4, 2, 2), 18, 9, 5)
    lat, grid, nb, nw, nfroz = case
    st, M, U = w.synthetic.make_model_arrays(lat, grid, nb, nw, eps=0.02)
    nk, nn = st["kpb_k"].shape
    frozen = w.synthetic.make_frozen(nk, nb, nfroz)
    ctx = w.Context(nb, nw, nk, nn, st["kpb_k"], st["bvec"], st["wb"], w.api.to_abi_M(M))
    ctx.set_frozen(frozen)
    X, Y = so.U_to_X_Y(U, frozen)
    XY1 = np.ascontiguousarray(so.X_Y_to_XY(X, Y))
    f1, it1, nfg1 = ctx.disentangle(XY1, 1e-10, 1e-7, 10, 3)
    ctx.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), case, out), nprocs=2, join=True)
    (fa, ita, nfga, Xa), (fb, itb, nfgb, Xb) = out[0], out[1]
    assert fa == fb and ita == itb and nfga == nfgb
    assert ita == it1 and abs(fa - f1) < 1e-9 * abs(f1)
    XYsh = np.concatenate([Xa, Xb], axis=0)
    assert np.abs(XYsh - XY1).max() < 1e-8
    Xs, Ys = so.XY_to_X_Y(XYsh, nb, nw)
    f_check = so.omega(M, st["kpb_k"], st["bvec"], st["wb"], so.X_Y_to_U(Xs, Ys))["Omega"]
    assert abs(f_check - fa) < 1e-9 * abs(fa)


def test_multi_device_context_two_gpus(w):
    """wb200_multi_* with two DISTINCT devices in one process (peer access over NVLink, no IPC): fg, spread
    and the k-sharded max_localize against the single-device results."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import spread_oracle as so
    st, M, U = w.synthetic.make_model_arrays("cubic", (4, 2, 2), 72, 72, eps=0.02)
    nk, nn = st["kpb_k"].shape
    m = w.MultiContext(72, 72, nk, nn, st["kpb_k"], st["bvec"], st["wb"], w.api.to_abi_M(M), devices=[0, 1])
    G = np.empty((nk, 72, 72), dtype=np.complex128)
    f = m.fg(w.api.to_abi_U(U), G=G)
    f_ref, G_ref = so.fg_maxloc(M, st["kpb_k"], st["bvec"], st["wb"], U)
    assert abs(f - f_ref) < 1e-10 * abs(f_ref)
    assert np.abs(G.transpose(0, 2, 1) - G_ref).max() < 1e-10 * np.abs(G_ref).max()
    U2 = w.api.to_abi_U(U)
    f2, it2, nfg2 = m.max_localize(U2, f_tol=1e-10, g_tol=1e-7, max_iter=5)
    m.close()
    ctx = w.Context(72, 72, nk, nn, st["kpb_k"], st["bvec"], st["wb"], w.api.to_abi_M(M))
    U1 = w.api.to_abi_U(U)
    f1, it1, nfg1 = ctx.max_localize(U1, f_tol=1e-10, g_tol=1e-7, max_iter=5)
    ctx.close()
    assert (it1, nfg1) == (it2, nfg2) and abs(f1 - f2) < 1e-9 * abs(f1) and np.abs(U1 - U2).max() < 1e-8
