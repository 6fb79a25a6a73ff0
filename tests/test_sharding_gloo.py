"""Multi-process (gloo, world_size 2, CPU) test of the N > 1 host logic: slab bounds, halo
exchange of neighbour gauges, all-reduce of the spread partial sums.  The per-slab arithmetic is
the oracle's here (no GPU); on the GPU box the same plan drives wb200_fg_phase1/2_dev."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import spread_oracle as so


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import __graft_entry__ as ge
    w = ge.load_package()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        st, M, U = w.synthetic.make_model_arrays("bcc", (4, 3, 3), 6, 5)
        nk, nn = st["kpb_k"].shape
        bounds = w.sharding.slab_bounds(nk, world)
        k0, k1 = bounds[rank], bounds[rank + 1]
        plan = w.sharding.HaloPlan(st["kpb_k"], bounds, rank)
        # this rank only holds its own rows of U (others poisoned), then exchanges the halo
        Ut = torch.full((nk, 6, 5), complex(np.nan, np.nan), dtype=torch.complex128)
        Ut[k0:k1] = torch.from_numpy(U[k0:k1])
        plan.exchange(Ut, dist)
        need = np.unique(np.concatenate([np.arange(k0, k1), st["kpb_k"][k0:k1].ravel()]))
        got = Ut.numpy()
        assert np.array_equal(got[need], U[need])
        assert plan.n_halo == len(need) - (k1 - k0)
        # phase 1 on the slab, all-reduce, phase 2
        MU, N = so.compute_MU_UtMU_gathered(M[k0:k1], got[k0:k1], got[st["kpb_k"][k0:k1]])
        ps = so.partial_sums_from_N(N, st["bvec"][k0:k1], st["wb"])
        vec = torch.from_numpy(np.concatenate([ps["wbphi"].ravel(), ps["r2"], [ps["OI"], ps["OOD"]]]))
        dist.all_reduce(vec)
        v = vec.numpy()
        r = -v[:15].reshape(5, 3) / nk
        omega = v[15:20] / nk - (r ** 2).sum(axis=1)
        G = so.grad_from_MU_N_r(MU, N, st["bvec"][k0:k1], st["wb"], r, nk)
        f_ref, G_ref = so.fg_maxloc(M, st["kpb_k"], st["bvec"], st["wb"], U)
        sp = so.omega(M, st["kpb_k"], st["bvec"], st["wb"], U)
        assert abs(omega.sum() - f_ref) < 1e-12 * abs(f_ref)
        assert abs(v[20] / nk - sp["OmegaI"]) < 1e-12 and abs(v[21] / nk - sp["OmegaOD"]) < 1e-12
        assert np.abs(G - G_ref[k0:k1]).max() < 1e-13
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_with_gloo(w):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_slab_bounds_and_plan_symmetry(w):
    st = w.synthetic.make_stencil(w.synthetic.LATTICES["bcc"](), (16, 16, 16))
    for world in (1, 2, 4, 8):
        b = w.sharding.slab_bounds(4096, world)
        assert b[0] == 0 and b[-1] == 4096 and all(b[i + 1] - b[i] == 4096 // world for i in range(world))
        plans = [w.sharding.HaloPlan(st["kpb_k"], b, r) for r in range(world)]
        for r in range(world):
            for q in range(world):   # what r sends to q is what q expects from r
                assert np.array_equal(plans[r].send_idx[q], plans[q].recv_idx[r])
        if world == 8:   # two planes per rank, one halo plane on each side: 2 x 256 k-points
            assert all(p.n_halo == 512 for p in plans)
        if world == 1:
            assert plans[0].n_halo == 0
