"""k-point sharding across processes (one process per GPU): slab bounds and the gauge halo exchange.

The reference is single-process; what shards is the (k, b) pair loop of compute_MU_UtMU! /
omega! / omega_grad! (src/spread.jl:164-195, 254-360, 398-481).  Rank r owns a contiguous
k-slab: its overlaps M, its gradient G and its rows of U never move.  Per evaluation:
  1. halo exchange: rank r receives U_{k+b} for the neighbours of its slab that other ranks own
     (`HaloPlan.exchange`, one NCCL all-to-all with per-peer index lists), or an all-gather of
     the whole gauge when slabs are thin;
  2. wb200_fg_phase1_dev on the slab -> 4 nw + 2 partial sums; all-reduce(sum) over ranks;
  3. wb200_fg_phase2_dev -> spread scalars (identical on every rank) and the slab's gradient.
Pure host logic (torch.distributed); works with the gloo backend on CPU tensors for testing.
"""
from __future__ import annotations

import numpy as np


def slab_bounds(nk, world):
    """[k_begin of rank r] + [nk]: contiguous, as even as possible (k index = (i*n2 + j)*n3 + l, so
    slabs are planes of the slowest grid axis when world divides n1)."""
    return [(nk * r) // world for r in range(world + 1)]


def owner_of(k, bounds):
    return np.searchsorted(np.asarray(bounds[1:]), k, side="right")


class HaloPlan:
    """Index lists for the neighbour-gauge exchange, computed identically on every rank from the
    global kpb_k table [nk, nn] (it is tiny: nk * nn int32)."""

    def __init__(self, kpb_k, bounds, rank):
        self.rank = rank
        self.world = len(bounds) - 1
        self.bounds = list(bounds)
        kpb_k = np.asarray(kpb_k)
        need_from = []   # need_from[dst][src] = sorted k owned by src that dst reads
        for dst in range(self.world):
            k0, k1 = bounds[dst], bounds[dst + 1]
            nb_k = np.unique(kpb_k[k0:k1])
            own = owner_of(nb_k, bounds)
            need_from.append([nb_k[(own == src)] if src != dst else nb_k[:0] for src in range(self.world)])
        self.recv_idx = [np.asarray(a, dtype=np.int64) for a in need_from[rank]]
        self.send_idx = [np.asarray(need_from[dst][rank], dtype=np.int64) for dst in range(self.world)]
        self.recv_counts = [len(a) for a in self.recv_idx]
        self.send_counts = [len(a) for a in self.send_idx]
        self.n_halo = int(sum(self.recv_counts))
        # contiguous fast path: split every per-peer list into maximal runs of consecutive k; when
        # the lists are few long runs (slabs of a regular grid) the rows are sent and received in
        # place with point-to-point copies, with no gather / scatter kernels at all
        self.send_runs = [self._runs(a) for a in self.send_idx]
        self.recv_runs = [self._runs(a) for a in self.recv_idx]
        n_runs = sum(len(r) for r in self.send_runs) + sum(len(r) for r in self.recv_runs)
        self.contiguous = 0 < n_runs <= 8 * max(self.world - 1, 1)

    @staticmethod
    def _runs(idx):
        """[(start, stop)] of maximal consecutive runs of a sorted index array"""
        if len(idx) == 0:
            return []
        brk = np.nonzero(np.diff(idx) != 1)[0]
        starts = np.concatenate([[idx[0]], idx[brk + 1]])
        stops = np.concatenate([idx[brk] + 1, [idx[-1] + 1]])
        return [(int(a), int(b)) for a, b in zip(starts, stops)]

    def exchange(self, U, dist, scratch=None):
        """U: tensor [nk_total, ...] whose own slab is current; fills the halo rows in place.
        `dist` is torch.distributed (passed in so this module imports without torch)."""
        import torch
        if self.contiguous and dist.get_backend() != "gloo":
            ops = []
            Ur = torch.view_as_real(U) if U.is_complex() else U
            for peer in range(self.world):          # same order on both sides of every pair
                for a, b in self.recv_runs[peer]:
                    ops.append(dist.P2POp(dist.irecv, Ur[a:b], peer))
                for a, b in self.send_runs[peer]:
                    ops.append(dist.P2POp(dist.isend, Ur[a:b], peer))
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            return
        per = U[0].numel()
        flat = U.reshape(U.shape[0], per)
        send_rows = np.concatenate(self.send_idx) if self.world > 1 else np.zeros(0, dtype=np.int64)
        recv_rows = np.concatenate(self.recv_idx) if self.world > 1 else np.zeros(0, dtype=np.int64)
        if len(send_rows) == 0 and len(recv_rows) == 0:
            return
        key = str(U.device)
        if getattr(self, "_dev_key", None) != key:      # index tensors live on the device, built once
            self._si = torch.as_tensor(send_rows, device=U.device)
            self._ri = torch.as_tensor(recv_rows, device=U.device)
            self._dev_key = key
        si, ri = self._si, self._ri
        send = flat.index_select(0, si).contiguous()
        recv = torch.empty((len(recv_rows), per), dtype=U.dtype, device=U.device)
        # complex tensors travel as real pairs (not every backend has complex collectives)
        s_r = torch.view_as_real(send) if send.is_complex() else send
        r_r = torch.view_as_real(recv) if recv.is_complex() else recv
        dist.all_to_all_single(r_r, s_r, output_split_sizes=self.recv_counts,
                               input_split_sizes=self.send_counts)
        flat.index_copy_(0, ri, recv)


class SharedGauges:
    """Peer-memory plumbing of a one-process-per-GPU run: every rank allocates the FULL gauge array and a comm
    buffer (readiness flags + all-reduce slots) in CUDA-IPC shareable memory and exports its context's packed
    gauge image; the handles travel through torch.distributed once, every rank maps its peers' allocations and
    hands the pointer tables to its slab context (wb200_set_peer_gauges / _comm / _packed).  After that the
    evaluations and the optimiser need no torch.distributed / NCCL call at all.  `ptr` is this rank's gauge
    array; `close()` unmaps and frees.  comm=False keeps the round-1 scheme (caller-supplied all-reduce)."""

    def __init__(self, ctx, dist, device_index, rank, world, nk, nb, nw, bounds, comm=True):
        import torch
        from . import lib
        self._lib, self.ctx, self.dist = lib, ctx, dist
        self.device, self.rank, self.world = device_index, rank, world
        self.ptr, handle = lib.ipc_alloc(device_index, 16 * nk * nb * nw)
        self.comm_ptr, chandle, phandle = None, None, None
        if comm:
            nbytes = lib.peer_comm_bytes(nw)
            self.comm_ptr, chandle = lib.ipc_alloc(device_index, nbytes)
            torch.as_tensor(lib.DeviceArray(self.comm_ptr, (nbytes // 8,), "<i8"),
                            device=torch.device("cuda", device_index)).zero_()
            torch.cuda.synchronize()
            try:
                phandle = lib.ipc_get_handle(device_index, ctx.packed_gauge()[0])
            except lib.WB200Error:
                phandle = None                      # shape outside the tensor path: raw rows are pulled instead
        handles = [None] * world
        dist.all_gather_object(handles, (handle, chandle, phandle))     # also orders the zeroing before any signal
        opened = lambda r, i: lib.ipc_open(device_index, handles[r][i])
        self.ptrs = [self.ptr if r == rank else opened(r, 0) for r in range(world)]
        ctx.set_peer_gauges(self.ptrs, bounds)
        self.comm_ptrs = self.packed_ptrs = None
        if comm:
            self.comm_ptrs = [self.comm_ptr if r == rank else opened(r, 1) for r in range(world)]
            ctx.set_peer_comm(self.comm_ptrs)
            if all(h[2] is not None for h in handles):
                self.packed_ptrs = [ctx.packed_gauge()[0] if r == rank else opened(r, 2) for r in range(world)]
                ctx.set_peer_packed(self.packed_ptrs)

    def close(self):
        import torch
        lib = self._lib
        torch.cuda.synchronize()
        self.dist.barrier()                      # nobody still reads a peer's memory
        self.ctx.set_peer_packed(None)
        self.ctx.set_peer_comm(None)
        self.ctx.set_peer_gauges(None, None)
        for r in range(self.world):
            if r != self.rank:
                lib.ipc_close(self.device, self.ptrs[r])
                if self.comm_ptrs:
                    lib.ipc_close(self.device, self.comm_ptrs[r])
                if self.packed_ptrs:
                    lib.ipc_close(self.device, self.packed_ptrs[r])
        self.dist.barrier()
        lib.ipc_free(self.device, self.ptr)
        if self.comm_ptr:
            lib.ipc_free(self.device, self.comm_ptr)


def torch_allreduce(dist, device):
    """The callback wb200_set_allreduce wants, on top of torch.distributed (NCCL): wraps the raw
    device pointer as a tensor and reduces it in place on torch's current stream (the stream the
    ctx must have been given with set_stream)."""
    import torch
    from .lib import DeviceArray

    def fn(dptr, count, op, stream):
        t = torch.as_tensor(DeviceArray(dptr, (count,), "<f8"), device=device)
        dist.all_reduce(t, op=(dist.ReduceOp.SUM, dist.ReduceOp.MAX, dist.ReduceOp.MIN)[op])

    return fn
