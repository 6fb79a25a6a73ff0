"""ctypes binding of libwannier_b200.so (the C-ABI declared in include/wannier_b200.h).

This is the same thin layer a Julia `ccall` wrapper would be: plain
pointers and sizes in, status codes out.  There is NO fallback of any kind: if the shared library
is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WB200_LIB") or os.path.join(_HERE, "lib", "libwannier_b200.so")

# every symbol include/wannier_b200.h declares (tests check that the .so exports all of them)
SYMBOLS = [
    "wb200_create", "wb200_destroy", "wb200_last_error", "wb200_version", "wb200_set_stream",
    "wb200_synchronize", "wb200_rotation_kernel", "wb200_launch_count", "wb200_set_penalty",
    "wb200_set_frozen", "wb200_fg", "wb200_spread", "wb200_center", "wb200_rotate_overlaps",
    "wb200_fg_xy", "wb200_project_tangent", "wb200_retract", "wb200_max_localize",
    "wb200_disentangle", "wb200_nsums", "wb200_nres", "wb200_fg_phase1_dev", "wb200_fg_phase2_dev",
    "wb200_fg_dev", "wb200_min_abs_diag", "wb200_project_tangent_dev", "wb200_retract_dev",
    "wb200_profile", "wb200_profile_read", "wb200_probe_fp64_peak", "wb200_trim_pools", "wb200_pool_stats",
    "wb200_omega_local", "wb200_fg_rotate", "wb200_opt_rotate", "wb200_position_op", "wb200_berry_connection",
    "wb200_set_allreduce",
    "wb200_ipc_alloc", "wb200_ipc_open", "wb200_ipc_close", "wb200_ipc_free", "wb200_set_peer_gauges",
    "wb200_peer_comm_bytes", "wb200_set_peer_comm", "wb200_peer_allreduce_dev", "wb200_packed_gauge",
    "wb200_ipc_get_handle", "wb200_set_peer_packed",
    "wb200_multi_create", "wb200_multi_destroy", "wb200_multi_last_error", "wb200_multi_ndev",
    "wb200_multi_slab_begin", "wb200_multi_launch_count", "wb200_multi_set_penalty", "wb200_multi_set_frozen",
    "wb200_multi_fg", "wb200_multi_spread", "wb200_multi_fg_xy", "wb200_multi_max_localize",
    "wb200_multi_disentangle",
    "wb200_u_to_xy", "wb200_orthonorm_freeze", "wb200_create_stiefel",
    "wb200_mmn_header", "wb200_mmn_read", "wb200_amn_header", "wb200_amn_read", "wb200_eig_read",
    "wb200_io_last_error",
    "wb200_host_register", "wb200_host_unregister",
    "wb200_mag_create", "wb200_mag_destroy", "wb200_mag_last_error", "wb200_mag_overlap",
    "wb200_mag_omega_updn_grad", "wb200_mag_fg_xy", "wb200_mag_disentangle",
]

FLAG_DEFAULT = 0
FLAG_NO_TENSOR = 1
FLAG_FORCE_TENSOR = 2
FLAG_NO_FUSED_TRIAL = 4
FLAG_M_ROW_MAJOR = 8


class WB200Error(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"libwannier_b200 status {status}: {msg}")
        self.status = status


_lib = None


def load():
    """dlopen the library (once).  Raises if it has not been built: the product path never
    degrades to a CPU implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
            " (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, dbl = C.c_void_p, C.c_int, C.c_double
    pd = C.c_void_p  # all array arguments are passed as raw addresses
    lib.wb200_create.argtypes = [C.POINTER(vp), i32, i32, i32, i32, i32, i32, i32, pd, pd, pd, pd,
                                 C.c_uint]
    lib.wb200_create.restype = i32
    lib.wb200_destroy.argtypes = [vp]
    lib.wb200_destroy.restype = None
    lib.wb200_last_error.argtypes = [vp]
    lib.wb200_last_error.restype = C.c_char_p
    lib.wb200_version.argtypes = []
    lib.wb200_version.restype = C.c_char_p
    lib.wb200_set_stream.argtypes = [vp, vp]
    lib.wb200_synchronize.argtypes = [vp]
    lib.wb200_rotation_kernel.argtypes = [vp]
    lib.wb200_launch_count.argtypes = [vp]
    lib.wb200_launch_count.restype = C.c_longlong
    lib.wb200_set_penalty.argtypes = [vp, pd, dbl]
    lib.wb200_set_frozen.argtypes = [vp, pd]
    lib.wb200_fg.argtypes = [vp, pd, pd, pd]
    lib.wb200_spread.argtypes = [vp, pd, pd, pd, pd, pd]
    lib.wb200_center.argtypes = [vp, pd, pd]
    lib.wb200_rotate_overlaps.argtypes = [vp, pd, pd, pd]
    lib.wb200_fg_xy.argtypes = [vp, pd, pd, pd]
    lib.wb200_project_tangent.argtypes = [vp, pd, pd, i32, i32, i32]
    lib.wb200_retract.argtypes = [vp, pd, i32, i32, i32]
    lib.wb200_max_localize.argtypes = [vp, pd, dbl, dbl, i32, i32, pd, pd, pd]
    lib.wb200_disentangle.argtypes = [vp, pd, dbl, dbl, i32, i32, pd, pd, pd]
    lib.wb200_nsums.argtypes = [vp]
    lib.wb200_nres.argtypes = [vp]
    lib.wb200_fg_phase1_dev.argtypes = [vp, pd, pd, i32]
    lib.wb200_fg_phase2_dev.argtypes = [vp, pd, pd, pd]
    lib.wb200_fg_dev.argtypes = [vp, pd, pd, pd, i32]
    lib.wb200_min_abs_diag.argtypes = [vp, pd]
    lib.wb200_project_tangent_dev.argtypes = [vp, pd, pd, i32, i32, i32]
    lib.wb200_retract_dev.argtypes = [vp, pd, i32, i32, i32, pd]
    lib.wb200_omega_local.argtypes = [vp, pd, pd]
    lib.wb200_fg_rotate.argtypes = [vp, pd, pd, pd]
    lib.wb200_opt_rotate.argtypes = [vp, pd, dbl, dbl, i32, i32, pd, pd, pd]
    lib.wb200_position_op.argtypes = [vp, pd, pd]
    lib.wb200_berry_connection.argtypes = [vp, pd, pd, i32, i32]
    lib.wb200_set_allreduce.argtypes = [vp, vp, vp]
    lib.wb200_ipc_alloc.argtypes = [i32, C.c_size_t, C.POINTER(vp), pd]
    lib.wb200_ipc_open.argtypes = [i32, pd, C.POINTER(vp)]
    lib.wb200_ipc_close.argtypes = [i32, vp]
    lib.wb200_ipc_free.argtypes = [i32, vp]
    lib.wb200_set_peer_gauges.argtypes = [vp, i32, pd, pd]
    lib.wb200_peer_comm_bytes.argtypes = [i32]
    lib.wb200_peer_comm_bytes.restype = C.c_size_t
    lib.wb200_set_peer_comm.argtypes = [vp, i32, pd]
    lib.wb200_peer_allreduce_dev.argtypes = [vp, pd, i32, i32]
    lib.wb200_packed_gauge.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    lib.wb200_ipc_get_handle.argtypes = [i32, vp, pd]
    lib.wb200_set_peer_packed.argtypes = [vp, i32, pd]
    lib.wb200_multi_create.argtypes = [C.POINTER(vp), pd, i32, i32, i32, i32, i32, pd, pd, pd, pd, C.c_uint]
    lib.wb200_multi_destroy.argtypes = [vp]
    lib.wb200_multi_destroy.restype = None
    lib.wb200_multi_last_error.argtypes = [vp]
    lib.wb200_multi_last_error.restype = C.c_char_p
    lib.wb200_multi_ndev.argtypes = [vp]
    lib.wb200_multi_slab_begin.argtypes = [vp, i32]
    lib.wb200_multi_launch_count.argtypes = [vp]
    lib.wb200_multi_launch_count.restype = C.c_longlong
    lib.wb200_multi_set_penalty.argtypes = [vp, pd, dbl]
    lib.wb200_multi_set_frozen.argtypes = [vp, pd]
    lib.wb200_multi_fg.argtypes = [vp, pd, pd, pd]
    lib.wb200_multi_spread.argtypes = [vp, pd, pd, pd, pd, pd]
    lib.wb200_multi_fg_xy.argtypes = [vp, pd, pd, pd]
    lib.wb200_multi_max_localize.argtypes = [vp, pd, dbl, dbl, i32, i32, pd, pd, pd]
    lib.wb200_multi_disentangle.argtypes = [vp, pd, dbl, dbl, i32, i32, pd, pd, pd]
    lib.wb200_u_to_xy.argtypes = [vp, pd, pd]
    lib.wb200_orthonorm_freeze.argtypes = [vp, pd]
    lib.wb200_create_stiefel.argtypes = [C.POINTER(vp), i32, i32, i32, i32]
    lib.wb200_mmn_header.argtypes = [C.c_char_p, pd, pd, pd]
    lib.wb200_mmn_read.argtypes = [C.c_char_p, i32, i32, i32, pd, pd, pd, i32]
    lib.wb200_amn_header.argtypes = [C.c_char_p, pd, pd, pd]
    lib.wb200_amn_read.argtypes = [C.c_char_p, i32, i32, i32, pd]
    lib.wb200_eig_read.argtypes = [C.c_char_p, i32, i32, pd]
    lib.wb200_io_last_error.argtypes = []
    lib.wb200_io_last_error.restype = C.c_char_p
    lib.wb200_host_register.argtypes = [pd, C.c_size_t]
    lib.wb200_host_unregister.argtypes = [pd]
    lib.wb200_mag_create.argtypes = [C.POINTER(vp), vp, vp, pd]
    lib.wb200_mag_destroy.argtypes = [vp]
    lib.wb200_mag_destroy.restype = None
    lib.wb200_mag_last_error.argtypes = [vp]
    lib.wb200_mag_last_error.restype = C.c_char_p
    lib.wb200_mag_overlap.argtypes = [vp, pd, pd, pd, pd]
    lib.wb200_mag_omega_updn_grad.argtypes = [vp, pd, pd, pd, pd]
    lib.wb200_mag_fg_xy.argtypes = [vp, pd, dbl, pd, pd, pd]
    lib.wb200_mag_disentangle.argtypes = [vp, pd, dbl, dbl, dbl, i32, i32, pd, pd, pd]
    lib.wb200_profile.argtypes = [vp, i32]
    lib.wb200_profile_read.argtypes = [vp, pd, pd]
    lib.wb200_probe_fp64_peak.argtypes = [i32, pd, pd]
    lib.wb200_trim_pools.argtypes = []
    lib.wb200_trim_pools.restype = None
    lib.wb200_pool_stats.argtypes = [i32, pd, pd]
    lib.wb200_pool_stats.restype = C.c_longlong
    for name in SYMBOLS:
        f = getattr(lib, name)
        if name not in ("wb200_destroy", "wb200_last_error", "wb200_version", "wb200_launch_count", "wb200_trim_pools",
                        "wb200_pool_stats",
                        "wb200_multi_destroy", "wb200_multi_last_error", "wb200_multi_launch_count",
                        "wb200_mag_destroy", "wb200_mag_last_error", "wb200_io_last_error"):
            f.restype = i32
    _lib = lib
    return lib


def _addr(a):
    """address of a numpy array / int pointer / None"""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return int(a)
    return a.ctypes.data


def _check_array(a, dtype, size, name):
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags["C_CONTIGUOUS"]:
        raise TypeError(f"{name}: need a C-contiguous numpy array of {dtype}")
    if a.size != size:
        raise ValueError(f"{name}: expected {size} elements, got {a.size}")


class Context:
    """Owns one wb200_ctx (one k-slab on one B200).  Arrays are passed in the C-ABI's memory
    layout (= Julia's): U [nk][nw][nb] complex128, M [nk][nn][nb(l)][nb(m)], G like U."""

    def __init__(self, nb, nw, nk_total, nn, kpb_k, bvec, bweights, M, *, device=0, k_begin=0,
                 nk=None, flags=FLAG_DEFAULT):
        self._h = None
        lib = load()
        nk = nk_total if nk is None else nk
        kpb_k = np.ascontiguousarray(kpb_k, dtype=np.int32)
        bvec = np.ascontiguousarray(bvec, dtype=np.float64)
        bweights = np.ascontiguousarray(bweights, dtype=np.float64)
        _check_array(kpb_k, np.int32, nk * nn, "kpb_k")
        _check_array(bvec, np.float64, nk * nn * 3, "bvec")
        _check_array(bweights, np.float64, nn, "bweights")
        if not isinstance(M, (int, np.integer)):      # an int is a raw (device) address
            _check_array(M, np.complex128, nk * nn * nb * nb, "M")
        self.nb, self.nw, self.nk_total, self.nn, self.nk, self.k_begin = nb, nw, nk_total, nn, nk, k_begin
        self.device = device
        h = C.c_void_p()
        st = lib.wb200_create(C.byref(h), device, nb, nw, nk_total, k_begin, nk, nn, _addr(kpb_k),
                              _addr(bvec), _addr(bweights), _addr(M), flags)
        if st != 0:
            raise WB200Error(st, lib.wb200_last_error(None).decode())
        self._h = h
        self._lib = lib

    # -- plumbing ---------------------------------------------------------------------------
    def close(self):
        if self._h is not None:
            self._lib.wb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        cb_error = getattr(self, "_cb_error", None)
        if cb_error is not None:
            self._cb_error = None
            raise cb_error
        if st != 0:
            raise WB200Error(st, self._lib.wb200_last_error(self._h).decode())

    @property
    def rotation_kernel(self):
        return ("cuda-core-fp64", "dmma-fp64")[self._lib.wb200_rotation_kernel(self._h)]

    @property
    def launch_count(self):
        return int(self._lib.wb200_launch_count(self._h))

    @property
    def nsums(self):
        return self._lib.wb200_nsums(self._h)

    @property
    def nres(self):
        return self._lib.wb200_nres(self._h)

    def set_stream(self, stream_ptr):
        self._ck(self._lib.wb200_set_stream(self._h, stream_ptr))

    def synchronize(self):
        self._ck(self._lib.wb200_synchronize(self._h))

    def set_penalty(self, r0, lam):
        if r0 is None:
            self._ck(self._lib.wb200_set_penalty(self._h, None, 0.0))
            return
        r0 = np.ascontiguousarray(r0, dtype=np.float64)
        _check_array(r0, np.float64, 3 * self.nw, "r0")
        self._ck(self._lib.wb200_set_penalty(self._h, _addr(r0), float(lam)))

    def set_frozen(self, frozen):
        if frozen is None:
            self._ck(self._lib.wb200_set_frozen(self._h, None))
            return
        fr = np.ascontiguousarray(frozen, dtype=np.uint8)
        _check_array(fr, np.uint8, self.nk * self.nb, "frozen")
        self._ck(self._lib.wb200_set_frozen(self._h, _addr(fr)))

    def set_peer_gauges(self, ptrs, bounds):
        """ptrs[r]: device address of rank r's full gauge array as mapped on this device; None resets."""
        if ptrs is None:
            self._ck(self._lib.wb200_set_peer_gauges(self._h, 0, None, None))
            return
        p = np.asarray(ptrs, dtype=np.uint64)
        b = np.ascontiguousarray(bounds, dtype=np.int32)
        self._ck(self._lib.wb200_set_peer_gauges(self._h, len(p), _addr(p), _addr(b)))

    def set_peer_comm(self, ptrs):
        """ptrs[r]: device address of rank r's comm buffer (peer_comm_bytes(nw) bytes, zeroed) as mapped on this
        device; None removes.  Readiness flags replace the "rows are in place" all-reduce and, unless a callback
        is registered with set_allreduce, the library's own peer-memory all-reduce serves every collective."""
        if ptrs is None:
            self._ck(self._lib.wb200_set_peer_comm(self._h, 0, None))
            self._comm = False
            self._sharded = getattr(self, "_cb", None) is not None
            return
        p = np.asarray(ptrs, dtype=np.uint64)
        self._ck(self._lib.wb200_set_peer_comm(self._h, len(p), _addr(p)))
        self._comm = True
        self._sharded = True

    def set_peer_packed(self, ptrs):
        """ptrs[r]: device address of rank r's packed-gauge image (packed_gauge()) as mapped here; None removes."""
        if ptrs is None:
            self._ck(self._lib.wb200_set_peer_packed(self._h, 0, None))
            return
        p = np.asarray(ptrs, dtype=np.uint64)
        self._ck(self._lib.wb200_set_peer_packed(self._h, len(p), _addr(p)))

    def packed_gauge(self):
        """(device address, bytes) of this ctx's packed-gauge image; raises for shapes outside the tensor path."""
        ptr, n = C.c_void_p(), C.c_size_t()
        self._ck(self._lib.wb200_packed_gauge(self._h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def peer_allreduce_dev(self, dptr, count, op=0):
        self._ck(self._lib.wb200_peer_allreduce_dev(self._h, dptr, count, op))

    def set_allreduce(self, fn):
        """fn(dptr: int, count: int, op: int (0 sum, 1 max), stream: int) all-reduces `count` doubles
        in place over the ranks on `stream`; None unregisters.  The ctypes thunk is kept alive here."""
        if fn is None:
            self._cb = None
            self._sharded = getattr(self, "_comm", False)    # the library's own peer-memory collective remains
            self._ck(self._lib.wb200_set_allreduce(self._h, None, None))
            return
        self._sharded = True
        proto = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p)

        def thunk(dptr, count, op, stream, user):
            # an exception must not be swallowed by ctypes: latch it, return a failure status (the C
            # side then aborts the call with WB200_ERR_CUDA) and re-raise it from _ck
            try:
                fn(dptr or 0, count, op, stream or 0)
                return 0
            except BaseException as e:      # noqa: BLE001
                self._cb_error = e
                return 1

        self._cb = proto(thunk)
        self._ck(self._lib.wb200_set_allreduce(self._h, C.cast(self._cb, C.c_void_p), None))

    def profile(self, enable):
        self._ck(self._lib.wb200_profile(self._h, int(enable)))

    def profile_read(self):
        """{class: (ms, launches)} since the last read; classes: pack, rotate, reduce, grad"""
        ms = np.zeros(4)
        cnt = np.zeros(4, dtype=np.int32)
        self._ck(self._lib.wb200_profile_read(self._h, _addr(ms), _addr(cnt)))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(("pack", "rotate", "reduce", "grad"))}

    # -- host entry points ------------------------------------------------------------------
    def _u(self, U):
        # single-slab ctx: the whole model's gauges; k-slab ctx of a sharded run (set_allreduce +
        # set_peer_gauges): the slab's rows
        n = (self.nk if getattr(self, "_sharded", False) else self.nk_total) * self.nb * self.nw
        _check_array(U, np.complex128, n, "U")
        return _addr(U)

    def fg(self, U, want_f=True, G=None):
        """fg!(F, G, U): returns f (or None); writes G in place when given."""
        f = C.c_double(0.0)
        if G is not None:
            _check_array(G, np.complex128, self.nk * self.nb * self.nw, "G")
        self._ck(self._lib.wb200_fg(self._h, self._u(U), C.addressof(f) if want_f else None, _addr(G)))
        return f.value if want_f else None

    def spread(self, U):
        out5 = np.zeros(5)
        om = np.zeros(self.nw)
        r = np.zeros((self.nw, 3))
        mn = C.c_double(0.0)
        self._ck(self._lib.wb200_spread(self._h, self._u(U), _addr(out5), _addr(om), _addr(r),
                                        C.addressof(mn)))
        return out5, om, r, mn.value

    def center(self, U):
        r = np.zeros((self.nw, 3))
        self._ck(self._lib.wb200_center(self._h, self._u(U), _addr(r)))
        return r

    def rotate_overlaps(self, U, want_N=True, want_MU=False):
        N = np.empty((self.nk, self.nn, self.nw, self.nw), dtype=np.complex128) if want_N else None
        MU = np.empty((self.nk, self.nn, self.nw, self.nb), dtype=np.complex128) if want_MU else None
        self._ck(self._lib.wb200_rotate_overlaps(self._h, self._u(U), _addr(N), _addr(MU)))
        return N, MU

    def fg_xy(self, XY, want_f=True, G=None):
        n = self.nk * (self.nw * self.nw + self.nb * self.nw)
        _check_array(XY, np.complex128, n, "XY")
        if G is not None:
            _check_array(G, np.complex128, n, "G_XY")
        f = C.c_double(0.0)
        self._ck(self._lib.wb200_fg_xy(self._h, _addr(XY), C.addressof(f) if want_f else None, _addr(G)))
        return f.value if want_f else None

    def project_tangent(self, X, G, nrow, ncol):
        count = X.size // (nrow * ncol)
        _check_array(X, np.complex128, count * nrow * ncol, "X")
        _check_array(G, np.complex128, count * nrow * ncol, "G")
        self._ck(self._lib.wb200_project_tangent(self._h, _addr(X), _addr(G), nrow, ncol, count))

    def retract(self, X, nrow, ncol):
        count = X.size // (nrow * ncol)
        _check_array(X, np.complex128, count * nrow * ncol, "X")
        self._ck(self._lib.wb200_retract(self._h, _addr(X), nrow, ncol, count))

    def u_to_xy(self, U):
        """X_Y_to_XY(U_to_X_Y(U, frozen)) with the ctx's frozen bands -> XY [nk][nw*nw + nb*nw]"""
        _check_array(U, np.complex128, self.nk * self.nb * self.nw, "U")
        XY = np.empty((self.nk, self.nw * self.nw + self.nb * self.nw), dtype=np.complex128)
        self._ck(self._lib.wb200_u_to_xy(self._h, _addr(U), _addr(XY)))
        return XY

    def orthonorm_freeze(self, U):
        """in place on U [nk][nw][nb]"""
        _check_array(U, np.complex128, self.nk * self.nb * self.nw, "U")
        self._ck(self._lib.wb200_orthonorm_freeze(self._h, _addr(U)))

    def max_localize(self, U, f_tol=1e-7, g_tol=1e-5, max_iter=200, history=3):
        # U is the ctx's k-slab (= everything for a single-slab ctx)
        _check_array(U, np.complex128, self.nk * self.nb * self.nw, "U")
        f = C.c_double(0.0)
        it = C.c_int(0)
        nfg = C.c_int(0)
        self._ck(self._lib.wb200_max_localize(self._h, _addr(U), f_tol, g_tol, max_iter, history,
                                              C.addressof(f), C.addressof(it), C.addressof(nfg)))
        return f.value, it.value, nfg.value

    def disentangle(self, XY, f_tol=1e-7, g_tol=1e-5, max_iter=200, history=3):
        n = self.nk * (self.nw * self.nw + self.nb * self.nw)
        _check_array(XY, np.complex128, n, "XY")
        f = C.c_double(0.0)
        it = C.c_int(0)
        nfg = C.c_int(0)
        self._ck(self._lib.wb200_disentangle(self._h, _addr(XY), f_tol, g_tol, max_iter, history,
                                             C.addressof(f), C.addressof(it), C.addressof(nfg)))
        return f.value, it.value, nfg.value

    def omega_local(self, U):
        loc = np.zeros(self.nk)
        self._ck(self._lib.wb200_omega_local(self._h, self._u(U), _addr(loc)))
        return loc

    def fg_rotate(self, W, want_f=True, G=None):
        _check_array(W, np.complex128, self.nw * self.nw, "W")
        if G is not None:
            _check_array(G, np.complex128, self.nw * self.nw, "G")
        f = C.c_double(0.0)
        self._ck(self._lib.wb200_fg_rotate(self._h, _addr(W), C.addressof(f) if want_f else None, _addr(G)))
        return f.value if want_f else None

    def opt_rotate(self, W, f_tol=1e-7, g_tol=1e-5, max_iter=200, history=3):
        _check_array(W, np.complex128, self.nw * self.nw, "W")
        f, it, nfg = C.c_double(0.0), C.c_int(0), C.c_int(0)
        self._ck(self._lib.wb200_opt_rotate(self._h, _addr(W), f_tol, g_tol, max_iter, history,
                                            C.addressof(f), C.addressof(it), C.addressof(nfg)))
        return f.value, it.value, nfg.value

    def position_op(self, U):
        R = np.empty((self.nw, self.nw, 3), dtype=np.complex128)      # [n][m][a]
        self._ck(self._lib.wb200_position_op(self._h, self._u(U), _addr(R)))
        return R

    def berry_connection(self, U, imlog_diag=True, force_hermiticity=True):
        A = np.empty((self.nk, self.nw, self.nw, 3), dtype=np.complex128)   # [k][n][m][a]
        self._ck(self._lib.wb200_berry_connection(self._h, self._u(U), _addr(A), int(imlog_diag),
                                                  int(force_hermiticity)))
        return A

    # -- device-pointer entry points (raw addresses, e.g. torch.Tensor.data_ptr()) -----------
    def fg_phase1_dev(self, dU, dsums, exact_split=False):
        self._ck(self._lib.wb200_fg_phase1_dev(self._h, dU, dsums, int(exact_split)))

    def fg_phase2_dev(self, dsums, dres, dG):
        self._ck(self._lib.wb200_fg_phase2_dev(self._h, dsums, dres, dG))

    def fg_dev(self, dU, dres, dG, exact_split=False):
        self._ck(self._lib.wb200_fg_dev(self._h, dU, dres, dG, int(exact_split)))

    def min_abs_diag(self):
        v = C.c_double(0.0)
        self._ck(self._lib.wb200_min_abs_diag(self._h, C.addressof(v)))
        return v.value

    def project_tangent_dev(self, dX, dG, nrow, ncol, count):
        self._ck(self._lib.wb200_project_tangent_dev(self._h, dX, dG, nrow, ncol, count))

    def retract_dev(self, dX, nrow, ncol, count):
        it = C.c_int(0)
        self._ck(self._lib.wb200_retract_dev(self._h, dX, nrow, ncol, count, C.addressof(it)))
        return it.value


class StiefelContext(Context):
    """wb200_create_stiefel: no stencil, no overlaps — the per-k Stiefel operations and the setup step
    (U_to_X_Y / orthonorm_freeze) for callers that have no model at hand."""

    def __init__(self, nb, nw, nk, *, device=0):
        self._h = None
        lib = load()
        self.nb, self.nw, self.nk_total, self.nn, self.nk, self.k_begin = nb, nw, nk, 0, nk, 0
        self.device = device
        h = C.c_void_p()
        st = lib.wb200_create_stiefel(C.byref(h), device, nb, nw, nk)
        if st != 0:
            raise WB200Error(st, "wb200_create_stiefel failed (bad sizes or no sm_100 device; there is no CPU fallback)")
        self._h = h
        self._lib = lib


class MultiContext:
    """One call, several GPUs (wb200_multi_*): the k-points are sharded in contiguous slabs over `devices`
    (a device may be listed more than once).  Same array conventions and method names as Context; all
    arrays are the full model's."""

    def __init__(self, nb, nw, nk, nn, kpb_k, bvec, bweights, M, *, devices, flags=FLAG_DEFAULT):
        self._h = None
        lib = load()
        kpb_k = np.ascontiguousarray(kpb_k, dtype=np.int32)
        bvec = np.ascontiguousarray(bvec, dtype=np.float64)
        bweights = np.ascontiguousarray(bweights, dtype=np.float64)
        _check_array(kpb_k, np.int32, nk * nn, "kpb_k")
        _check_array(bvec, np.float64, nk * nn * 3, "bvec")
        _check_array(bweights, np.float64, nn, "bweights")
        _check_array(M, np.complex128, nk * nn * nb * nb, "M")
        dv = np.ascontiguousarray(devices, dtype=np.int32)
        self.nb, self.nw, self.nk, self.nk_total, self.nn = nb, nw, nk, nk, nn
        self.devices = [int(d) for d in dv]
        h = C.c_void_p()
        st = lib.wb200_multi_create(C.byref(h), _addr(dv), len(dv), nb, nw, nk, nn, _addr(kpb_k), _addr(bvec),
                                    _addr(bweights), _addr(M), flags)
        if st != 0:
            raise WB200Error(st, lib.wb200_multi_last_error(None).decode())
        self._h = h
        self._lib = lib

    def close(self):
        if self._h is not None:
            self._lib.wb200_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != 0:
            raise WB200Error(st, self._lib.wb200_multi_last_error(self._h).decode())

    @property
    def launch_count(self):
        return int(self._lib.wb200_multi_launch_count(self._h))

    @property
    def bounds(self):
        return [self._lib.wb200_multi_slab_begin(self._h, r) for r in range(len(self.devices) + 1)]

    def set_penalty(self, r0, lam):
        if r0 is None:
            self._ck(self._lib.wb200_multi_set_penalty(self._h, None, 0.0))
            return
        r0 = np.ascontiguousarray(r0, dtype=np.float64)
        _check_array(r0, np.float64, 3 * self.nw, "r0")
        self._ck(self._lib.wb200_multi_set_penalty(self._h, _addr(r0), float(lam)))

    def set_frozen(self, frozen):
        fr = None if frozen is None else np.ascontiguousarray(frozen, dtype=np.uint8)
        if fr is not None:
            _check_array(fr, np.uint8, self.nk * self.nb, "frozen")
        self._ck(self._lib.wb200_multi_set_frozen(self._h, _addr(fr)))

    def fg(self, U, want_f=True, G=None):
        _check_array(U, np.complex128, self.nk * self.nb * self.nw, "U")
        if G is not None:
            _check_array(G, np.complex128, self.nk * self.nb * self.nw, "G")
        f = C.c_double(0.0)
        self._ck(self._lib.wb200_multi_fg(self._h, _addr(U), C.addressof(f) if want_f else None, _addr(G)))
        return f.value if want_f else None

    def spread(self, U):
        _check_array(U, np.complex128, self.nk * self.nb * self.nw, "U")
        out5, om, r, mn = np.zeros(5), np.zeros(self.nw), np.zeros((self.nw, 3)), C.c_double(0.0)
        self._ck(self._lib.wb200_multi_spread(self._h, _addr(U), _addr(out5), _addr(om), _addr(r), C.addressof(mn)))
        return out5, om, r, mn.value

    def fg_xy(self, XY, want_f=True, G=None):
        n = self.nk * (self.nw * self.nw + self.nb * self.nw)
        _check_array(XY, np.complex128, n, "XY")
        if G is not None:
            _check_array(G, np.complex128, n, "G_XY")
        f = C.c_double(0.0)
        self._ck(self._lib.wb200_multi_fg_xy(self._h, _addr(XY), C.addressof(f) if want_f else None, _addr(G)))
        return f.value if want_f else None

    def max_localize(self, U, f_tol=1e-7, g_tol=1e-5, max_iter=200, history=3):
        _check_array(U, np.complex128, self.nk * self.nb * self.nw, "U")
        f, it, nfg = C.c_double(0.0), C.c_int(0), C.c_int(0)
        self._ck(self._lib.wb200_multi_max_localize(self._h, _addr(U), f_tol, g_tol, max_iter, history,
                                                    C.addressof(f), C.addressof(it), C.addressof(nfg)))
        return f.value, it.value, nfg.value

    def disentangle(self, XY, f_tol=1e-7, g_tol=1e-5, max_iter=200, history=3):
        n = self.nk * (self.nw * self.nw + self.nb * self.nw)
        _check_array(XY, np.complex128, n, "XY")
        f, it, nfg = C.c_double(0.0), C.c_int(0), C.c_int(0)
        self._ck(self._lib.wb200_multi_disentangle(self._h, _addr(XY), f_tol, g_tol, max_iter, history,
                                                   C.addressof(f), C.addressof(it), C.addressof(nfg)))
        return f.value, it.value, nfg.value


def probe_fp64_peak(device=0):
    """(dmma_tflops, dfma_tflops) measured live on `device`."""
    lib = load()
    a, b = C.c_double(0.0), C.c_double(0.0)
    st = lib.wb200_probe_fp64_peak(device, C.addressof(a), C.addressof(b))
    if st != 0:
        raise WB200Error(st, "wb200_probe_fp64_peak failed")
    return a.value, b.value


def trim_pools():
    """Give the device / page-locked buffers and streams kept from destroyed contexts back to the driver."""
    load().wb200_trim_pools()


def pool_stats(device=0):
    """(bytes kept on `device` (< 0: page-locked host bytes), hits, misses) of the caching allocator."""
    h, m = C.c_longlong(0), C.c_longlong(0)
    kept = load().wb200_pool_stats(device, C.addressof(h), C.addressof(m))
    return kept, h.value, m.value


def ipc_alloc(device, nbytes):
    """(device address, 64-byte handle) of a cudaMalloc allocation that peers can map."""
    lib = load()
    ptr = C.c_void_p()
    h = (C.c_ubyte * 64)()
    st = lib.wb200_ipc_alloc(device, nbytes, C.byref(ptr), C.addressof(h))
    if st != 0:
        raise WB200Error(st, "wb200_ipc_alloc failed")
    return ptr.value, bytes(h)


def ipc_get_handle(device, ptr):
    """64-byte CUDA-IPC handle of an existing cudaMalloc allocation (its base address)"""
    lib = load()
    h = (C.c_ubyte * 64)()
    st = lib.wb200_ipc_get_handle(device, ptr, C.addressof(h))
    if st != 0:
        raise WB200Error(st, "wb200_ipc_get_handle failed")
    return bytes(h)


def peer_comm_bytes(nw):
    return int(load().wb200_peer_comm_bytes(nw))


def ipc_open(device, handle):
    lib = load()
    ptr = C.c_void_p()
    buf = (C.c_ubyte * 64).from_buffer_copy(handle)
    st = lib.wb200_ipc_open(device, C.addressof(buf), C.byref(ptr))
    if st != 0:
        raise WB200Error(st, "wb200_ipc_open failed (CUDA IPC not available between these processes?)")
    return ptr.value


def ipc_close(device, ptr):
    load().wb200_ipc_close(device, ptr)


def ipc_free(device, ptr):
    load().wb200_ipc_free(device, ptr)


class DeviceArray:
    """Lets torch (or cupy) wrap a raw device allocation: torch.as_tensor(DeviceArray(...), device=...)."""

    def __init__(self, ptr, shape, typestr="<c16"):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=2)


def host_register(a):
    """page-lock a numpy array that is reused across host-pointer calls (cudaHostRegister)"""
    lib = load()
    st = lib.wb200_host_register(_addr(a), a.nbytes)
    if st != 0:
        raise WB200Error(st, lib.wb200_last_error(None).decode())


def host_unregister(a):
    lib = load()
    st = lib.wb200_host_unregister(_addr(a))
    if st != 0:
        raise WB200Error(st, lib.wb200_last_error(None).decode())


def _io_ck(lib, st):
    if st != 0:
        raise WB200Error(st, lib.wb200_io_last_error().decode())


def read_mmn_native(path, nthreads=0):
    """-> (M [nk][nn][nb(l)][nb(m)] in the C-ABI image, kpb_k [nk][nn] 0-based int32, kpb_G [nk][nn][3] int32)"""
    lib = load()
    nb, nk, nn = C.c_int(0), C.c_int(0), C.c_int(0)
    bp = os.fsencode(path)
    _io_ck(lib, lib.wb200_mmn_header(bp, C.addressof(nb), C.addressof(nk), C.addressof(nn)))
    nb, nk, nn = nb.value, nk.value, nn.value
    M = np.empty((nk, nn, nb, nb), dtype=np.complex128)
    kpb_k = np.empty((nk, nn), dtype=np.int32)
    kpb_G = np.empty((nk, nn, 3), dtype=np.int32)
    _io_ck(lib, lib.wb200_mmn_read(bp, nb, nk, nn, _addr(M), _addr(kpb_k), _addr(kpb_G), nthreads))
    return M, kpb_k, kpb_G


def read_amn_native(path):
    """-> A [nk][nw][nb] complex (C-ABI image)"""
    lib = load()
    nb, nk, nw = C.c_int(0), C.c_int(0), C.c_int(0)
    bp = os.fsencode(path)
    _io_ck(lib, lib.wb200_amn_header(bp, C.addressof(nb), C.addressof(nk), C.addressof(nw)))
    A = np.empty((nk.value, nw.value, nb.value), dtype=np.complex128)
    _io_ck(lib, lib.wb200_amn_read(bp, nb.value, nk.value, nw.value, _addr(A)))
    return A


def read_eig_native(path, nb, nk):
    lib = load()
    E = np.empty((nk, nb), dtype=np.float64)
    _io_ck(lib, lib.wb200_eig_read(os.fsencode(path), nb, nk, _addr(E)))
    return E


class MagContext:
    """One wb200_mag (src/localization/coopt.jl): borrows the up and down single-slab contexts (same
    device) and owns Mud [nk][nb(l)][nb(m)] (Julia M[ik], nb x nb column-major per k)."""

    def __init__(self, up: Context, dn: Context, Mud):
        self._h = None
        lib = load()
        if not isinstance(Mud, (int, np.integer)):
            _check_array(Mud, np.complex128, up.nk * up.nb * up.nb, "Mud")
        self.up, self.dn = up, dn
        self.nb, self.nw, self.nk = up.nb, up.nw, up.nk
        self.n_inner = self.nw * self.nw + self.nb * self.nw
        h = C.c_void_p()
        st = lib.wb200_mag_create(C.byref(h), up._h, dn._h, _addr(Mud))
        if st != 0:
            raise WB200Error(st, lib.wb200_mag_last_error(None).decode())
        self._h = h
        self._lib = lib

    def close(self):
        if self._h is not None:
            self._lib.wb200_mag_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != 0:
            raise WB200Error(st, self._lib.wb200_mag_last_error(self._h).decode())

    def overlap(self, Uup, Udn, want_M=True):
        """-> (Mw2 [nw(n)][nw(m)] column-major image or None, omega_updn)"""
        n = self.nk * self.nb * self.nw
        _check_array(Uup, np.complex128, n, "Uup")
        _check_array(Udn, np.complex128, n, "Udn")
        Mw2 = np.empty((self.nw, self.nw), dtype=np.float64) if want_M else None
        o = C.c_double(0.0)
        self._ck(self._lib.wb200_mag_overlap(self._h, _addr(Uup), _addr(Udn), _addr(Mw2), C.addressof(o)))
        return Mw2, o.value

    def omega_updn_grad(self, Uup, Udn):
        n = self.nk * self.nb * self.nw
        _check_array(Uup, np.complex128, n, "Uup")
        _check_array(Udn, np.complex128, n, "Udn")
        Gu = np.empty((self.nk, self.nw, self.nb), dtype=np.complex128)
        Gd = np.empty_like(Gu)
        self._ck(self._lib.wb200_mag_omega_updn_grad(self._h, _addr(Uup), _addr(Udn), _addr(Gu), _addr(Gd)))
        return Gu, Gd

    def fg_xy(self, XY, lam, want_f=True, G=None):
        """-> (f or None, out4 = [Omega_updn, Omega_t, Omega_up, Omega_dn]); G [nk][2 n_inner] written in place"""
        n = self.nk * 2 * self.n_inner
        _check_array(XY, np.complex128, n, "XY")
        if G is not None:
            _check_array(G, np.complex128, n, "G")
        f = C.c_double(0.0)
        out4 = np.zeros(4, dtype=np.float64)
        self._ck(self._lib.wb200_mag_fg_xy(self._h, _addr(XY), float(lam), C.addressof(f) if want_f else None,
                                           _addr(G), _addr(out4)))
        return (f.value if want_f else None), out4

    def disentangle(self, XY, lam, f_tol=1e-7, g_tol=1e-5, max_iter=200, history=3):
        _check_array(XY, np.complex128, self.nk * 2 * self.n_inner, "XY")
        f, it, nfg = C.c_double(0.0), C.c_int(0), C.c_int(0)
        self._ck(self._lib.wb200_mag_disentangle(self._h, _addr(XY), float(lam), f_tol, g_tol, max_iter, history,
                                                 C.addressof(f), C.addressof(it), C.addressof(nfg)))
        return f.value, it.value, nfg.value
