/* wannier_b200.h — C-ABI of libwannier_b200.so
 *
 * B200 (sm_100a) implementation of the Marzari-Vanderbilt spread / gradient hot path of
 * Wannier.jl.  The reference has NO FFI for this path (it is plain Julia generic functions),
 * so every entry point below names the Julia function (file:line under the reference tree) it
 * replaces; INTEGRATION.md shows the `ccall` wrapper a maintainer adds on the Julia side.
 *
 * Conventions
 *   - complex = double[2] interleaved (re, im), matrices column-major: byte-identical to
 *     Julia `Array{ComplexF64}` / `Matrix{ComplexF64}`.
 *   - U    : [nk_total][nw][nb]   (Julia U[:, :, ik], nb x nw col-major, k slowest)
 *   - M    : [nk][nn][nb][nb]     (Julia M[ik][ib], nb x nb col-major) for the k-slab of the ctx
 *   - G    : [nk][nw][nb]         gradient for the k-slab of the ctx
 *   - indices are 0-based int32; `kpb_k` holds GLOBAL k indices (0 <= idx < nk_total)
 *   - every function returns 0 on success and a negative wb200_status otherwise; the message is
 *     available from wb200_last_error().  There is no CPU fallback: wb200_create fails without
 *     an sm_100 device.
 *   - a ctx is bound to ONE device and is not thread-safe (same contract as the reference's
 *     `Cache`, src/spread.jl:127-162, which is captured mutably by the fg! closure).
 *   - "host" entry points take host pointers, copy in/out and block until done.
 *     "_dev" entry points take device pointers valid on the ctx's device, enqueue on the ctx's
 *     stream (wb200_set_stream) and do not synchronise unless documented.
 */
#ifndef WANNIER_B200_H
#define WANNIER_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wb200_ctx wb200_ctx;

enum wb200_status {
    WB200_OK = 0,
    WB200_ERR_ARG = -1,       /* bad argument (shape, null pointer, index out of range) */
    WB200_ERR_CUDA = -2,      /* CUDA runtime error; see wb200_last_error */
    WB200_ERR_NODEVICE = -3,  /* no sm_100 device */
    WB200_ERR_NOCONV = -4,    /* iterative routine (polar factor, line search) did not converge */
    WB200_ERR_SINGULAR = -5   /* |N_nn| < 1e-10: the reference's "N^kb too small" (opt_rotate.jl:60-62) */
};

/* flags for wb200_create */
#define WB200_FLAG_DEFAULT 0u
#define WB200_FLAG_NO_TENSOR 1u   /* force the CUDA-core FP64 rotation kernel even where DMMA applies */
#define WB200_FLAG_FORCE_TENSOR 2u /* fail instead of falling back when the DMMA kernel cannot be used */
#define WB200_FLAG_M_ROW_MAJOR 8u  /* wb200_create: every M_kb is given row-major (C order, M[k][b][m][l] — what numpy
                                      holds); it is transposed on the device instead of by the caller */
#define WB200_FLAG_NO_FUSED_TRIAL 4u /* optimiser drivers: never use the one-launch trial-point kernel for tiny shapes
                                        (n_bands <= 32 and little arithmetic, tiny.cu); the general kernels run instead (tests compare the two) */

/* ---- context --------------------------------------------------------------------------------
 * Replaces `Cache(model)` (src/spread.jl:139-162) + the stencil fields read by the hot loops
 * (src/kpoints/kstencil.jl:19-60).  Uploads M once (it is never written by the optimiser,
 * src/localization/max_localize.jl:10-26) and owns all work buffers.
 *   device     CUDA device ordinal
 *   nb nw      n_bands, n_wann          nk_total  number of k-points of the model
 *   k_begin,nk the contiguous k-slab [k_begin, k_begin+nk) this ctx owns (single GPU: 0, nk_total)
 *   nn         number of b-vectors per k-point
 *   kpb_k      [nk][nn]  global index of k+b                      (stencil.kpb_k, 0-based)
 *   bvec_cart  [nk][nn][3] recip_lattice*(k[kpb]+G-k[ik]) in 1/A  (src/spread.jl:293)
 *   bweights   [nn] w_b in A^2, indexed by position ib             (stencil.bweights)
 *   M          [nk][nn][nb*nb] complex                             (model.overlaps)
 *              M alone may also be a DEVICE pointer (unified addressing decides; a 13 GB overlap
 *              set produced on the GPU need not bounce through the host).                        */
int wb200_create(wb200_ctx** out, int device, int nb, int nw, int nk_total, int k_begin, int nk,
                 int nn, const int32_t* kpb_k, const double* bvec_cart, const double* bweights,
                 const double* M, unsigned flags);
void wb200_destroy(wb200_ctx* ctx);
const char* wb200_last_error(const wb200_ctx* ctx); /* ctx may be NULL: last create error */
const char* wb200_version(void);
/* use the given cudaStream_t (e.g. torch's current stream) for all subsequent work.  NULL is CUDA's
 * legacy default stream.  A fresh ctx works on a private non-blocking stream. */
int wb200_set_stream(wb200_ctx* ctx, void* cuda_stream);
int wb200_synchronize(wb200_ctx* ctx);
/* which rotation kernel the ctx selected: 0 = CUDA-core FP64, 1 = FP64 DMMA tensor path */
int wb200_rotation_kernel(const wb200_ctx* ctx);
/* number of kernel launches issued by this ctx since creation (bench.py's gpu_launches) */
long long wb200_launch_count(const wb200_ctx* ctx);

/* CenterSpreadPenalty(r0, lambda) (src/spread.jl:214-227); r0 = NULL restores SpreadPenalty. */
int wb200_set_penalty(wb200_ctx* ctx, const double* r0 /*[nw][3]*/, double lambda);
/* frozen_bands of the disentanglement (src/localization/disentangle.jl:481-501), [nk][nb] 0/1 */
int wb200_set_frozen(wb200_ctx* ctx, const uint8_t* frozen);

/* ---- host-pointer entry points (the drop-in calls) -------------------------------------------
 * wb200_fg      = fg!(F, G, U) of get_fg!_maxloc (src/localization/max_localize.jl:13-23):
 *                 compute_MU_UtMU! (spread.jl:164-195) -> omega_grad! (398-481) -> omega! (254-360).
 *                 f and/or G may be NULL (Optim's only_fg! convention).  f receives Omega, or
 *                 Omega_t when a centre penalty is set (spread.jl:220).
 * wb200_spread  = omega(stencil, M, U) (spread.jl:377-381): out5 = {Omega, OmegaI, OmegaOD,
 *                 OmegaD, OmegaTilde}; omega_n [nw]; r [nw][3]; min_abs_Nnn = min |N_nn| (the
 *                 guard the reference has commented out, spread.jl:452-457).  Always forms the
 *                 full N = U^H M U, so the split is exact for any U (rectangular, non-unitary).
 * wb200_center  = center(stencil, M, U) (spread.jl:560-594).
 * wb200_rotate_overlaps = transform_gauge(M, kpb_k, U) (src/localization/gauge.jl:83-98) and
 *                 compute_MU_UtMU! (spread.jl:164-195): N [nk][nn][nw*nw]; MU [nk][nn][nb*nw]
 *                 (either may be NULL).
 * wb200_fg_xy   = fg!(Omega, G, XY) of get_fg!_disentangle (disentangle.jl:586-614); XY and G_XY
 *                 are (nw*nw + nb*nw) x nk column-major, column ik = [vec(X_k); vec(Y_k)].
 * A single-slab ctx (k_begin = 0, nk = nk_total) takes the whole model's arrays.  A k-slab ctx of a
 * sharded run (wb200_set_allreduce + wb200_set_peer_gauges) takes the SLAB's rows of U / XY / G in
 * wb200_fg, wb200_spread, wb200_center and wb200_fg_xy; every rank calls collectively.          */
int wb200_fg(wb200_ctx* ctx, const double* U, double* f, double* G);
int wb200_spread(wb200_ctx* ctx, const double* U, double out5[5], double* omega_n, double* r,
                 double* min_abs_Nnn);
int wb200_center(wb200_ctx* ctx, const double* U, double* r);
int wb200_rotate_overlaps(wb200_ctx* ctx, const double* U, double* N, double* MU);
int wb200_fg_xy(wb200_ctx* ctx, const double* XY, double* f, double* G_XY);

/* Optim.Stiefel_SVD (third-party; call sites max_localize.jl:62-63, disentangle.jl:687-690):
 *   project_tangent!: G <- G - X (X^H G + G^H X)/2      per matrix
 *   retract!:         X <- polar factor of X (= U V^H of svd(X)), computed GEMM-only
 * X, G: [count][ncol][nrow] complex, nrow x ncol column-major.                                   */
int wb200_project_tangent(wb200_ctx* ctx, const double* X, double* G, int nrow, int ncol, int count);
int wb200_retract(wb200_ctx* ctx, double* X, int nrow, int ncol, int count);

/* max_localize(model; f_tol, g_tol, max_iter, history_size) (max_localize.jl:44-101) with the
 * optimiser loop device-resident (L-BFGS on the power Stiefel manifold, strong-Wolfe line search,
 * Optim's stopping rules |df| <= f_tol |f| or |g|_inf <= g_tol).  U_inout [nk][nw][nb], nb == nw.
 * Single-slab ctx, or a k-slab ctx with wb200_set_allreduce + wb200_set_peer_gauges (see below).
 * iters / n_fg may be NULL.                                                                      */
int wb200_max_localize(wb200_ctx* ctx, double* U_inout, double f_tol, double g_tol, int max_iter,
                       int history, double* f_final, int* iters, int* n_fg);
/* disentangle(model; ...) (disentangle.jl:631-728) on the product manifold (X, Y); frozen set by
 * wb200_set_frozen; XY_inout as in wb200_fg_xy (initial point from U_to_X_Y on the host).       */
int wb200_disentangle(wb200_ctx* ctx, double* XY_inout, double f_tol, double g_tol, int max_iter,
                      int history, double* f_final, int* iters, int* n_fg);

/* ---- device-pointer entry points (resident data, k-sharding across processes) ----------------
 * One evaluation on a k-slab is split where the reference has its global dependency
 * (omega_grad! needs r from center! over ALL k, spread.jl:411):
 *   phase1: rotation of the slab's overlaps + per-slab partial sums  -> sums [wb200_nsums]
 *           (caller all-reduces `sums` over the ranks: ncclAllReduce(sum) / torch.distributed)
 *   phase2: spread scalars from the reduced sums and the slab's gradient.
 *   dU    device, [nk_total][nw][nb] (all k: neighbours k+b may live in other slabs)
 *   dsums device, wb200_nsums(ctx) doubles (layout: sum w b phi [nw][3], sum w(1-|N|^2+phi^2)
 *         [nw], sum w ||N||_F^2, sum w sum_n |N_nn|^2)
 *   dres  device, wb200_nres(ctx) doubles: {Omega, OmegaI, OmegaOD, OmegaD, OmegaTilde, f,
 *         0, 0, omega_n[nw], r[nw][3]}
 *   dG    device, [nk][nw][nb] or NULL
 * `exact_split` != 0 forms the full N (second GEMM) so OmegaI/OmegaOD are exact for any U;
 * 0 uses ||N||_F = ||M U_{k+b}||_F, valid for square unitary U_k (max_localize) — and on the tensor path with
 * n_bands == n_wann the gauge invariance of that norm, ||U_k^H M_kb U_{k+b}||_F = ||M_kb||_F, computed from
 * M once at creation (the spread functional's Omega_I is gauge invariant for an isolated manifold;
 * WB200_CONST_NORM=0 in the environment keeps the per-evaluation norm); Omega, r, omega_n and G never
 * depend on this choice.                                                                          */
int wb200_nsums(const wb200_ctx* ctx);
int wb200_nres(const wb200_ctx* ctx);
int wb200_fg_phase1_dev(wb200_ctx* ctx, const double* dU, double* dsums, int exact_split);
int wb200_fg_phase2_dev(wb200_ctx* ctx, const double* dsums, double* dres, double* dG);
/* phase1 + phase2 for a single-slab ctx, no host synchronisation */
int wb200_fg_dev(wb200_ctx* ctx, const double* dU, double* dres, double* dG, int exact_split);
/* min over the slab of |N_nn| from the last phase1 (synchronises) */
int wb200_min_abs_diag(wb200_ctx* ctx, double* out);
int wb200_project_tangent_dev(wb200_ctx* ctx, const double* dX, double* dG, int nrow, int ncol,
                              int count);
int wb200_retract_dev(wb200_ctx* ctx, double* dX, int nrow, int ncol, int count, int* iters);

/* ---- by-products of the rotation (same U_k^H M_kb U_{k+b}, different epilogue) -------------------
 * wb200_omega_local      = omega_local(stencil, M, U) (src/spread.jl:522-548): loc [nk].
 * wb200_fg_rotate        = f / g! of get_fg!_rotate (src/localization/opt_rotate.jl:12-81): one global
 *                          rotation W [nw*nw] applied at every k (the caller pre-rotates M and uses
 *                          identity gauges, opt_rotate.jl:113-116); G [nw*nw] = sum_k G_k.
 * wb200_opt_rotate       = opt_rotate(model; ...) (opt_rotate.jl:97-154): L-BFGS on one Stiefel matrix.
 * wb200_position_op      = position_op(stencil, M, U) (src/spread.jl:674-716): R [nw][nw][3] complex
 *                          (Julia Matrix{Vec3{ComplexF64}} image: component fastest, then m, then n).
 * wb200_berry_connection = compute_berry_connection_kspace (src/spread.jl:747-793):
 *                          A [nk][nw][nw][3] complex, same per-k image.                           */
int wb200_omega_local(wb200_ctx* ctx, const double* U, double* loc);
int wb200_fg_rotate(wb200_ctx* ctx, const double* W, double* f, double* G);
int wb200_opt_rotate(wb200_ctx* ctx, double* W_inout, double f_tol, double g_tol, int max_iter,
                     int history, double* f_final, int* iters, int* n_fg);
int wb200_position_op(wb200_ctx* ctx, const double* U, double* R);
int wb200_berry_connection(wb200_ctx* ctx, const double* U, double* A, int imlog_diag,
                           int force_hermiticity);

/* ---- peer memory (NVLink) for k-sharded runs, one process per GPU -------------------------------
 * Instead of a separate halo exchange, the gauge-packing kernel of the tensor path reads the rows
 * U_{k+b} that another rank owns DIRECTLY from that rank's memory over NVLink (plain loads on a
 * CUDA-IPC mapped pointer), fusing the neighbour-gauge exchange into the kernel that consumes it.
 *   wb200_ipc_alloc : cudaMalloc + cudaIpcGetMemHandle (every rank allocates a FULL [nk_total][nw][nb]
 *                     gauge array and owns the rows of its slab)
 *   wb200_ipc_open  : map a peer's allocation (handle exchanged by the host, e.g. all_gather_object)
 *   wb200_set_peer_gauges : U_ptrs[r] = device pointer of rank r's array as visible from this
 *                     device (own rank: the local pointer), bounds[r] = first k of rank r's slab.
 *                     Afterwards phase1 reads row k from the array of the rank that owns k.
 * The caller orders the accesses: peers must have finished writing their rows before phase1 is
 * enqueued (any collective after the update does), and must not overwrite them before every
 * rank's phase1 has completed (the all-reduce of the sums between phase1 and phase2 does).       */
int wb200_ipc_alloc(int device, size_t bytes, void** dptr, unsigned char handle[64]);
int wb200_ipc_open(int device, const unsigned char handle[64], void** dptr);
int wb200_ipc_close(int device, void* dptr);
int wb200_ipc_free(int device, void* dptr);
int wb200_set_peer_gauges(wb200_ctx* ctx, int world, const void* const* U_ptrs, const int32_t* bounds);
/* Halo overlap.  Once peer gauges are set, phase 1 of the tensor path (32 < nb <= 256) can split the neighbours
 * of every k-point into those whose gauge the rank owns and those a peer owns: the peer rows arrive on a
 * second stream WHILE the main stream packs the local rows and rotates the local-neighbour pairs; the
 * remote-neighbour pairs follow.  dU passed to wb200_fg_phase1_dev must be U_ptrs[rank].
 * WB200_HALO_SPLIT=0 (environment, read here) keeps the one-pass packing.
 * (The split needs the comm buffers and the peers' packed arrays below; without them phase 1 packs local and
 * peer rows in one pass, reading the peer rows over NVLink inside the pack kernel.)                         */

/* ---- the library's own collectives over peer memory (peer.cu), one slab per GPU ---------------------------
 * wb200_peer_comm_bytes(nw): size of a rank's comm buffer (readiness flags + all-reduce slots).  Every rank
 *   allocates one (wb200_ipc_alloc, or cudaMalloc with peer access in one process), ZEROES it, and all ranks
 *   synchronise before the first evaluation.
 * wb200_set_peer_comm(ctx, world, comm_ptrs): comm_ptrs[q] = rank q's buffer as mapped on this device; must
 *   follow wb200_set_peer_gauges; NULL removes.  From then on
 *     - an evaluation no longer opens with a one-element "rows are in place" all-reduce: phase 1 publishes the
 *       rank's evaluation counter into its peers' buffers (remote stores) and waits for the owners of its halo;
 *     - unless the caller registers its own wb200_set_allreduce callback, every all-reduce of the library
 *       (spread sums, min |N_nn|, the optimiser's scalar products) is ONE small kernel per rank that pushes the
 *       contribution into a slot of every peer's buffer, waits for the peers and sums the slots in rank order:
 *       bitwise identical on every rank, no NCCL on the hot path.  wb200_peer_allreduce_dev exposes it
 *       (op 0 sum, 1 max, 2 min; count <= 4 nw + 64) for callers that drive phase1 / phase2 themselves.
 *   Every rank must make the same sequence of calls.  One slab per device only (a spinning wait kernel must
 *   never share a hardware queue with the signal it waits for); bounded spins trap after ~20 s.
 * wb200_packed_gauge / wb200_ipc_get_handle / wb200_set_peer_packed: the tensor path (32 < nb <= 256) keeps the
 *   gauge as a packed B-operand image; with the peers' images mapped (Up_ptrs[q]) a slab packs only ITS rows and
 *   pulls the packed rows of its halo from the owners with the copy engines on a second stream while the
 *   local-neighbour pairs are being rotated (src/spread.jl:164-195 has no counterpart: single process).      */
size_t wb200_peer_comm_bytes(int nw);
int wb200_set_peer_comm(wb200_ctx* ctx, int world, void* const* comm_ptrs);
int wb200_peer_allreduce_dev(wb200_ctx* ctx, double* dptr, int count, int op);
int wb200_packed_gauge(wb200_ctx* ctx, void** dptr, size_t* bytes);
int wb200_ipc_get_handle(int device, void* dptr, unsigned char handle[64]);
int wb200_set_peer_packed(wb200_ctx* ctx, int world, void* const* Up_ptrs);

/* ---- sharded optimiser: caller-supplied collective ------------------------------------------------
 * With one process per GPU the library does not own a communicator.  The host registers a callback
 * that all-reduces `count` doubles at device pointer `dptr` IN PLACE over all ranks, enqueued on
 * `cuda_stream` (op 0 = sum, 1 = max, 2 = min) — ncclAllReduce from Julia/C, torch.distributed from Python.
 * With the callback and wb200_set_peer_gauges in place, wb200_max_localize and wb200_disentangle run
 * on a k-slab ctx: U_inout / XY is then the SLAB ([nk][nw][nb] / [nk][nw*nw + nb*nw]; frozen bands
 * per slab); every rank calls them collectively with the same arguments.
 * Vectors of the L-BFGS live slab-local; scalar products, |g|_inf, the spread sums and a one-element
 * "gauges are in place" reduction before each evaluation go through the callback.
 * The callback returns 0 on success; any other value aborts the call with WB200_ERR_CUDA (the ranks
 * are then out of step and the contexts must be destroyed).  Every branch of the optimiser depends
 * only on reduced scalars, so all ranks stay in step; a rank-local failure of a retraction turns
 * that evaluation's reduced sums into NaN, i.e. the trial point is rejected on every rank.       */
typedef int (*wb200_allreduce_fn)(void* dptr, int count, int op, void* cuda_stream, void* user);
int wb200_set_allreduce(wb200_ctx* ctx, wb200_allreduce_fn fn, void* user);

/* ---- one call, several GPUs (devices[], ndev) -----------------------------------------------
 * The reference's callers are single calls from one process: max_localize(model)
 * (src/localization/max_localize.jl:44-101), disentangle(model) (disentangle.jl:631-728), omega(model)
 * (src/spread.jl:370-381), fg!(F, G, U) (max_localize.jl:13-23).  A wb200_multi shards the k-points in
 * contiguous slabs over `devices` (one slab context per entry; a device may be listed more than once),
 * keeps one host thread per slab for the duration of a call, lets the pack kernels read neighbour gauges
 * from the owning device through peer access (NVLink) and reduces the spread sums / optimiser scalars in
 * rank order in pinned host memory (deterministic, identical on every slab).  All arrays are the FULL
 * model's, host pointers, laid out as for the single-device calls; M is uploaded slab by slab.
 * The same slab contexts can be driven one process per GPU: wb200_fg / wb200_spread / wb200_center /
 * wb200_fg_xy / wb200_max_localize / wb200_disentangle accept a k-slab ctx once wb200_set_allreduce and
 * wb200_set_peer_gauges are in place (arrays are then the SLAB's rows and every rank calls collectively). */
typedef struct wb200_multi wb200_multi;
int wb200_multi_create(wb200_multi** out, const int* devices, int ndev, int nb, int nw, int nk, int nn,
                       const int32_t* kpb_k, const double* bvec_cart, const double* bweights,
                       const double* M, unsigned flags);
void wb200_multi_destroy(wb200_multi* m);
const char* wb200_multi_last_error(const wb200_multi* m);   /* m may be NULL: last create error */
int wb200_multi_ndev(const wb200_multi* m);
int wb200_multi_slab_begin(const wb200_multi* m, int r);   /* first k of slab r (r = ndev: nk) */
long long wb200_multi_launch_count(const wb200_multi* m);
int wb200_multi_set_penalty(wb200_multi* m, const double* r0, double lambda);
int wb200_multi_set_frozen(wb200_multi* m, const uint8_t* frozen /*[nk][nb]*/);
int wb200_multi_fg(wb200_multi* m, const double* U, double* f, double* G);
int wb200_multi_spread(wb200_multi* m, const double* U, double out5[5], double* omega_n, double* r,
                       double* min_abs_Nnn);
int wb200_multi_fg_xy(wb200_multi* m, const double* XY, double* f, double* G_XY);
int wb200_multi_max_localize(wb200_multi* m, double* U_inout, double f_tol, double g_tol, int max_iter,
                             int history, double* f_final, int* iters, int* n_fg);
int wb200_multi_disentangle(wb200_multi* m, double* XY_inout, double f_tol, double g_tol, int max_iter,
                            int history, double* f_final, int* iters, int* n_fg);

/* ---- initial point of the disentanglement (src/localization/disentangle.jl:223-281, 369-402) --------------
 * wb200_u_to_xy          = X_Y_to_XY(U_to_X_Y(U, frozen)...) (disentangle.jl:369-402, 454-466): the starting
 *                          point `disentangle` builds at :666-670.  U [nk][nw][nb], frozen from wb200_set_frozen
 *                          (none: nothing frozen), XY as in wb200_fg_xy.  One kernel completes the frozen rows of
 *                          U_k to an orthonormal basis; the two polar factors come from the batched retraction.
 *                          The free block of Y_k is AN orthonormal basis of range(Ur); the reference's is the
 *                          LAPACK eigenbasis of a degenerate projector, i.e. defined up to a unitary mixing, so
 *                          X_k and Y_k agree with the reference up to that mixing while Y_k X_k =
 *                          orthonorm_freeze(U_k), the frozen rows of Y_k and X_k^H X_k = I agree exactly.
 * wb200_orthonorm_freeze = orthonorm_freeze(U_k, frozen_k) for every k (disentangle.jl:223-281), in place.
 * Errors: WB200_ERR_ARG (more frozen bands than n_wann), WB200_ERR_SINGULAR (frozen rows of U linearly
 * dependent), WB200_ERR_NOCONV (the free rows do not span n_wann - n_frozen dimensions: the reference's
 * assertion at :265).  Work on k-slab contexts too (slab rows, no communication).
 * wb200_create_stiefel   : a context WITHOUT stencil and overlaps for callers that have no model at hand
 *                          (U_to_X_Y(U, frozen) takes none): valid for wb200_set_frozen, wb200_u_to_xy,
 *                          wb200_orthonorm_freeze, wb200_retract, wb200_project_tangent; every evaluation entry
 *                          point returns WB200_ERR_ARG.                                                        */
int wb200_u_to_xy(wb200_ctx* ctx, const double* U, double* XY);
int wb200_orthonorm_freeze(wb200_ctx* ctx, double* U_inout);
int wb200_create_stiefel(wb200_ctx** out, int device, int nb, int nw, int nk);

/* ---- Wannier90 text ingest (the setup step in front of the path) ---------------------------------------------
 * The reference reads these files through WannierIO.jl (third party; `read_w90`, src/io/w90/model.jl:16-94).
 * The parsers here are native and multithreaded (a 16^3 x 12 x 128^2 .mmn is ~13 GB of text) and write the
 * C-ABI memory images directly, so their output can be handed to wb200_create unchanged:
 *   wb200_mmn_read: M [nk][nn][nb*nb] complex column-major, kpb_k [nk][nn] 0-based, kpb_G [nk][nn][3]; the b index
 *                   is the order of appearance of a k-point's pairs (src/kpoints/kstencil.jl:68-85).
 *   wb200_amn_read: A [nk][nw][nb] complex (the projections; Lowdin via wb200_retract gives the initial gauges,
 *                   src/io/w90/amn.jl:12-17).
 *   wb200_eig_read: E [nk][nb].
 * Sizes come from the *_header calls (.eig has no header: nb, nk from the .mmn).  nthreads <= 0: all cores
 * (at most 32).  Errors: WB200_ERR_ARG with the reason in wb200_io_last_error (thread-local).            */
int wb200_mmn_header(const char* path, int* nb, int* nk, int* nn);
int wb200_mmn_read(const char* path, int nb, int nk, int nn, double* M, int32_t* kpb_k, int32_t* kpb_G, int nthreads);
int wb200_amn_header(const char* path, int* nb, int* nk, int* nw);
int wb200_amn_read(const char* path, int nb, int nk, int nw, double* A);
int wb200_eig_read(const char* path, int nb, int nk, double* E);
const char* wb200_io_last_error(void);

/* ---- MagModel co-optimisation (src/localization/coopt.jl) ---------------------------------------------
 * Spin-up and spin-down models coupled by the overlap of up and down Wannier functions.  A wb200_mag
 * borrows two single-slab contexts on the same device (the up and down `Model`s of `MagModel`,
 * coopt.jl:8-17; equal n_bands, n_wann, n_kpts as `disentangle(::MagModel)` asserts, :355-357) and owns
 * Mud = <u_nk^up | u_mk^dn>, [nk][nb*nb] complex column-major (host or device pointer).
 *   wb200_mag_overlap         = overlap_updn(model, Uup, Udn) (coopt.jl:114-135): Mw2 [nw*nw] real column-major
 *                               = |sum_k Uup_k^H Mud_k Udn_k|^2 / nk^2 (may be NULL), and omega_updn (:146-160)
 *                               = n_wann - tr(Mw2) (may be NULL).  Uup, Udn [nk][nw][nb].
 *   wb200_mag_omega_updn_grad = omega_updn_grad(model, Uup, Udn) (coopt.jl:222-229): GUup, GUdn [nk][nw][nb].
 *   wb200_mag_fg_xy           = f / g! of get_fg!_disentangle(model::MagModel, lambda) (coopt.jl:252-321):
 *                               XY and G are 2 (nw*nw + nb*nw) x nk column-major, column ik =
 *                               [vec(Xup_k); vec(Yup_k); vec(Xdn_k); vec(Ydn_k)]; frozen bands from each context
 *                               (wb200_set_frozen).  f, G, out4 may be NULL; out4 = {Omega_updn, Omega_t,
 *                               Omega_up, Omega_dn} (the scalar fields of SpreadMag, coopt.jl:46-64).
 *   wb200_mag_disentangle     = disentangle(model::MagModel, lambda; f_tol, g_tol, max_iter, history_size)
 *                               (coopt.jl:339-420), device-resident like wb200_disentangle; XY_inout from
 *                               U_to_X_Y of both spins (:369-374).
 * omega(model::MagModel, Uup, Udn, lambda) (coopt.jl:66-73) is wb200_spread on each context plus
 * wb200_mag_overlap.  wb200_mag_destroy leaves the two contexts alive.                             */
typedef struct wb200_mag wb200_mag;
int wb200_mag_create(wb200_mag** out, wb200_ctx* up, wb200_ctx* dn, const double* Mud);
void wb200_mag_destroy(wb200_mag* m);
const char* wb200_mag_last_error(const wb200_mag* m);   /* m may be NULL: last create error */
int wb200_mag_overlap(wb200_mag* m, const double* Uup, const double* Udn, double* Mw2, double* omega_updn);
int wb200_mag_omega_updn_grad(wb200_mag* m, const double* Uup, const double* Udn, double* GUup, double* GUdn);
int wb200_mag_fg_xy(wb200_mag* m, const double* XY, double lambda, double* f, double* G, double out4[4]);
int wb200_mag_disentangle(wb200_mag* m, double* XY_inout, double lambda, double f_tol, double g_tol,
                          int max_iter, int history, double* f_final, int* iters, int* n_fg);

/* ---- pinned host memory ----------------------------------------------------------------------------------------
 * The host-pointer entry points (wb200_fg, wb200_multi_fg, ...) overlap PCIe copies with compute; that needs
 * page-locked memory.  Julia `Array`s and numpy arrays are pageable: register the U / G / XY buffers that are
 * reused across calls once (cudaHostRegister, portable across devices) and unregister them before they are freed.
 * Registering twice / unregistering an unknown pointer is not an error.  Message: wb200_last_error(NULL).   */
int wb200_host_register(void* ptr, size_t bytes);
int wb200_host_unregister(void* ptr);

/* ---- resource pools ---------------------------------------------------------------------------------------------
 * The reference's one-shot calls (omega(bvectors, M, U), omega_grad(...), max_localize(model): src/spread.jl:370-381,
 * 496-510, benchmark/bench_spread.jl, bench_grad.jl, bench_maxloc.jl) build a Cache per call; here that is a device
 * context per call.  The library therefore keeps the device buffers, page-locked host buffers and streams of
 * destroyed contexts (per device, blocks of at most 64 MB, at most 1 GB) and hands them out again, so that a one-shot
 * call costs its arithmetic and copies, not ~60 driver allocations.  WB200_POOL=0 disables this.
 *   wb200_trim_pools   gives everything kept back to the driver.
 *   wb200_pool_stats   bytes kept on `device` (device < 0: page-locked host bytes); hits / misses (nullable) count
 *                      device allocations served from / not served from the kept blocks since the process started. */
void wb200_trim_pools(void);
long long wb200_pool_stats(int device, long long* hits, long long* misses);

/* ---- measurement helpers (bench.py) -------------------------------------------------------------
 * wb200_profile: when enabled, every launch of the four kernel classes below is bracketed by CUDA
 * events on the ctx's stream; wb200_profile_read synchronises, returns the summed milliseconds and
 * launch counts since the last read and resets them.
 *   class 0: gauge packing   1: rotation (the dominant kernel)   2: spread reductions   3: gradient
 * wb200_probe_fp64_peak: short live measurement of this device's FP64 peaks (TFLOP/s): tensor-pipe
 * DMMA.8x8x4 and CUDA-core DFMA (MEASURED_PEAKS.json holds only HBM and bf16 figures).           */
int wb200_profile(wb200_ctx* ctx, int enable);
int wb200_profile_read(wb200_ctx* ctx, double ms[4], int counts[4]);
int wb200_probe_fp64_peak(int device, double* dmma_tflops, double* dfma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* WANNIER_B200_H */
