// common.cuh — shared declarations of libwannier_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "../../include/wannier_b200.h"

typedef double2 cplx;  // (x = re, y = im) == Julia ComplexF64

__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
// c += a*b
__device__ __forceinline__ void cfma(cplx& c, cplx a, cplx b) {
    c.x = fma(a.x, b.x, c.x);
    c.x = fma(-a.y, b.y, c.x);
    c.y = fma(a.x, b.y, c.y);
    c.y = fma(a.y, b.x, c.y);
}
// c += conj(a)*b
__device__ __forceinline__ void cfma_conj(cplx& c, cplx a, cplx b) {
    c.x = fma(a.x, b.x, c.x);
    c.x = fma(a.y, b.y, c.x);
    c.y = fma(a.x, b.y, c.y);
    c.y = fma(-a.y, b.x, c.y);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// generic batched complex GEMM   C[b] = alpha * op(A[ia(b)]) * op(B[ib(b)]) + beta * C[b]
// ---------------------------------------------------------------------------------------------
struct ZgemmArgs {
    int m, n, k;             // op(A): m x k, op(B): k x n, C: m x n
    const cplx* A;
    long long strideA;       // in complex elements
    int lda;
    const int32_t* idxA;     // optional batch -> matrix index
    const cplx* B;
    long long strideB;
    int ldb;
    const int32_t* idxB;
    cplx* C;
    long long strideC;
    int ldc;
    double alpha, beta;
    int batch;
    int transA, transB;      // 0 = N, 1 = conjugate transpose
};

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct DmmaPlan;  // rotate_dmma.cu

// Slabs of ONE process (wb200_multi_*, multi.cu): the host threads of the slabs can rendezvous and CUDA events order
// streams across devices, so the peer-memory collectives need no spinning kernels there (a spinning kernel and the
// runtime's lazy module loading / pageable copies of another host thread of the same process can wait for each other).
struct WbInproc {
    int n = 0;
    std::vector<cudaEvent_t> ev_ready, ev_packed;   // [slab]: "gauge rows written" / "rows packed", recorded on the slab's stream
    std::vector<cudaEvent_t> ev_ar[2];              // [parity][slab]: contribution pushed into every slab's slot
    std::function<bool()> barrier;                  // host-thread rendezvous of the slabs; false: a slab failed
};

struct wb200_ctx {
    int device = 0;
    int num_sms = 148;
    int nb = 0, nw = 0, nk_total = 0, k_begin = 0, nk = 0, nn = 0;
    unsigned flags = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;   // pipelined host entry point
    std::vector<cudaEvent_t> ev_up, ev_grad;
    std::vector<int> sub_bounds;                  // sub-slab bounds of the pipelined wb200_fg
    std::vector<std::vector<int>> sub_deps;       // sub-slabs of U each sub-slab reads
    cudaStream_t stream = nullptr;
    long long launches = 0;
    std::string err;

    // constant data (uploaded once)
    int32_t* d_kpb = nullptr;    // [nk][nn] global index of k+b
    int32_t* d_kself = nullptr;  // [nk*nn] global index of k for every pair
    int32_t* d_need = nullptr;   // [n_need] sorted global k indices the slab reads (own + k+b)
    int n_need = 0;
    std::vector<int32_t> h_need;
    // peer gauges (wb200_set_peer_gauges): row k is read from the array of the rank owning k
    const cplx** d_peerU = nullptr;   // [world] device pointers (device array)
    int32_t* d_need_owner = nullptr;  // [n_need] owner rank of need[i]
    int peer_world = 0;
    int peer_rank = -1;
    cplx* peer_self = nullptr;        // this rank's full gauge array (U_ptrs[rank])
    // halo split (set up with the peer gauges): need list reordered local rows first, per-k neighbour order
    // with the local neighbours first, and the second stream the peer-memory half of the packing runs on
    int32_t* d_need_split = nullptr;  // [n_need] local rows, then remote rows
    int32_t* d_owner_split = nullptr; // [n_need]
    int n_need_local = 0;
    uint8_t* d_border = nullptr;      // [nk][nn] neighbour indices, local ones first
    uint8_t* d_nloc = nullptr;        // [nk] number of local neighbours
    int halo_split = 0;               // 1: phase 1 overlaps the halo packing with the local-neighbour rotation
    cudaStream_t halo_stream = nullptr;
    cudaEvent_t ev_rows = nullptr, ev_halo = nullptr;
    // peer-memory comm buffers (wb200_set_peer_comm, peer.cu): readiness flags and all-reduce slots
    unsigned long long** d_comm = nullptr;        // [world] device array: rank q's comm buffer
    unsigned long long* my_comm = nullptr;        // comm_ptrs[rank]
    int comm_maxc = 0;                            // doubles per all-reduce slot
    unsigned long long spin_timeout_ns = 20000000000ull;
    WbInproc* inproc = nullptr;                   // set by wb200_multi_create: events + host barrier instead of flags
    unsigned long long epoch = 0, ar_epoch = 0;   // evaluation / all-reduce counters (identical on every rank)
    unsigned peer_mask = 0;                       // ranks owning rows this slab reads
    std::vector<void*> h_peerUp;                  // packed-gauge arrays of the ranks (wb200_set_peer_packed)
    int peer_pull_kernel = 0;                     // pull with a kernel instead of the copy engines (WB200_PEER_PULL=kernel)
    struct HaloRun { int owner, k0, count; };
    std::vector<HaloRun> halo_runs;               // contiguous runs of remote rows with one owner
    wb200_allreduce_fn allreduce = nullptr;   // caller-supplied collective (sharded optimiser)
    void* allreduce_user = nullptr;
    int retract_deferred = 0;  // optimiser, single slab: the small retraction does not wait for its convergence flag ...
    int retract_pending = 0;   // ... which is read after the evaluation's wait (stiefel_small.cu)
    int publish_eval = 0;      // the next wb_cross_launch also publishes f, min |N_nn| and the retraction flag (vecops.cu)
    int poison_next = 0;   // the next evaluation's sums are replaced by NaN before they are reduced (a
                           // rank-local failure, e.g. of a retraction, must be seen by every rank)
    double* d_bvec = nullptr;    // [nk][nn][3]
    double* d_wb = nullptr;      // [nn]
    cplx* d_M = nullptr;         // [nk][nn][nb*nb] col-major (generic path / exact N)
    cplx* d_M_given = nullptr;   // WB200_FLAG_M_ROW_MAJOR: the overlaps as uploaded, until wb200_create has transposed them
    int release_M_after_create = 0;   // the tensor plan has re-tiled M: wb200_create releases d_M after its final wait
    double wsum = 0.0;           // sum_b w_b
    std::vector<double> h_wb;
    std::vector<int32_t> h_kpb;  // host copy of kpb_k (sub-slab dependency plan)

    // penalty / frozen
    int has_penalty = 0;
    double lambda = 0.0;
    double* d_r0 = nullptr;       // [nw][3]
    uint8_t* d_frozen = nullptr;  // [nk][nb]
    int32_t* d_nfrozen = nullptr; // [nk]

    // work buffers
    cplx* d_U = nullptr;     // [nk_total][nb*nw] staging copy for host entry points
    cplx* d_G = nullptr;     // [nk][nb*nw]
    cplx* d_MU = nullptr;    // [nk][nn][nb*nw]
    cplx* d_N = nullptr;     // [nk][nn][nw*nw]   (lazy; exact split / rotate_overlaps)
    cplx* d_dN = nullptr;    // [nk][nn][nw]      diag N
    double* d_fro = nullptr; // [nk][nn]          ||N||_F^2 (or ||MU||_F^2)
    const double* fro_src = nullptr;  // where pair_reduce finds ||.||^2: nullptr = d_fro, else partials
    int fro_parts = 1;
    const cplx* dn_src = nullptr;     // diag N left as partials by the rotation (summed by pair_reduce); nullptr = d_dN is final
    int dn_parts = 1;
    double* d_part = nullptr;   // [nk][nsums]
    double* d_pmin = nullptr;   // [nk]
    double* d_sums = nullptr;   // [nsums]
    double* d_res = nullptr;    // [nres]
    double* d_min = nullptr;    // [1]
    double* d_scal = nullptr;   // [2048] scratch scalars (errors, flags)
    double* d_red = nullptr;    // block partials + results of the fused inner-product kernel (lazy)
    double* d_zpart = nullptr; size_t zpart_elems = 0;   // per-warp norm partials of the tensor GEMM epilogues (lazy)
    // pinned host scratch (nres + 64 doubles).  Slots: [0, nres] results of the host entry points (res, then min |N_nn|);
    // inside the optimiser [8] retraction flag (int), [16] f, [17] min |N_nn|, [20] step length read by a captured
    // trial-point graph, [WB_PIN_CROSS, +22) inner products, [48, 55) phase times of tiny.cu (-DWB_TINY_TIMING builds)
    double* h_pinned = nullptr;
    // generic scratch for stiefel / xy (lazy, sized on demand)
    cplx* d_tmp1 = nullptr; size_t tmp1_elems = 0;
    cplx* d_tmp2 = nullptr; size_t tmp2_elems = 0;
    cplx* d_XY = nullptr;  cplx* d_GXY = nullptr;   // [nk][nw*nw+nb*nw]

    // optional per-class kernel timing (wb200_profile)
    int profiling = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events[4];
    size_t prof_used[4] = {0, 0, 0, 0};

    // CUDA graphs of one evaluation (single-slab ctx): the ~9 launches of a small problem cost more than
    // its arithmetic; a graph replays them in one submission.  Keyed by everything the captured launches
    // bake in; a few entries because the optimiser alternates between trial buffers.
    struct EvalGraph {
        const void* dU = nullptr; const void* dres = nullptr; const void* dG = nullptr;
        int exact_split = 0, has_penalty = 0; double lambda = 0.0; const void* frozen = nullptr;
        cudaGraphExec_t exec = nullptr; long long nlaunch = 0; unsigned long long stamp = 0; int seen = 0;
    };
    EvalGraph graphs[6];
    // the optimiser's whole trial point (step, retraction, evaluation, projection, inner product) for small problems
    struct TrialGraph {
        const void *x = nullptr, *s = nullptr, *xn = nullptr, *gn = nullptr;
        int mode = 0, has_penalty = 0; double lambda = 0.0; const void* frozen = nullptr;
        cudaGraphExec_t exec = nullptr; long long nlaunch = 0; unsigned long long stamp = 0; int seen = 0;
    };
    TrialGraph tgraphs[4];
    // tiny.cu: the whole trial point in one cooperative launch (n_bands <= 16, one CTA per k-point)
    int tiny_ok = 0;           // shape served (decided at creation: the plain overlaps d_M are kept then)
    int tiny_blocks = 0;       // grid of the fused kernel (0: not planned yet); the whole grid must be co-resident
    int tiny_kper = 1;         // k-points per CTA
    int tiny_mu_global = 0;    // the rotated overlaps of a k-point wait for the gradient in d_MU instead of shared memory
    double* d_tiny = nullptr;  // [nk] per-k inner products, then the retraction flag word
    int tiny_epoch = 0;
    unsigned long long graph_clock = 0;
    int use_graphs = 1;        // WB200_GRAPH=0 disables
    unsigned zd_attr_mask = 0; // zgemm_dmma.cu: shared-memory attribute set for these instantiations
    int rot_attr_done = 0;     // rotate_dmma.cu: the same for the rotation kernel of this ctx's shape
    unsigned attr_mask = 0;    // other kernels that need more than 48 KB of dynamic shared memory (bit per kernel)
    int split_valid = 1;       // the last phase 1 produced a valid OmegaI / OmegaOD / OmegaD split

    DmmaPlan* dmma = nullptr;  // tensor-path state (nullptr: CUDA-core rotation)
    int rotation_kernel = 0;

    int nsums() const { return 4 * nw + 2; }
    int nres() const { return 8 + 4 * nw; }
};

#define WB_CUDA(ctx, call)                                                                      \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                    \
            return WB200_ERR_CUDA;                                                              \
        }                                                                                       \
    } while (0)

#define WB_TRY(expr)                  \
    do {                              \
        int s_ = (expr);              \
        if (s_ != WB200_OK) return s_; \
    } while (0)

#define WB_LAUNCH_CHECK(ctx)                                                 \
    do {                                                                     \
        (ctx)->launches++;                                                   \
        cudaError_t e_ = cudaGetLastError();                                 \
        if (e_ != cudaSuccess) {                                             \
            (ctx)->err = std::string("kernel launch: ") + cudaGetErrorString(e_) + " at " + \
                         __FILE__ + ":" + std::to_string(__LINE__);          \
            return WB200_ERR_CUDA;                                           \
        }                                                                    \
    } while (0)

// RAII bracket: records start/stop events around a kernel class when profiling is on
struct ProfScope {
    wb200_ctx* c; int cls; cudaEvent_t stop = nullptr;
    ProfScope(wb200_ctx* ctx, int k) : c(ctx), cls(k) {
        if (!c->profiling) return;
        auto& v = c->prof_events[cls];
        if (c->prof_used[cls] == v.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            v.emplace_back(a, b);
        }
        auto& pr = v[c->prof_used[cls]++];
        cudaEventRecord(pr.first, c->stream);
        stop = pr.second;
    }
    ~ProfScope() { if (stop) cudaEventRecord(stop, c->stream); }
};

// ---- mempool.cu: process-wide caching of device / page-locked buffers and streams ----
cudaError_t wb_dev_alloc_raw(void** p, size_t bytes);          // on the current device
template <class T> inline cudaError_t wb_dev_alloc(T** p, size_t bytes) { return wb_dev_alloc_raw((void**)p, bytes); }
void wb_dev_free(void* p);                                     // drains the block's device first, like cudaFree
void wb_dev_free_drained(void* p);                             // the caller has drained the device
cudaError_t wb_pinned_alloc_raw(void** p, size_t bytes);
template <class T> inline cudaError_t wb_pinned_alloc(T** p, size_t bytes) { return wb_pinned_alloc_raw((void**)p, bytes); }
void wb_pinned_free(void* p);
cudaError_t wb_stream_acquire(cudaStream_t* s);                // non-blocking stream of the current device
void wb_stream_release(cudaStream_t s);                        // idle streams only
int wb_device_info(int device, int* major, int* minor, int* num_sms);

// ---- api.cu ----
int wb_allreduce(wb200_ctx* ctx, double* dptr, int count, int op);   // checked call of the registered collective
int wb_ensure_pipeline(wb200_ctx* ctx);   // copy streams / events of the host-pointer entry points

// ---- peer.cu ----
#define WB_COMM_MAXR 32
int wb_peer_allreduce(wb200_ctx* ctx, double* dptr, int count, int op, int count2, int op2);
bool wb_allreduce_is_builtin(const wb200_ctx* ctx);
void wb_allreduce_install_builtin(wb200_ctx* ctx);
int wb_peer_signal(wb200_ctx* ctx, int which, unsigned long long epoch);
int wb_peer_wait(wb200_ctx* ctx, int which, unsigned long long epoch, cudaStream_t on);
int wb_peer_pull_packed(wb200_ctx* ctx);

// ---- kernels_generic.cu ----
int wb_zgemm_batched(wb200_ctx* ctx, const ZgemmArgs& a);
int wb_rotate_generic(wb200_ctx* ctx, const cplx* dU, int ka, int kb);  // MU, dN, fro(=||MU||^2) of local k in [ka, kb)
int wb_form_N(wb200_ctx* ctx, const cplx* dU);                      // N = U_k^H MU, fro = ||N||^2
int wb_pair_reduce(wb200_ctx* ctx, double* dsums);                  // part -> sums (+ min)
int wb_pair_partial(wb200_ctx* ctx, int ka, int kb);                // per-k partial sums of local k in [ka, kb)
int wb_sum_over_k(wb200_ctx* ctx, double* dsums);
int wb_finalize(wb200_ctx* ctx, const double* dsums, double* dres);
int wb_grad(wb200_ctx* ctx, const double* dsums, cplx* dG, int ka, int kb);
int wb_ensure_tmp(wb200_ctx* ctx, size_t e1, size_t e2);
int wb_ensure_N(wb200_ctx* ctx);
int wb_halo_fetch(wb200_ctx* ctx, cplx* dU);   // copy the remote rows of `need` into the local array (CUDA-core path)

// ---- stiefel.cu ----
int wb_project_tangent_dev(wb200_ctx* ctx, const cplx* dX, cplx* dG, int nrow, int ncol, int count);
// tangent_step: X = X0 + a s with X0 on the manifold and s tangent at X0, hence X^H X = I + a^2 s^H s: every
// singular value is >= 1, which the scaled Newton-Schulz iteration of the large-shape path exploits
struct WbTangentGram { const cplx* S; double alpha, fro_max; };   // S_k = s_k^H s_k of the search direction, max_k ||S_k||_F
int wb_retract_dev(wb200_ctx* ctx, cplx* dX, int nrow, int ncol, int count, int* iters, bool tangent_step = false,
                   const WbTangentGram* gram = nullptr);
int wb_tangent_gram(wb200_ctx* ctx, const cplx* ds, int nrow, int ncol, int count, cplx* dS, double* fro_max, bool* ok);
// st: distance in complex elements between the [X_k; Y_k] blocks of consecutive k (0: nw*nw + nb*nw, the
// packed XY array; the MagModel array interleaves the up and down blocks, st = 2 (nw*nw + nb*nw))
int wb_retract_xy(wb200_ctx* ctx, cplx* dXY, bool tangent_step = false, long long st = 0);
int wb_project_xy(wb200_ctx* ctx, const cplx* dXY, cplx* dG, long long st = 0);
int wb_xy_to_u(wb200_ctx* ctx, const cplx* dXY, cplx* dU, long long st = 0);
int wb_gu_to_gxy(wb200_ctx* ctx, const cplx* dXY, const cplx* dGU, cplx* dGXY, long long st = 0);

// ---- stiefel_small.cu ----
bool wb_small_applicable(int nrow, int ncol);
int wb_retract_small(wb200_ctx* ctx, cplx* dX, int nrow, int ncol, int ld, long long stride, int count);
bool wb_retract_pending_failed(wb200_ctx* ctx);
int wb_project_small(wb200_ctx* ctx, const cplx* dX, cplx* dG, int nrow, int ncol, int ld, long long stride,
                     int count);

// ---- tiny.cu ----
bool wb_tiny_shape(int nb, int nw, int nk_total, int nk, int nn);
bool wb_tiny_usable(wb200_ctx* ctx, int mode);
int wb_tiny_trial(wb200_ctx* ctx, int mode, const cplx* x, const cplx* s, double alpha, cplx* xn, cplx* gn);

// ---- vecops.cu ----
#define WB_CROSS_NA 3
#define WB_CROSS_NB 7
#define WB_CROSS_NOUT (WB_CROSS_NA * WB_CROSS_NB + 1)   // sums [na][7] row-major, then max |a_0|^2
#define WB_PIN_CROSS 24                                 // offset of the results in ctx->h_pinned
struct WbCrossArgs {
    const double* a[WB_CROSS_NA];
    const double* b[WB_CROSS_NB];
    int na, nb;
    int a_in_b;   // a[i] is b[i] for i < na: read once
};
struct WbLinArgs {
    const double* v[WB_CROSS_NB];
    double c[WB_CROSS_NB];
    int nv;
    double beta;
    const double* cdev;   // optional: c[cdev_idx] is read from device memory (a captured graph's step length)
    int cdev_idx;
};
int wb_cross_launch(wb200_ctx* ctx, const WbCrossArgs& p, size_t n);   // results in h_pinned after a sync
int wb_lincomb(wb200_ctx* ctx, const WbLinArgs& p, double* out, size_t n);   // out = beta out + sum c_j v_j
int wb_secant(wb200_ctx* ctx, double alpha, const double* d, const double* gnew, const double* g, double* s,
              double* y, size_t n);                                    // s = alpha d, y = gnew - g
int wb_axpby(wb200_ctx* ctx, double alpha, const double* x, double beta, double* y, size_t n);  // y = a x + b y

// ---- rotate_dmma.cu ----
int wb_dmma_plan_create(wb200_ctx* ctx, const double* hM);   // may leave ctx->dmma == nullptr if not applicable
void wb_dmma_plan_destroy(wb200_ctx* ctx);
int wb_pack_dmma(wb200_ctx* ctx, const cplx* dU, int ja, int jb);     // pack need[ja:jb) of the gauge
// MU, dN, fro(=||MU||^2) of local k in [ka, kb); phase 1 / 2: only the pairs whose neighbour is local / remote
int wb_rotate_dmma(wb200_ctx* ctx, const cplx* dU, int ka, int kb, int phase = 0);
bool wb_dmma_phases_supported(const wb200_ctx* ctx);
int wb_pack_dmma_local(wb200_ctx* ctx, const cplx* dU);   // the rows this rank owns (split need list)
int wb_dmma_packed_info(wb200_ctx* ctx, void** dptr, size_t* bytes, size_t* row_bytes);

// ---- zgemm_dmma.cu: batched complex GEMM on the FP64 tensor pipe from plain column-major operands ----
bool wb_zgemm_tensor_applicable(int m, int n, int k);
long long wb_zgemm_tensor_items(int m, int n, int batch);
// epi 0: C = alpha op(A) op(B) + beta C;  1: also per-warp ||AB||_F^2 partials into part;
// 2: C = alpha I - beta AB and per-warp ||AB - I||_F^2 partials (Newton-Schulz step matrix);
// 3: C = AB and the same partials
int wb_zgemm_tensor(wb200_ctx* ctx, const ZgemmArgs& a, int epi, double* part);
int wb_zgemm_tensor_err(wb200_ctx* ctx, int m, int n, int batch, const double* part, double* derr);
int wb_zgemm_tensor_fro(wb200_ctx* ctx, int m, int n, int batch, const double* part, double* dfro);
int wb_ensure_zpart(wb200_ctx* ctx, size_t elems);
// tensor pipe where the shape allows it (and WB200_FLAG_NO_TENSOR is not set), CUDA cores otherwise
int wb_zgemm(wb200_ctx* ctx, const ZgemmArgs& a);

// ---- optimizer.cu ----
int wb_lbfgs(wb200_ctx* ctx, int mode /*0 maxloc, 1 disentangle*/, cplx* dx, double f_tol,
             double g_tol, int max_iter, int history, double* f_final, int* iters, int* n_fg);
// the same driver on a caller-defined manifold problem (mag.cu): vectors of `ncplx` complex numbers on
// ctx's device and stream; eval leaves f in ctx->d_res[5] and writes the Euclidean gradient
struct WbCustomProblem {
    size_t ncplx = 0;
    std::function<int(cplx*, bool)> retract;             // (x, tangent_step)
    std::function<int(const cplx*, cplx*)> project;      // (x, g): g <- projection of g on the tangent space at x
    std::function<int(const cplx*, cplx*)> eval;         // (x, g)
};
int wb_lbfgs_custom(wb200_ctx* ctx, const WbCustomProblem& cp, cplx* dx, double f_tol, double g_tol,
                    int max_iter, int history, double* f_final, int* iters, int* n_fg);
