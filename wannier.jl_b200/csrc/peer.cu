// peer.cu — the collectives of a k-sharded run done over peer memory (NVLink / NVSwitch) by the library
// itself, one slab per GPU:
//   * readiness flags: "my gauge rows are in place" / "my rows are packed", published with remote stores
//     into every peer's comm buffer and awaited on the consumer's own memory (no collective in front of an
//     evaluation);
//   * the halo of the tensor path: a slab copies the PACKED rows it needs (the B-operand image the owner
//     built for itself anyway) out of the owners' packed arrays with the copy engines, on a second stream,
//     under the rotation of the pairs whose neighbour is local — no SM ever packs a row twice and no SM
//     waits for NVLink;
//   * the all-reduce of the spread sums and of the optimiser's scalars (4 nw + 2 or a few dozen doubles):
//     one small kernel per rank pushes its contribution into a slot of every peer's buffer, waits for the
//     peers' flags and sums the slots in rank order — bitwise identical on every rank and independent of
//     timing, ~10 us instead of an NCCL launch.
// The reference has no counterpart (single process; the sums are plain loops over k, src/spread.jl:284-322);
// this replaces torch.distributed / NCCL calls of the round-1 design on the hot path.
//
// Comm buffer of a rank (wb200_peer_comm_bytes(nw) bytes, zeroed by the caller, mapped by every peer):
//   words [0, 32)        ready flags   [src rank]   evaluation counter of src: its gauge rows are written
//   words [32, 64)       packed flags  [src rank]   evaluation counter of src: its rows are packed
//   words [64, 128)      all-reduce flags [parity][src rank]
//   words [128, ...)     all-reduce slots [parity][src rank][maxc] doubles, maxc = 4 nw + 64
#include "common.cuh"

namespace {

constexpr int MAXR = WB_COMM_MAXR;
constexpr int OFF_READY = 0, OFF_PACKED = MAXR, OFF_ARFLAG = 2 * MAXR, OFF_SLOTS = 4 * MAXR;

__device__ __forceinline__ void spin_until(const volatile unsigned long long* f, unsigned long long epoch,
                                           unsigned long long timeout_ns) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*f < epoch) {   // bounded (20 s unless WB200_SPIN_TIMEOUT_MS says otherwise): a dead peer must not hang the device
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) __trap();
        __nanosleep(100);
    }
}

// one thread per peer: publish this rank's counter in the peer's buffer (a remote store over NVLink)
__global__ void flag_signal_kernel(int world, int rank, unsigned long long* const* __restrict__ comm, int off,
                                   unsigned long long epoch) {
    const int q = threadIdx.x;
    if (q >= world || q == rank) return;
    __threadfence_system();
    *(volatile unsigned long long*)(comm[q] + off + rank) = epoch;
}
// one thread per peer this slab reads from: spin on the LOCAL flag word until that peer has published `epoch`
__global__ void flag_wait_kernel(int world, unsigned mask, const unsigned long long* mine, int off,
                                 unsigned long long epoch, unsigned long long timeout_ns) {
    const int q = threadIdx.x;
    if (q >= world || !((mask >> q) & 1u)) return;
    spin_until(mine + off + q, epoch, timeout_ns);
    __threadfence_system();
}

// op: 0 sum, 1 max, 2 min.  count2 > 0: elements [count, count + count2) are reduced with op2 (one launch for
// the optimiser's "sums + max" pair).
__global__ void __launch_bounds__(256) peer_allreduce_kernel(int world, int rank, unsigned long long* const* __restrict__ comm,
                                                             int maxc, unsigned long long epoch, double* __restrict__ x,
                                                             int count, int op, int count2, int op2,
                                                             unsigned long long timeout_ns) {
    const int par = (int)(epoch & 1ull);
    const int n = count + count2;
    // 1. my contribution into slot [par][rank] of every rank's buffer (my own included)
    for (int q = 0; q < world; q++) {
        double* slot = reinterpret_cast<double*>(comm[q] + OFF_SLOTS) + ((size_t)par * MAXR + rank) * maxc;
        for (int i = threadIdx.x; i < n; i += blockDim.x) slot[i] = x[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        __threadfence_system();
        *(volatile unsigned long long*)(comm[threadIdx.x] + OFF_ARFLAG + par * MAXR + rank) = epoch;
    }
    // 2. every contribution has landed in MY buffer
    if (threadIdx.x < world) spin_until(comm[rank] + OFF_ARFLAG + par * MAXR + threadIdx.x, epoch, timeout_ns);
    __syncthreads();
    __threadfence_system();
    // 3. reduce in rank order (identical on every rank)
    const double* slots = reinterpret_cast<const double*>(comm[rank] + OFF_SLOTS) + (size_t)par * MAXR * maxc;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int o = i < count ? op : op2;
        double acc = __ldcg(slots + i);
        for (int q = 1; q < world; q++) {
            const double v = __ldcg(slots + (size_t)q * maxc + i);
            acc = o == 0 ? acc + v : (o == 1 ? fmax(acc, v) : fmin(acc, v));
        }
        x[i] = acc;
    }
}

// the pull as a kernel (a few CTAs) instead of a copy-engine transfer, for slabs that share a device: copy engines
// run their queues in order, so a pull parked behind its wait kernel would hold up the OTHER slab's host copies
__global__ void __launch_bounds__(256) peer_copy_kernel(double2* __restrict__ dst, const double2* __restrict__ src, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// the same all-reduce in two kernels with a stream dependency (CUDA events) between them instead of a spin — for the
// slabs of one process
__global__ void __launch_bounds__(256) peer_push_kernel(int world, int rank, unsigned long long* const* __restrict__ comm,
                                                        int maxc, int par, const double* __restrict__ x, int n) {
    for (int q = 0; q < world; q++) {
        double* slot = reinterpret_cast<double*>(comm[q] + OFF_SLOTS) + ((size_t)par * MAXR + rank) * maxc;
        for (int i = threadIdx.x; i < n; i += blockDim.x) slot[i] = x[i];
    }
    __threadfence_system();
}
__global__ void __launch_bounds__(256) peer_sum_kernel(int world, const unsigned long long* __restrict__ mine, int maxc, int par,
                                                       double* __restrict__ x, int count, int op, int count2, int op2) {
    const double* slots = reinterpret_cast<const double*>(mine + OFF_SLOTS) + (size_t)par * MAXR * maxc;
    for (int i = threadIdx.x; i < count + count2; i += blockDim.x) {
        const int o = i < count ? op : op2;
        double acc = __ldcg(slots + i);
        for (int q = 1; q < world; q++) {
            const double v = __ldcg(slots + (size_t)q * maxc + i);
            acc = o == 0 ? acc + v : (o == 1 ? fmax(acc, v) : fmin(acc, v));
        }
        x[i] = acc;
    }
}

int inproc_barrier(wb200_ctx* ctx) {
    if (!ctx->inproc->barrier()) {
        ctx->err = "all-reduce callback failed: another slab of this call reported an error";
        return WB200_ERR_CUDA;
    }
    return WB200_OK;
}

int builtin_allreduce(void* dptr, int count, int op, void* /*stream*/, void* user) {
    return wb_peer_allreduce((wb200_ctx*)user, (double*)dptr, count, op, 0, 0) == WB200_OK ? 0 : 1;
}

}  // namespace

int wb_peer_allreduce(wb200_ctx* ctx, double* dptr, int count, int op, int count2, int op2) {
    if (!ctx->d_comm) {
        ctx->err = "peer all-reduce: no comm buffers (wb200_set_peer_comm)";
        return WB200_ERR_ARG;
    }
    if (count + count2 > ctx->comm_maxc) {
        ctx->err = "peer all-reduce: more than 4 nw + 64 doubles";
        return WB200_ERR_ARG;
    }
    ctx->ar_epoch++;
    if (ctx->inproc) {
        // push my contribution into every slab's slot; once EVERY slab has recorded its push (host rendezvous of the
        // slab threads — the streams are not drained), this stream waits for those events and sums in rank order
        const int par = (int)(ctx->ar_epoch & 1ull);
        peer_push_kernel<<<1, 256, 0, ctx->stream>>>(ctx->peer_world, ctx->peer_rank, ctx->d_comm, ctx->comm_maxc, par, dptr,
                                                     count + count2);
        WB_LAUNCH_CHECK(ctx);
        WB_CUDA(ctx, cudaEventRecord(ctx->inproc->ev_ar[par][ctx->peer_rank], ctx->stream));
        WB_TRY(inproc_barrier(ctx));
        for (int q = 0; q < ctx->peer_world; q++)
            if (q != ctx->peer_rank) WB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->inproc->ev_ar[par][q], 0));
        peer_sum_kernel<<<1, 256, 0, ctx->stream>>>(ctx->peer_world, ctx->my_comm, ctx->comm_maxc, par, dptr, count, op, count2, op2);
        WB_LAUNCH_CHECK(ctx);
        return WB200_OK;
    }
    peer_allreduce_kernel<<<1, 256, 0, ctx->stream>>>(ctx->peer_world, ctx->peer_rank, ctx->d_comm, ctx->comm_maxc, ctx->ar_epoch,
                                                      dptr, count, op, count2, op2, ctx->spin_timeout_ns);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}

bool wb_allreduce_is_builtin(const wb200_ctx* ctx) { return ctx->allreduce == builtin_allreduce; }
void wb_allreduce_install_builtin(wb200_ctx* ctx) {
    if (ctx->d_comm && !ctx->allreduce) { ctx->allreduce = builtin_allreduce; ctx->allreduce_user = ctx; }
}

// publish flag word `which` (0 ready, 1 packed) on ctx->stream; wait for the owners of the halo rows on `on`
int wb_peer_signal(wb200_ctx* ctx, int which, unsigned long long epoch) {
    if (ctx->inproc) {
        auto& ev = which ? ctx->inproc->ev_packed : ctx->inproc->ev_ready;
        WB_CUDA(ctx, cudaEventRecord(ev[ctx->peer_rank], ctx->stream));
        return WB200_OK;
    }
    flag_signal_kernel<<<1, 32, 0, ctx->stream>>>(ctx->peer_world, ctx->peer_rank, ctx->d_comm, which ? OFF_PACKED : OFF_READY, epoch);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}
int wb_peer_wait(wb200_ctx* ctx, int which, unsigned long long epoch, cudaStream_t on) {
    if (ctx->inproc) {   // every slab has recorded its event of this evaluation once the slab threads have met
        WB_TRY(inproc_barrier(ctx));
        auto& ev = which ? ctx->inproc->ev_packed : ctx->inproc->ev_ready;
        for (int q = 0; q < ctx->peer_world; q++)
            if ((ctx->peer_mask >> q) & 1u) WB_CUDA(ctx, cudaStreamWaitEvent(on, ev[q], 0));
        return WB200_OK;
    }
    if (!ctx->peer_mask) return WB200_OK;
    flag_wait_kernel<<<1, 32, 0, on>>>(ctx->peer_world, ctx->peer_mask, ctx->my_comm, which ? OFF_PACKED : OFF_READY, epoch,
                                      ctx->spin_timeout_ns);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}

// the packed rows this slab reads from peers: copy-engine pulls out of the owners' packed arrays (contiguous
// runs of k-points with one owner), on the halo stream
int wb_peer_pull_packed(wb200_ctx* ctx) {
    void* mine = nullptr;
    size_t bytes = 0, row = 0;
    WB_TRY(wb_dmma_packed_info(ctx, &mine, &bytes, &row));
    for (const auto& r : ctx->halo_runs) {
        const size_t off = (size_t)r.k0 * row;
        if (ctx->peer_pull_kernel) {
            peer_copy_kernel<<<32, 256, 0, ctx->halo_stream>>>((double2*)((char*)mine + off),
                                                               (const double2*)((const char*)ctx->h_peerUp[r.owner] + off),
                                                               (size_t)r.count * row / 16);
            WB_LAUNCH_CHECK(ctx);
            continue;
        }
        WB_CUDA(ctx, cudaMemcpyAsync((char*)mine + off, (const char*)ctx->h_peerUp[r.owner] + off, (size_t)r.count * row,
                                     cudaMemcpyDeviceToDevice, ctx->halo_stream));
    }
    return WB200_OK;
}

extern "C" size_t wb200_peer_comm_bytes(int nw) {
    return sizeof(double) * ((size_t)OFF_SLOTS + 2 * (size_t)MAXR * (4 * (size_t)(nw > 0 ? nw : 0) + 64));
}

extern "C" int wb200_set_peer_comm(wb200_ctx* ctx, int world, void* const* comm_ptrs) {
    if (!ctx) return WB200_ERR_ARG;
    cudaSetDevice(ctx->device);
    WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_comm) { wb_dev_free((void*)ctx->d_comm); ctx->d_comm = nullptr; }
    if (wb_allreduce_is_builtin(ctx)) { ctx->allreduce = nullptr; ctx->allreduce_user = nullptr; }
    ctx->my_comm = nullptr;
    ctx->epoch = 0;
    ctx->ar_epoch = 0;
    if (!comm_ptrs || world <= 0) return WB200_OK;
    if (world != ctx->peer_world || world > MAXR) {
        ctx->err = "wb200_set_peer_comm: call wb200_set_peer_gauges first, with the same world size (at most 32)";
        return WB200_ERR_ARG;
    }
    WB_CUDA(ctx, wb_dev_alloc((void**)&ctx->d_comm, sizeof(void*) * world));
    WB_CUDA(ctx, cudaMemcpy((void*)ctx->d_comm, comm_ptrs, sizeof(void*) * world, cudaMemcpyHostToDevice));
    ctx->my_comm = (unsigned long long*)comm_ptrs[ctx->peer_rank];
    ctx->comm_maxc = 4 * ctx->nw + 64;
    { const char* e = getenv("WB200_SPIN_TIMEOUT_MS"); ctx->spin_timeout_ns = 1000000ull * (unsigned long long)(e && atoll(e) > 0 ? atoll(e) : 20000); }
    wb_allreduce_install_builtin(ctx);   // the library's own collective, unless the caller registered one
    // streams and events of the host-pointer entry points now, not lazily in the middle of a collective evaluation:
    // cudaStreamCreate can wait for the device to drain, and a peer's kernel may be spinning on this slab's next launch
    if (ctx->nn > 0) WB_TRY(wb_ensure_pipeline(ctx));
    return WB200_OK;
}

extern "C" int wb200_peer_allreduce_dev(wb200_ctx* ctx, double* dptr, int count, int op) {
    if (!ctx || !dptr || count <= 0 || op < 0 || op > 2) return WB200_ERR_ARG;
    cudaSetDevice(ctx->device);
    return wb_peer_allreduce(ctx, dptr, count, op, 0, 0);
}

extern "C" int wb200_packed_gauge(wb200_ctx* ctx, void** dptr, size_t* bytes) {
    if (!ctx || !dptr || !bytes) return WB200_ERR_ARG;
    size_t row = 0;
    return wb_dmma_packed_info(ctx, dptr, bytes, &row);
}

extern "C" int wb200_set_peer_packed(wb200_ctx* ctx, int world, void* const* Up_ptrs) {
    if (!ctx) return WB200_ERR_ARG;
    cudaSetDevice(ctx->device);
    WB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->halo_stream) WB_CUDA(ctx, cudaStreamSynchronize(ctx->halo_stream));
    ctx->h_peerUp.clear();
    if (!Up_ptrs || world <= 0) return WB200_OK;
    if (world != ctx->peer_world) {
        ctx->err = "wb200_set_peer_packed: call wb200_set_peer_gauges first, with the same world size";
        return WB200_ERR_ARG;
    }
    ctx->h_peerUp.assign(Up_ptrs, Up_ptrs + world);
    { const char* e = getenv("WB200_PEER_PULL"); ctx->peer_pull_kernel = (e && e[0] == 'k') ? 1 : 0; }   // =kernel
    return WB200_OK;
}

extern "C" int wb200_ipc_get_handle(int device, void* dptr, unsigned char handle[64]) {
    if (!dptr || !handle || cudaSetDevice(device) != cudaSuccess) return WB200_ERR_ARG;
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, dptr) != cudaSuccess) { cudaGetLastError(); return WB200_ERR_CUDA; }
    memcpy(handle, &h, 64);
    return WB200_OK;
}
