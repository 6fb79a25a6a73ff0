// vecops.cu — flat-vector kernels of the device-resident optimiser (optimizer.cu).
//
// The L-BFGS driver that replaces Optim.optimize(... LBFGS(manifold, HagerZhang, m)) at
// src/localization/max_localize.jl:73-84 / disentangle.jl:702-713 / opt_rotate.jl:118-141 works on
// vectors of nk*nb*nw complex numbers (1.07 GB each at the north-star shape), so its cost is HBM
// passes and host round trips, not flops.  Three fused kernels keep both small:
//   * cross:   every inner product <a_i, b_j> of up to 3 x 7 vectors in ONE pass over memory (the
//              Gram rows the two-loop recursion needs) plus max |a_0| (Optim's g_tol test), reduced
//              in a fixed order (bitwise deterministic) and, in k-sharded runs, all-reduced;
//   * lincomb: out = beta out + sum_j c_j v_j of up to 7 vectors in one pass (the search direction
//              straight from the two-loop coefficients, the trial point x + alpha d);
//   * secant:  s = alpha d and y = g_new - g in one pass.
// Complex arrays are treated as arrays of double2; inner products are the real ones Optim uses on
// complex arrays, real(dot(a, b)).
#include "common.cuh"

namespace {

constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = 148 * 8;

// What the host reads after an evaluation of the optimiser, written straight into page-locked host memory by the
// last kernel of the chain (zero-copy stores) instead of four 8-byte device-to-host copies behind it: each copy is a
// node of its own in the stream and costs more than the kernels of a 64-k-point problem (profiles/r4_summary.md).
struct WbPublish {
    double* hout;          // h_pinned + WB_PIN_CROSS (nullptr: nothing is published)
    const double* f;       // d_res + 5   -> h_pinned[16]
    const double* dmin;    // d_min       -> h_pinned[17]
    const int* flag;       // retraction flag -> h_pinned[8] (as int)
    double* hbase;         // h_pinned
};

__device__ __forceinline__ void publish_result(const WbPublish& pub, int slot, double r) {
    if (!pub.hout) return;
    pub.hout[slot] = r;
    if (slot == 0) {
        if (pub.f) pub.hbase[16] = *pub.f;
        if (pub.dmin) pub.hbase[17] = *pub.dmin;
        if (pub.flag) *reinterpret_cast<int*>(pub.hbase + 8) = *pub.flag;
    }
}

// NA x NB inner products in one pass; A_IN_B: a[i] is b[i] (read once).  Launched with ONE block (short vectors: the
// reference's own problem sizes) the block's result is the final one: it is written to `final_out` and published,
// and the second stage is not launched.
template <int NA, int NB, bool A_IN_B>
__global__ void __launch_bounds__(RED_THREADS) cross_stage1(WbCrossArgs p, size_t n2,
                                                            double* __restrict__ partial,
                                                            double* __restrict__ final_out, WbPublish pub) {
    constexpr int NOUT = NA * NB + 1;
    double acc[NA][NB];
    double mx = 0.0;
#pragma unroll
    for (int i = 0; i < NA; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) acc[i][j] = 0.0;
    for (size_t e = blockIdx.x * (size_t)RED_THREADS + threadIdx.x; e < n2;
         e += (size_t)gridDim.x * RED_THREADS) {
        double2 b[NB], a[NA];
#pragma unroll
        for (int j = 0; j < NB; j++) b[j] = reinterpret_cast<const double2*>(p.b[j])[e];
#pragma unroll
        for (int i = 0; i < NA; i++) a[i] = A_IN_B ? b[i < NB ? i : 0] : reinterpret_cast<const double2*>(p.a[i])[e];
        mx = fmax(mx, a[0].x * a[0].x + a[0].y * a[0].y);
#pragma unroll
        for (int i = 0; i < NA; i++)
#pragma unroll
            for (int j = 0; j < NB; j++) acc[i][j] = fma(a[i].x, b[j].x, fma(a[i].y, b[j].y, acc[i][j]));
    }
    __shared__ double sh[RED_THREADS / 32][NOUT];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < NA; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) {
            const double v = warp_sum(acc[i][j]);
            if (lane == 0) sh[warp][i * NB + j] = v;
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) sh[warp][NOUT - 1] = mx;
    __syncthreads();
    if (threadIdx.x < NOUT) {
        const int q = threadIdx.x;
        double t = 0.0;
        for (int w = 0; w < RED_THREADS / 32; w++) t = (q == NOUT - 1) ? fmax(t, sh[w][q]) : t + sh[w][q];
        // results in the fixed layout [na][WB_CROSS_NB] + max, whatever NB is
        const int slot = (q == NOUT - 1) ? WB_CROSS_NOUT - 1 : (q / NB) * WB_CROSS_NB + (q % NB);
        if (final_out) {
            final_out[slot] = t;
            publish_result(pub, slot, t);
        } else {
            partial[(size_t)slot * gridDim.x + blockIdx.x] = t;
        }
    }
}

// one CTA per output: sums (or maximises) the block partials of that output in a fixed order
__global__ void __launch_bounds__(RED_THREADS) cross_stage2(const double* __restrict__ partial, int blocks,
                                                            double* __restrict__ out, WbPublish pub) {
    const int q = blockIdx.x;
    const bool is_max = q == WB_CROSS_NOUT - 1;
    const double* src = partial + (size_t)q * blocks;
    double t = 0.0;
    for (int b = threadIdx.x; b < blocks; b += RED_THREADS) t = is_max ? fmax(t, src[b]) : t + src[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double v = __shfl_xor_sync(0xffffffffu, t, o);
        t = is_max ? fmax(t, v) : t + v;
    }
    __shared__ double sh[RED_THREADS / 32];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0.0;
        for (int w = 0; w < RED_THREADS / 32; w++) r = is_max ? fmax(r, sh[w]) : r + sh[w];
        out[q] = r;
        publish_result(pub, q, r);
    }
}

__global__ void __launch_bounds__(256) lincomb_kernel(WbLinArgs p, double2* out, size_t n2) {
    if (p.cdev) p.c[p.cdev_idx] = *p.cdev;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n2;
         e += (size_t)gridDim.x * blockDim.x) {
        double2 r = make_double2(0.0, 0.0);
        if (p.beta != 0.0) {
            const double2 o = out[e];
            r = make_double2(p.beta * o.x, p.beta * o.y);
        }
#pragma unroll
        for (int j = 0; j < WB_CROSS_NB; j++)
            if (j < p.nv) {
                const double2 v = reinterpret_cast<const double2*>(p.v[j])[e];
                r.x = fma(p.c[j], v.x, r.x);
                r.y = fma(p.c[j], v.y, r.y);
            }
        out[e] = r;
    }
}

__global__ void __launch_bounds__(256) secant_kernel(double alpha, const double2* __restrict__ d,
                                                     const double2* __restrict__ gnew,
                                                     const double2* __restrict__ g, double2* __restrict__ s,
                                                     double2* __restrict__ y, size_t n2) {
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n2;
         e += (size_t)gridDim.x * blockDim.x) {
        const double2 dv = d[e], a = gnew[e], b = g[e];
        s[e] = make_double2(alpha * dv.x, alpha * dv.y);
        y[e] = make_double2(a.x - b.x, a.y - b.y);
    }
}

int stream_blocks(size_t n2) {
    const size_t want = (n2 + 255) / 256;
    return (int)(want < 148 * 16 ? (want ? want : 1) : 148 * 16);
}

}  // namespace

// Launches the cross products; the WB_CROSS_NOUT results (row-major [na][7] sums, then max |a_0|^2)
// are copied to ctx->h_pinned + WB_PIN_CROSS and are valid after the next stream synchronisation.
int wb_cross_launch(wb200_ctx* ctx, const WbCrossArgs& p, size_t n) {
    if (p.na < 1 || p.na > WB_CROSS_NA || p.nb < 1 || p.nb > WB_CROSS_NB || (n & 1) ||
        (p.a_in_b && p.na > p.nb)) {
        ctx->err = "wb_cross_launch: bad arguments";
        return WB200_ERR_ARG;
    }
    const size_t n2 = n / 2;
    const size_t want = (n2 + RED_THREADS - 1) / RED_THREADS;
    // up to 16 elements per thread: one block, one launch (a 64-k-point, 4-band vector has 1024 elements)
    const bool one_block = n2 <= (size_t)16 * RED_THREADS;
    const int blocks = one_block ? 1 : (int)(want < (size_t)RED_MAX_BLOCKS ? want : RED_MAX_BLOCKS);
    if (!ctx->d_red) {
        const size_t bytes = sizeof(double) * WB_CROSS_NOUT * (RED_MAX_BLOCKS + 1);
        WB_CUDA(ctx, wb_dev_alloc(&ctx->d_red, bytes));
        WB_CUDA(ctx, cudaMemsetAsync(ctx->d_red, 0, bytes, ctx->stream));   // slots a small instantiation never writes
    }
    double* out = ctx->d_red + (size_t)WB_CROSS_NOUT * RED_MAX_BLOCKS;
    // single slab: the results (and, for the optimiser's evaluations, f / min |N_nn| / the retraction flag) go to the
    // page-locked scratch by zero-copy stores; k-sharded runs reduce them first and copy
    WbPublish pub{nullptr, nullptr, nullptr, nullptr, ctx->h_pinned};
    if (!ctx->allreduce) {
        pub.hout = ctx->h_pinned + WB_PIN_CROSS;
        if (ctx->publish_eval) {
            pub.f = ctx->d_res + 5;
            pub.dmin = ctx->d_min;
            if (ctx->retract_pending) pub.flag = reinterpret_cast<const int*>(ctx->d_scal + 4);
        }
    }
    double* fin = one_block ? out : nullptr;
    // the shapes the optimiser uses get exact-size instantiations (no predicates, fewer registers);
    // anything else is padded to 3 x 7 with repeats of the first vector (extra products are ignored)
    WbCrossArgs a = p;
    if (p.na == 1 && p.nb == 1 && !p.a_in_b) {
        cross_stage1<1, 1, false><<<blocks, RED_THREADS, 0, ctx->stream>>>(a, n2, ctx->d_red, fin, pub);
    } else if (p.na == 1 && p.nb == 1) {
        cross_stage1<1, 1, true><<<blocks, RED_THREADS, 0, ctx->stream>>>(a, n2, ctx->d_red, fin, pub);
    } else if (p.na == 1 && p.nb == 3 && p.a_in_b) {
        cross_stage1<1, 3, true><<<blocks, RED_THREADS, 0, ctx->stream>>>(a, n2, ctx->d_red, fin, pub);
    } else if (p.na == 3 && p.nb == 7 && p.a_in_b) {
        cross_stage1<3, 7, true><<<blocks, RED_THREADS, 0, ctx->stream>>>(a, n2, ctx->d_red, fin, pub);
    } else {
        for (int i = p.na; i < WB_CROSS_NA; i++) a.a[i] = p.b[0];
        for (int j = p.nb; j < WB_CROSS_NB; j++) a.b[j] = p.b[0];
        if (p.a_in_b) {
            // the prefix property is lost by the padding unless a is read separately
            for (int i = 0; i < p.na; i++) a.a[i] = p.b[i];
            for (int i = p.na; i < WB_CROSS_NA; i++) a.a[i] = p.b[0];
        } else {
            for (int i = p.na; i < WB_CROSS_NA; i++) a.a[i] = p.a[0];
        }
        cross_stage1<WB_CROSS_NA, WB_CROSS_NB, false><<<blocks, RED_THREADS, 0, ctx->stream>>>(a, n2, ctx->d_red, fin, pub);
    }
    WB_LAUNCH_CHECK(ctx);
    if (!one_block) {
        cross_stage2<<<WB_CROSS_NOUT, RED_THREADS, 0, ctx->stream>>>(ctx->d_red, blocks, out, pub);
        WB_LAUNCH_CHECK(ctx);
    }
    if (pub.hout) return WB200_OK;
    if (ctx->allreduce) {   // k-sharded optimiser: the vectors are slab-local, the scalars global
        if (wb_allreduce_is_builtin(ctx)) {   // sums and the max in one peer-memory launch
            WB_TRY(wb_peer_allreduce(ctx, out, WB_CROSS_NOUT - 1, 0, 1, 1));
        } else {
            WB_TRY(wb_allreduce(ctx, out, WB_CROSS_NOUT - 1, 0));
            WB_TRY(wb_allreduce(ctx, out + WB_CROSS_NOUT - 1, 1, 1));
        }
    }
    WB_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + WB_PIN_CROSS, out, sizeof(double) * WB_CROSS_NOUT,
                                 cudaMemcpyDeviceToHost, ctx->stream));
    return WB200_OK;
}

int wb_lincomb(wb200_ctx* ctx, const WbLinArgs& p, double* out, size_t n) {
    if (p.nv < 0 || p.nv > WB_CROSS_NB || (n & 1)) {
        ctx->err = "wb_lincomb: bad arguments";
        return WB200_ERR_ARG;
    }
    lincomb_kernel<<<stream_blocks(n / 2), 256, 0, ctx->stream>>>(p, reinterpret_cast<double2*>(out), n / 2);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}

int wb_secant(wb200_ctx* ctx, double alpha, const double* d, const double* gnew, const double* g, double* s,
              double* y, size_t n) {
    secant_kernel<<<stream_blocks(n / 2), 256, 0, ctx->stream>>>(
        alpha, reinterpret_cast<const double2*>(d), reinterpret_cast<const double2*>(gnew),
        reinterpret_cast<const double2*>(g), reinterpret_cast<double2*>(s), reinterpret_cast<double2*>(y), n / 2);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}

// y = alpha x + beta y   (kept for the Stiefel projection)
int wb_axpby(wb200_ctx* ctx, double alpha, const double* x, double beta, double* y, size_t n) {
    WbLinArgs p{};
    p.v[0] = x; p.c[0] = alpha; p.nv = 1; p.beta = beta;
    return wb_lincomb(ctx, p, y, n);
}
