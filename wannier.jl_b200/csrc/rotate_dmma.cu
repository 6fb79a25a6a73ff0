// rotate_dmma.cu — the hot kernel: batched complex-FP64 rotation MU_kb = M_kb U_{k+b} on the FP64
// tensor pipe (DMMA.8x8x4, the only FP64 MMA sm_100a has; tcgen05 has no f64 kind), fused with
// the diag(N_kb) = diag(U_k^H MU_kb) and ||MU_kb||_F^2 reductions.
//
// Replaces compute_MU_UtMU! (src/spread.jl:164-195) for 32 < nb <= 256: one GEMM per (k,b) pair
// instead of two — everything omega!/omega_grad! need from N is its diagonal (spread.jl:307-312,
// 459-470), which is an O(nb nw) epilogue on the accumulator registers.
//
// Arithmetic: the kernel is bound by the FP64 tensor pipe (measured 36.9 TFLOP/s, profiles/), so
// the complex product uses Gauss' three real multiplications instead of four:
//     P1 = (Ar + Ai) Br,  P2 = Ar (Bi - Br),  P3 = Ai (Br + Bi);   Cr = P1 - P3,  Ci = P1 + P2
// A = M is constant, so the plane Ar + Ai is built ONCE with the tiled copy of M; the B planes
// (Br, Bi - Br, Br + Bi) are built by the per-evaluation gauge packing kernel.  24 instead of 32
// DMMA per k-step; the extra additions are exact to one rounding each (error stays O(eps |A||B|)).
//
// Structure (one CTA per (k, column tile), looping over the nn neighbours of k):
//   * M is re-laid-out ONCE at ctx creation into the exact shared-memory image of the A operand
//     (DMMA fragment order, three planes), so a pipeline stage is ONE contiguous 1-D bulk async
//     copy (cp.async.bulk -> UBLKCP, completion on an mbarrier); likewise the packed gauge.
//   * warps 8-11 are the producer warpgroup (setmaxnreg hands its registers to the consumers):
//     one elected lane drives a STAGES-deep ring of full/empty mbarriers
//     that never drains between the nn pairs of a k-point, and is primed BEFORE the U_k tile is
//     staged so the first chunks are in flight while the CTA sets up;
//     warps 0-7 are consumers: warp tile 32 x 16 complex = 4 x 2 DMMA tiles x (P1, P2, P3).
//   * U_k's column tile stays in shared memory (accumulator-fragment order) for the whole CTA, so
//     the epilogue of every pair is registers + LDS only; per-warp partial diagonals go straight
//     to global (summed in fixed order by pair_reduce_kernel) — no CTA-wide barrier anywhere
//     in the main loop, and results are bitwise deterministic.
#include "common.cuh"

struct DmmaPlan {
    bool small = false;           // nb <= 32: warp pair per (k,b) pair
    bool small_ring = true;       // ... operands through a bulk-copy ring (WB200_SMALL_RING=0: straight from global)
    long long small_ring_min = 0;   // fewer pairs than this: the launch-bound problem takes the plain kernel
    int num_sms = 148;
    int WM = 0, WN = 0, KC = 0;   // warp grid (8 consumer warps), k-chunk
    int NB = 0, NC = 0;           // padded rows, columns per CTA
    int nct = 0;                  // column tiles = ceil(nw / NC)
    int nch = 0;                  // k-chunks = ceil(nb / KC)
    size_t a_stage = 0, b_stage = 0;  // doubles per stage
    double* d_Mt = nullptr;       // [pairs][nch][a_stage]
    double* d_Up = nullptr;       // [nk_total][nct][nch][b_stage]
    cplx* d_dNp = nullptr;        // [pairs][WM][nw]  per-warp-row partial diagonals
    double* d_frop = nullptr;     // [pairs][nct][8]
    size_t smem = 0;
    double* d_froM = nullptr;     // [pairs] ||M_kb||_F^2 (nb == nw: the gauge-invariant norm of N_kb for unitary U)
};

namespace {

#ifdef WB_TIMING
__device__ long long* wb_dbg = nullptr;
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WB_DONE;\n"
        "bra WB_WAIT;\n"
        "WB_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
// the same with an L2 eviction policy (createpolicy): the overlaps are streamed once per evaluation
// (evict first), the packed gauges are re-read by every neighbour (evict last)
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- one-off / per-evaluation layout kernels ------------------------------------------------
// A-operand image of M.  Chunk c = [wm][ks][j(6)][lane] double2; the 12 doubles of a lane are
//   (re, im, re+im) of mt = 0..3 with row = wm*32 + mt*8 + lane/4, col = c*KC + ks*4 + lane%4.
__global__ void tile_overlaps_kernel(int nb, int WM, int KC, int nch, const cplx* __restrict__ M,
                                     double* __restrict__ Mt) {
    const long long pair = blockIdx.x;
    const int c = blockIdx.y;
    const int KS = KC / 4;
    const int items = WM * KS * 32;
    const cplx* src = M + pair * (long long)nb * nb;
    double* dst = Mt + (pair * nch + c) * (long long)(items * 12);
    for (int o = threadIdx.x; o < items; o += blockDim.x) {
        const int lane = o & 31, ks = (o >> 5) % KS, wm = (o >> 5) / KS;
        const int col = c * KC + ks * 4 + (lane & 3);
        double q[12];
#pragma unroll
        for (int mt = 0; mt < 4; mt++) {
            const int row = wm * 32 + mt * 8 + (lane >> 2);
            cplx v = make_double2(0.0, 0.0);
            if (row < nb && col < nb) v = src[row + (long long)col * nb];
            q[3 * mt] = v.x;
            q[3 * mt + 1] = v.y;
            q[3 * mt + 2] = v.x + v.y;
        }
        double2* d2 = reinterpret_cast<double2*>(dst) + ((long long)(wm * KS + ks) * 6) * 32 + lane;
#pragma unroll
        for (int j = 0; j < 6; j++) d2[j * 32] = make_double2(q[2 * j], q[2 * j + 1]);
    }
}
// B-operand image of U.  Chunk c = [wn][ks][j(3)][lane] double2; the 6 doubles of a lane are
//   (Br, Bi-Br, Br+Bi) of nt = 0..1 with l = c*KC + ks*4 + lane%4, n = ct*NC + wn*16 + nt*8 + lane/4.
__device__ __forceinline__ void pack_gauge_row(int nb, int nw, int WN, int KC, int nch, int nct, long long kg, int ct,
                                               const cplx* __restrict__ U, double* __restrict__ Up) {
    const int KS = KC / 4;
    const int items = WN * KS * 32;              // per chunk
    const cplx* src = U + kg * (long long)nb * nw;
    double* dst0 = Up + ((kg * nct + ct) * nch) * (long long)(items * 6);
#pragma unroll 2
    for (int oo = threadIdx.x; oo < items * nch; oo += blockDim.x) {
        const int c = oo / items, o = oo - c * items;
        const int lane = o & 31, ks = (o >> 5) % KS, wn = (o >> 5) / KS;
        const int l = c * KC + ks * 4 + (lane & 3);
        double q[6];
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
            const int n = ct * (16 * WN) + wn * 16 + nt * 8 + (lane >> 2);
            const cplx v = (l < nb && n < nw) ? src[l + (long long)n * nb] : make_double2(0.0, 0.0);
            q[3 * nt] = v.x;
            q[3 * nt + 1] = v.y - v.x;
            q[3 * nt + 2] = v.x + v.y;
        }
        double2* d2 = reinterpret_cast<double2*>(dst0 + (long long)c * (items * 6)) +
                      ((long long)(wn * KS + ks) * 3) * 32 + lane;
#pragma unroll
        for (int j = 0; j < 3; j++) d2[j * 32] = make_double2(q[2 * j], q[2 * j + 1]);
    }
}
__global__ void __launch_bounds__(256) pack_gauge_kernel(int nb, int nw, int WN, int KC, int nch, int nct,
                                                         const int32_t* __restrict__ need,
                                                         const int32_t* __restrict__ owner,
                                                         const cplx* const* __restrict__ peerU,
                                                         const cplx* __restrict__ U,
                                                         double* __restrict__ Up) {
    const long long kg = need ? need[blockIdx.x] : blockIdx.x;   // only the k-points this slab reads
    // k-sharded runs: a row another rank owns is read straight from that rank's memory over NVLink
    // (peer-mapped pointer) — the halo exchange is fused into the packing
    if (peerU) U = peerU[owner[blockIdx.x]];
    pack_gauge_row(nb, nw, WN, KC, nch, nct, kg, blockIdx.y, U, Up);
}
// ---- the rotation kernel ---------------------------------------------------------------------
// NORM = false: ||MU_kb||_F^2 is not accumulated (square gauges: the norm of N_kb is gauge invariant for unitary U and
// was computed from M once; a third of the epilogue's CUDA-core FP64 work, which competes with the DMMAs for the pipe)
template <int WM, int WN, int KC, int STAGES, bool NORM>
__global__ void __launch_bounds__(384, 1)
rotate_dmma_kernel(int nb, int nw, int nn, int nch, int nct, int k_begin, int k_off,
                   const int32_t* __restrict__ kpb, const double* __restrict__ Mt,
                   const double* __restrict__ Up, const cplx* __restrict__ U,
                   cplx* __restrict__ MU, cplx* __restrict__ dNp, double* __restrict__ frop,
                   const uint8_t* __restrict__ border, const uint8_t* __restrict__ nloc, int phase) {
    constexpr int KS = KC / 4;
    constexpr int NC = 16 * WN;
    constexpr int A_D2 = WM * KS * 6 * 32;   // double2 per A stage
    constexpr int B_D2 = WN * KS * 3 * 32;   // double2 per B stage
    constexpr uint32_t A_BYTES = A_D2 * 16, B_BYTES = B_D2 * 16;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* sUk = reinterpret_cast<cplx*>(smem_raw);                     // [8 warps][16][32]
    cplx* sA = sUk + 8 * 16 * 32;                                      // [STAGES][A_D2]
    cplx* sB = sA + STAGES * A_D2;                                     // [STAGES][B_D2]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sB + STAGES * B_D2);
    // bars[0..STAGES) full, bars[STAGES..2 STAGES) empty
    int* sPair = reinterpret_cast<int*>(bars + 2 * STAGES);   // [nn <= 255] pair index of the b-th neighbour this CTA rotates

#ifdef WB_TIMING
    const long long t_cta0 = clock64();
#endif
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_base = smem_u32(bars), sA_base = smem_u32(sA), sB_base = smem_u32(sB);
    const int kl = blockIdx.x / nct + k_off, ct = blockIdx.x % nct;
    const int n0 = ct * NC;
    // k-sharded runs split the neighbours of a k-point into those whose gauge this rank owns and those read
    // from a peer: `border` lists the local ones first, phase 1 rotates the first nloc[kl] pairs (while the
    // halo rows are still being fetched and packed on another stream), phase 2 the rest.  phase 0: all, in order.
    int b_lo = 0, b_cnt = nn;
    if (phase) {
        const int nl = nloc[kl];
        if (phase == 1) b_cnt = nl; else { b_lo = nl; b_cnt = nn - nl; }
        if (b_cnt == 0) return;   // uniform over the CTA, before any barrier
        border += (long long)kl * nn + b_lo;
    }
    const int total = b_cnt * nch;  // chunks streamed by this CTA

    // producer state (only meaningful in warp 8 lane 0)
    auto produce = [&](int i) {
        const int s = i % STAGES;
        const uint32_t round = (uint32_t)(i / STAGES);
        const int bi = i / nch, c = i - bi * nch;
        const long long pair = (long long)kl * nn + (phase ? (int)border[bi] : bi);
        const long long kp = kpb[pair];
        mbar_wait(bar_base + 8u * (STAGES + s), (round & 1u) ^ 1u);
        const uint32_t full = bar_base + 8u * s;
        mbar_arrive_expect_tx(full, A_BYTES + B_BYTES);
        // (L2 eviction policies on these copies were measured and make no difference here: evict-first
        // on the overlaps costs DRAM reads, 28.5 -> 33.0 GB — the four column-tile CTAs of a k-point
        // re-fetch them — and evict-last on the gauges changes nothing, 28.3 GB; profiles/r1f_summary.md)
        bulk_g2s(sA_base + s * A_BYTES, Mt + (pair * nch + c) * (long long)(A_D2 * 2), A_BYTES, full);
        bulk_g2s(sB_base + s * B_BYTES, Up + ((kp * nct + ct) * nch + c) * (long long)(B_D2 * 2),
                 B_BYTES, full);
    };
    const int npre = total < STAGES ? total : STAGES;

    // consumer warp coordinates (unused by the producer warpgroup)
    const int wm = (warp & 7) / WN, wn = (warp & 7) % WN;
    const int g = lane >> 2, t = lane & 3;
    // element e = (mt*2 + nt)*2 + j of this thread's accumulator fragment:
    //   m = wm*32 + mt*8 + g,  n = n0 + wn*16 + nt*8 + 2t + j
    cplx v[16];  // U_k values of this thread's 16 elements (prefetched, parked in its own smem slots)

    if (warp >= 8) {
        if (tid == 256) {
            for (int s = 0; s < STAGES; s++) {
                mbar_init(bar_base + 8u * s, 1);
                mbar_init(bar_base + 8u * (STAGES + s), 8);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            for (int i = 0; i < npre; i++) produce(i);   // prime the ring while the CTA sets up
        }
    } else {
        if (tid < b_cnt) sPair[tid] = kl * nn + (phase ? (int)border[tid] : tid);
        const cplx* Uk = U + (long long)(k_begin + kl) * nb * nw;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const int m = wm * 32 + (e >> 2) * 8 + g, n = n0 + wn * 16 + ((e >> 1) & 1) * 8 + 2 * t + (e & 1);
            v[e] = (m < nb && n < nw) ? Uk[m + (long long)n * nb] : make_double2(0.0, 0.0);
        }
    }
    __syncthreads();   // barrier inits visible to the consumers

    if (warp >= 8) {
        // ===== producer warpgroup: hand its registers to the consumers, one lane drives the ring =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (tid == 256)
            for (int i = npre; i < total; i++) produce(i);
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
#ifdef WB_TIMING
    const long long t_setup = clock64() - t_cta0;
#endif

    // ===== consumers =====
    // each thread parks its U_k values in smem slots only it ever touches (no barrier needed)
    cplx* uk = sUk + (warp * 16) * 32 + lane;
#pragma unroll
    for (int e = 0; e < 16; e++) uk[e * 32] = v[e];

    double p1[4][2][2], p2[4][2][2], p3[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) p1[i][j][e] = p2[i][j][e] = p3[i][j][e] = 0.0;

    // Deferred epilogue.  At the end of pair b the accumulators are combined into x[16] (the MU
    // values of this thread) and cleared; the expensive part of the epilogue — 16 STG.128 of MU,
    // the diagonal and norm accumulations — is then issued in 8 slices of two elements BETWEEN the
    // DMMA groups of the first 8 k-steps of pair b+1, so the store burst (64 KB per CTA) and the
    // FP64 CUDA-core work hide under tensor-pipe time instead of stalling all eight warps at once.
    cplx x[16];
    cplx d[2][2];
    double f = 0.0;
    long long epi_pair = 0;
    constexpr int EPI_CHUNKS = 8 / KS;   // chunks over which the 8 slices are spread

    auto epi_begin = [&](long long pair) {
        epi_pair = pair;
        f = 0.0;
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
#pragma unroll
            for (int j = 0; j < 2; j++) d[nt][j] = make_double2(0.0, 0.0);
#pragma unroll
        for (int mt = 0; mt < 4; mt++)
#pragma unroll
            for (int nt = 0; nt < 2; nt++)
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    x[(mt * 2 + nt) * 2 + j] = make_double2(p1[mt][nt][j] - p3[mt][nt][j],
                                                            p1[mt][nt][j] + p2[mt][nt][j]);
                    p1[mt][nt][j] = 0.0;
                    p2[mt][nt][j] = 0.0;
                    p3[mt][nt][j] = 0.0;
                }
    };
    auto epi_slice = [&](int sl) {   // elements 2*sl, 2*sl+1 (sl static after unrolling)
        cplx* mu = MU + epi_pair * (long long)nb * nw;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int e = 2 * sl + q;
            const int mt = e >> 2, nt = (e >> 1) & 1, j = e & 1;
            const int m = wm * 32 + mt * 8 + g, n = n0 + wn * 16 + nt * 8 + 2 * t + j;
            const cplx xv = x[e];
#ifndef WB_EPI_NOSTG
            if (m < nb && n < nw) mu[m + (long long)n * nb] = xv;
#else
            if (xv.x == 1.2345e300) mu[m + (long long)n * nb] = xv;   // ablation build only
#endif
#ifndef WB_EPI_NOMATH
            cfma_conj(d[nt][j], uk[e * 32], xv);
            if (NORM) {
                f = fma(xv.x, xv.x, f);
                f = fma(xv.y, xv.y, f);
            }
#endif
        }
    };
    auto epi_end = [&]() {
        // reduce over the 8 row groups g (lanes with equal t)
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    d[nt][j].x += __shfl_xor_sync(0xffffffffu, d[nt][j].x, o);
                    d[nt][j].y += __shfl_xor_sync(0xffffffffu, d[nt][j].y, o);
                }
            }
        if (NORM) f = warp_sum(f);
        if (g == 0) {
#pragma unroll
            for (int nt = 0; nt < 2; nt++)
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int n = n0 + wn * 16 + nt * 8 + 2 * t + j;
                    if (n < nw) dNp[(epi_pair * WM + wm) * nw + n] = d[nt][j];
                }
        }
        if (NORM && lane == 0) frop[(epi_pair * nct + ct) * 8 + warp] = f;
    };

#ifdef WB_TIMING
    long long t_wait = 0, t_epi = 0, t_first = 0;
    const long long t_start = clock64();
    unsigned long long g_start;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g_start));
#define WB_T0 const long long tt0_ = clock64();
#define WB_T1(acc) acc += clock64() - tt0_;
#else
#define WB_T0
#define WB_T1(acc)
#endif

    int s = 0;
    uint32_t ph = 0;
    // one chunk: wait for the stage, KS k-steps of 24 DMMA, release the stage.
    // SLICES: issue deferred-epilogue slice (c*KS + ks) after the DMMA group of each k-step.
    auto chunk = [&](int c, bool slices) {
        {
            WB_T0
            mbar_wait(bar_base + 8u * s, ph);
            WB_T1(t_wait)
        }
        const cplx* Ap = sA + s * A_D2 + (wm * KS * 6) * 32 + lane;
        const cplx* Bp = sB + s * B_D2 + (wn * KS * 3) * 32 + lane;
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
            double qa[12], qb[6];
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const cplx y = Ap[(ks * 6 + j) * 32];
                qa[2 * j] = y.x;
                qa[2 * j + 1] = y.y;
            }
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const cplx y = Bp[(ks * 3 + j) * 32];
                qb[2 * j] = y.x;
                qb[2 * j + 1] = y.y;
            }
            // qa[3 mt + {0,1,2}] = Ar, Ai, Ar+Ai ; qb[3 nt + {0,1,2}] = Br, Bi-Br, Br+Bi
#pragma unroll
            for (int mt = 0; mt < 4; mt++)
#pragma unroll
                for (int nt = 0; nt < 2; nt++) {
                    dmma884(p1[mt][nt][0], p1[mt][nt][1], qa[3 * mt + 2], qb[3 * nt]);
                    dmma884(p2[mt][nt][0], p2[mt][nt][1], qa[3 * mt], qb[3 * nt + 1]);
                    dmma884(p3[mt][nt][0], p3[mt][nt][1], qa[3 * mt + 1], qb[3 * nt + 2]);
                }
            if (slices) epi_slice(c * KS + ks);
        }
        // WAR across proxies: our generic-proxy LDS reads of this stage must be ordered before
        // the producer's next async-proxy bulk write into it (without this fence a stage was
        // observed being overwritten under a slow warp; see tests: multi-wave determinism)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_base + 8u * (STAGES + s));
        if (++s == STAGES) { s = 0; ph ^= 1; }
    };

    // (Measured, tools/dmma_mix.cu: next to a saturating DMMA stream a CUDA-core FP64 instruction gets
    // a pipe slot only every ~10 cycles.  Staggering the slices of the two warps of an SMSP, or
    // skewing the warps by whole chunks, was tried and is slower: the warps drift apart and stall on
    // the 5-stage ring.)
    for (int b = 0; b < b_cnt; b++) {
        // first EPI_CHUNKS chunks of the pair carry the deferred epilogue of the previous pair
        if (b > 0) {
#pragma unroll
            for (int c = 0; c < EPI_CHUNKS; c++) chunk(c, true);
            WB_T0
            epi_end();
            WB_T1(t_epi)
        } else {
#pragma unroll 1
            for (int c = 0; c < EPI_CHUNKS; c++) chunk(c, false);
        }
#pragma unroll 1
        for (int c = EPI_CHUNKS; c < nch; c++) chunk(c, false);
        WB_T0
        epi_begin(sPair[b]);
        WB_T1(t_epi)
    }
    {   // the last pair has no successor to hide under
        WB_T0
#pragma unroll
        for (int sl = 0; sl < 8; sl++) epi_slice(sl);
        epi_end();
        WB_T1(t_epi)
    }
#ifdef WB_TIMING
    if (lane == 0 && wb_dbg) {
        unsigned long long g_end;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g_end));
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        long long* dd = wb_dbg + ((long long)blockIdx.x * 8 + warp) * 8;
        dd[0] = clock64() - t_start; dd[1] = t_wait; dd[2] = t_epi; dd[3] = t_first;
        dd[4] = (long long)g_start; dd[5] = (long long)g_end; dd[6] = smid; dd[7] = t_setup;
    }
#endif
}

// ---- small shapes (nb <= 32): one pair per WARP PAIR, operands straight from global ------------
// A 32 x 32 (zero padded) pair needs no shared memory at all: the fragment-ordered images of M and
// of the packed gauge are read with coalesced LDG.128 directly into DMMA fragments (one k-step
// ahead), warp 2i computes columns 0-15 and warp 2i+1 columns 16-31 of pair i (the second read of
// the A image hits L1), and each warp covers all 32 rows, so diag(N) is complete inside the warp.
// 8 warps (4 pairs) per CTA, no barriers, no producer.
__global__ void __launch_bounds__(256, 1)
rotate_small_kernel(int nb, int nw, int nn, long long pair0, long long pair_end, int k_begin,
                    const int32_t* __restrict__ kpb, const double* __restrict__ Mt,
                    const double* __restrict__ Up, const cplx* __restrict__ U, cplx* __restrict__ MU,
                    cplx* __restrict__ dN, double* __restrict__ frop) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long p = pair0 + (long long)blockIdx.x * 4 + (warp >> 1);
    const int half = warp & 1;
    if (p >= pair_end) return;
    const int g = lane >> 2, t = lane & 3;
    const long long kl = p / nn, kp = kpb[p];
    const cplx* A = reinterpret_cast<const cplx*>(Mt) + p * 1536 + lane;               // [(ks*6 + j)*32]
    const cplx* B = reinterpret_cast<const cplx*>(Up) + kp * 1536 + (half * 24) * 32 + lane;  // [(ks*3 + j)*32]

    double p1[4][2][2], p2[4][2][2], p3[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) p1[i][j][e] = p2[i][j][e] = p3[i][j][e] = 0.0;
    double qa[2][12], qb[2][6];
    auto load = [&](int ks, double (&a)[12], double (&b)[6]) {
#pragma unroll
        for (int j = 0; j < 6; j++) {
            const cplx y = __ldg(A + (ks * 6 + j) * 32);
            a[2 * j] = y.x;
            a[2 * j + 1] = y.y;
        }
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const cplx y = __ldg(B + (ks * 3 + j) * 32);
            b[2 * j] = y.x;
            b[2 * j + 1] = y.y;
        }
    };
    load(0, qa[0], qb[0]);
#pragma unroll
    for (int ks = 0; ks < 8; ks++) {
        const int cur = ks & 1;
        if (ks + 1 < 8) load(ks + 1, qa[cur ^ 1], qb[cur ^ 1]);
#pragma unroll
        for (int mt = 0; mt < 4; mt++)
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
                dmma884(p1[mt][nt][0], p1[mt][nt][1], qa[cur][3 * mt + 2], qb[cur][3 * nt]);
                dmma884(p2[mt][nt][0], p2[mt][nt][1], qa[cur][3 * mt], qb[cur][3 * nt + 1]);
                dmma884(p3[mt][nt][0], p3[mt][nt][1], qa[cur][3 * mt + 1], qb[cur][3 * nt + 2]);
            }
    }
    // epilogue: MU, diag(U_k^H MU) (complete: the warp holds all rows), ||MU||_F^2 of this half
    cplx* mu = MU + p * (long long)nb * nw;
    const cplx* Uk = U + (long long)(k_begin + kl) * nb * nw;
    cplx d[2][2];
    double f = 0.0;
#pragma unroll
    for (int nt = 0; nt < 2; nt++)
#pragma unroll
        for (int j = 0; j < 2; j++) d[nt][j] = make_double2(0.0, 0.0);
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const cplx x = make_double2(p1[mt][nt][j] - p3[mt][nt][j], p1[mt][nt][j] + p2[mt][nt][j]);
                const int m = mt * 8 + g, n = half * 16 + nt * 8 + 2 * t + j;
                if (m < nb && n < nw) {
                    mu[m + (long long)n * nb] = x;
                    cfma_conj(d[nt][j], Uk[m + (long long)n * nb], x);
                }
                f = fma(x.x, x.x, f);
                f = fma(x.y, x.y, f);
            }
#pragma unroll
    for (int nt = 0; nt < 2; nt++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                d[nt][j].x += __shfl_xor_sync(0xffffffffu, d[nt][j].x, o);
                d[nt][j].y += __shfl_xor_sync(0xffffffffu, d[nt][j].y, o);
            }
        }
    f = warp_sum(f);
    if (g == 0) {
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int n = half * 16 + nt * 8 + 2 * t + j;
                if (n < nw) dN[p * nw + n] = d[nt][j];
            }
    }
    if (lane == 0) frop[p * 2 + half] = f;
}

// ---- small shapes, ring version ---------------------------------------------------------------
// rotate_small_kernel above keeps only one k-step of loads in flight per warp (2 warps per SMSP, one
// CTA per SM because of the 96 accumulator registers), so it runs at DRAM latency, not bandwidth.
// Here a persistent CTA owns a contiguous range of pairs; a producer warp streams the operand images
// with 1-D bulk async copies into per-warp-pair rings of 4 slots x 2 k-steps (12 KB per slot, 192 KB
// per CTA), so every warp pair always has three chunks (~37 KB per warp pair, ~150 KB per SM) in
// flight while it computes on the fourth.  Warp 2w / 2w+1 compute columns 0-15 / 16-31 of the pairs
// i = w (mod 4) of the CTA's range; the U_k values of the epilogue are loaded at the start of the
// pair so their latency hides under the DMMAs.
constexpr int SR_KS = 2;                                  // k-steps per chunk
constexpr int SR_NCH = 8 / SR_KS;                         // chunks per pair
constexpr int SR_SLOTS = 4;                               // slots per warp pair
constexpr uint32_t SR_A_BYTES = SR_KS * 6 * 32 * 16;      // 6144
constexpr uint32_t SR_B_BYTES = SR_KS * 3 * 32 * 16;      // 3072 per column half
constexpr uint32_t SR_SLOT_BYTES = SR_A_BYTES + 2 * SR_B_BYTES;   // 12288
constexpr size_t SR_RING_BYTES = (size_t)4 * SR_SLOTS * SR_SLOT_BYTES;
constexpr size_t SR_SMEM = SR_RING_BYTES + 8 * 2 * 4 * SR_SLOTS;

// FULL: nb > 28 and nw > 24, nothing to skip: no predicates in the main loop.  NORM as in rotate_dmma_kernel.
template <bool FULL, bool NORM>
__global__ void __launch_bounds__(384, 1)
rotate_small_ring_kernel(int nb, int nw, int nn, long long pair0, long long pair_end, int k_begin,
                         const int32_t* __restrict__ kpb, const double* __restrict__ Mt,
                         const double* __restrict__ Up, const cplx* __restrict__ U, cplx* __restrict__ MU,
                         cplx* __restrict__ dN, double* __restrict__ frop) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + SR_RING_BYTES);
    // bars[(wp*SR_SLOTS + s)*2 + {0 full, 1 empty}]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t ring_base = smem_u32(smem_raw), bar_base = smem_u32(bars);
    const long long total = pair_end - pair0;
    const long long per = (total + gridDim.x - 1) / gridDim.x;
    const long long my0 = pair0 + (long long)blockIdx.x * per;
    const long long my1 = my0 + per < pair_end ? my0 + per : pair_end;
    const int cnt = my1 > my0 ? (int)(my1 - my0) : 0;
    if (cnt == 0) return;

    if (warp == 8 && lane == 0) {
        for (int i = 0; i < 4 * SR_SLOTS; i++) {
            mbar_init(bar_base + 16u * i, 1);
            mbar_init(bar_base + 16u * i + 8u, 2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // zero padding is never multiplied: k-steps, row tiles and column tiles beyond nb / nw are skipped
    // (at nb = 18, nw = 9 that leaves 5 of 8 k-steps, 3 of 4 row tiles and one of the two column
    // halves), and chunks that hold only padding are neither copied nor waited for
    const int ksteps = (nb + 3) / 4, mtiles = (nb + 7) / 8;
    const int nch_act = FULL ? SR_NCH : (ksteps + SR_KS - 1) / SR_KS;

    // warp pair w owns the contiguous quarter [q0(w), q0(w + 1)) of the CTA's pairs: consecutive pairs
    // share k (nn neighbours each), so the U_k values of the epilogue are reloaded once per k-point
    auto q0 = [&](int w) { return (int)((long long)cnt * w / 4); };

    if (warp >= 8) {
        // producer warpgroup (warps 8-11): hands its registers to the consumers; warp 8 drives the rings
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp > 9) return;
        // warp 8 feeds the rings of warp pairs 0-1, warp 9 those of warp pairs 2-3.  The u-th chunk
        // of a warp pair goes to its slot u % SR_SLOTS.  Lane l keeps the neighbour index of pair
        // (l % 16) of the current group of 16 of stream l / 16 (one coalesced load per group), so
        // the issuing lane never waits on a dependent global load between copies.
        const uint64_t pol_a = l2_policy_evict_first(), pol_b = l2_policy_evict_last();
        const int wbase = 2 * (warp - 8);
        const int w_l = wbase + (lane >> 4), j_l = lane & 15;
        const int len_l = q0(w_l + 1) - q0(w_l);
        const int maxlen = max(q0(wbase + 1) - q0(wbase), q0(wbase + 2) - q0(wbase + 1));
        for (int qb = 0; qb < maxlen; qb += 16) {
            const int mine = qb + j_l < len_l ? kpb[my0 + q0(w_l) + qb + j_l] : 0;
            for (int j = 0; j < 16; j++)
                for (int ws = 0; ws < 2; ws++) {
                    const long long kp = __shfl_sync(0xffffffffu, mine, ws * 16 + j);
                    const int wp = wbase + ws, q = qb + j;
                    if (lane != 0 || q >= q0(wp + 1) - q0(wp)) continue;
                    const long long p = my0 + q0(wp) + q;
                    const char* a = reinterpret_cast<const char*>(Mt) + p * 24576;
                    const char* b = reinterpret_cast<const char*>(Up) + kp * 24576;
#pragma unroll
                    for (int c = 0; c < SR_NCH; c++) {
                        if (!FULL && c >= nch_act) break;
                        const int u = q * nch_act + c;
                        const int s = u % SR_SLOTS;
                        const uint32_t full = bar_base + 16u * (wp * SR_SLOTS + s), empty = full + 8u;
                        mbar_wait(empty, ((uint32_t)(u / SR_SLOTS) & 1u) ^ 1u);
                        mbar_arrive_expect_tx(full, SR_SLOT_BYTES);
                        const uint32_t dst = ring_base + (uint32_t)(wp * SR_SLOTS + s) * SR_SLOT_BYTES;
                        bulk_g2s_hint(dst, a + c * SR_A_BYTES, SR_A_BYTES, full, pol_a);
                        bulk_g2s_hint(dst + SR_A_BYTES, b + c * SR_B_BYTES, SR_B_BYTES, full, pol_b);
                        bulk_g2s_hint(dst + SR_A_BYTES + SR_B_BYTES, b + 12288 + c * SR_B_BYTES, SR_B_BYTES, full, pol_b);
                    }
                }
        }
        return;
    }

    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // column half of this warp; swapped in warp pairs 2-3 so that, when only half 0 has columns
    // (nw <= 16), the four working warps sit on four different SMSPs
    const int wp = warp >> 1, half = (warp & 1) ^ ((warp >> 2) & 1);
    const int g = lane >> 2, t = lane & 3;
    const int ntiles = nw - half * 16 >= 16 ? 2 : (nw - half * 16 + 7) / 8;   // <= 0: no columns at all
    int u = 0;   // chunk counter of this warp pair
    if (!FULL && ntiles <= 0) {   // nothing to compute: keep the ring protocol going, report a zero norm
        for (int i = q0(wp); i < q0(wp + 1); i++) {
            for (int c = 0; c < nch_act; c++, u++) {
                const uint32_t full = bar_base + 16u * (wp * SR_SLOTS + u % SR_SLOTS);
                mbar_wait(full, (uint32_t)(u / SR_SLOTS) & 1u);
                if (lane == 0) mbar_arrive(full + 8u);
            }
            if (NORM && lane == 0) frop[(my0 + i) * 2 + half] = 0.0;
        }
        return;
    }
    cplx uk[16];   // U_k values of this thread's 16 outputs
#ifdef WB_TIMING
    long long r_wait = 0, r_mma = 0, r_epi = 0;
    const long long r_start = clock64();
#endif
    long long p = my0 + q0(wp);
    long long kl = p / nn;
    int bi = (int)(p - kl * nn);
    bool fresh = true;
    for (int i = q0(wp); i < q0(wp + 1); i++, p++) {
        if (fresh) {   // new k-point (once per nn pairs): its gauge values are in flight while the DMMAs run
            const cplx* Uk = U + (long long)(k_begin + kl) * nb * nw;
#pragma unroll
            for (int e = 0; e < 16; e++) {
                const int m = (e >> 2) * 8 + g, n = half * 16 + ((e >> 1) & 1) * 8 + 2 * t + (e & 1);
                uk[e] = (m < nb && n < nw) ? __ldg(Uk + m + (long long)n * nb) : make_double2(0.0, 0.0);
            }
        }
        double p1[4][2][2], p2[4][2][2], p3[4][2][2];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 2; b++)
#pragma unroll
                for (int e = 0; e < 2; e++) p1[a][b][e] = p2[a][b][e] = p3[a][b][e] = 0.0;
#pragma unroll
        for (int c = 0; c < SR_NCH; c++) {
            if (!FULL && c >= nch_act) break;
            const int s = u % SR_SLOTS;
            const uint32_t full = bar_base + 16u * (wp * SR_SLOTS + s);
#ifdef WB_TIMING
            const long long tw0 = clock64();
#endif
            mbar_wait(full, (uint32_t)(u / SR_SLOTS) & 1u);
#ifdef WB_TIMING
            const long long tw1 = clock64();
            r_wait += tw1 - tw0;
#endif
            const cplx* Ap = reinterpret_cast<const cplx*>(smem_raw + (size_t)(wp * SR_SLOTS + s) * SR_SLOT_BYTES) + lane;
            const cplx* Bp = Ap + (SR_A_BYTES + half * SR_B_BYTES) / 16;
#pragma unroll
            for (int ks = 0; ks < SR_KS; ks++) {
                if (!FULL && c * SR_KS + ks >= ksteps) break;
                double qa[12], qb[6];
#pragma unroll
                for (int j = 0; j < 6; j++) {
                    const cplx y = Ap[(ks * 6 + j) * 32];
                    qa[2 * j] = y.x;
                    qa[2 * j + 1] = y.y;
                }
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const cplx y = Bp[(ks * 3 + j) * 32];
                    qb[2 * j] = y.x;
                    qb[2 * j + 1] = y.y;
                }
#pragma unroll
                for (int mt = 0; mt < 4; mt++)
#pragma unroll
                    for (int nt = 0; nt < 2; nt++) {
                        if (FULL || (mt < mtiles && nt < ntiles)) {
                            dmma884(p1[mt][nt][0], p1[mt][nt][1], qa[3 * mt + 2], qb[3 * nt]);
                            dmma884(p2[mt][nt][0], p2[mt][nt][1], qa[3 * mt], qb[3 * nt + 1]);
                            dmma884(p3[mt][nt][0], p3[mt][nt][1], qa[3 * mt + 1], qb[3 * nt + 2]);
                        }
                    }
            }
            // generic-proxy reads of the slot before the producer's next async-proxy write into it
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(full + 8u);
            u++;
#ifdef WB_TIMING
            r_mma += clock64() - tw1;
#endif
        }
#ifdef WB_TIMING
        const long long te0 = clock64();
#endif
        // epilogue: MU, diag(U_k^H MU) (complete: the warp holds all rows), ||MU||_F^2 of this half.
        // (Deferring it into the next pair's k-steps, as rotate_dmma_kernel does, was measured slower
        // here: 0.246 vs 0.196 ms at nb = 32, 12^3 k — eight k-steps cannot hide eight slices.)
        cplx* mu = MU + p * (long long)nb * nw;
        cplx d[2][2];
        double f4[2][2] = {{0.0, 0.0}, {0.0, 0.0}};   // independent chains; summed below
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
#pragma unroll
            for (int j = 0; j < 2; j++) d[nt][j] = make_double2(0.0, 0.0);
#pragma unroll
        for (int mt = 0; mt < 4; mt++)
#pragma unroll
            for (int nt = 0; nt < 2; nt++)
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const cplx x = make_double2(p1[mt][nt][j] - p3[mt][nt][j], p1[mt][nt][j] + p2[mt][nt][j]);
                    const int m = mt * 8 + g, n = half * 16 + nt * 8 + 2 * t + j;
                    if (m < nb && n < nw) {
                        mu[m + (long long)n * nb] = x;
                        cfma_conj(d[nt][j], uk[(mt * 2 + nt) * 2 + j], x);
                    }
                    if (NORM) f4[nt][j] = fma(x.x, x.x, fma(x.y, x.y, f4[nt][j]));
                }
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    d[nt][j].x += __shfl_xor_sync(0xffffffffu, d[nt][j].x, o);
                    d[nt][j].y += __shfl_xor_sync(0xffffffffu, d[nt][j].y, o);
                }
            }
        double f = 0.0;
        if (NORM) f = warp_sum((f4[0][0] + f4[0][1]) + (f4[1][0] + f4[1][1]));
        if (g == 0) {
#pragma unroll
            for (int nt = 0; nt < 2; nt++)
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int n = half * 16 + nt * 8 + 2 * t + j;
                    if (n < nw) dN[p * nw + n] = d[nt][j];
                }
        }
        if (NORM && lane == 0) frop[p * 2 + half] = f;
        fresh = ++bi == nn;
        if (fresh) { bi = 0; kl++; }
#ifdef WB_TIMING
        r_epi += clock64() - te0;
#endif
    }
#ifdef WB_TIMING
    if (lane == 0 && wb_dbg) {
        long long* dd = wb_dbg + ((long long)blockIdx.x * 8 + warp) * 8;
        dd[0] = clock64() - r_start; dd[1] = r_wait; dd[2] = r_mma; dd[3] = r_epi; dd[4] = q0(wp + 1) - q0(wp);
    }
#endif
}

template <int WM, int WN, int KC, int STAGES>
size_t smem_bytes() {
    constexpr int KS = KC / 4;
    return sizeof(cplx) * (8 * 16 * 32 + (size_t)STAGES * (WM * KS * 6 * 32 + WN * KS * 3 * 32)) +
           sizeof(unsigned long long) * 2 * STAGES + sizeof(int) * 256;
}

template <int WM, int WN, int KC, int STAGES>
int launch_rotate(wb200_ctx* ctx, const cplx* dU, int ka, int kb, int phase, bool norm) {
    DmmaPlan* p = ctx->dmma;
    if (!ctx->rot_attr_done) {   // once per ctx (contexts of different devices share the instantiations)
        WB_CUDA(ctx, cudaFuncSetAttribute(rotate_dmma_kernel<WM, WN, KC, STAGES, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
        WB_CUDA(ctx, cudaFuncSetAttribute(rotate_dmma_kernel<WM, WN, KC, STAGES, false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem));
        ctx->rot_attr_done = 1;
    }
    auto kern = norm ? rotate_dmma_kernel<WM, WN, KC, STAGES, true> : rotate_dmma_kernel<WM, WN, KC, STAGES, false>;
    kern<<<(kb - ka) * p->nct, 384, p->smem, ctx->stream>>>(ctx->nb, ctx->nw, ctx->nn, p->nch, p->nct, ctx->k_begin, ka,
                                                          ctx->d_kpb, p->d_Mt, p->d_Up, dU, ctx->d_MU, p->d_dNp, p->d_frop,
                                                          ctx->d_border, ctx->d_nloc, phase);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}

// ||M_kb||_F^2 per pair, one CTA per pair (fixed-order tree)
__global__ void __launch_bounds__(256) overlap_norm_kernel(int nb, const cplx* __restrict__ M, double* __restrict__ out) {
    const cplx* m = M + (long long)blockIdx.x * nb * nb;
    double s = 0.0;
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) s = fma(m[e].x, m[e].x, fma(m[e].y, m[e].y, s));
    __shared__ double sh[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += sh[w];
        out[blockIdx.x] = t;
    }
}

}  // namespace

// WM, WN, KC, STAGES per configuration.  Shared memory: 64 KB U_k tile + STAGES x (A + B) stage,
// A stage = WM*KC*6 KB/16... in bytes: WM*(KC/4)*6*512, B stage = WN*(KC/4)*3*512.
#define WB_CFG_128 4, 2, 16, 2   /* 64 < nb <= 128: CTA 128 x 32, stage 48+12 KB   -> 184 KB (measured: 2 x 16-deep beats 5 x 8-deep) */
#define WB_CFG_256 8, 1, 8, 3    /* 128 < nb <= 256: CTA 256 x 16, stage 48+3 KB   -> 217 KB */
#define WB_CFG_64 2, 4, 16, 3    /* 32 < nb <= 64:  CTA 64 x 64, stage 24+24 KB    -> 208 KB */

int wb_dmma_plan_create(wb200_ctx* ctx, const double* /*hM*/) {
    const int nb = ctx->nb, nw = ctx->nw;
    if (nb > 256 || (nb <= 32 && nw > 32) || ctx->nn > 255) return WB200_OK;  // CUDA-core path
    DmmaPlan* p = new DmmaPlan();
    if (nb <= 32) {
        p->small = true; p->WM = 1; p->WN = 2; p->KC = 32; p->smem = 0;
        const char* e = getenv("WB200_SMALL_RING");
        p->small_ring = !(e && e[0] == '0');
        cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, ctx->device);
        p->small_ring_min = (long long)p->num_sms * 4 * 2;   // two pairs per warp pair: below that a ring never fills
    }
    else if (nb <= 64) { p->WM = 2; p->WN = 4; p->KC = 16; p->smem = smem_bytes<WB_CFG_64>(); }
    else if (nb <= 128) { p->WM = 4; p->WN = 2; p->KC = 16; p->smem = smem_bytes<WB_CFG_128>(); }
    else { p->WM = 8; p->WN = 1; p->KC = 8; p->smem = smem_bytes<WB_CFG_256>(); }
    p->NB = 32 * p->WM;
    p->NC = 16 * p->WN;
    p->nct = (nw + p->NC - 1) / p->NC;
    p->nch = (nb + p->KC - 1) / p->KC;
    p->a_stage = (size_t)p->WM * (p->KC / 4) * 32 * 12;
    p->b_stage = (size_t)p->WN * (p->KC / 4) * 32 * 6;
    ctx->dmma = p;
    const size_t pairs = (size_t)ctx->nk * ctx->nn;
    WB_CUDA(ctx, wb_dev_alloc(&p->d_Mt, sizeof(double) * pairs * p->nch * p->a_stage));
    WB_CUDA(ctx, wb_dev_alloc(&p->d_Up, sizeof(double) * (size_t)ctx->nk_total * p->nct * p->nch * p->b_stage));
    if (!p->small) WB_CUDA(ctx, wb_dev_alloc(&p->d_dNp, sizeof(cplx) * pairs * p->WM * nw));
    WB_CUDA(ctx, wb_dev_alloc(&p->d_frop, sizeof(double) * pairs * (p->small ? 2 : p->nct * 8)));
    // re-lay-out M (already uploaded to d_M) into the A-operand image; d_M is then released:
    // the tensor path never reads the plain layout again.
    dim3 grid((unsigned)pairs, (unsigned)p->nch);
    tile_overlaps_kernel<<<grid, 128, 0, ctx->stream>>>(nb, p->WM, p->KC, p->nch, ctx->d_M, p->d_Mt);
    WB_LAUNCH_CHECK(ctx);
    if (nb == nw) {
        const char* e = getenv("WB200_CONST_NORM");
        if (!(e && e[0] == '0')) {
            WB_CUDA(ctx, wb_dev_alloc(&p->d_froM, sizeof(double) * pairs));
            overlap_norm_kernel<<<(unsigned)pairs, 256, 0, ctx->stream>>>(nb, ctx->d_M, p->d_froM);
            WB_LAUNCH_CHECK(ctx);
        }
    }
    // d_M is released when the tiling kernels have run: enqueue nothing after them that reads it, and let the
    // end-of-create wait cover the release (the block goes back to the pool only then: see wb200_create)
    // (tiny shapes keep it: the fused trial-point kernel of tiny.cu reads the plain layout)
    ctx->release_M_after_create = ctx->tiny_ok ? 0 : 1;
    return WB200_OK;
}

void wb_dmma_plan_destroy(wb200_ctx* ctx) {
    DmmaPlan* p = ctx->dmma;
    if (!p) return;
    // (wb200_destroy has drained the device)
    if (p->d_Mt) wb_dev_free_drained(p->d_Mt);
    if (p->d_Up) wb_dev_free_drained(p->d_Up);
    if (p->d_dNp) wb_dev_free_drained(p->d_dNp);
    if (p->d_frop) wb_dev_free_drained(p->d_frop);
    if (p->d_froM) wb_dev_free_drained(p->d_froM);
    delete p;
    ctx->dmma = nullptr;
}

int wb_pack_dmma(wb200_ctx* ctx, const cplx* dU, int ja, int jb) {
    DmmaPlan* p = ctx->dmma;
    dim3 g((unsigned)(jb - ja), (unsigned)p->nct);
    ProfScope ps(ctx, 0);
    pack_gauge_kernel<<<g, 256, 0, ctx->stream>>>(ctx->nb, ctx->nw, p->WN, p->KC, p->nch, p->nct,
                                                  ctx->d_need + ja,
                                                  ctx->d_need_owner ? ctx->d_need_owner + ja : nullptr,
                                                  ctx->d_peerU, dU, p->d_Up);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}

// the halo split (api.cu) applies to the three CTA-per-(k, column tile) kernels
bool wb_dmma_phases_supported(const wb200_ctx* ctx) { return ctx->dmma && !ctx->dmma->small; }

// the rows this rank owns (the first n_need_local entries of the split need list)
int wb_pack_dmma_local(wb200_ctx* ctx, const cplx* dU) {
    DmmaPlan* p = ctx->dmma;
    if (ctx->n_need_local <= 0) return WB200_OK;
    dim3 g((unsigned)ctx->n_need_local, (unsigned)p->nct);
    ProfScope ps(ctx, 0);
    pack_gauge_kernel<<<g, 256, 0, ctx->stream>>>(ctx->nb, ctx->nw, p->WN, p->KC, p->nch, p->nct, ctx->d_need_split,
                                                  ctx->d_owner_split, ctx->d_peerU, dU, p->d_Up);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}

// the packed-gauge image [nk_total][nct][nch][b_stage]: base pointer, size, bytes per k-point
int wb_dmma_packed_info(wb200_ctx* ctx, void** dptr, size_t* bytes, size_t* row_bytes) {
    DmmaPlan* p = ctx->dmma;
    if (!p || p->small) {
        ctx->err = "no packed gauge image for this shape (tensor path with 32 < nb <= 256 only)";
        return WB200_ERR_ARG;
    }
    *row_bytes = sizeof(double) * (size_t)p->nct * p->nch * p->b_stage;
    *bytes = *row_bytes * (size_t)ctx->nk_total;
    *dptr = p->d_Up;
    return WB200_OK;
}

int wb_rotate_dmma(wb200_ctx* ctx, const cplx* dU, int ka, int kb, int phase) {
    DmmaPlan* p = ctx->dmma;
    ctx->fro_src = nullptr;
    ctx->dn_src = nullptr;
    if (p->small) {
        const long long p0 = (long long)ka * ctx->nn, p1 = (long long)kb * ctx->nn;
        bool ring_used = false;
        {
            ProfScope ps(ctx, 1);
            if (p->small_ring && p1 - p0 >= p->small_ring_min) {
                if (!ctx->rot_attr_done) {
                    WB_CUDA(ctx, cudaFuncSetAttribute(rotate_small_ring_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SR_SMEM));
                    WB_CUDA(ctx, cudaFuncSetAttribute(rotate_small_ring_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SR_SMEM));
                    WB_CUDA(ctx, cudaFuncSetAttribute(rotate_small_ring_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SR_SMEM));
                    WB_CUDA(ctx, cudaFuncSetAttribute(rotate_small_ring_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SR_SMEM));
                    ctx->rot_attr_done = 1;
                }
                // persistent: one CTA per SM, each a contiguous range of pairs (4 in flight per CTA)
                const long long want = (p1 - p0 + 3) / 4;
                const int grid = (int)(want < p->num_sms ? want : p->num_sms);
                const bool full = ctx->nb > 28 && ctx->nw > 24, norm = p->d_froM == nullptr;
                auto kern = full ? (norm ? rotate_small_ring_kernel<true, true> : rotate_small_ring_kernel<true, false>)
                                 : (norm ? rotate_small_ring_kernel<false, true> : rotate_small_ring_kernel<false, false>);
                kern<<<grid, 384, SR_SMEM, ctx->stream>>>(ctx->nb, ctx->nw, ctx->nn, p0, p1, ctx->k_begin, ctx->d_kpb, p->d_Mt,
                                                          p->d_Up, dU, ctx->d_MU, ctx->d_dN, p->d_frop);
                ring_used = true;
            } else {
                rotate_small_kernel<<<(unsigned)((p1 - p0 + 3) / 4), 256, 0, ctx->stream>>>(
                    ctx->nb, ctx->nw, ctx->nn, p0, p1, ctx->k_begin, ctx->d_kpb, p->d_Mt, p->d_Up, dU, ctx->d_MU,
                    ctx->d_dN, p->d_frop);
            }
        }
        WB_LAUNCH_CHECK(ctx);
        if (ring_used && p->d_froM) { ctx->fro_src = p->d_froM; ctx->fro_parts = 1; }
        else { ctx->fro_src = p->d_frop; ctx->fro_parts = 2; }   // pair_reduce_kernel adds the two half-tile norms itself
        return WB200_OK;
    }
    {
        ProfScope ps(ctx, 1);
        const bool norm = p->d_froM == nullptr;
        if (p->WM == 2) WB_TRY((launch_rotate<WB_CFG_64>(ctx, dU, ka, kb, phase, norm)));
        else if (p->WM == 4) WB_TRY((launch_rotate<WB_CFG_128>(ctx, dU, ka, kb, phase, norm)));
        else WB_TRY((launch_rotate<WB_CFG_256>(ctx, dU, ka, kb, phase, norm)));
    }
    // per-warp partial diagonals and norms are summed (fixed order) by pair_reduce_kernel itself
    ctx->dn_src = p->d_dNp; ctx->dn_parts = p->WM;
    if (p->d_froM) { ctx->fro_src = p->d_froM; ctx->fro_parts = 1; }
    else { ctx->fro_src = p->d_frop; ctx->fro_parts = p->nct * 8; }
    return WB200_OK;
}

#ifdef WB_TIMING
extern "C" int wb200_debug_set_timing_buffer(long long* dptr) {
    return cudaMemcpyToSymbol(wb_dbg, &dptr, sizeof(dptr)) == cudaSuccess ? 0 : -2;
}
#endif
