// zgemm_dmma.cu — batched complex-FP64 GEMM on the FP64 tensor pipe from PLAIN column-major operands:
//     C[b] = alpha * op(A[ia(b)]) * op(B[ib(b)]) + beta * C[b],   op = identity or conjugate transpose
// for every GEMM of the path that is not the rotation of the constant overlaps:
//   * Stiefel retract! / project_tangent! above n = 64 (Optim.Stiefel_SVD; call sites
//     src/localization/max_localize.jl:62-63, src/localization/disentangle.jl:687-690): X^H X, X T,
//     X^H G, G - X S, square and rectangular (the Y factor of the disentanglement);
//   * the exact N_kb = U_k^H (M_kb U_{k+b}) (second mul! of compute_MU_UtMU!, src/spread.jl:174) behind
//     omega(model), transform_gauge (src/localization/gauge.jl:83-98), position_op and the Berry connection;
//   * X_Y_to_U! / GU_to_GX_GY (src/localization/disentangle.jl:351-356, 481-501) above nb = 64;
//   * the rotation M_kb U_{k+b} itself for nb > 256 (src/spread.jl:173).
//
// The rotation kernel (rotate_dmma.cu) owes its speed to operand images pre-tiled in fragment order;
// that pays for the constant overlaps but costs a re-layout pass in front of every GEMM whose
// operands change (10.7 of the 71 ms of a max_localize evaluation at n_wann = 128 in round 1).  Here the
// producer warp copies plain column segments with 1-D bulk async copies (cp.async.bulk -> UBLKCP,
// completion on an mbarrier) into shared-memory tiles whose leading dimensions are padded so that
// the DMMA fragment loads (LDS.128) are bank-conflict free:
//     A  (m x k, not transposed)  tile [KC cols][BM + 2]      one copy per column   (lane: col t, row g)
//     A^H (A is k x m)            tile [BM rows][KC + 4]      one copy per row of op(A)
//     B  (k x n, not transposed)  tile [BN cols][KC + 4]      one copy per column
//     B^H (B is n x k)            tile [KC rows][BN + 2]      one copy per k
// (a quarter warp reads 8 different 16-byte bank groups: (t LD + g) mod 8 with LD = 2 mod 8, or
// (g LD + t) mod 8 with LD = 4 mod 8).  The consumers form Gauss' three planes themselves
// ((Ar + Ai), Ar, Ai  x  Br, (Bi - Br), (Br + Bi): 8 DADD next to 24 DMMA per k-step), conjugation is
// a sign in those additions.  CTAs are persistent over a contiguous range of output tiles and the
// ring of full/empty mbarriers never drains between tiles.  Nothing is zero padded: k-steps beyond k
// are skipped, a partial last k-step is masked in registers, rows / columns beyond m / n are computed
// on whatever the tile holds and never stored (a DMMA output row depends on its own A row only).
#include "common.cuh"

#include <type_traits>

namespace {

__device__ __forceinline__ uint32_t zd_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void zd_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void zd_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void zd_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void zd_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ZD_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra ZD_DONE;\n"
        "bra ZD_WAIT;\n"
        "ZD_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void zd_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
// 16-byte LDGSTS with zero fill (src_bytes = 0 writes zeros) and its completion on an mbarrier
__device__ __forceinline__ void zd_cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void zd_cp_async_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void zd_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct ZdArgs {
    int m, n, k;
    const cplx* A; long long strideA; int lda; const int32_t* idxA;
    const cplx* B; long long strideB; int ldb; const int32_t* idxB;
    cplx* C; long long strideC; int ldc;
    double alpha, beta;
    int tiles_m, tiles_n;
    long long items;      // batch * tiles_m * tiles_n
    int epi;              // 0: C = alpha AB + beta C;  1: + per-warp ||AB||_F^2 partials;
                          // 2: C = alpha I - beta AB and per-warp ||AB - I||_F^2 partials (Newton-Schulz step
                          //    matrix, plain: alpha = 1.5, beta = 0.5);  3: C = AB and the same partials
    double* part;         // epi 1, 2: [items][8]
};

constexpr int ZD_KC = 16, ZD_KS = 4, ZD_STAGES = 4;

template <int WM, int WN, bool TA, bool TB>
struct ZdCfg {
    static constexpr int BM = 32 * WM, BN = 16 * WN;
    static constexpr int LDA = TA ? ZD_KC + 4 : BM + 2;
    static constexpr int LDB = TB ? BN + 2 : ZD_KC + 4;
    static constexpr int A_ELEMS = TA ? BM * LDA : ZD_KC * LDA;
    static constexpr int B_ELEMS = TB ? ZD_KC * LDB : BN * LDB;
    static constexpr size_t SMEM = sizeof(cplx) * (size_t)ZD_STAGES * (A_ELEMS + B_ELEMS) + 8 * 2 * ZD_STAGES;
};

template <int WM, int WN, bool TA, bool TB>
__global__ void __launch_bounds__(384, 1) zgemm_dmma_kernel(ZdArgs a) {
    using Cfg = ZdCfg<WM, WN, TA, TB>;
    constexpr int BM = Cfg::BM, BN = Cfg::BN, LDA = Cfg::LDA, LDB = Cfg::LDB;
    constexpr int KC = ZD_KC, KS = ZD_KS, STAGES = ZD_STAGES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cplx* sA = reinterpret_cast<cplx*>(smem_raw);                  // [STAGES][A_ELEMS]
    cplx* sB = sA + STAGES * Cfg::A_ELEMS;                         // [STAGES][B_ELEMS]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sB + STAGES * Cfg::B_ELEMS);
    const uint32_t bar_base = zd_smem_u32(bars);                   // [0, STAGES) full, [STAGES, 2 STAGES) empty
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 256) {
        for (int s = 0; s < STAGES; s++) {
            zd_mbar_init(bar_base + 8u * s, TA ? 129 : 1);   // bulk-copy thread (+ 128 LDGSTS threads for A^H)
            zd_mbar_init(bar_base + 8u * (STAGES + s), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long i0 = a.items * blockIdx.x / gridDim.x, i1 = a.items * (blockIdx.x + 1) / gridDim.x;
    const int tiles = a.tiles_m * a.tiles_n;
    const int nch = (a.k + KC - 1) / KC;

    if (warp >= 8) {
        // ===== producer warpgroup: its registers go to the consumers.  Warp 8 issues the bulk copies of the
        // operands whose contiguous runs are long (whole tile columns).  The rows of an A^H tile are only
        // KC elements long (256 B): a bulk copy costs the TMA unit ~50 cycles whatever its size, so 128 of
        // them per stage starve the tensor pipe (measured: 4.1 ms against 1.9 ms for the same flops).  Those
        // tiles are loaded by all four producer warps with 16-byte LDGSTS (zero filled beyond m / k),
        // completion counted on the same mbarrier. =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (!TA && warp != 8) return;
        const int ptid = tid - 256;
        const uint32_t sA_base = zd_smem_u32(sA), sB_base = zd_smem_u32(sB);
        uint32_t u = 0;
        for (long long item = i0; item < i1; item++) {
            const long long b = item / tiles;
            const int tile = (int)(item - b * tiles);
            const int tm = tile / a.tiles_n, tn = tile - tm * a.tiles_n;
            const int m0 = tm * BM, n0 = tn * BN;
            const int rows = min(BM, a.m - m0), cols = min(BN, a.n - n0);
            const cplx* Ab = a.A + (a.idxA ? (long long)a.idxA[b] : b) * a.strideA;
            const cplx* Bb = a.B + (a.idxB ? (long long)a.idxB[b] : b) * a.strideB;
            for (int c = 0; c < nch; c++, u++) {
                const int s = (int)(u % STAGES);
                const uint32_t round = u / STAGES;
                const int k0 = c * KC, kv = min(KC, a.k - k0);
                const uint32_t full = bar_base + 8u * s;
                zd_mbar_wait(bar_base + 8u * (STAGES + s), (round & 1u) ^ 1u);
                const uint32_t dA = sA_base + (uint32_t)s * (Cfg::A_ELEMS * 16), dB = sB_base + (uint32_t)s * (Cfg::B_ELEMS * 16);
                if (warp == 8) {
                    if (lane == 0) zd_mbar_expect_tx(full, (uint32_t)(((TA ? 0 : rows) + cols) * kv * 16));
                    __syncwarp();
                    if (!TA) {
                        for (int l = lane; l < kv; l += 32)
                            zd_bulk_g2s(dA + (uint32_t)l * (LDA * 16), Ab + (long long)(k0 + l) * a.lda + m0, (uint32_t)rows * 16, full);
                    }
                    if (!TB) {
                        for (int j = lane; j < cols; j += 32)
                            zd_bulk_g2s(dB + (uint32_t)j * (LDB * 16), Bb + (long long)(n0 + j) * a.ldb + k0, (uint32_t)kv * 16, full);
                    } else {
                        for (int l = lane; l < kv; l += 32)
                            zd_bulk_g2s(dB + (uint32_t)l * (LDB * 16), Bb + (long long)(k0 + l) * a.ldb + n0, (uint32_t)cols * 16, full);
                    }
                }
                if (TA) {
#pragma unroll 4
                    for (int e = ptid; e < BM * KC; e += 128) {
                        const int i = e / KC, l = e % KC;
                        const bool ok = i < rows && l < kv;
                        zd_cp_async16(dA + (uint32_t)(i * LDA + l) * 16, ok ? Ab + (long long)(m0 + i) * a.lda + k0 + l : Ab,
                                      ok ? 16u : 0u);
                    }
                    zd_cp_async_arrive(full);
                }
            }
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");

    // ===== consumers: warp tile 32 x 16 complex = 4 x 2 DMMA tiles x (P1, P2, P3) =====
    const int wm = warp / WN, wn = warp % WN;
    const int g = lane >> 2, t = lane & 3;
    // fragment base offsets (in complex elements) inside a stage
    const int offA = TA ? (wm * 32 + g) * LDA + t : t * LDA + wm * 32 + g;   // + mt*8 rows, + ks*4 cols
    const int offB = TB ? t * LDB + wn * 16 + g : (wn * 16 + g) * LDB + t;   // + nt*8 cols, + ks*4 rows
    constexpr int A_MT = TA ? 8 * LDA : 8, A_KS = TA ? 4 : 4 * LDA;
    constexpr int B_NT = TB ? 8 : 8 * LDB, B_KS = TB ? 4 * LDB : 4;

    uint32_t u = 0;
    for (long long item = i0; item < i1; item++) {
        const long long b = item / tiles;
        const int tile = (int)(item - b * tiles);
        const int tm = tile / a.tiles_n, tn = tile - tm * a.tiles_n;
        const int m0 = tm * BM, n0 = tn * BN;
        double p1[4][2][2], p2[4][2][2], p3[4][2][2];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 2; j++)
#pragma unroll
                for (int e = 0; e < 2; e++) p1[i][j][e] = p2[i][j][e] = p3[i][j][e] = 0.0;

        auto chunk = [&](auto masked_t, const int kv) {
            constexpr bool masked = decltype(masked_t)::value;
            const int s = (int)(u % STAGES);
            zd_mbar_wait(bar_base + 8u * s, (u / STAGES) & 1u);
            const cplx* Ap = sA + s * Cfg::A_ELEMS + offA;
            const cplx* Bp = sB + s * Cfg::B_ELEMS + offB;
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
                if (masked && ks * 4 >= kv) break;
                const bool dead = masked && ks * 4 + t >= kv;   // this lane's k index is beyond k
                double qa[12], qb[6];
#pragma unroll
                for (int mt = 0; mt < 4; mt++) {
                    cplx y = Ap[mt * A_MT + ks * A_KS];
                    if (dead) y = make_double2(0.0, 0.0);
                    const double ai = TA ? -y.y : y.y;
                    qa[3 * mt] = y.x;
                    qa[3 * mt + 1] = ai;
                    qa[3 * mt + 2] = y.x + ai;
                }
#pragma unroll
                for (int nt = 0; nt < 2; nt++) {
                    cplx y = Bp[nt * B_NT + ks * B_KS];
                    if (dead) y = make_double2(0.0, 0.0);
                    const double bi = TB ? -y.y : y.y;
                    qb[3 * nt] = y.x;
                    qb[3 * nt + 1] = bi - y.x;
                    qb[3 * nt + 2] = y.x + bi;
                }
                // P1 = (Ar + Ai) Br,  P2 = Ar (Bi - Br),  P3 = Ai (Br + Bi);  Cr = P1 - P3,  Ci = P1 + P2
#pragma unroll
                for (int mt = 0; mt < 4; mt++)
#pragma unroll
                    for (int nt = 0; nt < 2; nt++) {
                        zd_dmma(p1[mt][nt][0], p1[mt][nt][1], qa[3 * mt + 2], qb[3 * nt]);
                        zd_dmma(p2[mt][nt][0], p2[mt][nt][1], qa[3 * mt], qb[3 * nt + 1]);
                        zd_dmma(p3[mt][nt][0], p3[mt][nt][1], qa[3 * mt + 1], qb[3 * nt + 2]);
                    }
            }
            // generic-proxy reads of the stage before the producer's next async-proxy write into it
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) zd_mbar_arrive(bar_base + 8u * (STAGES + s));
            u++;
        };
#pragma unroll 1
        for (int c = 0; c + 1 < nch; c++) chunk(std::false_type{}, KC);
        {
            const int kv = a.k - (nch - 1) * KC;
            if (kv == KC) chunk(std::false_type{}, KC); else chunk(std::true_type{}, kv);
        }

        // ---- epilogue ----
        cplx* Cb = a.C + b * a.strideC;
        double f = 0.0;
#pragma unroll
        for (int mt = 0; mt < 4; mt++)
#pragma unroll
            for (int nt = 0; nt < 2; nt++)
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int mi = m0 + wm * 32 + mt * 8 + g, ni = n0 + wn * 16 + nt * 8 + 2 * t + j;
                    if (mi >= a.m || ni >= a.n) continue;
                    const cplx x = make_double2(p1[mt][nt][j] - p3[mt][nt][j], p1[mt][nt][j] + p2[mt][nt][j]);
                    cplx* dst = Cb + mi + (long long)ni * a.ldc;
                    if (a.epi >= 2) {
                        const double dg = mi == ni ? 1.0 : 0.0;
                        const double dx = x.x - dg;
                        f = fma(dx, dx, fma(x.y, x.y, f));
                        *dst = a.epi == 3 ? x : make_double2(a.alpha * dg - a.beta * x.x, -a.beta * x.y);
                    } else {
                        cplx v = make_double2(a.alpha * x.x, a.alpha * x.y);
                        if (a.beta != 0.0) {
                            const cplx o = *dst;
                            v.x = fma(a.beta, o.x, v.x);
                            v.y = fma(a.beta, o.y, v.y);
                        }
                        *dst = v;
                        if (a.epi == 1) f = fma(x.x, x.x, fma(x.y, x.y, f));
                    }
                }
        if (a.epi) {
            f = warp_sum(f);
            if (lane == 0) a.part[item * 8 + warp] = f;
        }
    }
}

// err[0] = max over matrices of sqrt(sum of the matrix's partials) (non-negative doubles order like their bits)
__global__ void __launch_bounds__(32) zd_err_kernel(int per, const double* __restrict__ part, double* __restrict__ err) {
    const double* p = part + (long long)blockIdx.x * per;
    double s = 0.0;
    for (int i = threadIdx.x; i < per; i += 32) s += p[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicMax((unsigned long long*)err, (unsigned long long)__double_as_longlong(sqrt(s)));
}

// fro[b] = sum of the item partials of batch entry b (fixed order)
__global__ void __launch_bounds__(32) zd_fro_kernel(int per, const double* __restrict__ part, double* __restrict__ fro) {
    const double* p = part + (long long)blockIdx.x * per;
    double s = 0.0;
    for (int i = threadIdx.x; i < per; i += 32) s += p[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) fro[blockIdx.x] = s;
}

template <int WM, int WN, bool TA, bool TB>
int zd_launch(wb200_ctx* ctx, const ZdArgs& a) {
    using Cfg = ZdCfg<WM, WN, TA, TB>;
    auto kern = zgemm_dmma_kernel<WM, WN, TA, TB>;
    const unsigned bit = 1u << ((WM == 4 ? 0 : 3) + (TA ? 1 : (TB ? 2 : 0)));   // once per ctx and instantiation
    if (!(ctx->zd_attr_mask & bit)) {
        WB_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        ctx->zd_attr_mask |= bit;
    }
    const long long grid = a.items < ctx->num_sms ? a.items : ctx->num_sms;
    kern<<<(unsigned)grid, 384, Cfg::SMEM, ctx->stream>>>(a);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}

}  // namespace

bool wb_zgemm_tensor_applicable(int m, int n, int k) {
    // below this the fused small-shape kernels / the CUDA-core ZGEMM are the better tools
    return m >= 32 && n >= 16 && k >= 16 && (long long)m * n >= 1024;
}

// C[b] = alpha op(A) op(B) + beta C for b < batch.  epi / part as in ZdArgs (part sized [items][8] by the caller
// through wb_zgemm_tensor_items).
long long wb_zgemm_tensor_items(int m, int n, int batch) {
    const bool wide = m <= 64;
    const int BM = wide ? 64 : 128, BN = wide ? 64 : 32;
    return (long long)batch * ((m + BM - 1) / BM) * ((n + BN - 1) / BN);
}

int wb_zgemm_tensor(wb200_ctx* ctx, const ZgemmArgs& g, int epi, double* part) {
    if (g.batch <= 0 || g.m <= 0 || g.n <= 0 || g.k <= 0) return WB200_OK;
    if (g.transA && g.transB) {
        ctx->err = "wb_zgemm_tensor: A^H B^H is not instantiated";
        return WB200_ERR_ARG;
    }
    const bool wide = g.m <= 64;
    const int BM = wide ? 64 : 128, BN = wide ? 64 : 32;
    ZdArgs a{};
    a.m = g.m; a.n = g.n; a.k = g.k;
    a.A = g.A; a.strideA = g.strideA; a.lda = g.lda; a.idxA = g.idxA;
    a.B = g.B; a.strideB = g.strideB; a.ldb = g.ldb; a.idxB = g.idxB;
    a.C = g.C; a.strideC = g.strideC; a.ldc = g.ldc;
    a.alpha = g.alpha; a.beta = g.beta;
    a.tiles_m = (g.m + BM - 1) / BM; a.tiles_n = (g.n + BN - 1) / BN;
    a.items = (long long)g.batch * a.tiles_m * a.tiles_n;
    a.epi = epi; a.part = part;
    if (wide) {
        if (g.transA) return zd_launch<2, 4, true, false>(ctx, a);
        if (g.transB) return zd_launch<2, 4, false, true>(ctx, a);
        return zd_launch<2, 4, false, false>(ctx, a);
    }
    if (g.transA) return zd_launch<4, 2, true, false>(ctx, a);
    if (g.transB) return zd_launch<4, 2, false, true>(ctx, a);
    return zd_launch<4, 2, false, false>(ctx, a);
}

// max over the batch of sqrt(sum of partials)  -> derr[0] (must be zeroed by the caller)
int wb_zgemm_tensor_err(wb200_ctx* ctx, int m, int n, int batch, const double* part, double* derr) {
    const int per = (int)(wb_zgemm_tensor_items(m, n, batch) / batch) * 8;
    zd_err_kernel<<<batch, 32, 0, ctx->stream>>>(per, part, derr);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}
int wb_zgemm_tensor_fro(wb200_ctx* ctx, int m, int n, int batch, const double* part, double* dfro) {
    const int per = (int)(wb_zgemm_tensor_items(m, n, batch) / batch) * 8;
    zd_fro_kernel<<<batch, 32, 0, ctx->stream>>>(per, part, dfro);
    WB_LAUNCH_CHECK(ctx);
    return WB200_OK;
}
