// i8_ozaki_probe.cu — how fast could the rotation's arithmetic run on the int8 tensor cores?
//
// The rotation MU_kb = M_kb U_{k+b} (src/spread.jl:173) is bound by the FP64 tensor pipe (DMMA).  The only faster
// arithmetic on sm_100a is int8 MMA: an Ozaki-style emulation slices both operands into 7 signed 7-bit pieces
// (tools/ozaki_accuracy.py: 3e-13 relative error with 7 slices) and needs, per (k, b) pair, the 28 slice products
// A_i B_j with i + j <= 6, accumulated exactly in int32 per significance level l = i + j.  With the complex product
// written as ONE real product [Ar | -Ai ; ...] the shapes per pair are M = 128, N = 256 (Re | Im), K = 256, and level
// l is a single GEMM with the slices concatenated along K:
//     D_l = [A_0 | A_1 | ... | A_l] * [B_l ; B_{l-1} ; ... ; B_0]          (128 x 256(l+1)) * (256(l+1) x 256)
// This probe times exactly those 7 GEMMs per pair with cuBLAS's batched int8 kernels (int32 accumulate and output) on a
// sample of pairs with the operand reuse of the real problem (every A slice set belongs to one pair, every B slice set
// to one k-point, 12 pairs per k-point) and scales to the 49 152 pairs of the north-star evaluation.  It is an UPPER
// bound on what a hand-written tcgen05 kernel must beat for its MMA phase and a LOWER bound on the whole emulation,
// which additionally has to slice U every evaluation, drain 7 int32 accumulators per pair into FP64, scale, store MU
// and form diag(U_k^H MU).  A library call in a measurement tool, not on the product path.  Prints one JSON object.
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
#define CB(x) do { cublasStatus_t s_ = (x); if (s_ != CUBLAS_STATUS_SUCCESS) { fprintf(stderr, "%s: cublas status %d\n", #x, (int)s_); exit(1); } } while (0)

__global__ void fill_i8(int8_t* p, size_t n, unsigned seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned x = (unsigned)i * 2654435761u + seed;
        x ^= x >> 13; x *= 0x5bd1e995u; x ^= x >> 15;
        p[i] = (int8_t)((int)(x % 127u) - 63);
    }
}

int main(int argc, char** argv) {
    const int nk_s = argc > 1 ? atoi(argv[1]) : 256;   // sampled k-points
    const int nn = 12, M = 128, N = 256, K = 256, S = 7;
    const int pairs = nk_s * nn;
    const long long pairs_full = 49152;
    const size_t a_pair = (size_t)M * K * S, b_k = (size_t)N * K * S;     // bytes per pair / per k-point
    int8_t *dA, *dB;
    int32_t* dC;
    CK(cudaMalloc(&dA, a_pair * pairs));
    CK(cudaMalloc(&dB, b_k * nk_s));
    CK(cudaMalloc(&dC, sizeof(int32_t) * (size_t)M * N * S * pairs));
    fill_i8<<<1024, 256>>>(dA, a_pair * pairs, 1u);
    fill_i8<<<1024, 256>>>(dB, b_k * nk_s, 2u);
    CK(cudaDeviceSynchronize());
    // cuBLAS int8: "TN" with K-major operands.  C (col-major, ldc = N') = op(A') op(B') with the roles swapped so that
    // both operands are read K-contiguously: C^T[n, m] = sum_k Bt[k, n] At[k, m]
    //   first operand  = B slices of the k-point, stored [n][k_concat] (k fastest), op = T  -> N x Kl
    //   second operand = A slices of the pair,    stored [m][k_concat] (k fastest), op = N  -> Kl x M
    std::vector<const void*> hA(pairs), hB(pairs);
    std::vector<void*> hC(pairs);
    const void **pA, **pB;
    void** pC;
    CK(cudaMalloc(&pA, sizeof(void*) * pairs)); CK(cudaMalloc(&pB, sizeof(void*) * pairs)); CK(cudaMalloc(&pC, sizeof(void*) * pairs));
    cublasHandle_t h;
    CB(cublasCreate(&h));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int32_t one = 1, zero = 0;
    double best = 1e30;
    for (int rep = 0; rep < 4; rep++) {
        float total = 0.f;
        for (int l = 0; l < S; l++) {
            const int Kl = K * (l + 1);
            for (int p = 0; p < pairs; p++) {
                const int k = p / nn, kb = (k * 7 + (p % nn) * 13 + 1) % nk_s;    // some other k-point's gauge
                hA[p] = dA + a_pair * p;                                          // [A_0 | ... | A_l]: the first Kl columns
                hB[p] = dB + b_k * kb + (size_t)(S - 1 - l) * K;                  // reversed slice order: [B_l ; ... ; B_0]
                hC[p] = dC + ((size_t)p * S + l) * M * N;
            }
            CK(cudaMemcpy(pA, hA.data(), sizeof(void*) * pairs, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(pB, hB.data(), sizeof(void*) * pairs, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(pC, hC.data(), sizeof(void*) * pairs, cudaMemcpyHostToDevice));
            CK(cudaEventRecord(e0));
            CB(cublasGemmBatchedEx(h, CUBLAS_OP_T, CUBLAS_OP_N, N, M, Kl, &one, pB, CUDA_R_8I, K * S, pA, CUDA_R_8I, K * S, &zero,
                                   pC, CUDA_R_32I, N, pairs, CUBLAS_COMPUTE_32I, CUBLAS_GEMM_DEFAULT));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            total += ms;
        }
        if (rep > 0 && total < best) best = total;
    }
    const double ops = 2.0 * M * N * K * 28.0 * pairs;
    const double ms_full = best * (double)pairs_full / pairs;
    printf("{\"sampled_pairs\": %d, \"ms_sample\": %.4f, \"int8_tops\": %.1f, \"ms_per_evaluation_scaled\": %.3f, "
           "\"int32_output_gbytes_per_evaluation\": %.1f, "
           "\"note\": \"7 level GEMMs per pair (slices concatenated along K), cublasGemmBatchedEx int8 -> int32; MMA phase of an "
           "Ozaki emulation of the 19.1 ms FP64 rotation, without slicing U, the FP64 drains, MU store and diag epilogue\"}\n",
           pairs, best, ops / (best * 1e-3) / 1e12, ms_full, 4.0 * M * N * S * pairs_full / 1e9);
    return 0;
}
