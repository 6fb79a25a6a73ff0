// fp64_peak.cu — measures the FP64 denominators of the roofline on this B200 (MEASURED_PEAKS.json
// has only HBM and bf16): CUDA-core DFMA peak, tensor-pipe DMMA.8x8x4 peak, and cuBLAS ZGEMM
// (one large, and strided-batched at the north-star tile shape) as the library reference point.
// Prints one JSON object.  Timing: CUDA events, 3 warm-ups, best of 5.
#include <cublas_v2.h>
#include <cuComplex.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(err_)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a0, double b0) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
    double a = a0 + threadIdx.x * 1e-12, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float best_ms(F f) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) f();
    float best = 1e30f;
    for (int i = 0; i < 5; i++) {
        CK(cudaEventRecord(e0));
        f();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256));
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    for (int bps = 1; bps <= 8; bps *= 2) {  // blocks of 256 threads per SM -> 8..64 warps/SM
        const int iters = 20000;
        float ms = best_ms([&] { dfma_kernel<<<sms * bps, 256>>>(out, iters, 1.0000001, 1e-9); });
        double tf = 2.0 * 16 * iters * 256.0 * sms * bps / (ms * 1e-3) / 1e12;
        printf(", \"dfma_tflops_%dwarps\": %.2f", bps * 8, tf);
    }
    for (int bps = 1; bps <= 4; bps *= 2) {
        const int iters = 4000;
        float ms = best_ms([&] { dmma_kernel<16><<<sms * bps, 256>>>(out, iters, 1.0000001, 1e-3); });
        // one DMMA.8x8x4 = 256 FMA = 512 flop per warp
        double tf = 512.0 * 16 * iters * 8.0 * sms * bps / (ms * 1e-3) / 1e12;
        printf(", \"dmma_tflops_%dwarps_16acc\": %.2f", bps * 8, tf);
    }
    {   // one warp per SMSP: can a single warp saturate the pipe?
        const int iters = 4000;
        float ms = best_ms([&] { dmma_kernel<16><<<sms, 128>>>(out, iters, 1.0000001, 1e-3); });
        printf(", \"dmma_tflops_4warps_16acc\": %.2f", 512.0 * 16 * iters * 4.0 * sms / (ms * 1e-3) / 1e12);
        ms = best_ms([&] { dmma_kernel<24><<<sms, 128>>>(out, iters, 1.0000001, 1e-3); });
        printf(", \"dmma_tflops_4warps_24acc\": %.2f", 512.0 * 24 * iters * 4.0 * sms / (ms * 1e-3) / 1e12);
        ms = best_ms([&] { dmma_kernel<24><<<sms, 384>>>(out, iters, 1.0000001, 1e-3); });
        printf(", \"dmma_tflops_12warps_24acc\": %.2f", 512.0 * 24 * iters * 12.0 * sms / (ms * 1e-3) / 1e12);
    }
    {
        const int iters = 4000;
        float ms = best_ms([&] { dmma_kernel<4><<<sms, 256>>>(out, iters, 1.0000001, 1e-3); });
        printf(", \"dmma_tflops_8warps_4acc\": %.2f", 512.0 * 4 * iters * 8.0 * sms / (ms * 1e-3) / 1e12);
        ms = best_ms([&] { dmma_kernel<1><<<sms, 32>>>(out, iters, 1.0000001, 1e-3); });
        // dependent chain on one warp per SM: latency of one DMMA in ns
        printf(", \"dmma_dependent_latency_ns\": %.2f", ms * 1e6 / iters);
    }
    cublasHandle_t h; cublasCreate(&h);
    {
        const int n = 4096;
        cuDoubleComplex *A, *B, *C;
        CK(cudaMalloc(&A, sizeof(cuDoubleComplex) * n * n)); CK(cudaMalloc(&B, sizeof(cuDoubleComplex) * n * n));
        CK(cudaMalloc(&C, sizeof(cuDoubleComplex) * n * n));
        CK(cudaMemset(A, 0, sizeof(cuDoubleComplex) * n * n)); CK(cudaMemset(B, 0, sizeof(cuDoubleComplex) * n * n));
        cuDoubleComplex one = make_cuDoubleComplex(1, 0), zero = make_cuDoubleComplex(0, 0);
        float ms = best_ms([&] { cublasZgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); });
        printf(", \"cublas_zgemm_4096_tflops\": %.2f", 8.0 * n * n * (double)n / (ms * 1e-3) / 1e12);
        double *dA = (double*)A, *dB = (double*)B, *dC = (double*)C; double done = 1, dzero = 0;
        ms = best_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &done, dA, n, dB, n, &dzero, dC, n); });
        printf(", \"cublas_dgemm_4096_tflops\": %.2f", 2.0 * n * n * (double)n / (ms * 1e-3) / 1e12);
        cudaFree(A); cudaFree(B); cudaFree(C);
    }
    {
        const int n = 128, batch = 8192;
        cuDoubleComplex *A, *B, *C;
        size_t e = (size_t)n * n * batch;
        CK(cudaMalloc(&A, sizeof(cuDoubleComplex) * e)); CK(cudaMalloc(&B, sizeof(cuDoubleComplex) * e));
        CK(cudaMalloc(&C, sizeof(cuDoubleComplex) * e));
        CK(cudaMemset(A, 0, sizeof(cuDoubleComplex) * e)); CK(cudaMemset(B, 0, sizeof(cuDoubleComplex) * e));
        cuDoubleComplex one = make_cuDoubleComplex(1, 0), zero = make_cuDoubleComplex(0, 0);
        float ms = best_ms([&] {
            cublasZgemmStridedBatched(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, (long long)n * n, B, n,
                                      (long long)n * n, &zero, C, n, (long long)n * n, batch);
        });
        printf(", \"cublas_zgemm_batched_128x8192_tflops\": %.2f", 8.0 * n * n * (double)n * batch / (ms * 1e-3) / 1e12);
        cudaFree(A); cudaFree(B); cudaFree(C);
    }
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf(", \"sm_clock_khz_max\": %d}\n", clk);
    return 0;
}
