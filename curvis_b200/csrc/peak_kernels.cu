// peak_kernels.cu — FMA-only micro-kernels: the measured ALU roofline denominators.
// MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only; the geodesic integrator is
// bound by the fp64 (parity mode) or fp32 (fast mode) FMA pipes, so bench.py measures those
// pipes on the same box, in the same process, with the same clocks.
#include <cuda_runtime.h>
#include "launch.h"

namespace curvis {

template <typename T>
__global__ void __launch_bounds__(256) fma_chain(T* out, int iters, T a, T b) {
    // 8 independent dependency chains per thread keep the pipe full at any occupancy.
    T x0 = (T)threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

template <typename T>
static cudaError_t time_fma(int sm_count, cudaStream_t stream, int iters, double* tflops) {
    const int blocks = sm_count * 8, threads = 256;
    T* d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(T) * (size_t)blocks * threads);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {  // first rep is the warm-up
        cudaEventRecord(e0, stream);
        fma_chain<T><<<blocks, threads, 0, stream>>>(d, iters, (T)0.999, (T)0.001);
        cudaEventRecord(e1, stream);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    if (e != cudaSuccess) return e;
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    *tflops = flops / ((double)best * 1e-3) / 1e12;
    return cudaGetLastError();
}

cudaError_t measure_fma_peak(int sm_count, cudaStream_t stream, double* fp64_tflops, double* fp32_tflops) {
    cudaError_t e = time_fma<double>(sm_count, stream, 4096, fp64_tflops);
    if (e != cudaSuccess) return e;
    return time_fma<float>(sm_count, stream, 8192, fp32_tflops);
}

}  // namespace curvis
