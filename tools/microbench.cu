// Dev microbenchmarks (GPU box): how the sm_100a fp64 pipe shares issue slots with the integer pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NINT, int MODE>
__global__ void __launch_bounds__(128) mix(double* out, int iters, double a, double b, unsigned m) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    double y0 = a, y1 = a + 1, y2 = a + 2, y3 = a + 3, z0 = b, z1 = b * 2, z2 = b * 3, z3 = b * 4;
    unsigned i0 = threadIdx.x, i1 = i0 * 3, i2 = i0 * 5, i3 = i0 * 7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE == 0) {  // 2 of 3 operands shared (a, b)
                x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
                x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
            } else if (MODE == 1) {  // three distinct register operands
                x0 = fma(y0, z0, x0); x1 = fma(y1, z1, x1); x2 = fma(y2, z2, x2); x3 = fma(y3, z3, x3);
                x4 = fma(y0, z1, x4); x5 = fma(y1, z2, x5); x6 = fma(y2, z3, x6); x7 = fma(y3, z0, x7);
            } else {  // one dependent chain per thread, ILP 1 (latency bound unless enough warps)
                x0 = fma(x0, a, b); x0 = fma(x0, a, b); x0 = fma(x0, a, b); x0 = fma(x0, a, b);
                x0 = fma(x0, a, b); x0 = fma(x0, a, b); x0 = fma(x0, a, b); x0 = fma(x0, a, b);
            }
#pragma unroll
            for (int k = 0; k < NINT; ++k) {
                if ((k & 3) == 0) i0 = (i0 ^ m) + 0x9e3779b9u;
                else if ((k & 3) == 1) i1 = (i1 ^ m) + 0x7f4a7c15u;
                else if ((k & 3) == 2) i2 = (i2 ^ m) + 0x85ebca6bu;
                else i3 = (i3 ^ m) + 0xc2b2ae35u;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7)) + (double)(i0 ^ i1 ^ i2 ^ i3) + y0 + z0;
}

template <int NINT, int MODE>
void run(const char* name, int blocks_per_sm) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * blocks_per_sm, threads = 128, iters = 20000;
    double* d;
    cudaMalloc(&d, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        mix<NINT, MODE><<<blocks, threads>>>(d, iters, 0.999, 0.001, 0x55aa);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    // cycles per warp-level DFMA per SMSP, assuming 1.965 GHz
    const double warps_per_smsp = blocks_per_sm * (threads / 32) / 4.0;
    const double dfma_per_warp = 32.0 * iters;
    const double cyc = best * 1e-3 * 1.965e9 / (dfma_per_warp * warps_per_smsp);
    printf("%-34s warps/SMSP=%4.1f  ms=%8.3f  cycles per warp-DFMA per SMSP=%.3f  (int ops per DFMA: %.2f)\n", name, warps_per_smsp, best, cyc,
           NINT * 2 / 8.0);
    cudaFree(d);
}

int main() {
    run<0, 0>("dfma shared-operands", 8);
    run<0, 1>("dfma 3 distinct regs", 8);
    run<4, 0>("dfma + 1.0 int/dfma", 8);
    run<8, 0>("dfma + 2.0 int/dfma", 8);
    run<16, 0>("dfma + 4.0 int/dfma", 8);
    run<4, 1>("dfma3 + 1.0 int/dfma", 8);
    run<8, 1>("dfma3 + 2.0 int/dfma", 8);
    run<0, 2>("dfma serial chain, 8 warps/SMSP", 8);
    run<0, 2>("dfma serial chain, 4 warps/SMSP", 4);
    run<0, 2>("dfma serial chain, 2 warps/SMSP", 2);
    run<0, 2>("dfma serial chain, 1 warp/SMSP", 1);
    run<0, 0>("dfma ILP8, 1 warp/SMSP", 1);
    return 0;
}
