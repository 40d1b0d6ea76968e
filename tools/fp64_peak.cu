// Dev tool (GPU box): FP64 FMA throughput of the CUDA cores (not the tensor cores), the roofline denominator SURVEY.md
// section 8(d) asks for next to the resident (shared-memory) kernels, whose HBM traffic is negligible.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_peak tools/fp64_peak.cu && /tmp/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void __launch_bounds__(256) fma_kernel(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x * 1e-3 + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true: keeps the loop alive
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double* out;
    cudaMalloc(&out, 1 << 20);
    const int iters = 20000, ILP = 8;
    const int blocks = p.multiProcessorCount * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fma_kernel<ILP><<<blocks, 256>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * (double)blocks * 256 * ILP * iters;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_fma_tflops\": %.2f, \"fma_per_clk_per_sm\": %.1f, \"sm_clock_mhz\": %.0f}\n", p.name,
           p.multiProcessorCount, best, best * 1e12 / 2.0 / p.multiProcessorCount / (clk * 1e3), clk / 1e3);
    return 0;
}
