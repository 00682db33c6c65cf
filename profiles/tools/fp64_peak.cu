// Micro-benchmark: FP64 FMA pipe throughput (independent DFMA chains) and DMMA m8n8k4 throughput on the
// current GPU.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double *out, int iters) {
    double a[8];
    for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 1e-3 + k;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = fma(a[k], m, c);
    }
    double s = 0;
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double *out, int iters) {
    double c[8][2];
    for (int k = 0; k < 8; ++k) c[k][0] = c[k][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int warps = 4; warps <= 32; warps *= 2) {
        const int iters = 20000, blocks = sms * 2, threads = warps * 32;
        for (int which = 0; which < 2; ++which) {
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                if (which == 0) dfma_kernel<<<blocks, threads>>>(out, iters);
                else dmma_kernel<<<blocks, threads>>>(out, iters);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            const double flop = which == 0 ? 2.0 * 8 * iters * (double)blocks * threads
                                           : 2.0 * 256 * 8 * iters * (double)blocks * warps;
            printf("%s warps/CTA=%2d (2 CTAs/SM): %.3f ms  %.2f TFLOP/s\n", which == 0 ? "DFMA" : "DMMA", warps, best,
                   flop / best / 1e9);
        }
    }
    printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
