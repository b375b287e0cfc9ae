// Micro-benchmark: FP64 FMA (DFMA) vs FP64 tensor-core (mma.sync m8n8k4 f64) peak on one GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void dmma_kernel(double* out, int iters) {
    double c0[ILP], c1[ILP];
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dfma_kernel(double* out, int iters) {
    double c[ILP];
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fma(a, c[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void ffma_kernel(float* out, int iters) {
    float c[ILP];
    float a = threadIdx.x * 1e-3f, b = threadIdx.x * 2e-3f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fmaf(a, c[i], b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            dmma_kernel<8><<<sms, warps * 32>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 2.0 * 256 * 8 * (double)iters * warps * sms;
            if (rep) printf("DMMA m8n8k4  warps/SM=%2d  %.2f TFLOP/s\n", warps, flops / ms * 1e-9);
        }
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            dfma_kernel<8><<<sms, warps * 32>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 2.0 * 32 * 8 * (double)iters * warps * sms;
            if (rep) printf("DFMA         warps/SM=%2d  %.2f TFLOP/s\n", warps, flops / ms * 1e-9);
        }
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            ffma_kernel<8><<<sms, warps * 32>>>((float*)out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 2.0 * 32 * 8 * (double)iters * warps * sms;
            if (rep) printf("FFMA         warps/SM=%2d  %.2f TFLOP/s\n", warps, flops / ms * 1e-9);
        }
    }
    return 0;
}
