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

// packed FP32: fma.rn.f32x2 (SASS FFMA2), two FMAs per issue slot
template <int ILP>
__global__ void ffma2_kernel(float* out, int iters) {
    unsigned long long c[ILP], a, b;
    float a0 = threadIdx.x * 1e-3f, b0 = threadIdx.x * 2e-3f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a0 + 1.f));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b0 + 1.f));
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(c[i]) : "f"((float)i), "f"((float)-i));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32x2 %0, %1, %0, %2;" : "+l"(c[i]) : "l"(a), "l"(b));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(c[i])); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// dependent chain of DMMAs in one warp: cycles per instruction = latency
__global__ void dmma_latency_kernel(double* out, long long* cyc, int iters) {
    double c0 = 1.0, c1 = 2.0, a = threadIdx.x * 1e-3, b = 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        dmma(c0, c1, a, b); dmma(c0, c1, a, b); dmma(c0, c1, a, b); dmma(c0, c1, a, b);
    }
    long long t1 = clock64();
    out[threadIdx.x] = c0 + c1;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void dfma_latency_kernel(double* out, long long* cyc, int iters) {
    double c = 1.0, a = 1.0000001, b = 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) { c = fma(a, c, b); c = fma(a, c, b); c = fma(a, c, b); c = fma(a, c, b); }
    long long t1 = clock64();
    out[threadIdx.x] = c;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    {
        long long* cyc; cudaMalloc(&cyc, 8); long long h;
        dmma_latency_kernel<<<1, 32>>>(out, cyc, 10000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DMMA dependent-chain latency: %.1f cycles\n", h / 40000.0);
        dfma_latency_kernel<<<1, 32>>>(out, cyc, 10000); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA dependent-chain latency: %.1f cycles\n", h / 40000.0);
    }
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
            ffma2_kernel<8><<<sms, warps * 32>>>((float*)out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 2.0 * 64 * 8 * (double)iters * warps * sms;
            if (rep) printf("FFMA2        warps/SM=%2d  %.2f TFLOP/s\n", warps, flops / ms * 1e-9);
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
