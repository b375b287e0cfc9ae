// Micro-benchmark: what the legacy warp-level tensor-core path (mma.sync) delivers on one B200 for the operand
// types a batched 30 x 30 factorisation could use: TF32 m16n8k8, BF16 m16n8k16, FP16 m16n8k16 (fp32 accumulate),
// throughput at several occupancies plus the dependent-chain latency of the TF32 form.
// the 3-term TF32 split needs >= ~250 TFLOP/s of mma.sync TF32 to beat the FP32 SIMT kernels.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_peak mma_sync_peak.cu   (NOT yet run on a GPU)
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int KIND, int ILP>
__global__ void mma_kernel(float* out, int iters) {
    float c[ILP][4];
    unsigned a[4], b[2];
    // small finite operand bit patterns (0.5 as tf32 / packed halves close to 0.5)
    const unsigned pat = KIND == 0 ? 0x3f000000u : KIND == 1 ? 0x3f003f00u : 0x38003800u;
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = pat + (threadIdx.x & 1);
    b[0] = pat; b[1] = pat + (threadIdx.x & 2);
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 0.f; c[i][3] = 1.f; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0) mma_tf32(c[i], a, b);
            else if (KIND == 1) mma_bf16(c[i], a, b);
            else mma_f16(c[i], a, b);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void tf32_latency_kernel(float* out, long long* cyc, int iters) {
    float c[4] = {1.f, 2.f, 3.f, 4.f};
    unsigned a[4] = {0x3a000000u, 0x3a000000u, 0x3a000000u, 0x3a000000u}, b[2] = {0x3a000000u, 0x3a000000u};
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) { mma_tf32(c, a, b); mma_tf32(c, a, b); mma_tf32(c, a, b); mma_tf32(c, a, b); }
    long long t1 = clock64();
    out[threadIdx.x] = c[0] + c[1] + c[2] + c[3];
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int KIND>
static void sweep(const char* name, double fma_per_mma, float* out, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            mma_kernel<KIND, 8><<<sms, warps * 32>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
        }
        const double flops = 2.0 * fma_per_mma * 8 * (double)iters * warps * sms;
        printf("%-22s warps/SM=%2d  %8.1f TFLOP/s\n", name, warps, flops / ms * 1e-9);
    }
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    float* out; cudaMalloc(&out, sizeof(float) * sms * 1024);
    long long* cyc; cudaMalloc(&cyc, 8); long long h = 0;
    tf32_latency_kernel<<<1, 32>>>(out, cyc, 10000);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%s, %d SMs\nmma.sync m16n8k8 tf32 dependent-chain latency: %.1f cycles\n", p.name, sms, h / 40000.0);
    sweep<0>("mma.sync m16n8k8 tf32", 16.0 * 8 * 8, out, sms);
    sweep<1>("mma.sync m16n8k16 bf16", 16.0 * 8 * 16, out, sms);
    sweep<2>("mma.sync m16n8k16 f16", 16.0 * 8 * 16, out, sms);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    return 0;
}
