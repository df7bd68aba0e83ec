// Microbenchmark: scalar FP32 mul/add (round-to-nearest, never fused) against the sm_100 packed f32x2 forms.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

__global__ void k_scalar(float* out, float m, float a) {
    float v[2 * CHAINS];
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; i++) v[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 2 * CHAINS; i++) v[i] = __fadd_rn(__fmul_rn(v[i], m), a);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long pk(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

__global__ void k_packed(float* out, float m, float a) {
    unsigned long long v[CHAINS];
    const unsigned long long mm = pk(m, m), aa = pk(a, a);
#pragma unroll
    for (int i = 0; i < CHAINS; i++) v[i] = pk(threadIdx.x * 0.001f + 2 * i, threadIdx.x * 0.001f + 2 * i + 1);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) v[i] = add2(mul2(v[i], mm), aa);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[i]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_fma_scalar(float* out, float m, float a) {
    float v[2 * CHAINS];
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; i++) v[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 2 * CHAINS; i++) v[i] = __fmaf_rn(v[i], m, a);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fma_packed(float* out, float m, float a) {
    unsigned long long v[CHAINS];
    const unsigned long long mm = pk(m, m), aa = pk(a, a);
#pragma unroll
    for (int i = 0; i < CHAINS; i++) v[i] = pk(threadIdx.x * 0.001f + 2 * i, threadIdx.x * 0.001f + 2 * i + 1);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) v[i] = fma2(v[i], mm, aa);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[i]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: packed fp32 math interleaved with integer ALU work (does the packed form free issue slots?)
__global__ void k_mix_scalar(float* out, float m, float a, unsigned* iout) {
    float v[2 * CHAINS];
    unsigned w[CHAINS];
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; i++) v[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) w[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 2 * CHAINS; i++) v[i] = __fadd_rn(__fmul_rn(v[i], m), a);
#pragma unroll
        for (int i = 0; i < CHAINS; i++) w[i] = (w[i] ^ (w[i] >> 3)) + 0x9E3779B9u;
    }
    float s = 0; unsigned t = 0;
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; i++) s += v[i];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) t += w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
__global__ void k_mix_packed(float* out, float m, float a, unsigned* iout) {
    unsigned long long v[CHAINS];
    unsigned w[CHAINS];
    const unsigned long long mm = pk(m, m), aa = pk(a, a);
#pragma unroll
    for (int i = 0; i < CHAINS; i++) v[i] = pk(threadIdx.x * 0.001f + 2 * i, threadIdx.x * 0.001f + 2 * i + 1);
#pragma unroll
    for (int i = 0; i < CHAINS; i++) w[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) v[i] = add2(mul2(v[i], mm), aa);
#pragma unroll
        for (int i = 0; i < CHAINS; i++) w[i] = (w[i] ^ (w[i] >> 3)) + 0x9E3779B9u;
    }
    float s = 0; unsigned t = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v[i]));
        s += lo + hi;
    }
#pragma unroll
    for (int i = 0; i < CHAINS; i++) t += w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; i++) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5;
}

int main() {
    const int blocks = 148 * 8, threads = 256;
    float* out; unsigned* iout;
    cudaMalloc(&out, blocks * threads * 4);
    cudaMalloc(&iout, blocks * threads * 4);
    const double lane_ops = (double)blocks * threads * ITERS * 2 * CHAINS;  // (mul+add) pairs or fmas per lane
    float t;
    t = timeit([&] { k_scalar<<<blocks, threads>>>(out, 0.999f, 0.5f); });
    printf("scalar mul+add : %.3f ms  %.1f G(mul+add)/s\n", t, lane_ops / t * 1e-6);
    t = timeit([&] { k_packed<<<blocks, threads>>>(out, 0.999f, 0.5f); });
    printf("packed mul+add : %.3f ms  %.1f G(mul+add)/s\n", t, lane_ops / t * 1e-6);
    t = timeit([&] { k_fma_scalar<<<blocks, threads>>>(out, 0.999f, 0.5f); });
    printf("scalar fma     : %.3f ms  %.1f Gfma/s\n", t, lane_ops / t * 1e-6);
    t = timeit([&] { k_fma_packed<<<blocks, threads>>>(out, 0.999f, 0.5f); });
    printf("packed fma     : %.3f ms  %.1f Gfma/s\n", t, lane_ops / t * 1e-6);
    t = timeit([&] { k_mix_scalar<<<blocks, threads>>>(out, 0.999f, 0.5f, iout); });
    printf("scalar mix     : %.3f ms\n", t);
    t = timeit([&] { k_mix_packed<<<blocks, threads>>>(out, 0.999f, 0.5f, iout); });
    printf("packed mix     : %.3f ms\n", t);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
